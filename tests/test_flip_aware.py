"""The mask-flip-aware comparison itself (tests/parity.py: flip_aware_compare), checked on the CPU oracle alone.

A 'product' result is fabricated from the float64 oracle by flipping chosen ReLU / LeakyReLU mask elements whose
pre-activation sits next to the kink: the comparison must find exactly those flips and hold 1e-4 on everything else,
and it must still reject an error that no flip explains."""
import pytest
import torch

import parity
from oracle import ffc_ref as R


def _setup(seed=0):
    torch.manual_seed(seed)
    cfg = R.FFCConfig(32, 16, 4, .25, .25, 2, 1, norm="bn", act="leaky_relu", upsampling=True)
    import fastfourierconvolution_b200 as ffc
    import torch.nn as nn
    mod = ffc.FFC_BN_ACT(32, 16, 4, .25, .25, 2, 1, upsampling=True, norm_layer=nn.BatchNorm2d, activation_layer=nn.LeakyReLU)
    sd = {k: v.detach().clone() for k, v in mod.state_dict().items()}
    xs = [torch.randn(2, 24, 8, 8), torch.randn(2, 8, 8, 8)]
    cots = [torch.randn(2, 12, 16, 16), torch.randn(2, 4, 16, 16)]

    def oracle_run(overrides):
        P = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone().double()
                 if v.is_floating_point() else v.clone()) for k, v in sd.items()}
        xd = [x.double().requires_grad_(True) for x in xs]
        with R.ActTape(overrides) as tape:
            out = R.ffc_bn_act(tuple(xd), P, "", cfg, True)
        sum((o * c.double()).sum() for o, c in zip(out, cots)).backward()
        res = {f"out{i}": o.detach() for i, o in enumerate(out)}
        res.update({f"din{i}": x.grad for i, x in enumerate(xd)})
        res.update({"grad/" + k: v.grad for k, v in P.items() if v.requires_grad and v.grad is not None})
        return res, tape
    return oracle_run


def _nearest_to_kink(tape, site, n):
    a = tape.pre[site].abs().reshape(-1)
    return torch.argsort(a)[:n].tolist()


def test_detects_exactly_the_flipped_elements():
    oracle_run = _setup()
    res0, tape = oracle_run({})
    assert len(tape.pre) >= 4          # SE ReLU, bn1 ReLU, FU ReLU, outer LeakyReLU x2
    # flip the pre-activation closest to zero of the two outer LeakyReLU sites in the "product".  (A real flip has a
    # pre-activation of ~0 and leaves every forward value alone; on these small planes no element is that close, so the
    # fabricated flips sit at the last activations, where nothing downstream depends on the forward value either.)
    n = len(tape.pre)
    overrides = {}
    for site in (n - 2, n - 1):
        j = _nearest_to_kink(tape, site, 1)[0]
        m = (tape.pre[site] > 0).clone()
        m.view(-1)[j] = ~m.view(-1)[j]
        overrides[site] = m
    fake, _ = oracle_run(overrides)
    # an FP32 "product" result with two flipped mask elements; forward values from the unflipped run (in a real flip the
    # pre-activation is ~0, so the forward value does not notice -- these small planes have no element that close)
    fake = {k: (res0[k] if k.startswith("out") else v).float() for k, v in fake.items()}
    plain = {k: parity.relerr(fake[k], res0[k]) for k in res0 if k.startswith(("din", "grad/"))}
    assert max(plain.values()) > 1e-4, "the fabricated flips must matter, otherwise this test shows nothing"
    errs, flips = parity.flip_aware_compare(fake, oracle_run, margins=(5.0e-2,), max_flips=12, what="fabricated flips")
    assert flips == 2 and max(errs.values()) < 1e-4, (flips, errs)


def test_rejects_an_error_that_is_not_a_flip():
    oracle_run = _setup(1)
    res0, _ = oracle_run({})
    fake = {k: v.float().clone() for k, v in res0.items()}
    fake["grad/ffc.convl2l.weight"] *= 1.001               # a 1e-3 relative error in one gradient
    with pytest.raises(AssertionError, match="no activation-mask flip explains"):
        parity.flip_aware_compare(fake, oracle_run, margins=(1e-2,), max_flips=8, what="scaled gradient")


def test_clean_result_needs_no_flips():
    oracle_run = _setup(2)
    res0, _ = oracle_run({})
    errs, flips = parity.flip_aware_compare({k: v.float() for k, v in res0.items()}, oracle_run)
    assert flips == 0 and max(errs.values()) < 1e-6
