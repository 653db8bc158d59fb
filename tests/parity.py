"""Shared helpers of the parity tests: fixtures, module/oracle runners, the error metric.

Metric: max|got - ref| / max|ref| per tensor.  Tolerance for the FP32 path is
1e-4 (BASELINE.json north_star).  A conv bias that feeds a BatchNorm has an exactly-zero true
gradient, so bias gradients are measured against the magnitude of the sibling weight gradient.
"""
import os

import numpy as np
import torch

import cases                      # tests/golden/cases.py
from oracle import ffc_ref as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4                        # north_star: FP32 path within 1e-4 max relative error


def load_fixture(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


def relerr(got, ref, floor=0.0):
    got = torch.as_tensor(got).double().cpu()
    ref = torch.as_tensor(ref).double().cpu()
    if ref.numel() == 0:
        return 0.0
    assert got.shape == ref.shape, (got.shape, ref.shape)
    den = max(ref.abs().max().item(), floor, 1e-30)
    return (got - ref).abs().max().item() / den


def fixture_state(fx, dtype=torch.float32):
    sd = {}
    for k, v in fx.items():
        if k.startswith("sd/"):
            t = torch.from_numpy(v.copy())
            sd[k[3:]] = t.to(dtype) if t.is_floating_point() else t
    return sd


def fixture_inputs(fx, dtype=torch.float32):
    xs = []
    i = 0
    while f"in{i}" in fx:
        xs.append(torch.from_numpy(fx[f"in{i}"].copy()).to(dtype))
        i += 1
    return xs


def _collect(outs, xs, named_params, named_buffers):
    res = {}
    for i, o in enumerate(outs):
        res[f"out{i}"] = o.detach()
    for i, x in enumerate(xs):
        res[f"din{i}"] = x.grad
    for k, p in named_params:
        if p.grad is not None:
            res["grad/" + k] = p.grad
    for k, b in named_buffers:
        res["post/" + k] = b.detach()
    return res


def run_module(mod, fx, device="cpu"):
    """Loads the fixture's state into ``mod`` (a fastfourierconvolution_b200 module), runs forward and
    backward with the fixture's inputs / cotangents on ``device``."""
    mod.load_state_dict(fixture_state(fx), strict=True)
    mod.to(device)
    mod.train(bool(fx["train"]))
    xs = [x.to(device).requires_grad_(True) for x in fixture_inputs(fx)]
    out = mod(xs[0] if len(xs) == 1 else tuple(xs))
    outs = [o for o in (out if isinstance(out, tuple) else (out,)) if torch.is_tensor(o)]
    loss = sum((o * torch.from_numpy(fx[f"cot{i}"]).to(device)).sum() for i, o in enumerate(outs))
    loss.backward()
    return _collect(outs, xs, mod.named_parameters(), mod.named_buffers())


def run_oracle(name, fx, dtype=torch.float64):
    """Runs oracle/ffc_ref.py on the fixture (CPU)."""
    oracle = cases.CASES[name][3](R)
    P = {}
    for k, v in fixture_state(fx, dtype).items():
        P[k] = v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(
            ("running_mean", "running_var", "weight_u", "weight_v")) else v.clone()
    xs = [x.requires_grad_(True) for x in fixture_inputs(fx, dtype)]
    out = oracle(P, xs, bool(fx["train"]))
    outs = [o for o in (out if isinstance(out, tuple) else (out,)) if torch.is_tensor(o)]
    loss = sum((o * torch.from_numpy(fx[f"cot{i}"]).to(dtype)).sum() for i, o in enumerate(outs))
    loss.backward()
    params = [(k, v) for k, v in P.items() if v.requires_grad]
    buffers = [(k, v) for k, v in P.items() if not v.requires_grad]
    return _collect(outs, xs, params, buffers)


def compare(got, ref, tol=TOL, what=""):
    """ref: dict of reference arrays (fixture or oracle result).  Returns {key: err}; asserts tol."""
    errs = {}
    for k, r in ref.items():
        if not (k.startswith(("out", "din", "grad/", "post/"))):
            continue
        r = torch.as_tensor(r)
        if not r.is_floating_point():
            assert k in got and torch.equal(torch.as_tensor(got[k]).cpu(), r), f"{what}: {k} differs"
            continue
        assert k in got, f"{what}: {k} missing (have {sorted(got)})"
        floor = 0.0
        if k.startswith("grad/") and k.endswith("bias"):
            sib = k[:-4] + "weight"
            sib = sib if sib in ref else k[:-4] + "weight_orig"
            if sib in ref:
                floor = float(np.abs(np.asarray(torch.as_tensor(ref[sib]).double())).max())
        errs[k] = relerr(got[k], r, floor)
    for k in got:
        if k.startswith("grad/"):
            assert k in ref, f"{what}: unexpected gradient for {k}"
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, f"{what}: above tolerance {tol}: {bad}"
    return errs
