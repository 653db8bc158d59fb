"""Host-logic tests: the product's nn.Modules and autograd Functions, executed on the host
*emulation* build of the CUDA sources (tests/emu_backend.py), against the reference's golden
outputs.  This checks the Python wiring and every kernel's index arithmetic without a GPU; the
GPU parity tests proper are tests/test_gpu_parity.py."""
import pytest
import torch
import torch.nn as nn

import cases
import parity
import fastfourierconvolution_b200 as ffc
from fastfourierconvolution_b200 import harness as H
from oracle import ffc_ref as R


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_module_matches_reference_golden(name, emu):
    fx = parity.load_fixture(name)
    mod = cases.CASES[name][0](ffc.layers)
    got = parity.run_module(mod, fx, "cpu")
    parity.compare(got, fx, tol=parity.TOL, what=name)


def _model_case(fn):
    return {"fgan32": lambda: H.FGenerator(128, 4, "fgan32"), "fd": lambda: H.FDiscriminator(True, 4),
            "cfg1": lambda: H.FFCGenerator(100, 1, 32), "d32": lambda: H.SNDiscriminator(True, 4, 7),
            "fgan64": lambda: H.FGenerator(128, 4, "fgan64"), "fd64": lambda: H.FDiscriminatorSN64(True, 4),
            "fgan128": lambda: H.FGenerator(128, 4, "fgan128")}[fn]()


# the float64 oracle of each model fixture: f(x, P) -> output (training mode)
_MODEL_ORACLE = {"fgan32": lambda x, P: R.fgenerator(x, P, True, "fgan32"), "fgan64": lambda x, P: R.fgenerator(x, P, True, "fgan64"),
                 "fgan128": lambda x, P: R.fgenerator(x, P, True, "fgan128"), "fd": lambda x, P: R.sngan_fdiscriminator(x, P, True),
                 "cfg1": lambda x, P: R.ffc_generator(x, P, True), "d32": lambda x, P: R.sn_discriminator(x, P, True, 7),
                 "fd64": lambda x, P: R.fdiscriminator_sn64(x, P, True)}


def run_model_fixture(name, fn, device):
    """Runs the product's model on the fixture's weights / input / cotangent.  Returns (got, fixture, oracle_run) where
    ``got`` holds out0, din0, grad/<key>, post/<key> for the keys the fixture stores and ``oracle_run(overrides)`` is the
    float64 oracle on the same data (for parity.flip_aware_compare)."""
    fx = parity.load_fixture(name)
    mod = _model_case(fn)
    sd = mod.state_dict()
    R.deterministic_fill(sd, int(fx["seed"]))
    for k in sd:
        if "_noise" in k:
            sd[k].zero_()
    mod.load_state_dict(sd)
    sd = {k: v.detach().clone() for k, v in mod.state_dict().items()}
    mod.to(device).train(True)
    x = torch.from_numpy(fx["in0"]).to(device).requires_grad_(True)
    out = mod(x)
    (out * torch.from_numpy(fx["cot0"]).to(device)).sum().backward()
    got = {"out0": out.detach(), "din0": x.grad}
    params, bufs = dict(mod.named_parameters()), dict(mod.named_buffers())
    for k in fx:
        if k.startswith("grad/"):
            got[k] = params[k[5:]].grad
        if k.startswith("post/"):
            got[k] = bufs[k[5:]]

    def oracle_run(overrides, dtype=torch.float64):
        P = {}
        for k, v in sd.items():
            leaf = v.is_floating_point() and not k.endswith(("running_mean", "running_var", "weight_u", "weight_v"))
            P[k] = v.clone().to(dtype).requires_grad_(True) if leaf else (v.clone().to(dtype) if v.is_floating_point() else v.clone())
        xd = torch.from_numpy(fx["in0"].copy()).to(dtype).requires_grad_(True)
        with R.ActTape(overrides) as tape:
            ref = _MODEL_ORACLE[fn](xd, P)
        (ref * torch.from_numpy(fx["cot0"]).to(dtype)).sum().backward()
        res = {"out0": ref.detach(), "din0": xd.grad}
        for k in fx:
            if k.startswith("grad/"):
                res[k] = P[k[5:]].grad
            if k.startswith("post/"):
                res[k] = P[k[5:]]
        return res, tape

    return got, fx, oracle_run


@pytest.mark.parametrize("name,fn", [("model_fgan32_G", "fgan32"), ("model_sngan_FD", "fd"), ("model_ffcgen_cfg1", "cfg1"),
                                     ("model_fgan32_D", "d32"), ("model_fgan64_G", "fgan64"),
                                     ("model_fgan64_FD", "fd64")])
def test_model_matches_reference_golden(name, fn, emu):
    """Host emulation of the kernels (index arithmetic, wiring) on the whole-model fixtures: 2e-4 in the max norm against
    the reference's FP32 run on every stored tensor (two FP32 evaluations), with activation-mask flips accounted for."""
    got, fx, oracle_run = run_model_fixture(name, fn, "cpu")
    errs, flips = parity.flip_aware_compare(got, oracle_run, ref={k: torch.from_numpy(v) for k, v in fx.items()}, tol=2e-4, what=name)
    assert max(errs.values()) < 2e-4, errs


def test_cpu_tensor_is_rejected_without_emulation():
    """No CPU fallback: the product path refuses CPU tensors."""
    m = ffc.FourierUnitSN(2, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(1, 2, 8, 8))


def test_unsupported_configs_raise():
    with pytest.raises(NotImplementedError):
        ffc.FFC(8, 8, 3, .5, .5, groups=2)
    with pytest.raises(NotImplementedError):
        ffc.FourierUnitSN(4, 4, groups=2)
    with pytest.raises(AssertionError):
        ffc.FFC(8, 8, 3, .5, .5, stride=3)


def test_eval_generator_emits_uint8(emu):
    g = H.FGenerator(128, 4, "fgan32").eval()
    with torch.no_grad():
        out = g(torch.randn(2, 128))
    assert out.dtype == torch.uint8 and out.shape == (2, 3, 32, 32)


def test_generator_forward_advances_every_batchnorm_counter_once(emu):
    """torch's BatchNorm bookkeeping (num_batches_tracked += 1 per training forward) through layers/_util.batched_bn_counters:
    every BatchNorm the forward runs advances by exactly one per forward, the constructed-but-unused lfu holders
    (spectral_transform.py:65-67 vs :94-105) stay at zero as in the reference, eval forwards leave all of them alone."""
    torch.manual_seed(0)
    g = H.FGenerator(128, 4, "fgan32").train()
    z = torch.randn(2, 128)
    g(z); g(z)
    counts = {k: int(v) for k, v in g.state_dict().items() if k.endswith("num_batches_tracked")}
    assert counts and all(v == (0 if ".lfu." in k else 2) for k, v in counts.items()), counts
    g.eval()
    with torch.no_grad():
        g(z)
    after = {k: int(v) for k, v in g.state_dict().items() if k.endswith("num_batches_tracked")}
    assert after == counts
    # a layer used outside the generator still counts for itself
    m = ffc.FourierUnitSN(4, 4).train()
    m(torch.randn(2, 4, 8, 8))
    assert int(m.bn.num_batches_tracked) == 1


def test_requires_grad_toggling_like_the_training_loop(emu):
    """fgan_complete.py:368-381 flips requires_grad on G and D every half step."""
    torch.manual_seed(0)
    m = ffc.FFC_BN_ACT(16, 8, 4, .25, .25, 2, 1, upsampling=True, norm_layer=torch.nn.BatchNorm2d,
                       activation_layer=torch.nn.GELU)
    x = (torch.randn(2, 12, 4, 4), torch.randn(2, 4, 4, 4))
    m.requires_grad_(False)
    out = m(x)
    assert not out[0].requires_grad and not out[1].requires_grad
    m.requires_grad_(True)
    out = m(x)
    (out[0].sum() + out[1].sum()).backward()
    grads = {k: p.grad is not None for k, p in m.named_parameters()}
    assert not any(v for k, v in grads.items() if ".lfu." in k)          # unused branch: no gradient
    assert all(v for k, v in grads.items() if ".lfu." not in k)


def sn_discriminator_flip_aware(device, n_convs, size, batch, seed=0, tol=parity.TOL):
    """The plain SN conv discriminator on the product's kernels against the float64 oracle (R.sn_discriminator: the
    reference's nn.Conv2d + LeakyReLU(0.1) arithmetic through torch.nn.utils.spectral_norm, fgan_complete.py:142-171) on
    identical weights and spectral-norm state: output, input gradient, every parameter gradient (weight_orig, bias) and the
    advanced power-iteration vectors in the MAX norm at ``tol``, LeakyReLU elements that land on the other side of the kink
    detected and accounted for (parity.flip_aware_compare).  Returns (errs, flips)."""
    torch.manual_seed(seed)
    ours = H.SNDiscriminator(True, 4, n_convs, backend="ffc_b200").train()
    ours.apply(H.weights_init)
    with torch.no_grad():
        for p in ours.parameters():              # N(0, 0.02) weights give ~1e-9 activations after 7 layers: rescale
            if p.dim() > 1:
                p.mul_(8.0)
            else:
                p.normal_(0, 0.1)
    sd = {k: v.detach().clone() for k, v in ours.state_dict().items()}
    x = torch.rand(batch, 3, size, size) * 2 - 1
    cot = torch.randn(batch, 1)

    def oracle_run(overrides, dtype=torch.float64):
        P = {k: (v.clone().to(dtype).requires_grad_(True) if not k.endswith(("weight_u", "weight_v")) else v.clone().to(dtype)) for k, v in sd.items()}
        xd = x.clone().to(dtype).requires_grad_(True)
        with R.ActTape(overrides) as tape:
            out = R.sn_discriminator(xd, P, True, n_convs)
        (out * cot.to(dtype)).sum().backward()
        res = {"out0": out.detach(), "din0": xd.grad}
        for k, v in P.items():
            if not v.requires_grad:
                res["post/" + k] = v
            elif v.grad is not None:
                res["grad/" + k] = v.grad
        return res, tape

    ours.to(device)
    xa = x.clone().to(device).requires_grad_(True)
    oa = ours(xa)
    (oa * cot.to(device)).sum().backward()
    got = {"out0": oa.detach(), "din0": xa.grad}
    got.update({"grad/" + k: p.grad for k, p in ours.named_parameters()})
    got.update({"post/" + k: b.detach() for k, b in ours.named_buffers()})
    cache = {}
    ref32, _ = oracle_run({}, torch.float32)
    noise, _, _ = parity.flip_aware_errors(ref32, oracle_run, cache=cache)
    return parity.flip_aware_compare(got, oracle_run, tol=tol, noise_floor=noise, cache=cache,
                                     what=f"SNDiscriminator({n_convs} convs, {size}x{size}, batch {batch}; reference FP32 itself {max(noise.values()):.1e})")


def test_sn_discriminator_on_product_kernels_matches_oracle(emu):
    errs, flips = sn_discriminator_flip_aware("cpu", 7, 32, 2)
    assert max(errs.values()) < parity.TOL, errs


# (B, cin, cout, H, k, stride): the SN discriminator stages of fgan / fgan64 / fgan128 (fgan_complete.py:150-166) at small batch
CONV_ACT_CASES = [(4, 3, 64, 32, 3, 1), (2, 64, 64, 32, 4, 2), (2, 64, 128, 16, 3, 1), (4, 128, 128, 16, 4, 2),
                  (4, 256, 512, 4, 3, 1), (3, 512, 512, 4, 4, 2)]


def conv2d_act_errs(device, case, seed=0):
    """LeakyReLU(0.1)(conv(x, w) + b) and its three gradients against float64, max norm.  Output elements whose float64
    pre-activation lies within 1e-4 * max of the kink get a zero cotangent, so a sign decided by rounding cannot matter."""
    import torch.nn.functional as F
    from fastfourierconvolution_b200 import ops
    B, cin, cout, Hs, k, stride = case
    torch.manual_seed(seed)
    x = torch.randn(B, cin, Hs, Hs)
    w = torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5
    b = torch.randn(cout) * 0.3
    xr, wr, br = (t.double().requires_grad_(True) for t in (x, w, b))
    pre = F.conv2d(xr, wr, br, stride=stride, padding=1)
    ref = F.leaky_relu(pre, 0.1)
    cot = torch.randn(ref.shape, dtype=torch.float64)
    cot[pre.detach().abs() < 1e-4 * pre.detach().abs().max()] = 0
    (ref * cot).sum().backward()
    xo, wo, bo = (t.to(device).requires_grad_(True) for t in (x, w, b))
    out = ops.conv2d_act(xo, wo, bo, stride, 1, ops.ACT_LEAKY, 0.1)
    (out * cot.float().to(device)).sum().backward()
    return {"out": parity.relerr(out, ref.detach()), "dx": parity.relerr(xo.grad, xr.grad),
            "dw": parity.relerr(wo.grad, wr.grad), "db": parity.relerr(bo.grad, br.grad)}


@pytest.mark.parametrize("case", CONV_ACT_CASES[:4])
def test_conv2d_act_matches_float64(case, emu):
    errs = conv2d_act_errs("cpu", case)
    assert max(errs.values()) < parity.TOL, errs


def spectral_norm_errs(device, shape, training, seed=0, transposed=False):
    """ops.spectral_norm_weight (three kernels) against torch.nn.utils.spectral_norm's own hook in float64: the weight,
    its gradient with respect to weight_orig, and the power-iteration vectors left in the buffers.  ``transposed``: an
    nn.ConvTranspose2d holder (SpectralNorm.dim == 1)."""
    import copy
    from fastfourierconvolution_b200 import ops
    torch.manual_seed(seed)
    cls = torch.nn.ConvTranspose2d if transposed else torch.nn.Conv2d
    conv = torch.nn.utils.spectral_norm(cls(shape[1], shape[0], shape[2], bias=False))
    conv.train(training)
    ref = copy.deepcopy(conv).double()
    conv.to(device)
    u, v = conv.weight_u.detach().clone(), conv.weight_v.detach().clone()
    w = ops.spectral_norm_weight(conv.weight_orig, u, v, training, 1e-12, 1 if transposed else 0)
    for hook in ref._forward_pre_hooks.values():
        hook(ref, (None,))
    cot = torch.randn(w.shape)
    (w * cot.to(device)).sum().backward()
    (ref.weight * cot.double()).sum().backward()
    return {"w": parity.relerr(w, ref.weight.detach()), "dw": parity.relerr(conv.weight_orig.grad, ref.weight_orig.grad),
            "u": parity.relerr(u, ref.weight_u), "v": parity.relerr(v, ref.weight_v)}


@pytest.mark.parametrize("shape,training", [((64, 3, 3), True), ((40, 24, 4), True), ((40, 24, 4), False), ((130, 70, 3), True)])
def test_spectral_norm_weight_matches_torch_hook(shape, training, emu):
    errs = spectral_norm_errs("cpu", shape, training)
    assert max(errs.values()) < 1e-5, errs


@pytest.mark.parametrize("shape,training", [((24, 40, 4), True), ((24, 40, 4), False), ((7, 130, 3), True)])
def test_spectral_norm_weight_dim1_matches_torch_hook(shape, training, emu):
    errs = spectral_norm_errs("cpu", shape, training, transposed=True)
    assert max(errs.values()) < 1e-5, errs


@pytest.mark.parametrize("B,C,N", [(2, 64, 16), (2, 24, 32), (1, 96, 16), (2, 8, 64)])
def test_fourier_unit_sweep_shapes_general_form(B, C, N, emu):
    """BASELINE configs[4] shapes that take the general form (rfft2 | 1x1 mix | BN statistics | BN+ReLU->irfft2), training
    mode, forward and backward against the float64 oracle.  Output in the max norm; gradients in the relative L2 norm
    (a ReLU element on the other side of its kink moves the max norm, SURVEY.md 8(c) caveat 1)."""
    torch.manual_seed(C + N)
    m = ffc.FourierUnitSN(C, C).train()
    with torch.no_grad():
        m.bn.weight.uniform_(0.5, 1.5); m.bn.bias.normal_(0, 0.2)
    P = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else
             (v.double().clone() if v.is_floating_point() else v.clone())) for k, v in m.state_dict().items()}
    x = torch.randn(B, C, N, N)
    xr = x.double().requires_grad_(True)
    ref = R.fourier_unit(xr, P, "", True)
    cot = torch.randn(ref.shape, dtype=torch.float64)
    (ref * cot).sum().backward()
    xo = x.clone().requires_grad_(True)
    out = m(xo)
    (out * cot.float()).sum().backward()

    def l2(a, b):
        return ((a.detach().double() - b.detach().double()).norm() / b.detach().double().norm()).item()
    assert parity.relerr(out, ref.detach()) < 1e-5
    assert l2(xo.grad, xr.grad) < 1e-3 and l2(m.conv_layer.weight.grad, P["conv_layer.weight"].grad) < 1e-3
    assert l2(m.bn.weight.grad, P["bn.weight"].grad) < 1e-3 and l2(m.bn.bias.grad, P["bn.bias"].grad) < 1e-3
    assert parity.relerr(m.bn.running_var, P["bn.running_var"]) < 1e-5


def any_size_fu_errs(dev, B, Cin, Cout, H, W, train):
    """FourierUnitSN on a plane that is not a square power of two (direct-DFT plane kernels, csrc/ffc_dft2.cu), forward and
    backward against the float64 oracle; the reference accepts every size ('works for odd H/W')."""
    torch.manual_seed(H * 131 + W)
    m = ffc.FourierUnitSN(Cin, Cout).train(train)
    with torch.no_grad():
        m.bn.weight.uniform_(0.5, 1.5); m.bn.bias.normal_(0, 0.2)
        m.bn.running_mean.uniform_(-0.2, 0.2); m.bn.running_var.uniform_(0.5, 1.5)
    P = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else
             (v.double().clone() if v.is_floating_point() else v.clone())) for k, v in m.state_dict().items()}
    x = torch.randn(B, Cin, H, W)
    xr = x.double().requires_grad_(True)
    ref = R.fourier_unit(xr, P, "", train)
    cot = torch.randn(ref.shape, dtype=torch.float64)
    (ref * cot).sum().backward()
    m = m.to(dev)
    xo = x.to(dev).requires_grad_(True)
    out = m(xo)
    (out * cot.float().to(dev)).sum().backward()

    def l2(a, b):
        return ((a.detach().double().cpu() - b.detach().double()).norm() / b.detach().double().norm()).item()
    return {"out": parity.relerr(out.cpu(), ref.detach()), "dx": l2(xo.grad, xr.grad),
            "dW": l2(m.conv_layer.weight.grad, P["conv_layer.weight"].grad),
            "dgamma": l2(m.bn.weight.grad, P["bn.weight"].grad), "dbeta": l2(m.bn.bias.grad, P["bn.bias"].grad),
            "running_var": parity.relerr(m.bn.running_var.cpu(), P["bn.running_var"])}


ANY_SIZE_CASES = [(2, 4, 6, 8, 8), (2, 4, 4, 7, 9), (2, 3, 3, 12, 12), (1, 2, 2, 48, 48), (2, 2, 2, 5, 16), (1, 2, 3, 24, 6), (2, 2, 2, 1, 3)]


@pytest.mark.parametrize("B,Cin,Cout,H,W", ANY_SIZE_CASES)
@pytest.mark.parametrize("train", [True, False])
def test_fourier_unit_any_plane_size(B, Cin, Cout, H, W, train, emu):
    errs = any_size_fu_errs("cpu", B, Cin, Cout, H, W, train)
    assert errs["out"] < 1e-5 and errs["running_var"] < 1e-5, errs
    assert max(errs["dx"], errs["dW"], errs["dgamma"], errs["dbeta"]) < 1e-3, errs      # relative L2 (ReLU kinks, see above)


def test_spectral_transform_non_power_of_two_plane(emu):
    """SpectralTransform at 12x12 (mg = 6 scale of fgan_cond_complete.py:325) against the oracle."""
    torch.manual_seed(3)
    m = ffc.SpectralTransform(8, 8, stride=1).train()
    P = {k: (v.double().clone() if v.is_floating_point() else v.clone()) for k, v in m.state_dict().items()}
    x = torch.randn(2, 8, 12, 12)
    ref = R.spectral_transform(x.double(), P, "", 1, False, True)
    assert parity.relerr(m(x), ref) < 1e-5


@pytest.mark.parametrize("B,C,N,chunk", [(3, 6, 64, 0), (5, 4, 32, 3 * 4 * 32 * 36 * 4), (2, 40, 16, 0), (1, 4, 128, 0)])
@pytest.mark.parametrize("train", [True, False])
def test_fourier_unit_l2_staged_form(B, C, N, chunk, train, emu):
    """The L2-staged Fourier unit (csrc/ffc_fu3.cu: plane rfft2 | channel mix + BN statistics | BN+ReLU -> plane irfft2 over
    chunks of images) in host emulation against the float64 oracle, forward with residual and running statistics; ``chunk``
    bytes force several chunks (one of them ragged).  The emulation build mixes in plain FP32; the tensor-core mix is checked
    on the GPU (tests/test_gpu_parity.py)."""
    from fastfourierconvolution_b200 import _C
    torch.manual_seed(C + N)
    m = ffc.FourierUnitSN(C, C).train(train)
    m.fused = "staged"
    with torch.no_grad():
        m.bn.weight.uniform_(0.5, 1.5); m.bn.bias.normal_(0, 0.2); m.bn.running_mean.normal_(0, 0.1); m.bn.running_var.uniform_(0.5, 1.5)
    P = {k: (v.double().clone() if v.is_floating_point() else v.clone()) for k, v in m.state_dict().items()}
    x, res = torch.randn(B, C, N, N), torch.randn(B, C, N, N)
    _C.lib().ffc_debug_fu3_chunk_bytes(chunk)
    try:
        with torch.no_grad():
            out = m._run(x, None, res)
    finally:
        _C.lib().ffc_debug_fu3_chunk_bytes(0)
    ref = R.fourier_unit(x.double(), P, "", train) + res.double()
    assert parity.relerr(out, ref) < 1e-6
    assert parity.relerr(m.bn.running_mean, P["bn.running_mean"]) < 1e-5 and parity.relerr(m.bn.running_var, P["bn.running_var"]) < 1e-5


@pytest.mark.parametrize("B,C,Co,N", [(3, 6, 6, 64), (5, 4, 7, 32), (2, 40, 36, 16), (1, 3, 4, 128)])
@pytest.mark.parametrize("train", [True, False])
def test_fourier_unit_l2_staged_backward(B, C, Co, N, train, emu):
    """Backward of the L2-staged Fourier unit (ffc_fu3_fwd_keep + ffc_fu3_bwd: adjoint transform with the ReLU mask and the
    BatchNorm-backward sums | constants | dY + dW | dS = dY W | adjoint transform) in host emulation against the float64
    oracle: dx, dW, dgamma, dbeta and the residual's gradient; channel counts that are not multiples of the 32-channel
    tile and Cin != Cout included.  Relative L2 norm for the gradients (a ReLU element on the other side of its kink moves
    the max norm, SURVEY.md 8(c) caveat 1)."""
    torch.manual_seed(C + N + Co)
    m = ffc.FourierUnitSN(C, Co).train(train)
    m.fused = "staged"
    with torch.no_grad():
        m.bn.weight.uniform_(0.5, 1.5); m.bn.bias.normal_(0, 0.2); m.bn.running_mean.normal_(0, 0.1); m.bn.running_var.uniform_(0.5, 1.5)
    P = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else
             (v.double().clone() if v.is_floating_point() else v.clone())) for k, v in m.state_dict().items()}
    x, res = torch.randn(B, C, N, N), torch.randn(B, Co, N, N)
    xr, rr = x.double().requires_grad_(True), res.double().requires_grad_(True)
    ref = R.fourier_unit(xr, P, "", train) + rr
    cot = torch.randn(ref.shape, dtype=torch.float64)
    (ref * cot).sum().backward()
    xo, ro = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
    out = m._run(xo, None, ro)
    (out * cot.float()).sum().backward()

    def l2(a, b):
        return ((a.detach().double() - b.detach().double()).norm() / b.detach().double().norm()).item()
    assert parity.relerr(out, ref.detach()) < 1e-5
    assert l2(xo.grad, xr.grad) < 1e-3 and l2(m.conv_layer.weight.grad, P["conv_layer.weight"].grad) < 1e-3
    assert l2(m.bn.weight.grad, P["bn.weight"].grad) < 1e-3 and l2(m.bn.bias.grad, P["bn.bias"].grad) < 1e-3
    assert torch.equal(ro.grad, cot.float())


@pytest.mark.parametrize("kind", ["adamw", "adam"])
def test_flat_adam_matches_torch_optimizer(kind, emu):
    """harness.FlatAdam (one table kernel per step over flat parameters / moments, gradients read in place) against
    torch.optim.AdamW / Adam over four steps on a small network with a parameter that never receives a gradient
    (the lfu.* case, spectral_transform.py:65-67): it must stay untouched and outside the flat buffers."""
    from fastfourierconvolution_b200.harness.train import FlatAdam
    torch.manual_seed(3)

    def make():
        torch.manual_seed(3)
        net = nn.Sequential(nn.Linear(9, 70), nn.Tanh(), nn.Linear(70, 5))
        net.unused = nn.Parameter(torch.randn(4, 4))
        return net
    a, b = make(), make()
    kw = dict(lr=2e-4, betas=(0.5, 0.999))
    ref = torch.optim.AdamW(a.parameters(), **kw) if kind == "adamw" else torch.optim.Adam(a.parameters(), **kw)
    mine = FlatAdam(b.parameters(), decoupled=(kind == "adamw"), weight_decay=0.01 if kind == "adamw" else 0.0, **kw)
    sched = torch.optim.lr_scheduler.LambdaLR(mine, lambda s: 1.0 - s / 10)
    sched_ref = torch.optim.lr_scheduler.LambdaLR(ref, lambda s: 1.0 - s / 10)
    for it in range(4):
        x = torch.randn(6, 9)
        for net, opt, sc in ((a, ref, sched_ref), (b, mine, sched)):
            opt.zero_grad()
            net(x).square().sum().backward()
            opt.step()
            sc.step()
    assert mine.adopted and b.unused.grad is None and torch.equal(a.unused, b.unused)
    assert all(p.data_ptr() >= mine.flat_p.data_ptr() for k, p in b.named_parameters() if k != "unused")
    for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert parity.relerr(pb.detach(), pa.detach()) < 1e-6, k
