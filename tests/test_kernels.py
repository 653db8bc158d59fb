"""C-ABI level checks of every kernel family against numpy / torch CPU math, on two backends:
  emu  -- the host emulation build of the CUDA sources (index arithmetic, edge cases; runs anywhere)
  cuda -- libffc_b200.so on the GPU (marked gpu): the same calls with device buffers."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import emu_backend
import parity


@pytest.fixture(scope="module", params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def lib(request):
    from fastfourierconvolution_b200 import _C
    if request.param == "emu":
        return _C.Library(emu_backend.build()), "cpu"
    return _C.lib(), "cuda"


def call(backend, name, *args):
    """Calls a C-ABI entry point with the tensors moved to the backend's device and copies every
    tensor back afterwards (outputs are written in place), so test bodies are backend agnostic."""
    library, device = backend
    moved = [a.to(device) if torch.is_tensor(a) else a for a in args]
    raw = [ctypes.c_void_p(a.data_ptr()) if torch.is_tensor(a) else a for a in moved]
    rc = getattr(library, name)(*raw)
    assert rc == 0, library.last_error()
    if device != "cpu":
        torch.cuda.synchronize()
        for a, m in zip(args, moved):
            if torch.is_tensor(a):
                a.copy_(m)


def upos(k, n):
    n1 = n if n <= 32 else n // 8
    n2 = n // n1
    return n2 * (k % n1) + k // n1


@pytest.mark.parametrize("n", [4, 8, 16, 32, 64, 128])
def test_rfft2_irfft2_all_sizes(lib, n):
    rng = np.random.default_rng(n)
    npl, wf = 5, n // 2 + 1
    perm = [upos(k, n) for k in range(n)]
    x = torch.from_numpy(rng.standard_normal((npl, n, n)).astype(np.float32))
    spec = torch.zeros(npl, 2, n, wf)
    call(lib, "ffc_rfft2", x, spec, npl, n, n, 0, None)
    ref = torch.fft.rfftn(x.double(), dim=(-2, -1), norm="ortho")
    got = torch.complex(spec[:, 0].double(), spec[:, 1].double())[:, perm, :]
    assert parity.relerr(torch.view_as_real(got), torch.view_as_real(ref)) < 1e-6
    # non-Hermitian spectrum through the inverse (torch c2r semantics), with residual
    z = torch.from_numpy(rng.standard_normal((npl, 2, n, wf)).astype(np.float32))
    refo = torch.fft.irfftn(torch.complex(z[:, 0].double(), z[:, 1].double()), s=(n, n), dim=(-2, -1), norm="ortho")
    zp = torch.empty_like(z)
    zp[:, :, perm, :] = z
    res = torch.from_numpy(rng.standard_normal((npl, n, n)).astype(np.float32))
    out = torch.zeros(npl, n, n)
    call(lib, "ffc_irfft2", zp, res, out, npl, n, n, 0, None)
    assert parity.relerr(out.double() - res.double(), refo) < 1e-6


@pytest.mark.parametrize("n", [8, 64])
def test_fft_adjoint_pairs(lib, n):
    """colscale=1 variants are the exact adjoints: <rfft2(x), s> == <x, irfft2_adj(s)> and vice versa."""
    rng = np.random.default_rng(1)
    npl, wf = 3, n // 2 + 1
    x = torch.from_numpy(rng.standard_normal((npl, n, n)).astype(np.float32))
    s = torch.from_numpy(rng.standard_normal((npl, 2, n, wf)).astype(np.float32))
    fx, ats = torch.zeros_like(s), torch.zeros_like(x)
    call(lib, "ffc_rfft2", x, fx, npl, n, n, 0, None)
    call(lib, "ffc_irfft2", s, None, ats, npl, n, n, 1, None)
    assert abs((fx.double() * s.double()).sum() - (x.double() * ats.double()).sum()) < 1e-3 * n
    ix, atx = torch.zeros_like(x), torch.zeros_like(s)
    call(lib, "ffc_irfft2", s, None, ix, npl, n, n, 0, None)
    call(lib, "ffc_rfft2", x, atx, npl, n, n, 1, None)
    assert abs((ix.double() * x.double()).sum() - (s.double() * atx.double()).sum()) < 1e-3 * n


CONV_CASES = [
    # B, cins, cout, Hi, k, s, p, transposed, bias, addend, out_pad
    (2, [5], 7, 8, 3, 1, 1, False, True, False, 0),
    (3, [6, 3], 70, 8, 4, 2, 1, False, True, True, 0),
    (2, [4], 3, 8, 1, 1, 0, False, False, True, 0),
    (2, [9, 5], 33, 4, 4, 2, 1, True, False, False, 0),
    (2, [10], 6, 1, 4, 1, 0, True, True, False, 0),
    (1, [3], 5, 5, 3, 2, 1, True, False, True, 1),
    (2, [3], 5, 7, 3, 2, 1, False, False, False, 0),
    (2, [20], 130, 6, 3, 1, 1, True, True, False, 0),
    (1, [17], 1, 4, 4, 1, 0, False, False, False, 0),
    (3, [40, 24], 100, 6, 4, 2, 1, True, True, True, 0),
    (2, [33], 20, 9, 3, 1, 1, False, True, False, 0),
    (1, [40], 400, 5, 3, 1, 1, False, True, False, 0),       # several output-channel tiles
    (2, [70, 10], 200, 4, 4, 2, 1, True, False, True, 0),
    (40, [36], 24, 6, 3, 1, 1, False, False, False, 0),      # several pixel tiles
    (9, [48], 40, 8, 4, 2, 1, True, False, False, 0),        # tcgen05 weight-gradient shapes (both directions)
    (9, [40], 48, 8, 4, 2, 1, False, True, False, 0),
    (33, [26], 30, 4, 3, 1, 1, False, False, False, 0),
    (5, [20, 12], 3, 12, 3, 1, 1, False, True, True, 0),     # <= 4 channels on one side: direct kernels
    (4, [3], 24, 8, 4, 2, 1, False, True, False, 0),
    (4, [4], 10, 6, 4, 2, 1, True, False, False, 0),
    (4, [10], 2, 6, 4, 2, 1, True, False, False, 0),
    (3, [9, 5], 3, 16, 3, 1, 1, False, False, True, 0),      # the 4-pixel strip forms (k3 s1 p1, width % 4 == 0)
    (3, [24], 8, 8, 1, 1, 0, False, False, True, 0),         # narrow 1x1 (SpectralTransform conv1 / conv2 and their dgrads)
    (2, [8], 16, 16, 1, 1, 0, False, True, False, 0),
    (2, [33], 20, 6, 1, 1, 0, True, False, False, 0),
    (2, [16], 23, 4, 1, 1, 0, False, True, True, 0),
    (3, [9], 3, 16, 3, 1, 1, True, True, False, 0),          # ... and the transposed one (mirrored taps): dgrad of a 3 -> C conv
    (2, [24, 8], 2, 8, 3, 1, 1, True, False, True, 0),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("reference_form", [0, 1, 2, 3, 4, 5])
def test_conv_forward_dgrad_wgrad(lib, case, reference_form):
    """0: tensor-core 3xTF32 kernels; 1: the simple single-buffered forms; 2: tuned FP32 SIMT kernels;
    3: tensor-core 3xTF32 with packed weights and a cp.async ring (ffc_conv2d_fwd_ws); 4: tcgen05 / TMEM implicit GEMM
    (device build only; the emulation build runs family 3 for it); 5: the default automatic choice, which adds the
    direct kernels for <= 4 channels and the tcgen05 weight gradient."""
    lib[0].ffc_debug_conv_reference(reference_form)
    _conv_mode[0] = reference_form if lib[1] != "cpu" else 1      # the emulation build computes every family in FP32
    try:
        _conv_case(lib, case)
    finally:
        lib[0].ffc_debug_conv_reference(5)


# 3xTF32 on the tensor cores: the dropped lo*lo term and the MMA's internal accumulation leave ~1e-5
# (still an order of magnitude inside the 1e-4 FP32 parity bound); the FP32 FMA families reach ~1e-7.
CONV_TOL = {0: 4e-5, 1: 3e-6, 2: 3e-6, 3: 4e-5, 4: 4e-5, 5: 4e-5}
_conv_mode = [0]


def _conv_fwd(lib, x0, w0, cin0, x1, w1, cin1, b, ad, y, B, cout, Hi, Wi, Ho, Wo, k, s, p, tr):
    nbytes = lib[0].ffc_conv2d_workspace_bytes(cin0, cin1, cout, k, s, p, tr)
    ws = torch.zeros(nbytes + 64, dtype=torch.uint8)
    call(lib, "ffc_conv2d_fwd_ws", x0, w0, cin0, x1, w1, cin1, b, ad, y, B, cout, Hi, Wi, Ho, Wo, k, s, p, tr, ws, ws.numel(), None)


def _conv_case(lib, case):
    B, cins, cout, Hi, k, s, p, tr, bias, addend, op = case
    torch.manual_seed(0)
    xs = [torch.randn(B, c, Hi, Hi) for c in cins]
    ws = [torch.randn(*((c, cout) if tr else (cout, c)), k, k) * 0.1 for c in cins]
    b = torch.randn(cout) if bias else None

    def op_ref(x, w):
        return F.conv_transpose2d(x, w, None, s, p, op) if tr else F.conv2d(x, w, None, s, p)

    ref = sum(op_ref(x.double(), w.double()) for x, w in zip(xs, ws))
    Ho = ref.shape[-1]
    ad = torch.randn(B, cout, Ho, Ho) if addend else None
    if b is not None:
        ref = ref + b.double().view(1, -1, 1, 1)
    if ad is not None:
        ref = ref + ad.double()
    y = torch.empty(B, cout, Ho, Ho)
    x1, w1 = (xs[1], ws[1]) if len(xs) > 1 else (None, None)
    _conv_fwd(lib, xs[0], ws[0], cins[0], x1, w1, cins[1] if x1 is not None else 0, b, ad, y,
              B, cout, Hi, Hi, Ho, Ho, k, s, p, int(tr))
    assert parity.relerr(y, ref) < CONV_TOL[_conv_mode[0]]
    dy = torch.randn(B, cout, Ho, Ho)
    for x, w in zip(xs, ws):
        xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
        op_ref(xd, wd).backward(dy.double())
        dW, dx = torch.empty_like(w), torch.empty_like(x)
        if tr:
            call(lib, "ffc_conv2d_wgrad", x, dy, dW, B, x.shape[1], cout, Hi, Hi, Ho, Ho, k, s, p, None)
        else:
            call(lib, "ffc_conv2d_wgrad", dy, x, dW, B, cout, x.shape[1], Ho, Ho, Hi, Hi, k, s, p, None)
        # wgrad is FP32 FMA in families 0-3; family 4 (and the default 5) use the tcgen05 3xTF32 kernel where it applies
        assert parity.relerr(dW, wd.grad) < (4e-5 if _conv_mode[0] >= 4 else 3e-6)
        _conv_fwd(lib, dy, w, cout, None, None, 0, None, None, dx, B, x.shape[1], Ho, Ho, Hi, Hi, k, s, p, int(not tr))
        assert parity.relerr(dx, xd.grad) < CONV_TOL[_conv_mode[0]]


# (B, cin0, cin1, cout0, cout1, Hi, k, stride, pad, transposed, bias)
BLOCK_CASES = [
    (3, 48, 16, 24, 8, 8, 4, 2, 1, True, False),      # an upsampling FFC stage: l2l | g2l and l2g
    (2, 36, 12, 40, 24, 8, 4, 2, 1, False, True),     # a downsampling stage with biases (sngan FDiscriminator)
    (2, 20, 0, 130, 70, 6, 3, 1, 1, False, True),     # no global input (first layer); several output tiles
    (2, 6, 3, 5, 4, 5, 3, 1, 1, True, True),          # narrow: falls back to two plain launches
    (2, 96, 32, 24, 8, 8, 4, 2, 1, True, True),       # deep K on a small grid: split-K with atomics into both outputs
    (1, 160, 0, 40, 24, 6, 3, 1, 1, False, True),
]


@pytest.mark.parametrize("case", BLOCK_CASES)
def test_conv_block_forward(lib, case):
    """ffc_conv2d_block_fwd_ws: y0 = conv(x0, w00) + conv(x1, w10) + bias[:cout0], y1 = conv(x0, w01) + bias[cout0:]."""
    B, cin0, cin1, cout0, cout1, Hi, k, s, p, tr, bias = case
    torch.manual_seed(1)
    shp = (lambda ci, co: (ci, co, k, k)) if tr else (lambda ci, co: (co, ci, k, k))
    x0 = torch.randn(B, cin0, Hi, Hi)
    w00, w01 = torch.randn(shp(cin0, cout0)) * 0.1, torch.randn(shp(cin0, cout1)) * 0.1
    x1 = torch.randn(B, cin1, Hi, Hi) if cin1 else None
    w10 = torch.randn(shp(cin1, cout0)) * 0.1 if cin1 else None
    b = torch.randn(cout0 + cout1) if bias else None
    op_ref = (lambda x, w: F.conv_transpose2d(x.double(), w.double(), None, s, p)) if tr else (lambda x, w: F.conv2d(x.double(), w.double(), None, s, p))
    r0, r1 = op_ref(x0, w00), op_ref(x0, w01)
    if cin1:
        r0 = r0 + op_ref(x1, w10)
    if bias:
        r0 = r0 + b[:cout0].double().view(1, -1, 1, 1)
        r1 = r1 + b[cout0:].double().view(1, -1, 1, 1)
    Ho = r0.shape[-1]
    y0, y1 = torch.empty(B, cout0, Ho, Ho), torch.empty(B, cout1, Ho, Ho)
    nbytes = lib[0].ffc_conv2d_workspace_bytes(cin0, cin1, cout0 + cout1, k, s, p, int(tr))
    ws = torch.zeros(nbytes + 64, dtype=torch.uint8)
    call(lib, "ffc_conv2d_block_fwd_ws", x0, w00, w01, cin0, x1, w10, cin1, b, y0, cout0, y1, cout1,
         B, Hi, Hi, Ho, Ho, k, s, p, int(tr), ws, ws.numel(), None)
    tol = 4e-5 if lib[1] != "cpu" else 3e-6
    assert parity.relerr(y0, r0) < tol and parity.relerr(y1, r1) < tol


@pytest.mark.parametrize("case", [(3, 40, 48, 8, 4, 2, 2, 0.1), (2, 24, 130, 6, 3, 1, 2, 0.2), (2, 3, 20, 8, 3, 1, 2, 0.1), (2, 33, 32, 6, 3, 1, 1, 0.0)])
def test_conv_act_forward(lib, case):
    """ffc_conv2d_act_fwd_ws: LeakyReLU / ReLU of conv + bias (fused epilogue on the tcgen05 path, two kernels otherwise)."""
    B, cin, cout, Hi, k, s, act, slope = case
    torch.manual_seed(2)
    x, w, b = torch.randn(B, cin, Hi, Hi), torch.randn(cout, cin, k, k) * 0.1, torch.randn(cout)
    ref = ACTS[act](F.conv2d(x.double(), w.double(), b.double(), s, 1)) if act == 1 else F.leaky_relu(F.conv2d(x.double(), w.double(), b.double(), s, 1), slope)
    Ho = ref.shape[-1]
    y = torch.empty(B, cout, Ho, Ho)
    nbytes = max(lib[0].ffc_conv2d_workspace_bytes(cin, 0, cout, k, s, 1, 0), 2 * cout * 8)
    ws = torch.zeros(nbytes + 64, dtype=torch.uint8)
    call(lib, "ffc_conv2d_act_fwd_ws", x, w, cin, b, y, B, cout, Hi, Hi, Ho, Ho, k, s, 1, act, slope, ws, ws.numel(), None)
    assert parity.relerr(y, ref) < (4e-5 if lib[1] != "cpu" else 3e-6)


ACTS = {0: lambda z: z, 1: F.relu, 2: lambda z: F.leaky_relu(z, 0.1), 3: F.gelu, 4: torch.tanh, 5: torch.sigmoid}


@pytest.mark.parametrize("case", [(4, 5, 4, 4, 1, 1, 3), (3, 7, 3, 5, 1, 1, 2), (2, 3, 8, 8, 0, 1, 4), (5, 6, 4, 4, 1, 0, 1),
                                  (2, 4, 2, 2, 0, 1, 5), (3, 300, 2, 2, 1, 1, 0), (2, 2, 1, 3, 1, 1, 1)])
def test_bn_act_forward_backward(lib, case):
    B, C, H, W, norm, training, act = case
    cf = ctypes.c_float
    torch.manual_seed(1)
    x = torch.randn(B, C, H, W) * 2 + 0.5
    g, b = torch.rand(C) + 0.5, torch.randn(C) * 0.1
    rm, rv = torch.randn(C) * 0.1, torch.rand(C) + 0.5
    xd, gd, bd = x.double().requires_grad_(True), g.double().requires_grad_(True), b.double().requires_grad_(True)
    rmd, rvd = rm.double().clone(), rv.double().clone()
    z = F.batch_norm(xd, rmd, rvd, gd, bd, bool(training), 0.1, 1e-5) if norm else xd
    yref = ACTS[act](z)
    dy = torch.randn(B, C, H, W)
    yref.backward(dy.double())
    y, sm, si = torch.empty_like(x), torch.empty(C), torch.empty(C)
    ws = torch.empty(2 * C, dtype=torch.float64)
    nb = ctypes.c_size_t(ws.numel() * 8)
    call(lib, "ffc_bn_act_fwd", x, y, g, b, rm, rv, sm, si, B, C, H * W, norm, training,
                               cf(1e-5), cf(0.1), act, cf(0.1), ws, nb, None)
    dx, dg, db = torch.empty_like(x), torch.zeros(C), torch.zeros(C)
    call(lib, "ffc_bn_act_bwd", x, dy, dx, g, b, sm, si, dg, db, B, C, H * W, norm, training,
                               act, cf(0.1), ws, nb, None)
    assert parity.relerr(y, yref.detach()) < 5e-6
    if B * H * W > 1:
        assert parity.relerr(dx, xd.grad) < 5e-6
    if norm:
        assert parity.relerr(dg, gd.grad, 1e-3) < 5e-6 and parity.relerr(db, bd.grad) < 5e-6
        assert parity.relerr(rm, rmd) < 5e-6 and parity.relerr(rv, rvd) < 5e-6


@pytest.mark.parametrize("case", [(3, 32, 2, 4, 0), (2, 16, 1, 4, 1), (2, 48, 3, 8, 2), (2, 8, 0, 4, 1)])
def test_se_forward_backward(lib, case):
    B, C, hid, H, mode = case
    torch.manual_seed(2)
    x, w1, w2 = torch.randn(B, C, H, H), torch.randn(hid, C) * 0.3, torch.randn(C, hid) * 0.3
    xd, w1d, w2d = x.double().requires_grad_(True), w1.double().requires_grad_(True), w2.double().requires_grad_(True)
    r = xd if mode == 0 else (F.interpolate(xd, scale_factor=2, mode="nearest") if mode == 1 else F.avg_pool2d(xd, 2, 2))
    gate = torch.sigmoid(F.linear(F.relu(F.linear(r.mean((2, 3)), w1d)), w2d))
    yref = r * gate.view(B, C, 1, 1)
    dy = torch.randn_like(yref).float()
    yref.backward(dy.double())
    y, sm, sh, sg = torch.empty(yref.shape), torch.empty(B, C), torch.empty(B, max(hid, 1)), torch.empty(B, C)
    ws = torch.empty(B * C * 4 + B * (C + hid) + 16, dtype=torch.float64)
    nb = ctypes.c_size_t(ws.numel() * 8)
    call(lib, "ffc_se_fwd", x, w1, w2, y, sm, sh, sg, B, C, hid, H, H, mode, ws, nb, None)
    dx, dw1, dw2 = torch.empty_like(x), torch.empty_like(w1), torch.empty_like(w2)
    call(lib, "ffc_se_bwd", x, dy, w1, w2, sm, sh, sg, dx, dw1, dw2, B, C, hid, H, H, mode, ws, nb, None)
    assert parity.relerr(y, yref.detach()) < 5e-6 and parity.relerr(dx, xd.grad) < 5e-6
    if hid:
        assert parity.relerr(dw1, w1d.grad) < 5e-6 and parity.relerr(dw2, w2d.grad) < 5e-6


def test_empty_batch_is_a_noop(lib):
    x = torch.zeros(0, 4, 8, 8)
    call(lib, "ffc_rfft2", torch.zeros(4), torch.zeros(4), 0, 8, 8, 0, None)
    y = torch.zeros(0, 3, 8, 8)
    w = torch.randn(3, 4, 3, 3)
    call(lib, "ffc_conv2d_fwd", torch.zeros(4), w, 4, None, None, 0, None, None, torch.zeros(4), 0, 3, 8, 8, 8, 8, 3, 1, 1, 0, None)
    # the entry points added for the block form and the discriminator stages
    ws = torch.zeros(1 << 16, dtype=torch.uint8)
    w32 = torch.randn(32, 4, 3, 3)
    call(lib, "ffc_conv2d_act_fwd_ws", torch.zeros(4), w32, 4, None, torch.zeros(4), 0, 32, 8, 8, 8, 8, 3, 1, 1, 2, ctypes.c_float(0.1),
         ws, ws.numel(), None)
    call(lib, "ffc_conv2d_block_fwd_ws", torch.zeros(4), w32, w32, 4, None, None, 0, None, torch.zeros(4), 32, torch.zeros(4), 32,
         0, 8, 8, 8, 8, 3, 1, 1, 0, ws, ws.numel(), None)


def test_new_entry_points_reject_bad_arguments(lib):
    """Error behaviour of the C ABI: non-zero status + message, nothing launched."""
    library = lib[0]
    ws = torch.zeros(1 << 12, dtype=torch.uint8)
    z = ctypes.c_void_p(0)
    p = ctypes.c_void_p(ws.data_ptr())
    # activation other than LeakyReLU / ReLU
    assert library.ffc_conv2d_act_fwd_ws(p, p, 4, z, p, 1, 8, 4, 4, 4, 4, 3, 1, 1, 3, ctypes.c_float(0.1), p, ws.numel(), z) != 0
    assert "LeakyReLU" in library.last_error()
    # LeakyReLU with a non-positive slope (the backward mask is read from the output's sign)
    assert library.ffc_conv2d_act_fwd_ws(p, p, 4, z, p, 1, 8, 4, 4, 4, 4, 3, 1, 1, 2, ctypes.c_float(0.0), p, ws.numel(), z) != 0
    # spectral norm: workspace too small, bad matrix view
    assert library.ffc_spectral_norm_fwd(p, p, p, z, z, p, p, 8, 16, 0, 1, ctypes.c_float(1e-12), p, 4, z) != 0
    assert "workspace" in library.last_error()
    assert library.ffc_spectral_norm_fwd(p, p, p, z, z, p, p, 8, 15, 4, 1, ctypes.c_float(1e-12), p, ws.numel(), z) != 0


def _fu_reference(x, w, gamma, beta, rmean, rvar, training, eps=1e-5):
    """FourierUnitSN.forward (fourier_unity.py:32-58) in float64 torch; returns out, batch mean, batch invstd."""
    B, Cin, H, W = x.shape
    X = torch.fft.rfftn(x.double(), dim=(-2, -1), norm="ortho")
    S = torch.stack((X.real, X.imag), dim=2).reshape(B, 2 * Cin, H, W // 2 + 1)
    Y = torch.einsum("oc,bchw->bohw", w.double(), S)
    if training:
        mean, var = Y.mean(dim=(0, 2, 3)), Y.var(dim=(0, 2, 3), unbiased=False)
    else:
        mean, var = rmean.double(), rvar.double()
    invstd = 1.0 / torch.sqrt(var + eps)
    Z = F.relu((Y - mean.view(1, -1, 1, 1)) * (invstd * gamma.double()).view(1, -1, 1, 1) + beta.double().view(1, -1, 1, 1))
    Cout = w.shape[0] // 2
    Zc = torch.complex(Z.view(B, Cout, 2, H, -1)[:, :, 0], Z.view(B, Cout, 2, H, -1)[:, :, 1])
    return torch.fft.irfftn(Zc, s=(H, W), dim=(-2, -1), norm="ortho"), mean, invstd, Y.var(dim=(0, 2, 3), unbiased=True)


FU_CASES = [
    # B, Cin, Cout, N, training, residual, two_pass
    (3, 8, 8, 32, 1, True, 0), (3, 8, 8, 32, 1, False, 1), (2, 8, 8, 32, 0, True, 0),
    (5, 16, 16, 16, 1, True, 0), (5, 16, 16, 16, 1, True, 1), (2, 16, 16, 16, 0, False, 0),
    (3, 32, 32, 8, 1, True, 0), (3, 32, 32, 8, 1, False, 1),
    (2, 5, 7, 16, 1, False, 0), (2, 7, 3, 8, 1, False, 1), (2, 12, 20, 8, 0, False, 0), (2, 3, 2, 32, 1, True, 0),
    (2, 16, 16, 32, 1, True, 0), (1, 32, 32, 32, 1, True, 1), (2, 24, 9, 16, 1, False, 0), (2, 4, 4, 4, 1, True, 0),
]


# 32x32 planes with <= 8 channels run the warp-private kernels of csrc/ffc_fu4.cu: more of them, with unequal channel counts,
# a batch beyond the co-resident capacity (two-pass form) and one image
FU_CASES += [(7, 8, 5, 32, 1, True, 0), (4, 1, 8, 32, 1, False, 0), (2, 5, 1, 32, 0, True, 0), (1, 8, 8, 32, 1, False, 0),
             (330, 2, 2, 32, 1, False, 0), (9, 4, 4, 32, 1, True, 1), (40, 8, 8, 32, 0, False, 0)]


@pytest.mark.parametrize("case", FU_CASES)
@pytest.mark.parametrize("ws_mode", ["min", "full"])
def test_fused_fourier_unit_forward(lib, case, ws_mode):
    """ffc_fu_fwd (cooperative single pass / two-pass / eval) against FourierUnitSN.forward in float64; with the minimal
    workspace (4*Cout doubles: zero fill + atomics) and with ffc_fu_workspace_bytes (per-image partial sums where the kernel
    has that form)."""
    B, Cin, Cout, N, training, has_res, two_pass = case
    torch.manual_seed(B * 1000 + Cin * 10 + N)
    x = torch.randn(B, Cin, N, N)
    w = torch.randn(2 * Cout, 2 * Cin) / (2 * Cin) ** 0.5
    gamma, beta = torch.rand(2 * Cout) + 0.5, torch.randn(2 * Cout) * 0.3
    rmean, rvar = torch.randn(2 * Cout) * 0.1, torch.rand(2 * Cout) + 0.5
    res = torch.randn(B, Cout, N, N) if has_res else None
    ref, mean, invstd, var_unb = _fu_reference(x, w, gamma, beta, rmean, rvar, training)
    if has_res:
        ref = ref + res.double()
    assert lib[0].ffc_fu_fused_supported(B, Cin, Cout, N, N) == 1
    out = torch.zeros(B, Cout, N, N)
    sm, si = torch.zeros(2 * Cout), torch.zeros(2 * Cout)
    rm, rv = rmean.clone(), rvar.clone()
    ws = torch.zeros(4 * Cout * 8 if ws_mode == "min" else lib[0].ffc_fu_workspace_bytes(B, Cout), dtype=torch.uint8)
    cf = ctypes.c_float
    lib[0].ffc_debug_fu_two_pass(two_pass)
    try:
        call(lib, "ffc_fu_fwd", x, w, gamma, beta, rm, rv, sm, si, res, out, B, Cin, Cout, N, N, training,
             cf(1e-5), cf(0.1), ws, ws.numel(), None)
    finally:
        lib[0].ffc_debug_fu_two_pass(0)
    assert parity.relerr(out, ref) < 2e-6
    assert parity.relerr(sm, mean) < 1e-5 and parity.relerr(si, invstd) < 1e-5
    if training:
        assert parity.relerr(rm, 0.9 * rmean.double() + 0.1 * mean) < 1e-5
        assert parity.relerr(rv, 0.9 * rvar.double() + 0.1 * var_unb) < 1e-5
    else:
        assert torch.equal(rm, rmean) and torch.equal(rv, rvar)


def _fu_reference_bwd(x, w, gamma, beta, rmean, rvar, training, dout, eps=1e-5):
    """Gradients of FourierUnitSN.forward by float64 autograd of the same op sequence."""
    xd = x.double().requires_grad_(True)
    wd, gd, bd = (t.double().requires_grad_(True) for t in (w, gamma, beta))
    B, Cin, H, W = x.shape
    X = torch.fft.rfftn(xd, dim=(-2, -1), norm="ortho")
    S = torch.stack((X.real, X.imag), dim=2).reshape(B, 2 * Cin, H, W // 2 + 1)
    Y = torch.einsum("oc,bchw->bohw", wd, S)
    if training:
        mean, var = Y.mean(dim=(0, 2, 3)), Y.var(dim=(0, 2, 3), unbiased=False)
    else:
        mean, var = rmean.double(), rvar.double()
    invstd = 1.0 / torch.sqrt(var + eps)
    Z = F.relu((Y - mean.view(1, -1, 1, 1)) * (invstd * gd).view(1, -1, 1, 1) + bd.view(1, -1, 1, 1))
    Cout = w.shape[0] // 2
    Zc = torch.complex(Z.view(B, Cout, 2, H, -1)[:, :, 0], Z.view(B, Cout, 2, H, -1)[:, :, 1])
    out = torch.fft.irfftn(Zc, s=(H, W), dim=(-2, -1), norm="ortho")
    out.backward(dout.double())
    return xd.grad, wd.grad, gd.grad, bd.grad, mean.detach(), invstd.detach()


FU_BWD_CASES = [
    # B, Cin, Cout, N, training
    (3, 8, 8, 32, 1), (2, 8, 8, 32, 0), (5, 16, 16, 16, 1), (3, 32, 32, 8, 1), (2, 5, 7, 16, 1), (2, 7, 3, 8, 1),
    (2, 12, 20, 8, 0), (2, 3, 2, 32, 1), (2, 16, 16, 32, 1), (2, 24, 9, 16, 1), (1, 32, 32, 16, 1),
    (7, 8, 5, 32, 1), (4, 1, 8, 32, 1), (2, 5, 1, 32, 0), (1, 8, 8, 32, 1), (70, 4, 4, 32, 1),
]


@pytest.mark.parametrize("case", FU_BWD_CASES)
@pytest.mark.parametrize("ws_mode", ["min", "full"])
def test_fused_fourier_unit_backward(lib, case, ws_mode):
    """ffc_fu_bwd against float64 autograd of FourierUnitSN.forward; with ffc_fu_bwd_workspace_bytes the 32x32 / <= 8 channel
    shapes take the warp-private kernel (csrc/ffc_fu4.cu: per-image partial sums and weight-gradient tiles)."""
    B, Cin, Cout, N, training = case
    torch.manual_seed(B * 1000 + Cin * 10 + N + 1)
    x = torch.randn(B, Cin, N, N)
    w = torch.randn(2 * Cout, 2 * Cin) / (2 * Cin) ** 0.5
    gamma, beta = torch.rand(2 * Cout) + 0.5, torch.randn(2 * Cout) * 0.3
    rmean, rvar = torch.randn(2 * Cout) * 0.1, torch.rand(2 * Cout) + 0.5
    dout = torch.randn(B, Cout, N, N)
    dx_r, dw_r, dg_r, db_r, mean, invstd = _fu_reference_bwd(x, w, gamma, beta, rmean, rvar, training, dout)
    assert lib[0].ffc_fu_bwd_supported(B, Cin, Cout, N, N) == 1
    dx, dw = torch.zeros_like(x), torch.full_like(w, 7.0)
    dg, db = torch.zeros(2 * Cout), torch.zeros(2 * Cout)
    ws = torch.zeros(4 * Cout * 8 if ws_mode == "min" else lib[0].ffc_fu_bwd_workspace_bytes(B, Cin, Cout), dtype=torch.uint8)
    call(lib, "ffc_fu_bwd", x, dout, w, gamma, beta, mean.float(), invstd.float(), dx, dw, dg, db,
         B, Cin, Cout, N, N, training, ws, ws.numel(), None)
    assert parity.relerr(dx, dx_r) < 5e-6
    assert parity.relerr(dw, dw_r) < 5e-6
    assert parity.relerr(dg, dg_r) < 5e-6 and parity.relerr(db, db_r) < 5e-6


# ---- glue kernels (csrc/ffc_glue.cu) and the spectral-norm backward -------------------------------------------------
@pytest.mark.parametrize("B,C,H", [(3, 5, 8), (2, 48, 7), (4, 16, 32)])
def test_noise_add_and_its_weight_gradient(lib, B, C, H):
    """NoiseInjection (layers/noise_injection.py:20-32): x + weight * noise, and dweight = sum(dy * noise)."""
    torch.manual_seed(B * C)
    x, noise, w = torch.randn(B, C, H, H), torch.randn(B, 1, H, H), torch.randn(1, C, 1, 1)
    out, dw = torch.zeros_like(x), torch.full((C,), 7.0)
    call(lib, "ffc_noise_add_fwd", x, w, noise, out, B, C, H * H, None)
    assert torch.equal(out, torch.addcmul(x, w, noise)) or (out - torch.addcmul(x, w, noise)).abs().max() < 1e-6
    dy = torch.randn(B, C, H, H)
    call(lib, "ffc_noise_add_bwd_w", dy, noise, dw, B, C, H * H, None)
    ref = (dy.double() * noise.double()).sum((0, 2, 3))
    assert parity.relerr(dw, ref) < 1e-6


@pytest.mark.parametrize("lo,hi", [(-1.0, 1.0), (1.0, -1.0)])
def test_to_uint8_matches_the_reference_epilogue(lib, lo, hi):
    """fgan_complete.py:136-138: (255 * (clamp(x, -1, 1) * 0.5 + 0.5)).to(uint8); lo > hi: no clamp (fgan64's own-range clamp)."""
    torch.manual_seed(0)
    x = torch.tanh(torch.randn(3, 3, 9, 9) * 2) if lo > hi else torch.randn(3, 3, 9, 9) * 1.5
    out = torch.zeros(x.shape, dtype=torch.uint8)
    call(lib, "ffc_to_uint8", x, out, ctypes.c_longlong(x.numel()), ctypes.c_float(lo), ctypes.c_float(hi), None)
    ref = (255 * ((x.clamp(-1, 1) if lo < hi else x) * 0.5 + 0.5)).to(torch.uint8)
    assert (out.int() - ref.int()).abs().max() <= 1 and (out != ref).float().mean() < 0.01      # FMA contraction may move a value across an integer


@pytest.mark.parametrize("h,w,kk", [(24, 70, 0), (16, 9 * 8, 9)])
def test_spectral_norm_backward(lib, h, w, kk):
    """dW = g / sigma - (sum(g * W) / sigma^2) u v^T in the weight's own storage order (kk > 0: ConvTranspose2d layout)."""
    torch.manual_seed(h + w)
    Wm = torch.randn(h, w)                       # matrix view
    u, v = F.normalize(torch.randn(h), dim=0), F.normalize(torch.randn(w), dim=0)
    sigma = torch.dot(u, Wm @ v).reshape(1)
    gm = torch.randn(h, w)
    ref = gm.double() / sigma.double() - (gm.double() * Wm.double()).sum() / sigma.double() ** 2 * torch.outer(u.double(), v.double())
    if kk:                                       # storage (w / kk, h, kk) <-> matrix (h, w)
        store = lambda m: m.reshape(h, w // kk, kk).permute(1, 0, 2).contiguous()
        Ws, gs, refs = store(Wm), store(gm), store(ref)
    else:
        Ws, gs, refs = Wm, gm, ref
    dw = torch.zeros_like(Ws)
    ws = torch.zeros(16)
    call(lib, "ffc_spectral_norm_bwd", gs, Ws, u, v, sigma, dw, h, w, kk, ws, ctypes.c_size_t(64), None)
    assert parity.relerr(dw, refs) < 1e-5


# ---- Linear layers and the optimiser (csrc/ffc_glue.cu) --------------------------------------------------------------
LL = ctypes.c_longlong


@pytest.mark.parametrize("B,K,O", [(5, 7, 3), (66, 100, 130), (16, 700, 1), (3, 64, 64)])
def test_linear_forward_and_gradients_on_the_gemm_kernel(lib, B, K, O):
    """nn.Linear (fgan_complete.py:92-95 stem, :160-170 head): x W^T + b, dy W, dy^T x through ffc_gemm_f32's strides and the
    bias gradient through ffc_colsum_f32; (16, 700, 1) takes the split-K path on 148 SMs."""
    torch.manual_seed(B + K + O)
    x, w, b, dy = torch.randn(B, K), torch.randn(O, K), torch.randn(O), torch.randn(B, O)
    y, dx, dw, db = torch.full((B, O), 9.0), torch.full((B, K), 9.0), torch.full((O, K), 9.0), torch.full((O,), 9.0)
    call(lib, "ffc_gemm_f32", x, w, b, y, B, O, K, LL(K), LL(1), LL(1), LL(K), LL(O), LL(1), None)
    call(lib, "ffc_gemm_f32", dy, w, None, dx, B, K, O, LL(O), LL(1), LL(K), LL(1), LL(K), LL(1), None)
    call(lib, "ffc_gemm_f32", dy, x, None, dw, O, K, B, LL(1), LL(O), LL(K), LL(1), LL(K), LL(1), None)
    call(lib, "ffc_colsum_f32", dy, db, B, O, None)
    assert parity.relerr(y, x.double() @ w.double().T + b.double()) < 2e-6
    assert parity.relerr(dx, dy.double() @ w.double()) < 2e-6
    assert parity.relerr(dw, dy.double().T @ x.double()) < 2e-6
    assert parity.relerr(db, dy.double().sum(0)) < 2e-6


@pytest.mark.parametrize("decoupled", [1, 0])
def test_adam_step_flat_and_table_match_torch(lib, decoupled):
    """optim.AdamW (fgan_complete.py:315-319: lr 2e-4, betas (0.5, 0.999), torch's default decay 0.01) / optim.Adam
    (sngan_complete.py:247-248) over three steps: the flat kernel and the pointer-table kernel (one tensor without a
    gradient in step 2, which torch skips) against torch's own optimiser."""
    torch.manual_seed(decoupled)
    shapes = [(5000,), (3, 7), (64, 65), (1,)]
    ps = [torch.randn(s) for s in shapes]
    ref = [torch.nn.Parameter(p.clone()) for p in ps]
    wd = 0.01 if decoupled else 0.0
    opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)(ref, lr=2e-4, betas=(0.5, 0.999), weight_decay=wd)
    offs, n = [], 0
    for p in ps:
        offs.append(n); n += (p.numel() + 63) // 64 * 64
    flat = torch.zeros(n)
    for p, o in zip(ps, offs):
        flat[o:o + p.numel()] = p.flatten()
    flat2 = flat.clone()
    m, v, m2, v2 = torch.zeros(n), torch.zeros(n), torch.zeros(n), torch.zeros(n)
    lr, step, step2 = torch.tensor([2e-4]), torch.zeros(1), torch.zeros(len(ps))
    blk_t, blk_p = [], []
    for t, p in enumerate(ps):
        for piece in range((p.numel() + 4095) // 4096):
            blk_t.append(t); blk_p.append(piece)
    library, device = lib
    for it in range(3):
        grads = [torch.randn(s) for s in shapes]
        skip = 1 if it == 1 else None
        for i, (r, g) in enumerate(zip(ref, grads)):
            r.grad = None if i == skip else g.clone()
        opt.step()
        # flat kernel: a zero gradient is NOT a skipped tensor (moments decay, AdamW decays the weight), so use it on the full steps only
        gflat = torch.zeros(n)
        for g, o in zip(grads, offs):
            gflat[o:o + g.numel()] = g.flatten()
        if skip is None:
            call(lib, "ffc_adam_step", flat, gflat, m, v, LL(n), lr, step, ctypes.c_float(0.5), ctypes.c_float(0.999), ctypes.c_float(1e-8),
                 ctypes.c_float(wd), ctypes.c_float(1.0), decoupled, None)
        # table kernel on device-resident copies (the table holds device pointers)
        dev = [t.to(device) for t in (flat2, m2, v2, lr, step2)]
        gdev = [g.to(device).contiguous() for g in grads]
        ptrs = torch.tensor([0 if i == skip else g.data_ptr() for i, g in enumerate(gdev)], dtype=torch.int64, device=device)
        tabs = [torch.tensor(a, dtype=dt, device=device) for a, dt in ((offs, torch.int64), ([p.numel() for p in ps], torch.int64),
                                                                     (blk_t, torch.int32), (blk_p, torch.int32))]
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        rc = library.ffc_adam_step_table(P(dev[0]), P(dev[1]), P(dev[2]), P(ptrs), P(tabs[0]), P(tabs[1]), P(tabs[2]), P(tabs[3]), len(ps), len(blk_t),
                                         P(dev[3]), P(dev[4]), 0.5, 0.999, 1e-8, wd, 1.0, decoupled, None)
        assert rc == 0, library.last_error()
        if device != "cpu":
            torch.cuda.synchronize()
        for dst, src in zip((flat2, m2, v2, step2), (dev[0], dev[1], dev[2], dev[4])):
            dst.copy_(src)
    assert step2.tolist() == [3.0, 2.0, 3.0, 3.0]
    for r, o in zip(ref, offs):
        assert parity.relerr(flat2[o:o + r.numel()].view(r.shape), r.detach()) < 1e-6
    # the flat kernel saw steps 1 and 3 only: compare it with a torch optimiser that saw the same two
    assert step.item() == 2.0


def test_gather_table_packs_gradients(lib):
    library, device = lib
    gs = [torch.randn(5000, device=device), None, torch.randn(3, 3, device=device)]
    sizes, offs = [5000, 70, 9], [0, 5056, 5184]
    dst = torch.full((5248,), 7.0, device=device)
    blk_t, blk_p = [0, 0, 1, 2], [0, 1, 0, 0]
    tabs = [torch.tensor(a, dtype=dt, device=device) for a, dt in ((offs, torch.int64), (sizes, torch.int64), (blk_t, torch.int32), (blk_p, torch.int32))]
    ptrs = torch.tensor([g.data_ptr() if g is not None else 0 for g in gs], dtype=torch.int64, device=device)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    assert library.ffc_gather_table(P(dst), P(ptrs), P(tabs[0]), P(tabs[1]), P(tabs[2]), P(tabs[3]), 4, None) == 0
    if device != "cpu":
        torch.cuda.synchronize()
    dst = dst.cpu()
    assert torch.equal(dst[:5000], gs[0].cpu()) and torch.equal(dst[5056:5126], torch.zeros(70)) and torch.equal(dst[5184:5193], gs[2].cpu().flatten())
    assert torch.all(dst[5000:5056] == 7.0)         # padding between tensors is not touched
