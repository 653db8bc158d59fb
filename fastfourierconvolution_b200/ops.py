"""torch.autograd.Function wrappers around the C ABI of libffc_b200.so.

Each Function is one primitive of the FFC hot path with a hand-written forward and backward; the
nn.Modules in ``fastfourierconvolution_b200.layers`` compose them exactly where the reference
composes ATen ops.  PyTorch is used for memory, streams and the autograd tape only.

Every backward is ``once_differentiable``: the gradients are written by kernels into fresh tensors, so a double
backward (``create_graph=True``: gradient penalties such as benchmark_models/sagan/trainer.py:136) raises instead of
silently returning a cut graph.  The FFC hot path of the reference's own scripts (hinge losses) never needs it.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch.autograd.function import once_differentiable

from . import _C

ACT_IDENTITY, ACT_RELU, ACT_LEAKY, ACT_GELU, ACT_TANH, ACT_SIGMOID = range(6)
RESAMPLE_NONE, RESAMPLE_UP2, RESAMPLE_AVGPOOL2 = range(3)


def _c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.contiguous()


# ---------------------------------------------------------------------------------------------
# convolution (Conv2d / ConvTranspose2d, up to two summed segments, fused bias and addend)
# ---------------------------------------------------------------------------------------------
def conv_out_size(hi: int, k: int, stride: int, pad: int, transposed: bool, out_pad: int = 0) -> int:
    if transposed:
        return (hi - 1) * stride - 2 * pad + k + out_pad
    return (hi + 2 * pad - k) // stride + 1


def _conv_launch(x0, w0, cin0, x1, w1, cin1, bias, addend, y, B, cout, Hi, Wi, Ho, Wo, k, stride, pad, transposed):
    """ffc_conv2d_fwd_ws with the packed-weight scratch it needs (stream-ordered, reused by the next call)."""
    L = _C.lib()
    nbytes = L.ffc_conv2d_workspace_bytes(cin0, cin1, cout, k, stride, pad, int(transposed))
    ws = _C.workspace(nbytes, y.device)
    _C.check(L.ffc_conv2d_fwd_ws(_C.ptr(x0), _C.ptr(w0), cin0, _C.ptr(x1), _C.ptr(w1), cin1,
                                 _C.ptr(bias), _C.ptr(addend), _C.ptr(y),
                                 B, cout, Hi, Wi, Ho, Wo, k, stride, pad, int(transposed),
                                 _C.ptr(ws), ws.numel(), _C.current_stream(y.device)))


class Conv2dFn(torch.autograd.Function):
    """y = conv(x0, w0) [+ conv(x1, w1)] [+ bias] [+ addend]   (ffc.py:91-96, ffc_transpose.py:98-106)."""

    @staticmethod
    def forward(ctx, x0, w0, x1, w1, bias, addend, stride, pad, transposed, out_pad):
        _C.require_device(x0, w0, x1, w1, bias, addend)
        x0, w0, x1, w1, bias, addend = _c(x0), _c(w0), _c(x1), _c(w1), _c(bias), _c(addend)
        B, cin0, Hi, Wi = x0.shape
        k = w0.shape[-1]
        cout = w0.shape[1] if transposed else w0.shape[0]
        if (w0.shape[0] if transposed else w0.shape[1]) != cin0 or w0.shape[-2] != k:
            raise ValueError(f"weight {tuple(w0.shape)} does not match input channels {cin0} (groups=1, square kernels only)")
        cin1 = 0
        if x1 is not None:
            cin1 = x1.shape[1]
            if x1.shape[0] != B or x1.shape[2:] != x0.shape[2:] or w1.shape[-1] != k \
                    or (w1.shape[1] if transposed else w1.shape[0]) != cout:
                raise ValueError("second conv segment is inconsistent with the first")
        Ho = conv_out_size(Hi, k, stride, pad, transposed, out_pad)
        Wo = conv_out_size(Wi, k, stride, pad, transposed, out_pad)
        y = torch.empty((B, cout, Ho, Wo), device=x0.device, dtype=torch.float32)
        if addend is not None and addend.shape != y.shape:
            raise ValueError(f"addend {tuple(addend.shape)} != output {tuple(y.shape)}")
        _conv_launch(x0, w0, cin0, x1, w1, cin1, bias, addend, y, B, cout, Hi, Wi, Ho, Wo, k, stride, pad, transposed)
        ctx.save_for_backward(x0, w0, x1, w1)
        ctx.cfg = (stride, pad, bool(transposed), k, cout, bias is not None, addend is not None)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x0, w0, x1, w1 = ctx.saved_tensors
        return _conv_backward(ctx.cfg, x0, w0, x1, w1, dy, ctx.needs_input_grad)


_forked = {}          # device -> side streams forked from the current stream and not joined yet


class _Fork:
    """``with _Fork(device, i):`` runs the block on side stream ``i`` after everything already queued on the current stream;
    ``_join(device)`` makes the current stream wait for the forks made since the last join (only those: during a CUDA-graph
    capture a wait on a stream that holds uncaptured work would invalidate the capture).  Output tensors are allocated
    BEFORE the block (on the current stream's pool).  No-op on CPU tensors (host emulation), when the current stream IS
    the side stream (a backward node that autograd runs on its forward's side stream), and with FFC_B200_SINGLE_STREAM=1."""

    def __init__(self, device, index=0):
        self.ctx = None
        if device.type == "cuda":
            sides = _C.side_streams(device)
            if sides is not None:
                side = sides[index % len(sides)]
                cur = torch.cuda.current_stream(device)
                if side != cur:
                    side.wait_stream(cur)
                    self.ctx = torch.cuda.stream(side)
                    _forked.setdefault((str(device), cur.cuda_stream), []).append(side)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _join(device):
    if device.type == "cuda":
        cur = torch.cuda.current_stream(device)
        for s in _forked.pop((str(device), cur.cuda_stream), []):
            cur.wait_stream(s)


def fork_map(device, fns, nstreams: int = 4):
    """Runs every ``fn()`` of ``fns`` on a side stream (round robin over ``nstreams``), all of them after what the current stream
    holds now, and returns ``[(result, wait)]``: ``wait()`` makes the current stream wait for THAT fn only.  For work that depends
    on the parameters alone -- the spectral-norm power iterations of a discriminator's layers -- so that layer i's convolution
    waits for its own weight and not for the whole chain (the chain of seven layers is ~160 us, a batch-32 convolution ~20 us).
    Every ``wait`` must be called before a CUDA-graph capture of the caller ends (it is what joins the side streams)."""
    sides = _C.side_streams(device, nstreams) if device.type == "cuda" else None
    if not sides:
        return [(fn(), (lambda: None)) for fn in fns]
    cur = torch.cuda.current_stream(device)
    out, started = [], set()
    for j, fn in enumerate(fns):
        side = sides[j % nstreams]
        if side == cur:
            out.append((fn(), (lambda: None)))
            continue
        if id(side) not in started:
            side.wait_stream(cur)
            started.add(id(side))
        with torch.cuda.stream(side):
            r = fn()
            ev = torch.cuda.Event()
            ev.record(side)
        if torch.is_tensor(r) and not torch.cuda.is_current_stream_capturing():
            r.record_stream(cur)          # eager mode: allocated on the side stream's pool, consumed on the current stream
        out.append((r, (lambda ev=ev: cur.wait_event(ev))))
    return out


def _conv_backward(cfg, x0, w0, x1, w1, dy, needs):
    """Gradients of Conv2dFn's ten inputs (None where not needed).  The data gradients stay on the current stream; the weight
    and bias gradients -- independent of them -- are queued on side streams (fork / join)."""
    stride, pad, transposed, k, cout, has_bias, has_addend = cfg
    dy = dy.contiguous()
    B, _, Ho, Wo = dy.shape
    L = _C.lib()
    dev = dy.device
    grads = [None] * 10
    segs = [(slot, x, w) for slot, (x, w) in enumerate(((x0, w0), (x1, w1))) if x is not None]
    dws = {slot: torch.empty_like(w) for slot, x, w in segs if needs[2 * slot + 1]}
    dxs = {slot: torch.empty_like(x) for slot, x, w in segs if needs[2 * slot]}
    db = torch.empty(cout, device=dev, dtype=torch.float32) if (has_bias and needs[4]) else None
    if dws or db is not None:
        with _Fork(dev, 0):
            st = _C.current_stream(dev)
            for slot, x, w in segs:
                if slot not in dws:
                    continue
                cin, Hi, Wi = x.shape[1], x.shape[2], x.shape[3]
                if transposed:   # S = x (cin, Hi), L = dy (cout, Ho) -> [cin][cout][k][k]
                    _C.check(L.ffc_conv2d_wgrad(_C.ptr(x), _C.ptr(dy), _C.ptr(dws[slot]), B, cin, cout, Hi, Wi, Ho, Wo, k, stride, pad, st))
                else:            # S = dy, L = x -> [cout][cin][k][k]
                    _C.check(L.ffc_conv2d_wgrad(_C.ptr(dy), _C.ptr(x), _C.ptr(dws[slot]), B, cout, cin, Ho, Wo, Hi, Wi, k, stride, pad, st))
                grads[2 * slot + 1] = dws[slot]
            if db is not None:
                ws = _C.workspace(2 * cout * 8, dev)
                _C.check(L.ffc_bias_grad(_C.ptr(dy), _C.ptr(db), B, cout, Ho * Wo, _C.ptr(ws), ws.numel(), st))
                grads[4] = db
    for slot, x, w in segs:
        if slot in dxs:
            cin, Hi, Wi = x.shape[1], x.shape[2], x.shape[3]
            # data gradient = the opposite gather form over dy with the same weight tensor
            _conv_launch(dy, w, cout, None, None, 0, None, None, dxs[slot], B, cin, Ho, Wo, Hi, Wi, k, stride, pad, not transposed)
            grads[2 * slot] = dxs[slot]
    _join(dev)
    if has_addend and needs[5]:
        grads[5] = dy
    return tuple(grads)


def conv2d(x0, w0, x1=None, w1=None, bias=None, addend=None, stride=1, pad=0, transposed=False, out_pad=0):
    return Conv2dFn.apply(x0, w0, x1, w1, bias, addend, stride, pad, transposed, out_pad)


class ConvBlockFn(torch.autograd.Function):
    """The local branches of FFC.forward in one launch (ffc.py:91-96, ffc_transpose.py:98-106):
        y0 = conv(x0, w00) [+ conv(x1, w10)] [+ bias0]      convl2l(x_l) + convg2l(x_g)
        y1 = conv(x0, w01) [+ bias1]                         convl2g(x_l)
    On the tcgen05 path the operand gathered from x0 feeds one implicit GEMM over the concatenated output channels; in the
    backward, dx0 = conv^T(dy0, w00) + conv^T(dy1, w01) is one two-segment launch as well."""

    @staticmethod
    def forward(ctx, x0, w00, w01, x1, w10, bias0, bias1, stride, pad, transposed, out_pad):
        _C.require_device(x0, w00, w01, x1, w10, bias0, bias1)
        x0, w00, w01, x1, w10 = _c(x0), _c(w00), _c(w01), _c(x1), _c(w10)
        B, cin0, Hi, Wi = x0.shape
        k = w00.shape[-1]
        co_dim, ci_dim = (1, 0) if transposed else (0, 1)
        cout0, cout1 = w00.shape[co_dim], w01.shape[co_dim]
        if w00.shape[ci_dim] != cin0 or w01.shape[ci_dim] != cin0 or w01.shape[-1] != k or w00.shape[-2] != k:
            raise ValueError("conv2d_block: weights do not match the input channels (groups=1, square kernels only)")
        cin1 = 0
        if x1 is not None:
            cin1 = x1.shape[1]
            if x1.shape[0] != B or x1.shape[2:] != x0.shape[2:] or w10.shape[ci_dim] != cin1 or w10.shape[co_dim] != cout0 or w10.shape[-1] != k:
                raise ValueError("conv2d_block: second input segment is inconsistent with the first")
        Ho = conv_out_size(Hi, k, stride, pad, transposed, out_pad)
        Wo = conv_out_size(Wi, k, stride, pad, transposed, out_pad)
        y0 = torch.empty((B, cout0, Ho, Wo), device=x0.device, dtype=torch.float32)
        y1 = torch.empty((B, cout1, Ho, Wo), device=x0.device, dtype=torch.float32)
        bias = None
        if bias0 is not None or bias1 is not None:
            z = lambda n: torch.zeros(n, device=x0.device, dtype=torch.float32)
            bias = torch.cat([bias0 if bias0 is not None else z(cout0), bias1 if bias1 is not None else z(cout1)]).contiguous()
        L = _C.lib()
        nbytes = L.ffc_conv2d_workspace_bytes(cin0, cin1, cout0 + cout1, k, stride, pad, int(transposed))
        ws = _C.workspace(nbytes, y0.device)
        _C.check(L.ffc_conv2d_block_fwd_ws(_C.ptr(x0), _C.ptr(w00), _C.ptr(w01), cin0, _C.ptr(x1), _C.ptr(w10), cin1, _C.ptr(bias),
                                           _C.ptr(y0), cout0, _C.ptr(y1), cout1, B, Hi, Wi, Ho, Wo, k, stride, pad, int(transposed),
                                           _C.ptr(ws), ws.numel(), _C.current_stream(y0.device)))
        ctx.save_for_backward(x0, w00, w01, x1, w10)
        ctx.cfg = (stride, pad, bool(transposed), k, cout0, cout1, bias0 is not None, bias1 is not None)
        return y0, y1

    @staticmethod
    @once_differentiable
    def backward(ctx, dy0, dy1):
        x0, w00, w01, x1, w10 = ctx.saved_tensors
        stride, pad, transposed, k, cout0, cout1, has_b0, has_b1 = ctx.cfg
        dy0, dy1 = dy0.contiguous(), dy1.contiguous()
        B, _, Ho, Wo = dy0.shape
        cin0, Hi, Wi = x0.shape[1], x0.shape[2], x0.shape[3]
        L = _C.lib()
        need = ctx.needs_input_grad
        g = [None] * 11

        dev = dy0.device

        def wgrad(x, dy, dw, cin, cout, st):
            if transposed:
                _C.check(L.ffc_conv2d_wgrad(_C.ptr(x), _C.ptr(dy), _C.ptr(dw), B, cin, cout, Hi, Wi, Ho, Wo, k, stride, pad, st))
            else:
                _C.check(L.ffc_conv2d_wgrad(_C.ptr(dy), _C.ptr(x), _C.ptr(dw), B, cout, cin, Ho, Wo, Hi, Wi, k, stride, pad, st))

        def bgrad(dy, db, cout, st):
            ws = _C.workspace(2 * cout * 8, dev)
            _C.check(L.ffc_bias_grad(_C.ptr(dy), _C.ptr(db), B, cout, Ho * Wo, _C.ptr(ws), ws.numel(), st))

        # outputs are allocated on the current stream; the weight / bias gradients and the second segment's data gradient are
        # independent of dx0 and run on side streams (fork / join)
        cin1 = x1.shape[1] if x1 is not None else 0
        dw00 = torch.empty_like(w00) if need[1] else None
        dw01 = torch.empty_like(w01) if need[2] else None
        dx1 = torch.empty_like(x1) if (x1 is not None and need[3]) else None
        dw10 = torch.empty_like(w10) if (x1 is not None and need[4]) else None
        db0 = torch.empty(cout0, device=dev, dtype=torch.float32) if (has_b0 and need[5]) else None
        db1 = torch.empty(cout1, device=dev, dtype=torch.float32) if (has_b1 and need[6]) else None
        if dw00 is not None or dw01 is not None or dw10 is not None:
            with _Fork(dev, 0):
                st0 = _C.current_stream(dev)
                if dw00 is not None:
                    wgrad(x0, dy0, dw00, cin0, cout0, st0)
                if dw01 is not None:
                    wgrad(x0, dy1, dw01, cin0, cout1, st0)
                if dw10 is not None:
                    wgrad(x1, dy0, dw10, cin1, cout0, st0)
        if dx1 is not None or db0 is not None or db1 is not None:
            with _Fork(dev, 1):
                st1 = _C.current_stream(dev)
                if dx1 is not None:
                    _conv_launch(dy0, w10, cout0, None, None, 0, None, None, dx1, B, cin1, Ho, Wo, Hi, Wi, k, stride, pad, not transposed)
                if db0 is not None:
                    bgrad(dy0, db0, cout0, st1)
                if db1 is not None:
                    bgrad(dy1, db1, cout1, st1)
        if need[0]:     # both output blocks flow back into x0: one two-segment launch of the opposite gather form
            dx0 = torch.empty_like(x0)
            _conv_launch(dy0, w00, cout0, dy1, w01, cout1, None, None, dx0, B, cin0, Ho, Wo, Hi, Wi, k, stride, pad, not transposed)
            g[0] = dx0
        _join(dev)
        g[1], g[2], g[3], g[4], g[5], g[6] = dw00, dw01, dx1, dw10, db0, db1
        return tuple(g)


def conv2d_block(x0, w00, w01, x1=None, w10=None, bias0=None, bias1=None, stride=1, pad=0, transposed=False, out_pad=0):
    return ConvBlockFn.apply(x0, w00, w01, x1, w10, bias0, bias1, stride, pad, transposed, out_pad)


class ConvActFn(torch.autograd.Function):
    """a = act(conv(x, w) + bias) for act in {LeakyReLU, ReLU}: the SN-conv + LeakyReLU(0.1) stage of the fgan
    discriminators (fgan_complete.py:162-168).  Only the activated output is kept: both activations preserve the
    sign of their argument, so the backward mask is read from ``a`` itself."""

    @staticmethod
    def forward(ctx, x, w, bias, stride, pad, act, slope):
        if act not in (ACT_LEAKY, ACT_RELU) or (act == ACT_LEAKY and not slope > 0):
            raise ValueError("conv2d_act: LeakyReLU (slope > 0) or ReLU only")
        _C.require_device(x, w, bias)
        x, w, bias = _c(x), _c(w), _c(bias)
        B, cin, Hi, Wi = x.shape
        cout, k = w.shape[0], w.shape[-1]
        if w.shape[1] != cin or w.shape[-2] != k:
            raise ValueError(f"weight {tuple(w.shape)} does not match input channels {cin} (groups=1, square kernels only)")
        Ho, Wo = conv_out_size(Hi, k, stride, pad, False), conv_out_size(Wi, k, stride, pad, False)
        a = torch.empty((B, cout, Ho, Wo), device=x.device, dtype=torch.float32)
        L = _C.lib()
        nbytes = max(L.ffc_conv2d_workspace_bytes(cin, 0, cout, k, stride, pad, 0), 2 * cout * 8)
        ws = _C.workspace(nbytes, x.device)
        _C.check(L.ffc_conv2d_act_fwd_ws(_C.ptr(x), _C.ptr(w), cin, _C.ptr(bias), _C.ptr(a), B, cout, Hi, Wi, Ho, Wo, k, stride, pad,
                                         int(act), float(slope), _C.ptr(ws), ws.numel(), _C.current_stream(x.device)))
        ctx.save_for_backward(x, w, a)
        ctx.cfg = (stride, pad, False, k, cout, bias is not None, False)
        ctx.act = (int(act), float(slope))
        return a

    @staticmethod
    @once_differentiable
    def backward(ctx, da):
        x, w, a = ctx.saved_tensors
        act, slope = ctx.act
        da = da.contiguous()
        B, cout, Ho, Wo = a.shape
        dpre = torch.empty_like(a)
        L = _C.lib()
        ws = _C.workspace(2 * cout * 8, a.device)
        _C.check(L.ffc_bn_act_bwd(_C.ptr(a), _C.ptr(da), _C.ptr(dpre), None, None, None, None, None, None,
                                  B, cout, Ho * Wo, 0, 0, act, slope, _C.ptr(ws), ws.numel(), _C.current_stream(a.device)))
        n = ctx.needs_input_grad
        g = _conv_backward(ctx.cfg, x, w, None, None, dpre, (n[0], n[1], False, False, n[2], False))
        return g[0], g[1], g[4], None, None, None, None


def conv2d_act(x, w, bias=None, stride=1, pad=0, act=ACT_LEAKY, slope=0.1):
    return ConvActFn.apply(x, w, bias, stride, pad, act, slope)


# ---------------------------------------------------------------------------------------------
# spectral normalisation of a weight (torch.nn.utils.spectral_norm's pre-forward hook)
# ---------------------------------------------------------------------------------------------
class SpectralNormFn(torch.autograd.Function):
    """weight = weight_orig / sigma with sigma = u . (W v) after one power iteration in training mode
    (torch.nn.utils.spectral_norm.SpectralNorm.compute_weight, dim 0 or 1; layers/snffc/snffc.py:23-33,
    fgan_complete.py:147-156) in three kernels instead of ~14 PyTorch launches.  ``uv`` = (weight_u, weight_v) buffers,
    updated in place like the hook does; the vectors sigma was computed with are kept for the backward."""

    @staticmethod
    def forward(ctx, w_orig, uv, power_iteration, eps, dim):
        u, v = uv
        _C.require_device(w_orig, u, v)
        if dim not in (0, 1) or (dim == 1 and w_orig.dim() < 3):
            raise ValueError("spectral_norm: dim 0 (Conv2d / Linear) or dim 1 (ConvTranspose2d) only")
        w_orig = w_orig.contiguous()
        h = w_orig.shape[dim]
        wd = w_orig.numel() // h
        kk = w_orig[0, 0].numel() if dim == 1 else 0          # k * k
        if u.numel() != h or v.numel() != wd or not (u.is_contiguous() and v.is_contiguous()):
            raise ValueError("spectral_norm: weight_u / weight_v do not match weight_orig viewed as (out channels, rest)")
        w_eff = torch.empty_like(w_orig)
        u_s, v_s = torch.empty_like(u), torch.empty_like(v)
        sigma = torch.empty(1, device=w_orig.device, dtype=torch.float32)
        L = _C.lib()
        ws = _C.workspace(L.ffc_spectral_norm_workspace_bytes(h, wd), w_orig.device)
        _C.check(L.ffc_spectral_norm_fwd(_C.ptr(w_orig), _C.ptr(u), _C.ptr(v), _C.ptr(u_s), _C.ptr(v_s), _C.ptr(w_eff), _C.ptr(sigma),
                                         h, wd, kk, int(bool(power_iteration)), float(eps), _C.ptr(ws), ws.numel(),
                                         _C.current_stream(w_orig.device)))
        ctx.save_for_backward(w_orig, u_s, v_s, sigma)
        ctx.dim = dim
        return w_eff

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        w_orig, u_s, v_s, sigma = ctx.saved_tensors
        g = g.contiguous()
        h = w_orig.shape[ctx.dim]
        wd = w_orig.numel() // h
        kk = w_orig[0, 0].numel() if ctx.dim == 1 else 0
        dw = torch.empty_like(w_orig)                   # dW = g / sigma - (sum(g * W) / sigma^2) u v^T in two kernels
        ws = _C.workspace(64, w_orig.device)
        _C.check(_C.lib().ffc_spectral_norm_bwd(_C.ptr(g), _C.ptr(w_orig), _C.ptr(u_s), _C.ptr(v_s), _C.ptr(sigma), _C.ptr(dw),
                                                h, wd, kk, _C.ptr(ws), ws.numel(), _C.current_stream(w_orig.device)))
        return dw, None, None, None, None


def spectral_norm_weight(w_orig, u, v, power_iteration=True, eps=1e-12, dim=0):
    return SpectralNormFn.apply(w_orig, (u, v), power_iteration, eps, dim)


# ---------------------------------------------------------------------------------------------
# BatchNorm2d + activation
# ---------------------------------------------------------------------------------------------
class BnActFn(torch.autograd.Function):
    """y = act(BatchNorm2d(x)) or act(x)   (ffc_bn_act.py:73-81, spectral_transform.py:89, fourier_unity.py:49)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, norm, training, eps, momentum, act, slope):
        _C.require_device(x, gamma, beta, running_mean, running_var)
        x = x.contiguous()
        B, C = x.shape[0], x.shape[1]
        HW = x.numel() // max(B * C, 1)
        y = torch.empty_like(x)
        L = _C.lib()
        save_mean = save_invstd = None
        if norm:
            save_mean = torch.empty(C, device=x.device, dtype=torch.float32)
            save_invstd = torch.empty(C, device=x.device, dtype=torch.float32)
            if not training and (running_mean is None or running_var is None):
                raise ValueError("BatchNorm in eval mode needs running statistics")
        ws = _C.workspace(2 * C * 8, x.device)
        _C.check(L.ffc_bn_act_fwd(_C.ptr(x), _C.ptr(y), _C.ptr(gamma), _C.ptr(beta),
                                  _C.ptr(running_mean), _C.ptr(running_var), _C.ptr(save_mean), _C.ptr(save_invstd),
                                  B, C, HW, int(norm), int(training), float(eps), float(momentum), int(act), float(slope),
                                  _C.ptr(ws), ws.numel(), _C.current_stream(x.device)))
        # running_mean / running_var are updated in place by the kernel (buffers, never saved for
        # backward, so no version bump is needed)
        ctx.save_for_backward(x, gamma, beta, save_mean, save_invstd)
        ctx.cfg = (bool(norm), bool(training), int(act), float(slope))
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, gamma, beta, save_mean, save_invstd = ctx.saved_tensors
        norm, training, act, slope = ctx.cfg
        dy = dy.contiguous()
        B, C = x.shape[0], x.shape[1]
        HW = x.numel() // max(B * C, 1)
        dx = torch.empty_like(x)
        dgamma = dbeta = None
        if norm:
            dgamma = torch.empty_like(gamma)
            dbeta = torch.empty_like(beta)
        ws = _C.workspace(2 * C * 8, x.device)
        L = _C.lib()
        _C.check(L.ffc_bn_act_bwd(_C.ptr(x), _C.ptr(dy), _C.ptr(dx), _C.ptr(gamma), _C.ptr(beta),
                                  _C.ptr(save_mean), _C.ptr(save_invstd), _C.ptr(dgamma), _C.ptr(dbeta),
                                  B, C, HW, int(norm), int(training), act, slope,
                                  _C.ptr(ws), ws.numel(), _C.current_stream(x.device)))
        return dx, dgamma, dbeta, None, None, None, None, None, None, None, None


def bn_act(x, gamma=None, beta=None, running_mean=None, running_var=None, norm=False, training=True,
           eps=1e-5, momentum=0.1, act=ACT_IDENTITY, slope=0.1):
    return BnActFn.apply(x, gamma, beta, running_mean, running_var, norm, training, eps, momentum, act, slope)


# ---------------------------------------------------------------------------------------------
# rfft2 / irfft2 on planar (B, 2C, H, Wf) spectra
# ---------------------------------------------------------------------------------------------
def _rfft2(x, colscale):
    B, C, H, W = x.shape
    spec = torch.empty((B, 2 * C, H, W // 2 + 1), device=x.device, dtype=torch.float32)
    _C.check(_C.lib().ffc_rfft2(_C.ptr(x), _C.ptr(spec), B * C, H, W, colscale, _C.current_stream(x.device)))
    return spec


def _irfft2(spec, residual, colscale, width=None):
    B, C2, H, Wf = spec.shape
    W = 2 * (Wf - 1) if width is None else int(width)          # an odd width cannot be told from the spectrum (irfftn's s=)
    if W // 2 + 1 != Wf:
        raise ValueError(f"irfft2: width {W} does not match a spectrum of {Wf} columns")
    out = torch.empty((B, C2 // 2, H, W), device=spec.device, dtype=torch.float32)
    _C.check(_C.lib().ffc_irfft2(_C.ptr(spec), _C.ptr(residual), _C.ptr(out), B * (C2 // 2), H, W, colscale,
                                 _C.current_stream(spec.device)))
    return out


class Rfft2Fn(torch.autograd.Function):
    """rfftn(norm='ortho') + re/im interleave into channels (fourier_unity.py:38-42)."""

    @staticmethod
    def forward(ctx, x):
        _C.require_device(x)
        ctx.width = x.shape[-1]
        return _rfft2(x.contiguous(), 0)

    @staticmethod
    @once_differentiable
    def backward(ctx, dspec):
        return _irfft2(dspec.contiguous(), None, 1, ctx.width)


class Irfft2Fn(torch.autograd.Function):
    """channels -> complex + irfftn(s=(H,W), norm='ortho') [+ residual] (fourier_unity.py:51-56)."""

    @staticmethod
    def forward(ctx, spec, residual, width=None):
        _C.require_device(spec, residual)
        ctx.has_res = residual is not None
        return _irfft2(spec.contiguous(), _c(residual), 0, width)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        dout = dout.contiguous()
        dspec = _rfft2(dout, 1) if ctx.needs_input_grad[0] else None
        return dspec, (dout if ctx.has_res and ctx.needs_input_grad[1] else None), None


class BnReluIrfft2Fn(torch.autograd.Function):
    """irfft2(relu(batchnorm(spec))) [+ residual] with the BatchNorm + ReLU applied as the spectrum is loaded
    (fourier_unity.py:49 folded into :51-56): statistics kernel, then one kernel; the normalised spectrum is never
    written.  Backward: adjoint transform of dout, then the ordinary BN + ReLU backward on the saved spectrum."""

    @staticmethod
    def forward(ctx, spec, gamma, beta, running_mean, running_var, residual, training, eps, momentum, width=None):
        _C.require_device(spec, gamma, beta, running_mean, running_var, residual)
        spec, residual = spec.contiguous(), _c(residual)
        B, C2, H, Wf = spec.shape
        W = 2 * (Wf - 1) if width is None else int(width)
        if W // 2 + 1 != Wf:
            raise ValueError(f"bn_relu_irfft2: width {W} does not match a spectrum of {Wf} columns")
        if not training and (running_mean is None or running_var is None):
            raise ValueError("BatchNorm in eval mode needs running statistics")
        L = _C.lib()
        st = _C.current_stream(spec.device)
        save_mean = torch.empty(C2, device=spec.device, dtype=torch.float32)
        save_invstd = torch.empty(C2, device=spec.device, dtype=torch.float32)
        ws = _C.workspace(2 * C2 * 8, spec.device)
        _C.check(L.ffc_bn_stats(_C.ptr(spec), _C.ptr(running_mean), _C.ptr(running_var), _C.ptr(save_mean), _C.ptr(save_invstd),
                                B, C2, H * Wf, int(training), float(eps), float(momentum), _C.ptr(ws), ws.numel(), st))
        out = torch.empty((B, C2 // 2, H, W), device=spec.device, dtype=torch.float32)
        _C.check(L.ffc_irfft2_bn_relu(_C.ptr(spec), _C.ptr(residual), _C.ptr(out), B * (C2 // 2), C2 // 2, H, W,
                                      _C.ptr(save_mean), _C.ptr(save_invstd), _C.ptr(gamma), _C.ptr(beta), st))
        ctx.save_for_backward(spec, gamma, beta, save_mean, save_invstd)
        ctx.cfg = (bool(training), residual is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        spec, gamma, beta, save_mean, save_invstd = ctx.saved_tensors
        training, has_res = ctx.cfg
        dout = dout.contiguous()
        B, C2, H, Wf = spec.shape
        dact = _rfft2(dout, 1)                                  # gradient with respect to relu(bn(spec))
        dspec = torch.empty_like(spec)
        dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(beta)
        ws = _C.workspace(2 * C2 * 8, spec.device)
        _C.check(_C.lib().ffc_bn_act_bwd(_C.ptr(spec), _C.ptr(dact), _C.ptr(dspec), _C.ptr(gamma), _C.ptr(beta),
                                         _C.ptr(save_mean), _C.ptr(save_invstd), _C.ptr(dgamma), _C.ptr(dbeta),
                                         B, C2, H * Wf, 1, int(training), ACT_RELU, 0.0,
                                         _C.ptr(ws), ws.numel(), _C.current_stream(spec.device)))
        return dspec, dgamma, dbeta, None, None, (dout if has_res and ctx.needs_input_grad[5] else None), None, None, None, None


def bn_relu_irfft2(spec, gamma, beta, running_mean, running_var, residual, training, eps, momentum, width=None):
    return BnReluIrfft2Fn.apply(spec, gamma, beta, running_mean, running_var, residual, training, eps, momentum, width)


def rfft2(x):
    return Rfft2Fn.apply(x)


def irfft2(spec, residual=None, width=None):
    return Irfft2Fn.apply(spec, residual, width)


def fft2_supported(H, W) -> int:
    """2: tuned power-of-two plane kernels, 1: direct DFT kernels (any H, W up to 128), 0: unsupported."""
    return int(_C.lib().ffc_fft2_supported(int(H), int(W)))


# ---------------------------------------------------------------------------------------------
# fused Fourier unit (forward in one / two kernels; backward recomputes through the general form)
# ---------------------------------------------------------------------------------------------
FUSED_FU_BACKWARD = True      # tests switch this off to exercise the general-form backward of FusedFuFn


def fu_fused_supported(B, Cin, Cout, H, W) -> bool:
    return bool(_C.lib().ffc_fu_fused_supported(int(B), int(Cin), int(Cout), int(H), int(W)))


def fu_staged_supported(B, Cin, Cout, H, W) -> bool:
    """The L2-staged form (csrc/ffc_fu3.cu): 16..128 planes, channel counts whose packed mix weights fit shared memory."""
    return bool(_C.lib().ffc_fu3_supported(int(B), int(Cin), int(Cout), int(H), int(W)))


class FusedFuFn(torch.autograd.Function):
    """out = [residual +] irfft2(relu(bn(conv1x1(rfft2(x)))))   (fourier_unity.py:32-58 in one op).

    Forward: ffc_fu_fwd (spectrum stays in shared memory).  Backward: ffc_fu_bwd, one cooperative kernel that
    recomputes the spectrum of x, transforms dout and keeps both in shared memory; when the batch is too large for
    one resident CTA per image the general-form kernels are used (rfft2, 1x1 mix, BN+ReLU backward, wgrad/dgrad,
    irfft2).  Nothing spectral is saved between forward and backward."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, running_mean, running_var, residual, training, eps, momentum, staged=False):
        _C.require_device(x, weight, gamma, beta, running_mean, running_var, residual)
        x, weight, residual = x.contiguous(), weight.contiguous(), _c(residual)
        B, Cin, H, W = x.shape
        Cout = weight.shape[0] // 2
        out = torch.empty((B, Cout, H, W), device=x.device, dtype=torch.float32)
        save_mean = torch.empty(2 * Cout, device=x.device, dtype=torch.float32)
        save_invstd = torch.empty(2 * Cout, device=x.device, dtype=torch.float32)
        L = _C.lib()
        st = _C.current_stream(x.device)
        keep = staged and any(ctx.needs_input_grad[:4])
        if keep:
            # planes / channel counts beyond one CTA's shared memory, with a backward to follow: both spectra are kept in
            # the library's plane layout, so the backward (ffc_fu3_bwd) recomputes nothing
            s_keep = torch.empty((B, Cin, H, W + 4), device=x.device, dtype=torch.float32)
            y_keep = torch.empty((B, Cout, H, W + 4), device=x.device, dtype=torch.float32)
            ws = _C.workspace(L.ffc_fu3_workspace_bytes(B, Cin, Cout, H, W, int(training)), x.device)
            _C.check(L.ffc_fu3_fwd_keep(_C.ptr(x), _C.ptr(weight), _C.ptr(gamma), _C.ptr(beta),
                                        _C.ptr(running_mean), _C.ptr(running_var), _C.ptr(save_mean), _C.ptr(save_invstd),
                                        _C.ptr(residual), _C.ptr(out), _C.ptr(s_keep), _C.ptr(y_keep), B, Cin, Cout, H, W,
                                        int(training), float(eps), float(momentum), _C.ptr(ws), ws.numel(), st))
            ctx.save_for_backward(x, weight, gamma, beta, save_mean, save_invstd, s_keep, y_keep)
            ctx.cfg = (bool(training), residual is not None)
            return out
        if staged:       # spectrum staged through L2 in chunks of images
            ws = _C.workspace(L.ffc_fu3_workspace_bytes(B, Cin, Cout, H, W, int(training)), x.device)
            fwd = L.ffc_fu3_fwd
        else:
            ws = _C.workspace(L.ffc_fu_workspace_bytes(B, Cout), x.device)
            fwd = L.ffc_fu_fwd
        _C.check(fwd(_C.ptr(x), _C.ptr(weight), _C.ptr(gamma), _C.ptr(beta),
                                     _C.ptr(running_mean), _C.ptr(running_var), _C.ptr(save_mean), _C.ptr(save_invstd),
                                     _C.ptr(residual), _C.ptr(out), B, Cin, Cout, H, W, int(training), float(eps),
                                     float(momentum), _C.ptr(ws), ws.numel(), st))
        ctx.save_for_backward(x, weight, gamma, beta, save_mean, save_invstd)
        ctx.cfg = (bool(training), residual is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x, weight, gamma, beta, save_mean, save_invstd = ctx.saved_tensors[:6]
        training, has_res = ctx.cfg
        dout = dout.contiguous()
        B, Cin, H, W = x.shape
        C2o, C2i = weight.shape[0], weight.shape[1]
        Wf = W // 2 + 1
        L = _C.lib()
        st = _C.current_stream(x.device)
        if len(ctx.saved_tensors) == 8 and FUSED_FU_BACKWARD:
            # L2-staged form: adjoint transform + ReLU mask + BN sums | constants | dY + dW | dS (tensor cores) | adjoint transform
            s_keep, y_keep = ctx.saved_tensors[6:]
            dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
            dw = torch.empty_like(weight)
            dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(beta)
            ws = _C.workspace(L.ffc_fu3_bwd_workspace_bytes(B, Cin, C2o // 2, H, W), x.device)
            _C.check(L.ffc_fu3_bwd(_C.ptr(dout), _C.ptr(s_keep), _C.ptr(y_keep), _C.ptr(weight), _C.ptr(gamma), _C.ptr(beta),
                                   _C.ptr(save_mean), _C.ptr(save_invstd), _C.ptr(dx), _C.ptr(dw), _C.ptr(dgamma), _C.ptr(dbeta),
                                   B, Cin, C2o // 2, H, W, int(training), _C.ptr(ws), ws.numel(), st))
            dres = dout if (has_res and ctx.needs_input_grad[6]) else None
            return dx, dw, dgamma, dbeta, None, None, dres, None, None, None, None
        if FUSED_FU_BACKWARD and L.ffc_fu_bwd_supported(B, Cin, C2o // 2, H, W):
            # one cooperative kernel: both spectra stay in shared memory (csrc/ffc_fu2_bwd.cu)
            dx, dw = torch.empty_like(x), torch.empty_like(weight)
            dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(beta)
            ws = _C.workspace(L.ffc_fu_bwd_workspace_bytes(B, Cin, C2o // 2), x.device)
            _C.check(L.ffc_fu_bwd(_C.ptr(x), _C.ptr(dout), _C.ptr(weight), _C.ptr(gamma), _C.ptr(beta),
                                  _C.ptr(save_mean), _C.ptr(save_invstd), _C.ptr(dx), _C.ptr(dw), _C.ptr(dgamma), _C.ptr(dbeta),
                                  B, Cin, C2o // 2, H, W, int(training), _C.ptr(ws), ws.numel(), st))
            dres = dout if (has_res and ctx.needs_input_grad[6]) else None
            return dx, dw, dgamma, dbeta, None, None, dres, None, None, None, None
        w4 = weight.view(C2o, C2i, 1, 1)
        spec = _rfft2(x, 0)                                            # recompute S
        y = torch.empty((B, C2o, H, Wf), device=x.device, dtype=torch.float32)
        _C.check(L.ffc_conv2d_fwd(_C.ptr(spec), _C.ptr(w4), C2i, None, None, 0, None, None, _C.ptr(y),
                                  B, C2o, H, Wf, H, Wf, 1, 1, 0, 0, st))         # recompute Y = W S
        g = _rfft2(dout, 1)                                            # adjoint of irfft2
        dy = torch.empty_like(y)
        dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(beta)
        ws = _C.workspace(2 * C2o * 8, x.device)
        _C.check(L.ffc_bn_act_bwd(_C.ptr(y), _C.ptr(g), _C.ptr(dy), _C.ptr(gamma), _C.ptr(beta),
                                  _C.ptr(save_mean), _C.ptr(save_invstd), _C.ptr(dgamma), _C.ptr(dbeta),
                                  B, C2o, H * Wf, 1, int(training), ACT_RELU, 0.0, _C.ptr(ws), ws.numel(), st))
        dx = dw = None
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(weight)
            _C.check(L.ffc_conv2d_wgrad(_C.ptr(dy), _C.ptr(spec), _C.ptr(dw), B, C2o, C2i, H, Wf, H, Wf, 1, 1, 0, st))
        if ctx.needs_input_grad[0]:
            dspec = torch.empty_like(spec)
            _C.check(L.ffc_conv2d_fwd(_C.ptr(dy), _C.ptr(w4), C2o, None, None, 0, None, None, _C.ptr(dspec),
                                      B, C2i, H, Wf, H, Wf, 1, 1, 0, 1, st))     # dS = W^T dY
            dx = _irfft2(dspec, None, 1)                               # adjoint of rfft2
        dres = dout if (has_res and ctx.needs_input_grad[6]) else None
        return dx, dw, dgamma, dbeta, None, None, dres, None, None, None, None


def fourier_unit_fused(x, weight2d, gamma, beta, running_mean, running_var, residual, training, eps, momentum, staged=False):
    return FusedFuFn.apply(x, weight2d, gamma, beta, running_mean, running_var, residual, training, eps, momentum, staged)


# ---------------------------------------------------------------------------------------------
# resample + SE gate (SpectralTransform prologue)
# ---------------------------------------------------------------------------------------------
class SeFn(torch.autograd.Function):
    """y = r(x) * sigmoid(W2 relu(W1 mean(r(x))))   (spectral_transform.py:79, 87, 12-28)."""

    @staticmethod
    def forward(ctx, x, w1, w2, mode):
        _C.require_device(x, w1, w2)
        x, w1, w2 = x.contiguous(), w1.contiguous(), w2.contiguous()
        B, C, Hi, Wi = x.shape
        hid = w1.shape[0]
        Ho, Wo = (Hi * 2, Wi * 2) if mode == RESAMPLE_UP2 else ((Hi // 2, Wi // 2) if mode == RESAMPLE_AVGPOOL2 else (Hi, Wi))
        y = torch.empty((B, C, Ho, Wo), device=x.device, dtype=torch.float32)
        mean = torch.empty((B, C), device=x.device, dtype=torch.float32)
        hidden = torch.empty((B, max(hid, 1)), device=x.device, dtype=torch.float32)
        gate = torch.empty((B, C), device=x.device, dtype=torch.float32)
        ws = _C.workspace(B * C * 8, x.device)
        _C.check(_C.lib().ffc_se_fwd(_C.ptr(x), _C.ptr(w1) if hid else None, _C.ptr(w2) if hid else None, _C.ptr(y),
                                     _C.ptr(mean), _C.ptr(hidden), _C.ptr(gate), B, C, hid, Hi, Wi, int(mode),
                                     _C.ptr(ws), ws.numel(), _C.current_stream(x.device)))
        ctx.save_for_backward(x, w1, w2, mean, hidden, gate)
        ctx.mode = int(mode)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, w1, w2, mean, hidden, gate = ctx.saved_tensors
        dy = dy.contiguous()
        B, C, Hi, Wi = x.shape
        hid = w1.shape[0]
        dx = torch.empty_like(x)
        dw1 = torch.empty_like(w1)
        dw2 = torch.empty_like(w2)
        ws = _C.workspace((3 * B * C + B * hid) * 4, x.device)
        _C.check(_C.lib().ffc_se_bwd(_C.ptr(x), _C.ptr(dy), _C.ptr(w1) if hid else None, _C.ptr(w2) if hid else None,
                                     _C.ptr(mean), _C.ptr(hidden), _C.ptr(gate), _C.ptr(dx),
                                     _C.ptr(dw1) if hid else None, _C.ptr(dw2) if hid else None,
                                     B, C, hid, Hi, Wi, ctx.mode, _C.ptr(ws), ws.numel(), _C.current_stream(x.device)))
        return dx, dw1, dw2, None


def se_resample(x, w1, w2, mode=RESAMPLE_NONE):
    return SeFn.apply(x, w1, w2, mode)


# ---------------------------------------------------------------------------------------------
# glue of the generators: NoiseInjection, uint8 image epilogue
# ---------------------------------------------------------------------------------------------
class NoiseAddFn(torch.autograd.Function):
    """x + weight * noise with one noise plane per image (layers/noise_injection.py:20-32) in one kernel; the weight gradient
    sum(dy * noise) in one more (the activation gradient is dy itself)."""

    @staticmethod
    def forward(ctx, x, weight, noise):
        _C.require_device(x, weight, noise)
        x, noise = x.contiguous(), noise.contiguous()
        B, C = x.shape[0], x.shape[1]
        HW = x.numel() // max(B * C, 1)
        if weight.numel() != C or noise.numel() != B * HW:
            raise ValueError("noise_add: weight must have one entry per channel and noise one plane per image")
        out = torch.empty_like(x)
        _C.check(_C.lib().ffc_noise_add_fwd(_C.ptr(x), _C.ptr(weight.contiguous()), _C.ptr(noise), _C.ptr(out), B, C, HW, _C.current_stream(x.device)))
        ctx.save_for_backward(noise)
        ctx.shape = (B, C, HW, tuple(weight.shape))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (noise,) = ctx.saved_tensors
        B, C, HW, wshape = ctx.shape
        dw = None
        if ctx.needs_input_grad[1]:
            dy = dy.contiguous()
            dw = torch.empty(wshape, device=dy.device, dtype=torch.float32)
            _C.check(_C.lib().ffc_noise_add_bwd_w(_C.ptr(dy), _C.ptr(noise), _C.ptr(dw), B, C, HW, _C.current_stream(dy.device)))
        return (dy if ctx.needs_input_grad[0] else None), dw, None


def noise_add(x, weight, noise):
    return NoiseAddFn.apply(x, weight, noise)


def to_uint8(x, lo=-1.0, hi=1.0):
    """uint8(255 * (clamp(x, lo, hi) * 0.5 + 0.5)) in one kernel (fgan_complete.py:136-138); lo > hi: no clamp."""
    _C.require_device(x)
    x = x.contiguous()
    out = torch.empty(x.shape, device=x.device, dtype=torch.uint8)
    _C.check(_C.lib().ffc_to_uint8(_C.ptr(x), _C.ptr(out), x.numel(), float(lo), float(hi), _C.current_stream(x.device)))
    return out


# ---------------------------------------------------------------------------------------------
# Linear layers (generator stem, discriminator head) and the optimiser step
# ---------------------------------------------------------------------------------------------
def _gemm(A, B, bias, M, N, K, sa, sb, out):
    """out (M x N, dense row-major) = A (M x K, element strides sa) @ B (K x N, element strides sb) [+ bias]."""
    _C.check(_C.lib().ffc_gemm_f32(_C.ptr(A), _C.ptr(B), _C.ptr(bias), _C.ptr(out), M, N, K, sa[0], sa[1], sb[0], sb[1], N, 1,
                                   _C.current_stream(out.device)))
    return out


class LinearFn(torch.autograd.Function):
    """y = x W^T + b (nn.Linear: fgan_complete.py:92-95 stem, :160-170 head) and its three gradients on one FP32 kernel."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        _C.require_device(x, weight, bias)
        x, weight = x.contiguous(), weight.contiguous()
        Bn, K = x.shape
        O = weight.shape[0]
        if weight.shape[1] != K:
            raise ValueError(f"linear: x has {K} features, weight expects {weight.shape[1]}")
        out = torch.empty((Bn, O), device=x.device, dtype=torch.float32)
        _gemm(x, weight, _c(bias), Bn, O, K, (K, 1), (1, K), out)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        Bn, K = x.shape
        O = weight.shape[0]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _gemm(dy, weight, None, Bn, K, O, (O, 1), (K, 1), torch.empty_like(x))                # dy W
        if ctx.needs_input_grad[1]:
            dw = _gemm(dy, x, None, O, K, Bn, (1, O), (K, 1), torch.empty_like(weight))                # dy^T x
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.empty(O, device=dy.device, dtype=torch.float32)
            _C.check(_C.lib().ffc_colsum_f32(_C.ptr(dy), _C.ptr(db), Bn, O, _C.current_stream(dy.device)))
        return dx, dw, db


def linear(x, weight, bias=None):
    return LinearFn.apply(x, weight, bias)


def adam_step(p, g, m, v, lr, step, beta1, beta2, eps, weight_decay, grad_scale=1.0, decoupled=True):
    """One AdamW (decoupled) / Adam step over flat FP32 buffers; ``lr`` and ``step`` are one-element device tensors
    (``step`` is incremented by the kernel)."""
    _C.require_device(p, g, m, v, lr, step)
    _C.check(_C.lib().ffc_adam_step(_C.ptr(p), _C.ptr(g), _C.ptr(m), _C.ptr(v), p.numel(), _C.ptr(lr), _C.ptr(step),
                                    float(beta1), float(beta2), float(eps), float(weight_decay), float(grad_scale),
                                    int(bool(decoupled)), _C.current_stream(p.device)))
