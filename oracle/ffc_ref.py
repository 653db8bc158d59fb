"""CPU oracle for the FFC hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional restatement (plain functions over a flat ``{state_dict key: tensor}``
parameter dictionary) of the reference's Fast-Fourier-Convolution layer stack.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module; the product package
``fastfourierconvolution_b200`` never does.

Every function cites the reference lines it follows (paths relative to the reference
checkout).  The floating-point primitives (FFT, convolution, batch-norm) are the
reference's own third-party dependency -- PyTorch (``requirements.txt:5`` pins
torch==1.10.2; 2.11.0 is what is installed) -- called at the same call sites as the
reference calls them, so on CPU this oracle executes the same library kernels as the
reference does and doubles as the CPU baseline.  ``oracle/fu_dft.py`` holds an
independent explicit-DFT float64 restatement of the Fourier unit that does not use
``torch.fft`` at all.

Parity pinning: the reference ships no tests or golden vectors, so
this oracle is pinned against outputs of the reference itself, generated in the build
container by ``tests/golden/make_golden.py`` (which imports ``/root/reference``) and
committed under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them.

Works for float32 and float64 (dtype follows the inputs / parameters).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple, Union

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]
Branch = Union[Tensor, int]


# --------------------------------------------------------------------------------------
# activation tape (for the parity tests' mask-flip analysis, SURVEY.md section 8(c) caveat 1)
# --------------------------------------------------------------------------------------
class ActTape:
    """Records the pre-activation of every ReLU / LeakyReLU site of one oracle run, in call order, and lets a later run
    use a prescribed 0/1 mask at chosen sites instead of ``x > 0``.

    Two FP32-accurate evaluations of the same network round a pre-activation that is ~0 to different sides of the
    kink now and then; the forward value does not notice, the gradient does (one flipped element of ~400k moved dX by
    2e-2).  The tests use this tape to (1) list the elements close enough to the kink to flip and (2) obtain the exact
    gradient change each such flip causes, so that the 1e-4 bound can be held on everything else.

        with ActTape() as t:            # record
            out = fgenerator(z, P, True)
        with ActTape({3: mask}) as t2:  # replay with site 3's mask prescribed (bool tensor of the site's shape)
            ...
    """
    current: "Optional[ActTape]" = None

    def __init__(self, overrides=None):
        self.overrides = dict(overrides or {})
        self.pre = []                     # detached pre-activations, one per site

    def __enter__(self):
        assert ActTape.current is None, "ActTape is not re-entrant"
        ActTape.current = self
        return self

    def __exit__(self, *exc):
        ActTape.current = None
        return False


def _kinked(x: Tensor, slope: float) -> Tensor:
    """ReLU (slope 0) / LeakyReLU(slope): every non-smooth activation of the oracle goes through here."""
    t = ActTape.current
    if t is None:
        return F.relu(x) if slope == 0.0 else F.leaky_relu(x, slope)
    idx = len(t.pre)
    t.pre.append(x.detach())
    mask = t.overrides.get(idx)
    if mask is None:
        mask = x.detach() > 0
    m = mask.to(x.dtype)
    return x * (m + slope * (1.0 - m)) if slope != 0.0 else x * m


# --------------------------------------------------------------------------------------
# configuration records
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class FFCConfig:
    """Constructor arguments of FFC / FFCTranspose / FFC_BN_ACT that shape the math.

    Mirrors ``layers/ffc/ffc_bn_act.py:25-31`` (and ``ffc.py:21-24``,
    ``ffc_transpose.py:19-22``).  ``norm`` is "bn" or "identity", ``act`` one of
    "identity", "relu", "leaky_relu" (slope 0.1, ``ffc_bn_act.py:66-67``), "gelu"
    (erf form, the ``nn.GELU`` default), "tanh", "sigmoid".
    """
    in_channels: int
    out_channels: int
    kernel_size: int
    ratio_gin: float
    ratio_gout: float
    stride: int = 1
    padding: int = 0
    bias: bool = False
    norm: str = "identity"
    act: str = "identity"
    enable_lfu: bool = True
    upsampling: bool = False
    out_padding: int = 0
    spectral_norm: bool = False   # SNFFC / SNFFCTranspose twins (layers/snffc/*.py)

    # channel split, ffc.py:33-36 / ffc_transpose.py:37-40
    @property
    def in_cg(self) -> int:
        return int(self.in_channels * self.ratio_gin)

    @property
    def in_cl(self) -> int:
        return self.in_channels - self.in_cg

    @property
    def out_cg(self) -> int:
        return int(self.out_channels * self.ratio_gout)

    @property
    def out_cl(self) -> int:
        return self.out_channels - self.out_cg


# --------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------
def batch_norm(x: Tensor, P: Params, pre: str, training: bool,
               momentum: float = 0.1, eps: float = 1e-5) -> Tensor:
    """``nn.BatchNorm2d`` forward incl. running-stat update (torch _BatchNorm.forward)."""
    rm, rv = P.get(pre + "running_mean"), P.get(pre + "running_var")
    if training and (pre + "num_batches_tracked") in P:
        P[pre + "num_batches_tracked"] += 1
    return F.batch_norm(x, rm, rv, P[pre + "weight"], P[pre + "bias"], training, momentum, eps)


def activation(x: Tensor, act: str) -> Tensor:
    if act == "identity":
        return x
    if act == "relu":
        return _kinked(x, 0.0)
    if act == "leaky_relu":            # ffc_bn_act.py:66-67 -> LeakyReLU(0.1)
        return _kinked(x, 0.1)
    if act == "gelu":
        return F.gelu(x)
    if act == "tanh":
        return torch.tanh(x)
    if act == "sigmoid":
        return torch.sigmoid(x)
    raise ValueError(act)


def spectral_norm_weight(P: Params, pre: str, training: bool, dim: int = 0,
                         eps: float = 1e-12, n_power_iterations: int = 1) -> Tensor:
    """``torch.nn.utils.spectral_norm`` as used at ``layers/snffc/snffc.py:23-33``.

    ``weight = weight_orig / sigma`` with ``sigma = u^T W v`` after one power iteration
    per training forward (u, v updated in place in ``P``; no update in eval mode).
    ``dim`` is 0 for Conv2d/Linear and 1 for ConvTranspose2d (torch's default rule).
    """
    w = P[pre + "weight_orig"]
    u, v = P[pre + "weight_u"], P[pre + "weight_v"]
    wm = w
    if dim != 0:
        wm = w.permute(dim, *[d for d in range(w.dim()) if d != dim])
    wm = wm.reshape(wm.size(0), -1)
    if training:
        with torch.no_grad():
            for _ in range(n_power_iterations):
                v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=eps))
                u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=eps))
        u, v = u.clone(), v.clone()
    sigma = torch.dot(u, torch.mv(wm, v))
    return w / sigma


def _weight(P: Params, pre: str, training: bool, sn: bool, dim: int = 0) -> Tensor:
    if sn and (pre + "weight_orig") in P:
        return spectral_norm_weight(P, pre, training, dim)
    return P[pre + "weight"]


# --------------------------------------------------------------------------------------
# row a2: FourierUnitSN.forward  (layers/ffc/fourier_unity.py:32-58)
# --------------------------------------------------------------------------------------
def fourier_unit(x: Tensor, P: Params, pre: str, training: bool) -> Tensor:
    b, c, h, w = x.shape
    # :38  rfft2, ortho
    s = torch.fft.rfftn(x, dim=(-2, -1), norm="ortho")                  # (b,c,h,wf) complex
    # :40-42  channel 2c+0 = Re, 2c+1 = Im
    s = torch.stack((s.real, s.imag), dim=2).reshape(b, 2 * c, h, s.shape[-1])
    # :45  1x1 channel mix, no bias
    y = F.conv2d(s, P[pre + "conv_layer.weight"])
    # :49  BN (batch stats over b,h,wf in training) + ReLU
    y = _kinked(batch_norm(y, P, pre + "bn.", training), 0.0)
    # :51-53  back to complex
    y = y.reshape(b, -1, 2, h, y.shape[-1])
    yc = torch.complex(y[:, :, 0].contiguous(), y[:, :, 1].contiguous())
    # :56  irfft2 to the input's spatial size, ortho
    return torch.fft.irfftn(yc, s=(h, w), dim=(-2, -1), norm="ortho")


# --------------------------------------------------------------------------------------
# row a4: SELayer.forward  (layers/ffc/spectral_transform.py:23-28)
# --------------------------------------------------------------------------------------
def se_layer(x: Tensor, P: Params, pre: str) -> Tensor:
    b, c = x.shape[:2]
    y = x.mean(dim=(2, 3))                                             # AdaptiveAvgPool2d(1)
    y = _kinked(F.linear(y, P[pre + "fc.0.weight"]), 0.0)
    y = torch.sigmoid(F.linear(y, P[pre + "fc.2.weight"]))
    return x * y.view(b, c, 1, 1)


# --------------------------------------------------------------------------------------
# row a6: SpectralTransform.forward  (layers/ffc/spectral_transform.py:77-110)
# --------------------------------------------------------------------------------------
def spectral_transform(x: Tensor, P: Params, pre: str, stride: int, upsample: bool,
                       training: bool, sn: bool = False) -> Tensor:
    if stride == 2 and upsample:                                       # :44-45, :79
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    elif stride == 2:                                                  # :46-47
        x = F.avg_pool2d(x, kernel_size=2, stride=2)
    x = se_layer(x, P, pre + "se_block.")                             # :87
    w1 = _weight(P, pre + "conv1.", training, sn)
    x = _kinked(batch_norm(F.conv2d(x, w1), P, pre + "bn1.", training), 0.0)  # :89
    f = fourier_unit(x, P, pre + "fu.", training)                     # :91
    w2 = _weight(P, pre + "conv2.", training, sn)
    return F.conv2d(x + f, w2)                                         # :108 (LFU term commented out)


# --------------------------------------------------------------------------------------
# rows a7 / a8 / a10 / a11: FFC.forward, FFCTranspose.forward and the SN twins
# (layers/ffc/ffc.py:84-99, layers/ffc/ffc_transpose.py:91-110, layers/snffc/*.py)
# --------------------------------------------------------------------------------------
def _local_conv(x: Branch, P: Params, pre: str, cfg: FFCConfig, training: bool) -> Branch:
    """One of convl2l / convl2g / convg2l; nn.Identity when a side has no channels."""
    has = (pre + "weight") in P or (pre + "weight_orig") in P
    if not has:
        return x                                                       # nn.Identity (ffc.py:46-47)
    bias = P.get(pre + "bias")
    if cfg.upsampling:                                                 # ffc_transpose.py:84-86
        w = _weight(P, pre, training, cfg.spectral_norm, dim=1)
        return F.conv_transpose2d(x, w, bias, cfg.stride, cfg.padding, cfg.out_padding)
    w = _weight(P, pre, training, cfg.spectral_norm, dim=0)
    return F.conv2d(x, w, bias, cfg.stride, cfg.padding)


def ffc(x: Union[Tensor, Tuple[Branch, Branch]], P: Params, pre: str, cfg: FFCConfig,
        training: bool) -> Tuple[Branch, Branch]:
    x_l, x_g = x if type(x) is tuple else (x, 0)
    out_l: Branch = 0
    out_g: Branch = 0
    if cfg.ratio_gout != 1:
        out_l = _local_conv(x_l, P, pre + "convl2l.", cfg, training) \
            + _local_conv(x_g, P, pre + "convg2l.", cfg, training)
    if cfg.ratio_gout != 0:
        out_g = _local_conv(x_l, P, pre + "convl2g.", cfg, training)
        if cfg.in_cg != 0 and cfg.out_cg != 0:                         # convg2g is a SpectralTransform
            out_g = out_g + spectral_transform(
                x_g, P, pre + "convg2g.", cfg.stride, cfg.upsampling, training, cfg.spectral_norm)
    return out_l, out_g


# --------------------------------------------------------------------------------------
# row a9: FFC_BN_ACT.forward  (layers/ffc/ffc_bn_act.py:70-83)
# --------------------------------------------------------------------------------------
def ffc_bn_act(x, P: Params, pre: str, cfg: FFCConfig, training: bool) -> Tuple[Branch, Branch]:
    x_l, x_g = ffc(x, P, pre + "ffc.", cfg, training)
    # norm/act are Identity on a side whose ratio makes it empty (:52-53, :63-64)
    if cfg.ratio_gout != 1:
        if cfg.norm == "bn":
            x_l = batch_norm(x_l, P, pre + "bn_l.", training)
        x_l = activation(x_l, cfg.act) if not isinstance(x_l, int) else x_l
    if cfg.ratio_gout != 0:
        if cfg.norm == "bn":
            x_g = batch_norm(x_g, P, pre + "bn_g.", training)
        x_g = activation(x_g, cfg.act) if not isinstance(x_g, int) else x_g
    return x_l, x_g


def resizer(x) -> Tensor:
    """layers/resizer.py:15-24: tuple -> tensor (concat on channels; drop an int-0 global)."""
    if type(x) is tuple:
        return x[0] if isinstance(x[1], int) else torch.cat(list(x), dim=1)
    return x


def noise_injection(x: Tensor, P: Params, pre: str, noise: Optional[Tensor] = None) -> Tensor:
    """layers/noise_injection.py:20-32: x + weight * N(0,1) with one noise plane per image."""
    if noise is None:
        b, _, h, w = x.shape
        noise = x.new_empty(b, 1, h, w).normal_()
    return x + P[pre + "weight"] * noise


# --------------------------------------------------------------------------------------
# callers of the hot path used by BASELINE.json's configs
# --------------------------------------------------------------------------------------
def ffc_generator_cfgs(nz: int = 100, nc: int = 1, ngf: int = 32, g: float = 0.5):
    """models/ffc_generator.py:24-28 (config 1)."""
    A = dict(act="leaky_relu", upsampling=True)
    return [
        FFCConfig(nz, ngf * 8, 4, 0, g, 1, 0, **A),
        FFCConfig(ngf * 8, ngf * 4, 4, g, g, 2, 1, **A),
        FFCConfig(ngf * 4, ngf * 2, 4, g, g, 2, 1, **A),
        FFCConfig(ngf * 2, ngf * 1, 4, g, g, 2, 1, **A),
        FFCConfig(ngf * 1, nc, 4, g, 0, 2, 1, act="tanh", upsampling=True),
    ]


def ffc_generator(z: Tensor, P: Params, training: bool, cfgs=None) -> Tensor:
    """models/ffc_generator.py:30-45."""
    cfgs = cfgs or ffc_generator_cfgs()
    x = z
    for i, cfg in enumerate(cfgs):
        x = ffc_bn_act(x, P, f"ffc{i}.", cfg, training)
    return resizer(x)


def fgenerator_cfgs(variant: str = "fgan32"):
    """FGenerator of fgan_complete.py:97-113 / fgan64_complete.py:101-122 /
    fgan128_complete.py:458-485 (sngan_complete.py:37-53 equals fgan32)."""
    if variant in ("fgan32", "sngan32"):
        ngf, r, n_up = 64, 0.25, 3
    elif variant == "fgan64":
        ngf, r, n_up = 64, 0.25, 4
    elif variant == "fgan128":
        ngf, r, n_up = 128, 0.5, 5
    else:
        raise ValueError(variant)
    up = dict(stride=2, padding=1, norm="bn", act="gelu", upsampling=True)
    chans = [ngf * 8, ngf * 4, ngf * 2, ngf] + [ngf] * (n_up - 3)
    cfgs = []
    for i in range(n_up):
        cfgs.append(FFCConfig(chans[i], chans[i + 1], 4, 0.0 if i == 0 else r, r, **up))
    cfgs.append(FFCConfig(ngf, 3, 3, r, 0.0, stride=1, padding=1, norm="identity", act="tanh"))
    return ngf, cfgs


def fgenerator(z: Tensor, P: Params, training: bool, variant: str = "fgan32", mg: int = 4,
               noises=None) -> Tensor:
    """FGenerator.forward, fgan_complete.py:115-140 (training: NoiseInjection after each
    upsampling stage; eval: uint8 conversion :136-138 is left to the caller)."""
    ngf, cfgs = fgenerator_cfgs(variant)
    x = F.linear(z, P["noise_to_feature.0.weight"], P["noise_to_feature.0.bias"])
    x = x.reshape(x.size(0), -1, mg, mg)
    for i, cfg in enumerate(cfgs):
        n = i + 2
        x = ffc_bn_act(x, P, f"conv{n}.", cfg, training)
        if training and i < len(cfgs) - 1:
            nl = None if noises is None else noises[i][0]
            ng = None if noises is None else noises[i][1]
            x = (noise_injection(x[0], P, f"lcl_noise{n}.", nl),
                 noise_injection(x[1], P, f"glb_noise{n}.", ng))
    return resizer(x)


def sn_discriminator(x: Tensor, P: Params, training: bool, n_convs: int = 7, mg: int = 4) -> Tensor:
    """Plain spectral-norm conv Discriminator, fgan_complete.py:142-171 (7 convs),
    fgan64_complete.py:159-191 (8), fgan128_complete.py:525-562 (9)."""
    m = x
    for i in range(1, n_convs + 1):
        pre = f"conv{i}."
        w = spectral_norm_weight(P, pre, training)
        k = w.shape[-1]
        m = _kinked(F.conv2d(m, w, P[pre + "bias"], 1 if k == 3 else 2, 1), 0.1)
    w = spectral_norm_weight(P, "fc.", training)
    return F.linear(m.reshape(-1, mg * mg * 512), w, P["fc.bias"])


def sngan_fdiscriminator_cfgs():
    """sngan_complete.py:125-136 (the only FFC discriminator that is trained)."""
    A = dict(bias=True, act="leaky_relu")
    return [
        FFCConfig(3, 64, 3, 0.0, 0.25, 1, 1, norm="identity", **A),
        FFCConfig(64, 128, 4, 0.25, 0.25, 2, 1, norm="bn", **A),
        FFCConfig(128, 256, 4, 0.25, 0.25, 2, 1, norm="bn", **A),
        FFCConfig(256, 512, 4, 0.25, 0.0, 2, 1, norm="bn", **A),
    ]


def sngan_fdiscriminator(x: Tensor, P: Params, training: bool, mg: int = 4) -> Tensor:
    """sngan_complete.py:148-157."""
    m = x
    for i, cfg in enumerate(sngan_fdiscriminator_cfgs()):
        m = ffc_bn_act(m, P, f"main.{i}.", cfg, training)
    m = resizer(m)
    w = spectral_norm_weight(P, "fc.", training)
    return F.linear(m.reshape(-1, mg * mg * 512), w, P["fc.bias"])


def fdiscriminator_sn64(x: Tensor, P: Params, training: bool, mg: int = 4) -> Tensor:
    """The 64x64 SNFFC discriminator of SURVEY.md 8(d) config 3 (harness FDiscriminatorSN64): sngan's FDiscriminator
    (sngan_complete.py:125-157) with one more stride-2 stage and SNFFC (layers/snffc/snffc.py:12-33) in place of FFC."""
    A = dict(bias=True, act="leaky_relu", spectral_norm=True)
    cfgs = [FFCConfig(3, 64, 3, 0.0, 0.25, 1, 1, norm="identity", **A), FFCConfig(64, 128, 4, 0.25, 0.25, 2, 1, norm="bn", **A),
            FFCConfig(128, 256, 4, 0.25, 0.25, 2, 1, norm="bn", **A), FFCConfig(256, 512, 4, 0.25, 0.25, 2, 1, norm="bn", **A),
            FFCConfig(512, 512, 4, 0.25, 0.0, 2, 1, norm="bn", **A)]
    m = x
    for i, cfg in enumerate(cfgs):
        m = ffc_bn_act(m, P, f"main.{i}.", cfg, training)
    m = resizer(m)
    w = spectral_norm_weight(P, "fc.", training)
    return F.linear(m.reshape(-1, mg * mg * 512), w, P["fc.bias"])


# hinge losses, fgan_complete.py:216-235
def hinge_loss_dis(fake: Tensor, real: Tensor) -> Tensor:
    return F.relu(1.0 - real).mean() + F.relu(1.0 + fake).mean()


def hinge_loss_gen(fake: Tensor) -> Tensor:
    return -fake.mean()


def deterministic_fill(P: Params, seed: int = 0) -> None:
    """Fill every tensor of a state-dict-shaped mapping from a closed-form, key-dependent
    sequence so the SAME weights can be regenerated anywhere (GPU box included) without
    shipping them: used for whole-model golden fixtures."""
    import zlib
    for k in sorted(P.keys()):
        t = P[k]
        if not torch.is_floating_point(t):
            t.zero_()
            continue
        h = zlib.crc32(k.encode()) % 9973
        n = t.numel()
        idx = torch.arange(n, dtype=torch.float64)
        base = torch.sin(idx * 0.7310585786 + 0.1 * h + 0.013 * seed) \
            + 0.5 * torch.cos(idx * 0.2689414214 * 1.7 + 0.37 * h)
        if k.endswith("running_var"):
            val = 1.0 + 0.25 * base.abs()
        elif k.endswith("running_mean"):
            val = 0.05 * base
        elif k.endswith(("weight_u", "weight_v")):
            val = base / base.norm()
        elif t.dim() <= 1 and k.endswith("weight") and "_noise" not in k:   # BN gamma
            val = 1.0 + 0.1 * base
        elif t.dim() <= 1 or k.endswith("bias"):
            val = 0.05 * base
        elif "_noise" in k:                                            # NoiseInjection.weight
            val = 0.05 * base
        else:
            fan_in = max(1, n // t.shape[0])
            val = base * (1.0 / math.sqrt(fan_in))
        with torch.no_grad():
            t.copy_(val.reshape(t.shape).to(t.dtype))
