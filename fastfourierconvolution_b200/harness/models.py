"""Model classes that call the hot path (SURVEY.md section 8(b) "who calls it").

The reference defines these inline in its training scripts, which run ``main()`` and download
datasets at import time, and its ``models/*`` classes do not construct (FFCModel.__init__ rejects
``inplanes=``); neither can be imported on the GPU box.  They are therefore restated here,
table-driven, on top of ``fastfourierconvolution_b200.layers`` with the reference's attribute names so
that a reference ``state_dict`` loads with ``strict=True`` (tests/test_layers_emu.py and tests/test_gpu_parity.py check the
key lists and the outputs against fixtures generated from the reference classes themselves).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..layers import FFC_BN_ACT, NoiseInjection, Resizer, Print


def weights_init(m):
    """fgan_complete.py:22-31: N(0, 0.02) on *Conv* weights, N(1, 0.02) / 0 on *BatchNorm*."""
    name = m.__class__.__name__
    if name.find("Conv") != -1:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif name.find("BatchNorm") != -1:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)


def _sn_linear(fc, x):
    """The spectral-norm Linear head of the discriminators (fgan_complete.py:157, 170): the hook's power iteration and
    W / sigma run in the library's three spectral-norm kernels (the same ones the SN convolutions use) instead of the stock
    hook's ~14 PyTorch launches, the product and its gradients in the library's FP32 matrix kernel (ops.linear).  CPU
    tensors (the host oracle side of a comparison) go through the module's own forward."""
    if not x.is_cuda:
        return fc(x)
    from .. import ops
    from ..layers import _util
    return ops.linear(x, _util.effective_weight(fc), fc.bias)


def hinge_loss_dis(fake, real):
    """fgan_complete.py:216-222."""
    return F.relu(1.0 - real).mean() + F.relu(1.0 + fake).mean()


def hinge_loss_gen(fake):
    """fgan_complete.py:231-235."""
    return -fake.mean()


# (ngf, ratio_g, number of upsampling stages, eval-mode clamp) per script
_FGEN = {
    "fgan32": (64, 0.25, 3, "unit"),     # fgan_complete.py:81-140, sngan_complete.py:23-80
    "fgan64": (64, 0.25, 4, "minmax"),   # fgan64_complete.py:85-156
    "fgan128": (128, 0.5, 5, "minmax"),  # fgan128_complete.py:442-522
}


class FGenerator(nn.Module):
    """FFC generator of the *_complete.py scripts: Linear stem -> n x FFC_BN_ACT(upsampling, BN, GELU)
    with NoiseInjection on both branches in training mode -> FFC_BN_ACT(k3, Tanh) -> Resizer."""

    def __init__(self, z_size=128, mg: int = 4, variant: str = "fgan32"):
        super().__init__()
        ngf, r, n_up, self._clamp = _FGEN[variant]
        self.z_size, self.ngf, self.mg, self.variant = z_size, ngf, mg, variant
        self.print_size = Print(False)
        self.resizer = Resizer()
        self.noise_to_feature = nn.Sequential(nn.Linear(z_size, mg * mg * ngf * 8))
        chans = [ngf * 8, ngf * 4, ngf * 2, ngf] + [ngf] * (n_up - 3)
        self._stages = []
        for i in range(n_up):
            n = i + 2
            cout = chans[i + 1]
            setattr(self, f"conv{n}", FFC_BN_ACT(chans[i], cout, 4, 0.0 if i == 0 else r, r, stride=2, padding=1,
                                                 activation_layer=nn.GELU, norm_layer=nn.BatchNorm2d, upsampling=True,
                                                 uses_noise=True, uses_sn=(variant != "fgan32")))
            setattr(self, f"lcl_noise{n}", NoiseInjection(int(cout * (1 - r))))
            setattr(self, f"glb_noise{n}", NoiseInjection(int(cout * r)))
            self._stages.append(n)
        self._last = n_up + 2
        setattr(self, f"conv{self._last}", FFC_BN_ACT(ngf, 3, 3, r, 0.0, stride=1, padding=1, activation_layer=nn.Tanh,
                                                      norm_layer=nn.Identity, upsampling=False, uses_noise=True,
                                                      uses_sn=(variant != "fgan32")))

    def forward(self, z):
        from ..layers import _util
        with _util.batched_bn_counters(self):            # the ten num_batches_tracked += 1 of a forward as one launch
            return self._forward(z)

    def _forward(self, z):
        from .. import ops
        lin = self.noise_to_feature[0]
        if z.is_cuda:    # the Linear stem (fgan_complete.py:92-95, 117-119) on the library's FP32 matrix kernel
            fake = ops.linear(z.reshape(z.size(0), -1), lin.weight, lin.bias)
        else:
            fake = self.noise_to_feature(z)
        fake = fake.reshape(fake.size(0), -1, self.mg, self.mg)
        for n in self._stages:
            fake = getattr(self, f"conv{n}")(fake)
            if self.training:
                fake = getattr(self, f"lcl_noise{n}")(fake[0]), getattr(self, f"glb_noise{n}")(fake[1])
        fake = self.resizer(getattr(self, f"conv{self._last}")(fake))
        if not self.training:                    # uint8 images for the metric code (fgan_complete.py:136-138)
            # fgan64_complete.py:150-153 clamps to the tensor's own min / max, which is the identity
            fake = ops.to_uint8(fake, -1.0, 1.0) if self._clamp == "unit" else ops.to_uint8(fake, 1.0, -1.0)
        return fake


class SNDiscriminator(nn.Module):
    """The plain spectral-norm conv discriminator the fgan scripts train against
    (fgan_complete.py:142-171 n_convs=7, fgan64_complete.py:159-191 n_convs=8,
    fgan128_complete.py:525-562 n_convs=9): SN Conv2d + LeakyReLU(0.1) stack, SN Linear head.

    SURVEY.md section 8(f) rank 1 -- it is more than half of a training step.  ``backend="ffc_b200"`` (default) runs
    every convolution, its data / weight / bias gradients and the LeakyReLU on the FP32-accurate sm_100a kernels of the
    local branches (same ``ops.conv2d`` as FFC.convl2l); the ``nn.Conv2d`` objects stay as parameter holders, so the
    state_dict (``convN.weight_orig / weight_u / weight_v / bias``) is the reference's.  ``backend="torch"`` is the
    reference's own arithmetic (nn.Conv2d.forward) and is what the parity tests compare against."""

    def __init__(self, sn=True, mg: int = 4, n_convs: int = 7, backend: str = "ffc_b200"):
        super().__init__()
        assert backend in ("ffc_b200", "torch")
        self.mg, self.n_convs, self.backend = mg, n_convs, backend
        sn_fn = torch.nn.utils.spectral_norm if sn else (lambda m: m)
        spec = [(3, 64, 3, 1), (64, 64, 4, 2), (64, 128, 3, 1), (128, 128, 4, 2), (128, 256, 3, 1),
                (256, 256, 4, 2), (256, 512, 3, 1), (512, 512, 4, 2), (512, 512, 4, 2)][:n_convs]
        for i, (ci, co, k, s) in enumerate(spec, 1):
            setattr(self, f"conv{i}", sn_fn(nn.Conv2d(ci, co, k, stride=s, padding=(1, 1))))
        self.fc = sn_fn(nn.Linear(mg * mg * 512, 1))
        self.act = nn.LeakyReLU(0.1)
        self.channels_last = False      # backend="torch" only: run cuDNN in NHWC (same math, no layout round trips)
        self.concurrent_forwards_ok = backend == "ffc_b200"      # GanTrainer._two_forwards: D(fake) and D(real) on two streams

    def forward(self, x):
        if self.backend == "ffc_b200":
            from .. import ops
            from ..layers import _util
            m = x
            convs = [getattr(self, f"conv{i}") for i in range(1, self.n_convs + 1)]
            # spectral_norm pre-forward hooks (W / sigma, one power iteration each) depend on the weights only: they run (4 small
            # launches per layer) on side streams, each joined right before its own convolution -- one serial chain joined
            # before the second convolution cost ~120 us of stall per forward
            # (the first layer's too: two forwards queued on different streams, GanTrainer._two_forwards, then still run the
            # iterations of a layer in program order, on that layer's side stream)
            mods = convs + ([self.fc] if x.is_cuda else [])
            ws = ops.fork_map(x.device, [(lambda c=c: _util.effective_weight(c)) for c in mods])
            for i, conv in enumerate(convs):
                ws[i][1]()
                m = ops.conv2d_act(m, ws[i][0], conv.bias, conv.stride[0], conv.padding[0], ops.ACT_LEAKY, self.act.negative_slope)
            w_fc = None
            if x.is_cuda:
                ws[-1][1]()
                w_fc = ws[-1][0]
            m = m.reshape(-1, self.mg * self.mg * 512)
            return ops.linear(m, w_fc, self.fc.bias) if w_fc is not None else self.fc(m)
        m = x.contiguous(memory_format=torch.channels_last) if self.channels_last else x
        for i in range(1, self.n_convs + 1):
            m = self.act(getattr(self, f"conv{i}")(m))
        return self.fc(m.reshape(-1, self.mg * self.mg * 512))


class FDiscriminator(nn.Module):
    """sngan_complete.py:116-157: the only FFC discriminator the reference trains."""

    def __init__(self, sn=True, mg: int = 4):
        super().__init__()
        self.mg = mg
        sn_fn = torch.nn.utils.spectral_norm if sn else (lambda m: m)
        common = dict(bias=True, uses_noise=False, uses_sn=True, activation_layer=nn.LeakyReLU)
        self.print_size = Print(False)
        self.resizer = Resizer()
        self.main = nn.Sequential(
            FFC_BN_ACT(3, 64, 3, 0.0, 0.25, stride=1, padding=1, norm_layer=nn.Identity, **common),
            FFC_BN_ACT(64, 128, 4, 0.25, 0.25, stride=2, padding=1, norm_layer=nn.BatchNorm2d, **common),
            FFC_BN_ACT(128, 256, 4, 0.25, 0.25, stride=2, padding=1, norm_layer=nn.BatchNorm2d, **common),
            FFC_BN_ACT(256, 512, 4, 0.25, 0.0, stride=2, padding=1, norm_layer=nn.BatchNorm2d, **common),
        )
        self.fc = sn_fn(nn.Linear(mg * mg * 512, 1))

    def forward(self, x):
        m = self.resizer(self.main(x))
        return _sn_linear(self.fc, m.view(-1, self.mg * self.mg * 512))


# stages of the SNFFC discriminator below: (in, out, kernel, ratio_gin, ratio_gout, stride, padding, norm)
FD_SN64_STAGES = [(3, 64, 3, 0.0, 0.25, 1, 1, nn.Identity), (64, 128, 4, 0.25, 0.25, 2, 1, nn.BatchNorm2d),
                  (128, 256, 4, 0.25, 0.25, 2, 1, nn.BatchNorm2d), (256, 512, 4, 0.25, 0.25, 2, 1, nn.BatchNorm2d),
                  (512, 512, 4, 0.25, 0.0, 2, 1, nn.BatchNorm2d)]


def build_fd_sn64(layers, sn_fn, mg=4):
    """(main, fc) of the 64x64 SNFFC discriminator from a ``layers`` package -- this one or the reference's own (the
    golden fixture tests/golden/model_fgan64_FD.npz is built from the reference's classes by this same function)."""
    stages = []
    for ci, co, k, rgi, rgo, s, p, norm in FD_SN64_STAGES:
        blk = layers.FFC_BN_ACT(ci, co, k, rgi, rgo, stride=s, padding=p, bias=True, uses_noise=False, uses_sn=True,
                                activation_layer=nn.LeakyReLU, norm_layer=norm)
        blk.ffc = layers.SNFFC(ci, co, k, rgi, rgo, s, p, 1, 1, True)       # FFC -> its spectral-norm twin (layers/snffc/snffc.py:12-33)
        stages.append(blk)
    return nn.Sequential(*stages), sn_fn(nn.Linear(mg * mg * 512, 1))


class FDiscriminatorSN64(nn.Module):
    """BASELINE configs[2] names a "spectral-norm snffc discriminator" for the 64x64 workload; the reference ships none
    (its only FFC discriminator, sngan_complete.py:116-157, takes 32x32 images and leaves ``uses_sn`` unused).  This is
    SURVEY.md 8(d)'s construction: sngan's FDiscriminator extended by one stride-2 stage, with every FFC replaced by
    the reference's SNFFC, so that ``layers/snffc`` runs at the CelebA shape and an oracle built from the reference's
    own classes exists."""

    def __init__(self, sn=True, mg: int = 4):
        super().__init__()
        from .. import layers
        self.mg = mg
        self.resizer = Resizer()
        self.main, self.fc = build_fd_sn64(layers, torch.nn.utils.spectral_norm if sn else (lambda m: m), mg)

    def forward(self, x):
        from ..layers import _util
        with _util.prefetch_spectral_norm(self, x.device):       # 17 power iterations, off the critical path of the convolutions
            m = self.resizer(self.main(x))
            return _sn_linear(self.fc, m.view(-1, self.mg * self.mg * 512))


class FFCGenerator(nn.Module):
    """models/ffc_generator.py:14-45 (config 1: nz=100, nc=1, ngf=32, g_factor=0.5)."""

    def __init__(self, nz: int, nc: int, ngf: int, g_factor: float = 0.5, debug: bool = False):
        super().__init__()
        self.print_size = Print(False)
        self.resizer = Resizer()
        a = dict(activation_layer=nn.LeakyReLU, upsampling=True)
        self.ffc0 = FFC_BN_ACT(nz, ngf * 8, 4, 0, g_factor, 1, 0, **a)
        self.ffc1 = FFC_BN_ACT(ngf * 8, ngf * 4, 4, g_factor, g_factor, 2, 1, **a)
        self.ffc2 = FFC_BN_ACT(ngf * 4, ngf * 2, 4, g_factor, g_factor, 2, 1, **a)
        self.ffc3 = FFC_BN_ACT(ngf * 2, ngf * 1, 4, g_factor, g_factor, 2, 1, **a)
        self.ffc4 = FFC_BN_ACT(ngf * 1, nc, 4, g_factor, 0, 2, 1, norm_layer=nn.Identity,
                               activation_layer=nn.Tanh, upsampling=True)

    def forward(self, x):
        for i in range(5):
            x = getattr(self, f"ffc{i}")(x)
        return self.resizer(x)


class FFCDiscriminator(nn.Module):
    """models/ffc_discriminator.py:14-60 (needs 64x64 inputs)."""

    def __init__(self, nc: int, ndf: int, debug: bool = False):
        super().__init__()
        self.print_size = Print(False)
        self.resizer = Resizer()
        a = dict(activation_layer=nn.LeakyReLU)
        self.ffc0 = FFC_BN_ACT(nc, ndf * 2, 4, 0, 0.5, 2, 1, **a)
        self.ffc1 = FFC_BN_ACT(ndf * 2, ndf * 4, 4, 0.5, 0.5, 2, 1, **a)
        self.ffc2 = FFC_BN_ACT(ndf * 4, ndf * 8, 4, 0.5, 0.5, 2, 1, **a)
        self.ffc3 = FFC_BN_ACT(ndf * 8, ndf * 16, 4, 0.5, 0.5, 2, 1, **a)
        self.ffc4 = FFC_BN_ACT(ndf * 16, 1, 4, 0.5, 0, 1, 0, norm_layer=nn.Identity, activation_layer=nn.Sigmoid)

    def forward(self, x):
        for i in range(5):
            x = getattr(self, f"ffc{i}")(x)
        return self.resizer(x)
