"""Generates tests/golden/*.npz by running the REFERENCE modules (imported from /root/reference,
CPU, float32) on seeded inputs.  Run once in the build container:

    python tests/golden/make_golden.py

The fixtures pin the oracle (oracle/ffc_ref.py, oracle/fu_dft.py) and the CUDA path to the
reference's own results; /root/reference is not available on the GPU box, these files are.
Layout of a fixture: in<i>, out<i>, din<i> (gradient of sum(out * cot<i>) w.r.t. in<i>), cot<i>,
sd/<key> (state_dict before the forward), grad/<key>, post/<key> (buffers after the forward).
"""
import contextlib
import io
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
for name in ("matplotlib", "matplotlib.pyplot"):           # util/data_loader.py:7 imports it; absent here
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, "/root/reference")
with contextlib.redirect_stdout(io.StringIO()):
    import layers as ref_layers                             # noqa: E402
    import models as ref_models                             # noqa: E402
import models.ffcmodel as _fm                               # noqa: E402

_orig_init = _fm.FFCModel.__init__
_fm.FFCModel.__init__ = lambda self, debug=False, **kw: _orig_init(self, debug=debug)   # models/ffc_generator.py:22


def load_script_classes(path):
    src = open(path).read()
    src = src[: src.index("\ndef train(args)")]
    ns = {"__name__": "ref_script", "math": math}
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(src, path, "exec"), ns)
    return ns


def quiet(fn):
    with contextlib.redirect_stdout(io.StringIO()):          # FFC.__init__ prints (ffc.py:38-39)
        return fn()


def randomize(mod, seed):
    """Non-trivial parameters AND buffers (running stats away from 0/1)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, p in mod.named_parameters():
            if p.numel() == 0:
                continue
            if p.dim() <= 1:
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g) if k.endswith("weight") and "noise" not in k
                        else 0.1 * torch.randn(p.shape, generator=g))
            else:
                fan_in = p[0].numel() if p.dim() > 1 else 1
                p.copy_(torch.randn(p.shape, generator=g) / math.sqrt(max(fan_in, 1)))
        for k, b in mod.named_buffers():
            if k.endswith("running_mean"):
                b.copy_(0.1 * torch.randn(b.shape, generator=g))
            elif k.endswith("running_var"):
                b.copy_(0.5 + torch.rand(b.shape, generator=g))


def record(name, mod, inputs, train=True, seed=0, extra=None):
    torch.manual_seed(seed)
    randomize(mod, seed + 1)
    mod.train(train)
    fx = {}
    for k, v in mod.state_dict().items():
        fx["sd/" + k] = v.detach().clone().numpy()
    xs = [x.clone().requires_grad_(True) for x in inputs]
    arg = xs[0] if len(xs) == 1 else tuple(xs)
    out = mod(arg)
    outs = [o for o in (out if isinstance(out, tuple) else (out,)) if torch.is_tensor(o)]
    g = torch.Generator().manual_seed(seed + 2)
    cots = [torch.randn(o.shape, generator=g) for o in outs]
    sum((o * c).sum() for o, c in zip(outs, cots)).backward()
    for i, x in enumerate(xs):
        fx[f"in{i}"] = x.detach().numpy()
        fx[f"din{i}"] = x.grad.numpy()
    for i, (o, c) in enumerate(zip(outs, cots)):
        fx[f"out{i}"] = o.detach().numpy()
        fx[f"cot{i}"] = c.numpy()
    for k, p in mod.named_parameters():
        if p.grad is not None:
            fx["grad/" + k] = p.grad.numpy()
    for k, b in mod.named_buffers():
        fx["post/" + k] = b.detach().clone().numpy()
    fx["train"] = np.array(int(train))
    if extra:
        fx.update(extra)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **fx)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, outputs {[tuple(o.shape) for o in outs]}")


def main():
    from cases import CASES
    t = torch.randn
    torch.manual_seed(1234)
    only = set(sys.argv[sys.argv.index("--only") + 1].split(",")) if "--only" in sys.argv else None
    for name, (ctor, shapes, train, _oracle) in CASES.items():
        inputs = [torch.randn(*s) for s in shapes]          # drawn for every case, so that --only leaves the others' inputs as they are
        if only is None or name in only:
            record(name, quiet(lambda: ctor(ref_layers)), inputs, train=train)
    if only is not None:
        return

    # whole models with closed-form weights (oracle.ffc_ref.deterministic_fill), so no weights are stored
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle.ffc_ref import deterministic_fill

    def record_model(name, mod, z, seed, keep_grads):
        sd = mod.state_dict()
        deterministic_fill(sd, seed)
        for k in sd:
            if "_noise" in k:
                sd[k].zero_()          # NoiseInjection draws fresh noise; zero weights keep training mode deterministic
        mod.train(True)
        zz = z.clone().requires_grad_(True)
        out = mod(zz)
        g = torch.Generator().manual_seed(seed + 2)
        cot = torch.randn(out.shape, generator=g)
        (out * cot).sum().backward()
        fx = {"in0": z.numpy(), "out0": out.detach().numpy(), "cot0": cot.numpy(), "din0": zz.grad.numpy(),
              "seed": np.array(seed)}
        params = dict(mod.named_parameters())
        for k in keep_grads:
            fx["grad/" + k] = params[k].grad.numpy()
        for k, b in mod.named_buffers():
            if (k.endswith("running_mean") and b.numel() <= 64) or (k.endswith("weight_u") and b.numel() <= 512):
                fx["post/" + k] = b.detach().clone().numpy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **fx)
        print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, output {tuple(out.shape)}")

    ns = load_script_classes("/root/reference/fgan_complete.py")
    record_model("model_fgan32_G", quiet(lambda: ns["FGenerator"](z_size=128, mg=4)), t(2, 128), 3,
                 ["conv3.ffc.convg2g.fu.conv_layer.weight", "conv4.ffc.convg2g.conv1.weight", "conv4.bn_g.weight",
                  "conv3.ffc.convg2g.se_block.fc.0.weight", "conv5.ffc.convg2l.weight"])
    # the plain SN conv discriminator the FFC generators are trained against
    record_model("model_fgan32_D", quiet(lambda: ns["Discriminator"](sn=True, mg=4)), t(4, 3, 32, 32) * 0.5, 9,
                 ["conv1.weight_orig", "conv1.bias", "conv2.weight_orig", "conv7.bias", "fc.weight_orig"])
    # fgan64: one more upsampling stage and a 64x64 Fourier unit (general form: rfft2 | mix | BN statistics | BN+ReLU->irfft2).
    # (uses_sn=True is stored and ignored by the reference's FFC_BN_ACT, ffc_bn_act.py:39: the generator has plain weights)
    ns64 = load_script_classes("/root/reference/fgan64_complete.py")
    record_model("model_fgan64_G", quiet(lambda: ns64["FGenerator"](z_size=128, mg=4)), t(2, 128), 11,
                 ["conv3.ffc.convg2g.fu.conv_layer.weight", "conv5.ffc.convg2g.fu.conv_layer.weight", "conv5.ffc.convl2g.weight",
                  "conv5.ffc.convg2g.conv1.weight", "conv6.ffc.convg2l.weight", "conv5.bn_g.weight"])
    # fgan128: ngf 128, ratio 0.5, five upsampling stages, Fourier units up to 32 channels @ 128x128 (the largest spectrum)
    ns128 = load_script_classes("/root/reference/fgan128_complete.py")
    record_model("model_fgan128_G", quiet(lambda: ns128["FGenerator"](z_size=128, mg=4)), t(1, 128), 21,
                 ["conv6.ffc.convg2g.fu.conv_layer.weight", "conv6.ffc.convg2g.conv1.weight", "conv6.bn_g.weight",
                  "conv7.ffc.convg2l.weight", "conv4.ffc.convg2g.fu.bn.weight"])
    # BASELINE configs[2] / SURVEY.md 8(d): the 64x64 SNFFC discriminator, assembled from the REFERENCE's FFC_BN_ACT and
    # SNFFC classes by the same builder the harness uses on its own layers
    from fastfourierconvolution_b200.harness.models import build_fd_sn64

    class RefFDiscriminatorSN64(nn.Module):
        def __init__(self):
            super().__init__()
            self.mg = 4
            self.resizer = ref_layers.Resizer()
            self.main, self.fc = build_fd_sn64(ref_layers, torch.nn.utils.spectral_norm)

        def forward(self, x):
            return self.fc(self.resizer(self.main(x)).view(-1, 16 * 512))

    # batch 4, and a seed for which neither FP32 evaluation order puts an activation on the other side of its kink: with
    # BatchNorm over 32-64 values per channel in the last stages about every third (seed, batch) choice does, and then the
    # reference's own FP32 gradients sit 1e-3..1e-2 away from its float64 ones (SURVEY.md 8(c) caveat 1)
    record_model("model_fgan64_FD", quiet(RefFDiscriminatorSN64), t(4, 3, 64, 64) * 0.5, 15,
                 ["main.0.ffc.convl2l.weight_orig", "main.1.ffc.convg2g.conv1.weight_orig", "main.1.ffc.convg2g.fu.conv_layer.weight",
                  "main.1.ffc.convl2g.weight_orig", "main.0.ffc.convl2l.bias", "main.0.ffc.convl2g.weight_orig"])   # (biases in front of a BatchNorm have gradient 0)
    ns = load_script_classes("/root/reference/sngan_complete.py")
    record_model("model_sngan_FD", quiet(lambda: ns["FDiscriminator"](sn=True, mg=4)), t(2, 3, 32, 32), 5,
                 ["main.1.ffc.convg2g.fu.conv_layer.weight", "main.2.ffc.convg2g.conv2.weight", "main.0.ffc.convl2g.weight"])
    record_model("model_ffcgen_cfg1", quiet(lambda: ref_models.FFCGenerator(100, 1, 32)), t(2, 100, 1, 1), 7,
                 ["ffc1.ffc.convg2g.fu.conv_layer.weight", "ffc3.ffc.convg2g.conv2.weight", "ffc4.ffc.convg2l.weight"])


if __name__ == "__main__":
    main()
