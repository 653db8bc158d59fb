"""Debug aid: L2-staged Fourier unit forward + backward on the GPU against the float64 oracle-free reference (PyTorch ops in
float64 on the GPU), tensor-core path vs the plain FP32 path (ffc_debug_fu3_simt_mix).  usage: python tools/debug_fu3_bwd.py B C N"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import fastfourierconvolution_b200 as ffc
from fastfourierconvolution_b200 import _C

B, C, N = (int(a) for a in sys.argv[1:4])
dev = "cuda:0"


def ref(x, w, gamma, beta, training, rm, rv):
    b, c, h, wd = x.shape
    s = torch.fft.rfftn(x, dim=(-2, -1), norm="ortho")
    s = torch.stack((s.real, s.imag), dim=2).reshape(b, 2 * c, h, s.shape[-1])
    y = F.conv2d(s, w)
    y = F.relu(F.batch_norm(y, rm.clone(), rv.clone(), gamma, beta, training, 0.1, 1e-5))
    y = y.reshape(b, -1, 2, h, y.shape[-1])
    return torch.fft.irfftn(torch.complex(y[:, :, 0].contiguous(), y[:, :, 1].contiguous()), s=(h, wd), dim=(-2, -1), norm="ortho")


for training in (True, False):
    for simt in (0, 1):
        torch.manual_seed(0)
        m = ffc.FourierUnitSN(C, C).to(dev).train(training)
        m.fused = "staged"
        with torch.no_grad():
            m.bn.weight.uniform_(0.5, 1.5); m.bn.bias.normal_(0, 0.2); m.bn.running_mean.normal_(0, 0.1); m.bn.running_var.uniform_(0.5, 1.5)
        x = torch.randn(B, C, N, N, device=dev)
        cot = torch.randn(B, C, N, N, device=dev)
        P = [t.detach().double().requires_grad_(True) for t in (x, m.conv_layer.weight, m.bn.weight, m.bn.bias)]
        r = ref(P[0], P[1], P[2], P[3], training, m.bn.running_mean.double(), m.bn.running_var.double())
        (r * cot.double()).sum().backward()
        _C.lib().ffc_debug_fu3_simt_mix(simt)
        xo = x.clone().requires_grad_(True)
        out = m(xo)
        (out * cot).sum().backward()
        _C.lib().ffc_debug_fu3_simt_mix(0)
        rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max()).item()
        print(f"train={training} simt={simt}: out {rel(out.detach(), r.detach()):.2e} dx {rel(xo.grad, P[0].grad):.2e} dw {rel(m.conv_layer.weight.grad, P[1].grad):.2e} "
              f"dgamma {rel(m.bn.weight.grad, P[2].grad):.2e} dbeta {rel(m.bn.bias.grad, P[3].grad):.2e}")
