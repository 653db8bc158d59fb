"""CPU restatement of the reference's GAN training step -- TEST / BASELINE INFRASTRUCTURE ONLY.

Follows fgan_complete.py:357-393 (generator update, discriminator update, AdamW lr 2e-4 betas
(0.5, 0.999), LambdaLR linear decay, hinge losses :216-235) on top of the functional oracle
(oracle/ffc_ref.py), i.e. on the same PyTorch CPU library kernels the reference itself executes.
Used by bench.py for the ``cpu_baseline`` leg and the ``--impl reference`` arm, and by tests.
"""
from __future__ import annotations

import torch

from . import ffc_ref as R

_NO_GRAD_SUFFIX = ("running_mean", "running_var", "num_batches_tracked", "weight_u", "weight_v")


def to_params(state_dict, dtype=torch.float32):
    P = {}
    for k, v in state_dict.items():
        v = v.detach().cpu()
        if v.is_floating_point():
            v = v.to(dtype).clone()
            if not k.endswith(_NO_GRAD_SUFFIX):
                v.requires_grad_(True)
        else:
            v = v.clone()
        P[k] = v
    return P


class RefTrainer:
    def __init__(self, sd_G, sd_D, variant="fgan32", n_convs=7, lr=2e-4, betas=(0.5, 0.999), num_total_steps=100000,
                 mg=4):
        self.PG, self.PD = to_params(sd_G), to_params(sd_D)
        self.variant, self.n_convs, self.mg = variant, n_convs, mg
        self.leaf_G = [v for v in self.PG.values() if v.requires_grad]
        self.leaf_D = [v for v in self.PD.values() if v.requires_grad]
        self.optim_G = torch.optim.AdamW(self.leaf_G, lr=lr, betas=betas)
        self.optim_D = torch.optim.AdamW(self.leaf_D, lr=lr, betas=betas)
        decay = lambda step: 1.0 - step / num_total_steps
        self.sched_G = torch.optim.lr_scheduler.LambdaLR(self.optim_G, decay)
        self.sched_D = torch.optim.lr_scheduler.LambdaLR(self.optim_D, decay)

    def G(self, z):
        return R.fgenerator(z, self.PG, True, self.variant, self.mg)

    def D(self, x):
        if self.n_convs == "fd64":                       # the 64x64 SNFFC discriminator (harness.FDiscriminatorSN64)
            return R.fdiscriminator_sn64(x, self.PD, True, self.mg)
        return R.sn_discriminator(x, self.PD, True, self.n_convs, self.mg)

    @staticmethod
    def _req(leaves, flag):
        for v in leaves:
            v.requires_grad_(flag)

    def step(self, z_g, z_d, real):
        self._req(self.leaf_G, True); self._req(self.leaf_D, False)          # fgan_complete.py:368-369
        self.optim_D.zero_grad(); self.optim_G.zero_grad()
        loss_G = R.hinge_loss_gen(self.D(self.G(z_g)))
        loss_G.backward()
        self.optim_G.step()
        self._req(self.leaf_G, False); self._req(self.leaf_D, True)          # :380-381
        self.optim_D.zero_grad(); self.optim_G.zero_grad()
        fake = self.G(z_d)
        loss_D = R.hinge_loss_dis(self.D(fake), self.D(real))
        loss_D.backward()
        self.optim_D.step()
        self.sched_G.step(); self.sched_D.step()
        return loss_G.detach(), loss_D.detach()
