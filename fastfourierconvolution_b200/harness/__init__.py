"""Callers of the hot path used for parity tests and benchmarks: the reference's generator /
discriminator classes restated on top of fastfourierconvolution_b200.layers (same attribute names
and state_dict keys), and the GAN training step of the *_complete.py scripts."""
from .models import (FGenerator, SNDiscriminator, FDiscriminator, FDiscriminatorSN64, FFCGenerator, FFCDiscriminator,
                     weights_init, hinge_loss_dis, hinge_loss_gen)
from .train import GanTrainer, FlatGradAllReduce
