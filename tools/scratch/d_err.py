import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tests/golden')
import torch
from test_layers_emu import _sn_discriminator_pair
for seed in (0, 1):
    for cfg in ((9, 128, 2), (9, 128, 8)):
        e = _sn_discriminator_pair("cuda:0", *cfg, seed=seed)
        print(cfg, seed, {k: f"{v:.1e}" for k, v in e.items() if not k.startswith("buf/")})
