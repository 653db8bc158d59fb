"""The reference's OWN model classes on the product layers ("models/*.py and the *_complete.py scripts run
unchanged"): tests/ref_classes_driver.py executes FGenerator / Discriminator / FDiscriminator of the four training scripts
and models.FFCGenerator / models.FFCDiscriminator first on the reference's ``layers`` package and then, through the shadow
``layers`` package (fastfourierconvolution_b200/dropin), on the sm_100a implementations, loads the reference state_dict with
strict=True and compares outputs, input gradients and every parameter gradient (two FP32 evaluations: 2e-4, max norm).

Needs the reference checkout (``/root/reference`` or ``$FFC_REFERENCE_ROOT``); skipped where it is absent (the GPU box).  Without
a GPU the kernels run in their host emulation build (index arithmetic + wiring); with one, on cuda:0."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FFC_REFERENCE_ROOT", "/root/reference")


def _shadow_exports():
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "fastfourierconvolution_b200", "dropin"))
    try:
        import importlib
        return importlib.import_module("layers")
    finally:
        sys.path.pop(0)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "layers")), reason="reference checkout not present")
def test_reference_model_classes_run_on_the_product_layers():
    r = subprocess.run([sys.executable, os.path.join(HERE, "ref_classes_driver.py")], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    res = json.loads(line[7:])
    assert set(res) == {"fgan32_G", "fgan32_D", "fgan64_G", "fgan64_D", "fgan128_G", "fgan128_D", "sngan_G", "sngan_FD", "sngan_mg6_G", "sngan_mg6_FD",
                        "cond_block", "cfg1_G", "cfg1_D"}
    for name, e in res.items():
        assert max(e["out"], e["din"], e["grad"]) < 2e-4, (name, e)
        if name.endswith(("_G", "_FD")) or name == "cfg1_D":
            assert e["product_ffc_modules"] >= 4, (name, e)        # the FFC stack really is the product's


def test_shadow_layers_package_exports_the_reference_names():
    """layers/__init__.py:2-22 of the reference: every name its models and scripts pick up through ``from layers import *``."""
    sub = subprocess.run([sys.executable, "-c",
                          "import sys; sys.path[:0]=[%r, %r]; import layers, json; print(json.dumps(sorted(layers.__all__)))"
                          % (os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "fastfourierconvolution_b200", "dropin"))],
                         capture_output=True, text=True, timeout=300)
    assert sub.returncode == 0, sub.stderr[-2000:]
    names = set(json.loads(sub.stdout.strip().splitlines()[-1]))
    assert names >= {"FFC", "FFCTranspose", "FFC_BN_ACT", "SpectralTransform", "FourierUnitSN", "SELayer", "SNFFC", "SNFFCTranspose",
                     "Resizer", "Print", "debug_print", "NoiseInjection", "GaussianNoise", "ConditionalBatchNorm2d", "aw_method"}
