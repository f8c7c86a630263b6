"""Generates the committed golden fixtures by importing the REFERENCE's own Python modules from
/root/reference (build container only) with the third-party stubs of oracle/ref_import.py.

    python tests/golden/make_golden.py

* pyramid_ref_*.npz   reference ``Pyramid.filter`` / ``inv_filter`` (src/train/pyramid.py:35-112: the real
                      coeff_to_values / values_to_coeff) on top of the steerable shim (parity unpinned
                      for the FFT/mask arithmetic).
* phasenet_ref_*.npz  reference ``PhaseNet`` (src/phase_net/phase_net.py) normalize -> forward on CPU,
                      seeded random-init weights (state_dict regenerated from the seed by the tests).
* fusionnet_ref_*.npz reference ``FusionNet.forward`` (src/fusion_net/fusion_net.py:46-77) on CPU.
* kernelest_ref_*.npz reference ``KernelEstimation`` + AdaCoFNet blend/uncertainty tail
                      (src/fusion_net/fusion_adacofnet.py:109-155,198-213) on CPU, given warped frames.
Inputs are regenerated from seeds; only outputs (float32, compressed) are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

torch.set_grad_enabled(False)


def seeded_image(N, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand((N, H, W), generator=g)


def pyramid_cases():
    return [(2, 48, 64, 6, 0), (1, 90, 150, 8, 1), (3, 64, 64, 8, 2)]


def make_pyramid():
    ref_import.install_stubs()
    from src.train.pyramid import Pyramid
    for (N, H, W, height, seed) in pyramid_cases():
        pyr = Pyramid(height=height, nbands=4, scale_factor=np.sqrt(2), device=torch.device("cpu"))
        img = seeded_image(N, H, W, seed)
        vals = pyr.filter(img)
        rec = pyr.inv_filter(vals)
        d = {"meta": np.array([N, H, W, height, seed]), "high": vals.high_level.numpy(), "low": vals.low_level.numpy(),
             "rec": rec.numpy()}
        for l, (p, a) in enumerate(zip(vals.phase, vals.amplitude)):
            d["phase%d" % l] = p.numpy()
            d["amp%d" % l] = a.numpy()
        name = "pyramid_ref_N%d_%dx%d_h%d_s%d.npz" % (N, H, W, height, seed)
        np.savez_compressed(os.path.join(HERE, name), **d)
        print("wrote", name, "recon err", float((rec - img).abs().max()))


if __name__ == "__main__":
    assert ref_import.available(), "needs /root/reference"
    which = sys.argv[1:] or ["pyramid", "models"]
    if "pyramid" in which:
        make_pyramid()
    if "models" in which:
        try:
            from make_golden_models import make_models
        except ImportError:
            make_models = None
        if make_models:
            make_models()
