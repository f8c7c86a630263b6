"""CPU tests: the oracle restatements of the networks / fusion recipe (oracle/nets.py,
oracle/fusion_pipeline.py) against the committed golden fixtures produced by the REAL reference
modules (tests/golden/make_golden_models.py), and -- when /root/reference is present -- against
those modules directly."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import fusion_pipeline as fp


def _fixtures(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "pipeline_ref_*.npz")))


def test_pipeline_oracle_matches_reference_golden(golden_dir):
    files = _fixtures(golden_dir)
    assert files
    f = files[0]  # 64x64, B=1 (a few seconds on CPU)
    z = np.load(f)
    B, H, W, seed = [int(v) for v in z["meta"]]
    state = fp.seeded_state(seed)
    chk = [float(sum(v.double().sum() for v in state[n].values())) for n in ("phase_net", "fusion_net", "adacof")]
    assert np.allclose(chk, z["checksum"], rtol=1e-9), "seeded init differs from the fixture's"
    rgb1, rgb2 = fp.seeded_frames(B, H, W, seed)
    st = {}
    fp.interp(fp.oracle_backend(state, hw=(H, W), threads=4), rgb1, rgb2, st)
    for k in ("lab1", "ada_pred", "flow_var_map", "lab_pred", "phase_pred", "phase_uncertainty", "ada_uncertainty",
              "base", "final"):
        assert np.abs(st[k].numpy() - z[k]).max() <= 2e-6, k


@pytest.mark.needs_reference
def test_oracle_nets_equal_reference_modules():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_models import reference_backend
    B, H, W, seed = 1, 64, 64, 3
    state = fp.seeded_state(seed)
    rgb1, rgb2 = fp.seeded_frames(B, H, W, seed)
    a, b = {}, {}
    fp.interp(reference_backend(state, H, W), rgb1, rgb2, a)
    fp.interp(fp.oracle_backend(state, hw=(H, W), threads=4), rgb1, rgb2, b)
    for k in a:
        assert float((a[k] - b[k]).abs().max()) <= 1e-6, k


def test_lab_round_trip_and_known_values():
    from oracle import lab
    rgb = np.random.default_rng(0).random((5, 7, 3))
    assert np.abs(lab.lab2rgb(lab.rgb2lab(rgb)) - rgb).max() < 1e-6
    # white, black, mid-grey (CIE L* of sRGB 0.5 is 53.389)
    assert np.allclose(lab.rgb2lab(np.ones((1, 1, 3)))[0, 0], [100, 0, 0], atol=2e-2)
    assert np.allclose(lab.rgb2lab(np.zeros((1, 1, 3)))[0, 0], [0, 0, 0], atol=1e-9)
    assert abs(lab.rgb2lab(np.full((1, 1, 3), 0.5))[0, 0, 0] - 53.389) < 1e-2
