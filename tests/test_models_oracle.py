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
    f = os.path.join(golden_dir, "pipeline_ref_B1_64x64_s0.npz")  # 64x64, B=1 (a few seconds on CPU)
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


def test_phasenet256_oracle_matches_reference_golden_and_fp64_budget(golden_dir):
    """configs[0] fixture: the oracle restatement (fp32) reproduces the reference modules' outputs, and its fp64 mode reproduces
    the stored arbiter (so the budgets the GPU tests use are what the fixture says)."""
    z = np.load(os.path.join(golden_dir, "phasenet_ref_256x256_s5.npz"))
    B, H, W, seed = [int(v) for v in z["meta"]]
    state = fp.seeded_state(seed)
    rgb1, rgb2 = fp.seeded_frames(B, H, W, seed)
    st32, st64 = {}, {}
    fp.interp_phasenet(fp.oracle_backend(state, hw=(H, W), threads=4), rgb1, rgb2, st32)
    fp.interp_phasenet(fp.oracle_backend(state, hw=(H, W), threads=4, precision="fp64"), rgb1, rgb2, st64)
    for k in ("lab_pred", "phase_pred", "low_level", "amp3", "amp9"):
        sl = int(z[k + "__stride"]) if k + "__stride" in z.files else 1
        a32 = st32[k].numpy()[..., ::sl, ::sl]
        a64 = st64[k].numpy()[..., ::sl, ::sl]
        assert np.abs(a32 - z[k]).max() <= 2e-6, k
        f64 = z[k].astype(np.float64) - z[k + "__d64"].astype(np.float64) / 1e4
        assert np.abs(a64 - f64).max() <= 2e-3 * float(z[k + "__budget"]) + 1e-7, k      # float16 storage of the difference
        assert float(np.abs(st32[k].numpy() - st64[k].numpy()).max()) == pytest.approx(float(z[k + "__budget"]), rel=1e-3, abs=1e-9)


def test_synth_generators_equal_oracle_generators():
    """fvfi.synth (what bench.py / tools use for seeded weights and frames) == the oracle's generators (what the fixtures use)."""
    from fvfi import synth
    a, b = fp.seeded_state(3), synth.seeded_state(3)
    for k in a:
        assert list(a[k].keys()) == list(b[k].keys())
        assert all(torch.equal(a[k][n], b[k][n]) for n in a[k])
    assert all(torch.equal(p, q) for p, q in zip(fp.seeded_frames(2, 64, 96, 1), synth.seeded_frames(2, 64, 96, 1)))


def test_lab_round_trip_and_known_values():
    from oracle import lab
    rgb = np.random.default_rng(0).random((5, 7, 3))
    assert np.abs(lab.lab2rgb(lab.rgb2lab(rgb)) - rgb).max() < 1e-6
    # white, black, mid-grey (CIE L* of sRGB 0.5 is 53.389)
    assert np.allclose(lab.rgb2lab(np.ones((1, 1, 3)))[0, 0], [100, 0, 0], atol=2e-2)
    assert np.allclose(lab.rgb2lab(np.zeros((1, 1, 3)))[0, 0], [0, 0, 0], atol=1e-9)
    assert abs(lab.rgb2lab(np.full((1, 1, 3), 0.5))[0, 0, 0] - 53.389) < 1e-2
