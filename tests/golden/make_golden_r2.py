"""Round-2 golden fixtures from the REAL reference modules (imported from /root/reference, build container only):

    python tests/golden/make_golden_r2.py [phasenet256] [pipelines] [ckpt]

Every fixture holds, per stage ``k`` of the recipe,
  ``k``        the output of the reference's own modules on CPU in fp32 (the parity target), and
  ``k__d64``   (k - the output of the oracle restatement run in fp64, oracle_backend(precision="fp64")) * 1e4 as float16 -- the
               ARBITER: how far the reference's own fp32 run is from the exact result of the same recipe.  The GPU tests assert,
               per stage,  |GPU - k| <= 1e-4  OR  |GPU - fp64| <= 2 max|k - fp64|  (no further from the truth than the reference).
The oracle restatement (fp32) is asserted equal to the reference modules on every stage before anything is written.

* phasenet_ref_256x256_s*.npz        BASELINE.json configs[0]: Pyramid(12, 4, sqrt 2) -> PhaseNet -> reconstruct on one 256x256 pair
                                     (height 12 => layers[7] serves levels 6..9, src/phase_net/phase_net.py:148).
* pipeline_ref_B1_256x256_s2.npz     full recipe at the training crop size (configs[4]).
* pipeline_ref_B1_184x328_s3.npz     full recipe at a size whose pyramid levels need Bluestein/Rader FFT lengths (23, 29, 41, 46,
                                     58, 82, 116, 164, 232) and whose AdaCoFNet input is reflect-padded to /32 in both axes.
* pipeline_ckpt_B1_256x256_s4.npz    full recipe with the SHIPPED checkpoints src/phase_net/phase_net.pt and
                                     src/fusion_net/fusion_net.pt (AdaCoF: seeded random init -- its checkpoint is a missing LFS blob).
``wrap_<call>_<level>_idx/_val``: the reference's phase at every coefficient within a window of +-pi, per decomposition call of the
recipe -- the branch the parity runs align the GPU's wrapped phases to (oracle/wrap_align.py explains why).
Stages above 30k elements are stored subsampled in the two image axes (stride in ``<k>__stride``); ``<k>__budget`` is
max|k - fp64| over ALL elements.  Inputs are regenerated from the seed.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import fusion_pipeline as fp, ref_import  # noqa: E402
from oracle.wrap_align import wrap_lists  # noqa: E402
from make_golden_models import reference_backend  # noqa: E402

torch.set_grad_enabled(False)
REF = ref_import.REF


def ckpt_state(seed):
    st = fp.seeded_state(seed)
    st["phase_net"] = torch.load(os.path.join(REF, "src/phase_net/phase_net.pt"), map_location="cpu")
    st["fusion_net"] = torch.load(os.path.join(REF, "src/fusion_net/fusion_net.pt"), map_location="cpu")
    return st


SKIP = ()


def run_case(recipe, state, B, H, W, seed, name, stride_over=30000):
    rgb1, rgb2 = fp.seeded_frames(B, H, W, seed)
    t0 = time.time()
    ref, o32, o64, dec = {}, {}, {}, {}
    recipe(reference_backend(state, H, W), rgb1, rgb2, ref, dec)
    recipe(fp.oracle_backend(state, hw=(H, W), threads=8), rgb1, rgb2, o32)
    err = {k: float((ref[k] - o32[k]).abs().max()) for k in ref}
    print(name, "oracle(fp32) vs reference modules:", {k: "%.1e" % v for k, v in err.items() if v > 0})
    assert max(err.values()) < 5e-6, err
    recipe(fp.oracle_backend(state, hw=(H, W), threads=8, precision="fp64"), rgb1, rgb2, o64)
    keep = {"meta": np.array([B, H, W, seed]),
            "checksum": np.array([float(sum(v.double().sum() for v in state[n].values())) for n in
                                  ("phase_net", "fusion_net", "adacof")])}
    budget = {}
    for k in ref:
        if k in SKIP:
            continue
        a, b = ref[k].numpy().astype(np.float32), o64[k].numpy()
        budget[k] = float(np.abs(a - b).max())
        keep[k + "__budget"] = np.array(budget[k])
        if a.size > stride_over and a.ndim >= 2:
            st = int(np.ceil(np.sqrt(a.size / float(stride_over))))
            a, b = a[..., ::st, ::st], b[..., ::st, ::st]
            keep[k + "__stride"] = np.array(st)
        keep[k] = np.ascontiguousarray(a)
        keep[k + "__d64"] = np.clip((a.astype(np.float64) - b) * 1e4, -6e4, 6e4).astype(np.float16)
    # the reference's branch at the coefficients within rounding of the negative real axis (oracle/wrap_align.py)
    wl = wrap_lists(dec)
    keep.update(wl)
    print(name, "coefficients stored for the branch alignment: %d" % sum(v.size for k, v in wl.items() if k.endswith("_idx")))
    print(name, "reference fp32 vs fp64 arbiter:", {k: "%.1e" % v for k, v in budget.items()})
    np.savez_compressed(os.path.join(HERE, name), **keep)
    print("wrote", name, "%.1f MB, %.0f s" % (os.path.getsize(os.path.join(HERE, name)) / 1e6, time.time() - t0))


if __name__ == "__main__":
    assert ref_import.available(), "needs /root/reference"
    which = sys.argv[1:] or ["phasenet256", "pipelines", "ckpt"]
    if "phasenet256" in which:
        run_case(fp.interp_phasenet, fp.seeded_state(5), 1, 256, 256, 5, "phasenet_ref_256x256_s5.npz")
    if "pipelines" in which:
        run_case(fp.interp, fp.seeded_state(0), 1, 64, 64, 0, "pipeline_ref_B1_64x64_s0.npz")       # replaces the round-1 fixtures
        run_case(fp.interp, fp.seeded_state(1), 2, 64, 96, 1, "pipeline_ref_B2_64x96_s1.npz")       # (same cases, new format)
        run_case(fp.interp, fp.seeded_state(2), 1, 256, 256, 2, "pipeline_ref_B1_256x256_s2.npz")
        run_case(fp.interp, fp.seeded_state(3), 1, 184, 328, 3, "pipeline_ref_B1_184x328_s3.npz")
    if "ckpt" in which:
        run_case(fp.interp, ckpt_state(4), 1, 256, 256, 4, "pipeline_ckpt_B1_256x256_s4.npz")
