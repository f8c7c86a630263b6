"""Generates tests/golden/adacof_ref_*.npz ON A B200 by running the REFERENCE's own AdaCoF
CUDA kernels (cubins in oracle/_ref, built by oracle/build_ref_kernels.py from
/root/reference/src/adacof/cupy_module/adacof.py:6-258) on seeded synthetic operands.

    gpurun -- python tests/golden/make_adacof_golden.py      # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/

Only the seed/shape and the OUTPUTS are stored (inputs are regenerated from the seed by
oracle.adacof.synth), float16-free, compressed.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import adacof as oa, ref_kernels  # noqa: E402

CASES = [(2, 3, 40, 56, 5, 1, 0), (1, 3, 33, 47, 5, 2, 1), (1, 3, 24, 40, 3, 1, 2)]


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for (B, C, H, W, F, d, seed) in CASES:
        inp, w, oi, oj, g = oa.synth(B, C, H, W, F, d, seed)
        dev = [torch.from_numpy(x).cuda() for x in (inp, w, oi, oj, g)]
        out = ref_kernels.forward(*dev[:4], d)
        gin, gw, gi, gj = ref_kernels.backward(dev[4], *dev[:4], d)
        torch.cuda.synchronize()
        assert float(gin.abs().max()) == 0.0
        name = "adacof_ref_B%d_C%d_H%d_W%d_F%d_D%d_s%d.npz" % (B, C, H, W, F, d, seed)
        np.savez_compressed(os.path.join(out_dir, name), shape=np.array([B, C, H, W, F, d, seed]),
                            out=out.cpu().numpy(), gw=gw.cpu().numpy(), gi=gi.cpu().numpy(),
                            gj=gj.cpu().numpy())
        print("wrote", name)


if __name__ == "__main__":
    main()
