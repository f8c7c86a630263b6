"""oracle/build_ref_kernels.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Recipe that compiles the REFERENCE's own four AdaCoF CUDA kernels for sm_100a.

The kernels are CUDA C strings inside /root/reference/src/adacof/cupy_module/adacof.py
(:6-258); tensor sizes and strides are baked in per shape by the reference's own
pure-Python expander cupy_kernel() (:261-299).  This script imports that module
where it lies, expands each kernel for the shapes the tests and bench use, and
compiles the expanded text with `nvcc -cubin` -- outputs go ONLY to oracle/_ref/
(git-ignored binaries that travel to the GPU box).  No reference source is
written into the repo: the expanded .cu text lives in a temp dir and is deleted.

oracle/ref_kernels.py loads the cubins with cuda-python and launches them with
the reference's launch geometry (grid ceil(n/512), block 512; adacof.py:349-351).
"""
import json
import os
import subprocess
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import  # noqa: E402

OUT = os.path.join(HERE, "_ref")

# (B, C, H, W, F, dilation): parity-test shapes + BASELINE.json configs[1] (bench)
SHAPES = [
    (2, 3, 40, 56, 5, 1),
    (1, 3, 33, 47, 5, 2),
    (1, 3, 24, 40, 3, 1),
    (2, 3, 96, 160, 5, 1),
    (1, 3, 256, 256, 5, 1),
    (8, 3, 1088, 1920, 5, 1),
]

KERNELS = {
    "kernel_AdaCoF_updateOutput": ["input", "weight", "offset_i", "offset_j", "output"],
    "kernel_AdaCoF_updateGradWeight": ["gradLoss", "input", "offset_i", "offset_j", "gradWeight"],
    "kernel_AdaCoF_updateGradAlpha": ["gradLoss", "input", "weight", "offset_i", "offset_j", "gradOffset_i"],
    "kernel_AdaCoF_updateGradBeta": ["gradLoss", "input", "weight", "offset_i", "offset_j", "gradOffset_j"],
}


def tag(shape):
    return "B%d_C%d_H%d_W%d_F%d_D%d" % shape


def main():
    if not ref_import.available():
        print("reference not present; keeping prebuilt oracle/_ref")
        return 0
    ref = ref_import.ref_adacof_module()
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    with tempfile.TemporaryDirectory() as tmp:
        for shape in SHAPES:
            B, C, H, W, F, d = shape
            pad = (F - 1) * d
            meta = lambda *s: torch.empty(*s, device="meta")
            tensors = {
                "input": meta(B, C, H + pad, W + pad), "weight": meta(B, F * F, H, W),
                "offset_i": meta(B, F * F, H, W), "offset_j": meta(B, F * F, H, W),
                "output": meta(B, C, H, W), "gradLoss": meta(B, C, H, W),
                "gradWeight": meta(B, F * F, H, W), "gradOffset_i": meta(B, F * F, H, W),
                "gradOffset_j": meta(B, F * F, H, W),
            }
            for name, names in KERNELS.items():
                src = ref.cupy_kernel(name, F, d, {k: tensors[k] for k in names})
                cu = os.path.join(tmp, "k.cu")
                with open(cu, "w") as f:
                    f.write(src)
                out = os.path.join(OUT, "%s_%s.cubin" % (name, tag(shape)))
                if not os.path.exists(out):
                    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-cubin",
                                           "-o", out, cu])
                manifest["%s/%s" % (name, tag(shape))] = os.path.basename(out)
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump({"shapes": [list(s) for s in SHAPES], "files": manifest}, f, indent=1)
    print("built %d reference cubins into %s" % (len(manifest), OUT))
    return 0


if __name__ == "__main__":
    sys.exit(main())
