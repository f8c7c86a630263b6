"""Pyramid reconstruct only (no high residual) for ncu: python tools/prof_pyramid_rec.py [N planes]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi.pyramid import Pyramid
N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
pyr = Pyramid(17, 4, np.sqrt(2), torch.device("cuda"))
x = torch.rand((N, 1080, 1920), device="cuda")
v = pyr.filter(x, want_high=False)
torch.cuda.synchronize()
for _ in range(2):
    y = pyr.inv_filter_sparse(v, use_high=False)
torch.cuda.synchronize()
print("ok", float((y - x).abs().max()))
