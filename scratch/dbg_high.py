import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi.pyramid import Pyramid
from fvfi import _lib
from fvfi.pyr_plan import ptr_array
N, H, W = 12, 1080, 1920
pyr = Pyramid(17, 4, np.sqrt(2), torch.device("cuda"))
x = torch.rand((N, H, W), device="cuda")
def wall(fn, name, reps=4):
    for i in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print("%-30s host %.2f ms  total %.2f ms" % (name, (t1 - t0) * 1e3, (t2 - t0) * 1e3))
wall(lambda: pyr.filter(x, want_high=False), "filter no high")
wall(lambda: pyr.filter(x, want_high=True), "filter high")
wall(lambda: torch.empty((N, 1, H, W), device="cuda"), "empty")
print(torch.cuda.memory_summary(abbreviated=True)[:1500])
