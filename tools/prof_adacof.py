"""AdaCoF kernels for ncu: python tools/prof_adacof.py [B]  (forward random offsets, backward, fused two-frame synthesis smooth offsets)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi import adacof
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H, W, F = 1088, 1920, 5
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda *s: torch.rand(s, device="cuda", generator=g)
inp = mk(B, 3, H + 4, W + 4)
w = torch.softmax(torch.randn((B, 25, H, W), device="cuda", generator=g), 1)
oi = (3 * torch.randn((B, 25, H, W), device="cuda", generator=g)).clamp_(-16, 16)
oj = (3 * torch.randn((B, 25, H, W), device="cuda", generator=g)).clamp_(-16, 16)
gout = torch.randn((B, 3, H, W), device="cuda", generator=g)
out = torch.empty((B, 3, H, W), device="cuda")
a1, b1 = mk(B, 25, H, W) - 0.5, mk(B, 25, H, W) - 0.5
occ = mk(B, 1, H, W)
w2 = torch.softmax(torch.randn((B, 25, H, W), device="cuda", generator=g), 1)     # frame 2: its own coefficient tensors (no L2 re-use)
a2, b2 = mk(B, 25, H, W) - 0.5, mk(B, 25, H, W) - 0.5
def t(fn, reps=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
px = B * H * W
bf = 4 * (75 * px + 3 * B * (H + 4) * (W + 4) + 3 * px)
bb = 4 * (3 * px + 3 * B * (H + 4) * (W + 4) + 150 * px)
bs = 4 * (150 * px + 6 * B * (H + 4) * (W + 4) + 5 * px)
tf = t(lambda: adacof.adacof_forward(inp, w, oi, oj, 1, out=out))
tfs = t(lambda: adacof.adacof_forward(inp, w, a1, b1, 1, out=out))
tb = t(lambda: adacof.adacof_backward(gout, inp, w, oi, oj, 1, "none"))
ts = t(lambda: adacof.adacofnet_warp_blend(inp, inp, w, a1, b1, w2, a2, b2, occ, 1, want_t=False))
print("fwd random  %.3f ms %.0f GB/s | fwd smooth %.3f ms %.0f GB/s | bwd random %.3f ms %.0f GB/s | fused smooth %.3f ms %.0f GB/s" %
      (tf, bf / tf / 1e6, tfs, bf / tfs / 1e6, tb, bb / tb / 1e6, ts, bs / ts / 1e6))
# true gradInput (extension): CTA-aggregated scatter (FVFI_GIN_SCATTER=warp: the warp-aggregated kernel), on top of the gradient kernel
for name, (x, y) in (("random", (oi, oj)), ("smooth", (a1, b1)), ("zero", (torch.zeros_like(a1), torch.zeros_like(a1)))):
    t_none = t(lambda: adacof.adacof_backward(gout, inp, w, x, y, 1, "none"), 3)
    t_true = t(lambda: adacof.adacof_backward(gout, inp, w, x, y, 1, "true"), 3)
    print("gradInput scatter, %s offsets: backward %.3f ms -> with true gradInput %.3f ms (scatter + memset %.3f ms for %d M taps)"
          % (name, t_none, t_true, t_true - t_none, 25 * px // 1000000))
