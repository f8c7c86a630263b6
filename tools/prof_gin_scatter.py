"""True-gradInput scatter for ncu / timing: python tools/prof_gin_scatter.py [B]   (FVFI_GIN_SCATTER=warp selects the warp-aggregated kernel)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi import adacof
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
H, W = 1088, 1920
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda *s: torch.rand(s, device="cuda", generator=g)
inp = mk(B, 3, H + 4, W + 4)
w = torch.softmax(torch.randn((B, 25, H, W), device="cuda", generator=g), 1)
gout = torch.randn((B, 3, H, W), device="cuda", generator=g)
for name, (oi, oj) in (("smooth", (mk(B, 25, H, W) - 0.5, mk(B, 25, H, W) - 0.5)),
                       ("iid", ((3 * torch.randn((B, 25, H, W), device="cuda", generator=g)).clamp_(-16, 16),
                                (3 * torch.randn((B, 25, H, W), device="cuda", generator=g)).clamp_(-16, 16)))):
    adacof.adacof_backward(gout, inp, w, oi, oj, 1, "true")
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); adacof.adacof_backward(gout, inp, w, oi, oj, 1, "none"); e1.record()
    adacof.adacof_backward(gout, inp, w, oi, oj, 1, "true"); e2.record(); torch.cuda.synchronize()
    print("%s offsets, B = %d: gradient kernel %.3f ms, with true gradInput %.3f ms" % (name, B, e0.elapsed_time(e1), e1.elapsed_time(e2)))
