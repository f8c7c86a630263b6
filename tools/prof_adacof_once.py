"""One launch of each AdaCoF kernel variant for `ncu --set full`: python tools/prof_adacof_once.py [B]
order: forward i.i.d. offsets, forward smooth, backward i.i.d., backward smooth, fused synthesis smooth, fused synthesis i.i.d."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi import adacof
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H, W = 1088, 1920
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda *s: torch.rand(s, device="cuda", generator=g)
inp = mk(B, 3, H + 4, W + 4)
w = torch.softmax(torch.randn((B, 25, H, W), device="cuda", generator=g), 1)
oi = (3 * torch.randn((B, 25, H, W), device="cuda", generator=g)).clamp_(-16, 16)
oj = (3 * torch.randn((B, 25, H, W), device="cuda", generator=g)).clamp_(-16, 16)
a1, b1 = mk(B, 25, H, W) - 0.5, mk(B, 25, H, W) - 0.5
gout = torch.randn((B, 3, H, W), device="cuda", generator=g)
out = torch.empty((B, 3, H, W), device="cuda")
occ = mk(B, 1, H, W)
w2 = torch.softmax(torch.randn((B, 25, H, W), device="cuda", generator=g), 1)     # frame 2: its own coefficient tensors (no L2 re-use)
a2, b2 = mk(B, 25, H, W) - 0.5, mk(B, 25, H, W) - 0.5
adacof.adacof_forward(inp, w, oi, oj, 1, out=out)
adacof.adacof_forward(inp, w, a1, b1, 1, out=out)
adacof.adacof_backward(gout, inp, w, oi, oj, 1, "none")
adacof.adacof_backward(gout, inp, w, a1, b1, 1, "none")
adacof.adacofnet_warp_blend(inp, inp, w, a1, b1, w2, a2, b2, occ, 1, want_t=False)
oi2 = (3 * torch.randn((B, 25, H, W), device="cuda", generator=g)).clamp_(-16, 16)
oj2 = (3 * torch.randn((B, 25, H, W), device="cuda", generator=g)).clamp_(-16, 16)
adacof.adacofnet_warp_blend(inp, inp, w, oi, oj, w2, oi2, oj2, occ, 1, want_t=False)
torch.cuda.synchronize()
print("ok")
