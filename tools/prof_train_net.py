"""FusionNet forward + backward alone (8 crops of 256x256, the trained part of configs[4]): ms per iteration and the kernel table.
    python tools/prof_train_net.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi.fusion_net import FusionNet
from fvfi import synth as fp

torch.manual_seed(0)
net = FusionNet().cuda()
net.load_state_dict(fp.seeded_state(0)["fusion_net"])
net.train()
for n, p in net.named_parameters():
    p.requires_grad_(not n.startswith("net."))
B, H, W = 8, 256, 256
ins = [torch.rand((B, c, H, W), device="cuda") for c in (3, 3, 3, 6, 3)]
target = torch.rand((B, 3, H, W), device="cuda")


def it():
    for p in net.parameters():
        p.grad = None
    loss = torch.nn.functional.l1_loss(target, torch.clip(net(*ins), 0, 1))
    loss.backward()
    return loss


for _ in range(3):
    it()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    it()
e1.record()
torch.cuda.synchronize()
print("FusionNet fwd+bwd, %d crops of %dx%d: %.3f ms / iteration" % (B, H, W, e0.elapsed_time(e1) / 10))
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    it()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
