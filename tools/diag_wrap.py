"""Developer tool (GPU box): after aligning the +-pi branch, where does the predicted pyramid still differ from the oracle?"""
import os, sys, math
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from oracle import fusion_pipeline as fp
from oracle.wrap_align import WrapAligner
from fvfi.pipeline import FusionPipeline
torch.backends.cudnn.allow_tf32 = False
H = W = 256
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 5
state = fp.seeded_state(seed)
rgb1, rgb2 = fp.seeded_frames(1, H, W, seed)
ost, odec = {}, {}
be = fp.oracle_backend(state, hw=(H, W), threads=8)
fp.interp_phasenet(be, rgb1, rgb2, ost, odec)
# oracle input decomposition
lab1, lab2 = fp.rgb2lab_planes(rgb1), fp.rgb2lab_planes(rgb2)
img = torch.cat((lab1.reshape(-1, H, W), lab2.reshape(-1, H, W)), 0)
ov = be.pyr.filter(img)
pipe = FusionPipeline(H, W, "cuda")
pipe.load_state(state)
captured = {}
al = WrapAligner.from_decomps(odec)
def hook(tag, planes, vals):
    captured["raw"] = vals
    v = al(tag, planes, vals)
    captured["aligned"] = v
    return v
pipe.filter_hook = hook
pipe.stages = {}
with torch.no_grad():
    pipe.phase_interp(rgb1.cuda(), rgb2.cuda())
print("flips", al.flips)
L = pipe.pyr.height - 2
gv = captured["aligned"]
for l in range(L):
    gp, ga = pipe.stages["phase%d" % l].cpu(), pipe.stages["amp%d" % l].cpu()
    op, oa = ost["phase%d" % l], ost["amp%d" % l]
    e = (ga.double() * torch.exp(1j * gp.double()) - oa.double() * torch.exp(1j * op.double())).abs()
    idx = np.unravel_index(int(e.argmax()), e.shape)
    # input differences at this level
    ip, ia = gv.phase[l].cpu(), gv.amplitude[l].cpu()
    dphi = (ip - ov.phase[l])
    dphi = torch.atan2(torch.sin(dphi), torch.cos(dphi)).abs()
    big = dphi > 0.05
    print("level %d: max pred coeff err %.2e at %s (pred amp there %.2e, level amp max %.2e); input: max wrapped |dphase| %.2e, #(>0.05 rad) %d, min input amp/max %.2e"
          % (l, float(e.max()), idx, float(oa[idx]), float(oa.max()), float(dphi.max()), int(big.sum()), float(ov.amplitude[l].min() / ov.amplitude[l].max())))
    if int(big.sum()):
        w = torch.nonzero(big)[:5]
        for t in w:
            t = tuple(int(x) for x in t)
            print("     input dphase %.3f at %s: oracle amp %.3e phase %.4f | gpu amp %.3e phase %.4f" % (float(dphi[t]), t, float(ov.amplitude[l][t]), float(ov.phase[l][t]), float(ia[t]), float(ip[t])))
print("low_level err", float((pipe.stages["low_level"].cpu() - ost["low_level"]).abs().max()), "lab_pred err", float((pipe.stages["lab_pred"].cpu() - ost["lab_pred"]).abs().max()))

# ---- bisect: which part of the GPU decomposition makes the fine-level predictions differ?
from fvfi import utils
from fvfi.pyramid import DecompValues
print("---- bisect (stepwise network on mixed inputs)")
dev = "cuda"
og = DecompValues(high_level=ov.high_level.to(dev), low_level=ov.low_level.to(dev), phase=[p.to(dev) for p in ov.phase], amplitude=[a.to(dev) for a in ov.amplitude])
def net(v):
    with torch.no_grad():
        vin = pipe.phase_net.normalize_vals(utils.get_concat_layers_inf(pipe.pyr, utils.separate_vals(v, 2)))
        return pipe.phase_net(vin)
def err0(pred, l=0):
    op, oa = ost["phase%d" % l], ost["amp%d" % l]
    e = (pred.amplitude[l].cpu().double() * torch.exp(1j * pred.phase[l].cpu().double()) - oa.double() * torch.exp(1j * op.double())).abs()
    return float(e.max())
hl = torch.zeros_like(og.high_level)
gvz = gv._replace(high_level=hl)
ogz = og._replace(high_level=hl)
print("all gpu inputs      : level0 err %.2e level1 err %.2e" % (err0(net(gvz)), err0(net(gvz), 1)))
print("all oracle inputs   : level0 err %.2e level1 err %.2e" % (err0(net(ogz)), err0(net(ogz), 1)))
print("oracle + gpu low    : %.2e" % err0(net(ogz._replace(low_level=gv.low_level))))
for l in range(L):
    ph = list(og.phase); am = list(og.amplitude)
    ph[l] = gv.phase[l]
    print("oracle + gpu phase[%d]: level0 %.2e level1 %.2e" % (l, err0(net(ogz._replace(phase=ph))), err0(net(ogz._replace(phase=ph)), 1)), end=" | ")
    ph = list(og.phase); am[l] = gv.amplitude[l]
    print("gpu amp[%d]: level0 %.2e level1 %.2e" % (l, err0(net(ogz._replace(amplitude=am))), err0(net(ogz._replace(amplitude=am)), 1)))

print("---- detail at the level-0 argmax")
l = 0
pred = net(gvz)
op, oa = ost["phase0"], ost["amp0"]
e = (pred.amplitude[l].cpu().double() * torch.exp(1j * pred.phase[l].cpu().double()) - oa.double() * torch.exp(1j * op.double())).abs()
c, _, y, x = [int(v) for v in np.unravel_index(int(e.argmax()), e.shape)]
plane, band = c // 4, c % 4
print("argmax channel %d (plane %d band %d) y %d x %d: pred gpu amp %.4e phase %.5f | oracle amp %.4e phase %.5f"
      % (c, plane, band, y, x, float(pred.amplitude[l][c, 0, y, x]), float(pred.phase[l][c, 0, y, x]), float(oa[c, 0, y, x]), float(op[c, 0, y, x])))
P = 3
for frame in range(2):
    for b in range(4):
        ch = (frame * P + plane) * 4 + b
        for dy in (-1, 0, 1):
            row = []
            for dx in (-1, 0, 1):
                yy, xx = min(max(y + dy, 0), H - 1), min(max(x + dx, 0), W - 1)
                gpv, opv = float(gv.phase[l][ch, 0, yy, xx]), float(ov.phase[l][ch, 0, yy, xx])
                gav, oav = float(gv.amplitude[l][ch, 0, yy, xx]), float(ov.amplitude[l][ch, 0, yy, xx])
                row.append("ph %.5f/%.5f amp %.2e/%.2e" % (gpv, opv, gav, oav))
            print("  frame %d band %d dy %+d: " % (frame, b, dy) + " | ".join(row))
# global statistics of the level-0 input phase difference
d = (gv.phase[0].cpu() - ov.phase[0]); d = torch.atan2(torch.sin(d), torch.cos(d)).abs()
a = ov.amplitude[0] / ov.amplitude[0].max()
for lo, hi in ((0, 1e-4), (1e-4, 1e-3), (1e-3, 1e-2), (1e-2, 1e-1), (1e-1, 2)):
    sel = (a >= lo) & (a < hi)
    if sel.any():
        print("level 0 input amp/max in [%g,%g): n=%d max dphase %.2e median %.2e" % (lo, hi, int(sel.sum()), float(d[sel].max()), float(d[sel].median())))
raw_d = (gv.phase[0].cpu() - ov.phase[0]).abs()
print("unwrapped |dphase| > 1: ", int((raw_d > 1).sum()), " in (0.05, 3]:", int(((raw_d > 0.05) & (raw_d <= 3)).sum()))
