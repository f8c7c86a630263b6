"""Developer tool (GPU box): the uncertainty branch (src/fusion_net/interpolate_twoframe.py:197-225) stage by stage vs the oracle."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from oracle import fusion_pipeline as fp
from oracle.wrap_align import WrapAligner
from fvfi.pipeline import FusionPipeline
from fvfi import utils
from fvfi.pyramid import DecompValues
H, W = int(sys.argv[1]), int(sys.argv[2])
state = fp.seeded_state(0)
r1, r2 = fp.seeded_frames(1, H, W, 0)
torch.set_num_threads(os.cpu_count())
ost, odec = {}, {}
be = fp.oracle_backend(state, hw=(H, W), threads=16)
fp.interp(be, r1, r2, ost, odec)
pipe = FusionPipeline(H, W, "cuda")
pipe.load_state(state)
cap = {}
al = WrapAligner.from_decomps(odec)
def hook(tag, planes, vals):
    cap[tag + "_planes"] = planes
    cap[tag + "_raw"] = vals
    v = al(tag, planes, vals)
    cap[tag] = v
    return v
pipe.filter_hook = hook
pipe.stages = {}
with torch.no_grad():
    pipe(r1.cuda(), r2.cuda())
print("flips", al.flips)
for k in ("ada_pred", "phase_pred", "freq_diff", "ada_uncertainty", "h_freq_diff"):
    print(k, "max abs err %.2e" % float((pipe.stages[k].cpu() - ost[k]).abs().max()))
ov, gv = odec["uncertainty"], cap["uncertainty"]
L = len(ov.phase)
pl = cap["uncertainty_planes"].cpu()
ref_planes = torch.cat((ost["ada_pred"].reshape(-1, H, W), ost["phase_pred"].reshape(-1, H, W)), 0)
print("input planes of the uncertainty call: max abs diff %.2e" % float((pl - ref_planes).abs().max()))
for l in range(L - 6, L):
    g = gv.amplitude[l].cpu().double() * torch.exp(1j * gv.phase[l].cpu().double())
    o = ov.amplitude[l].double() * torch.exp(1j * ov.phase[l].double())
    m = float(o.abs().max())
    dph = (gv.phase[l].cpu() - ov.phase[l]).abs()
    # the quantity the recipe uses: |phase_ph - phase_ada|
    P = gv.phase[l].shape[0] // 2
    dg = (gv.phase[l][P:] - gv.phase[l][:P]).abs().cpu()
    do = (ov.phase[l][P:] - ov.phase[l][:P]).abs()
    print("level %2d %3dx%-3d coef err/max %.2e | max |dphase| %.2e (n>0.5: %d) | max err of |phi_ph - phi_ada| %.2e | min amp/max %.2e"
          % (l, g.shape[-2], g.shape[-1], float((g - o).abs().max()) / m, float(dph.max()), int((dph > 0.5).sum()), float((dg - do).abs().max()),
             float(ov.amplitude[l].min()) / m))
# same call on the ORACLE's planes: isolates the decomposition from the input difference
with torch.no_grad():
    gv2 = pipe.pyr.filter(ref_planes.cuda(), want_high=False, levels=list(range(L - 6, L)))
for l in range(L - 6, L):
    g = gv2.amplitude[l].cpu().double() * torch.exp(1j * gv2.phase[l].cpu().double())
    o = ov.amplitude[l].double() * torch.exp(1j * ov.phase[l].double())
    print("  on the oracle's planes: level %2d coef err/max %.2e" % (l, float((g - o).abs().max()) / float(o.abs().max())))
# reconstruct the ORACLE's difference pyramid on the GPU
va, vp = be.separate_vals(ov, 2)
vd = be.get_first_value_levels(be.subtract_values(vp, va), use_levels=6)
ref_fd = be.pyr.inv_filter(vd).reshape(1, 3, H, W).mean(1) * 30
dv = DecompValues(high_level=vd.high_level.cuda(), low_level=vd.low_level.cuda(), phase=[p.cuda() for p in vd.phase], amplitude=[a.cuda() for a in vd.amplitude])
with torch.no_grad():
    got = pipe.pyr.inv_filter_sparse(dv, use_high=False, levels=range(L - 6, L)).reshape(1, 3, H, W).mean(1) * 30
print("GPU reconstruction of the oracle's difference pyramid: freq_diff max abs err %.2e (max|freq_diff| %.2e)" % (float((got.cpu() - ref_fd).abs().max()), float(ref_fd.abs().max())))
print("oracle freq_diff recomputed vs stage: %.2e" % float((ref_fd - ost["freq_diff"]).abs().max()))
