"""oracle/fusion_pipeline.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

CPU restatement of the reference's end-to-end fusion recipe ``interp``
(src/fusion_net/interpolate_twoframe.py:82-334) without file I/O, generalised from one frame pair
to a batch of B pairs (every stage of the reference is per-sample, so the batch is a stack).
``backend`` supplies the building blocks, so the SAME recipe runs with
  * the real reference modules imported from /root/reference (``reference_backend``; build
    container only -- used to generate tests/golden/pipeline_ref_*.npz), or
  * the oracle restatements in oracle/nets.py (``oracle_backend``; anywhere -- the CPU baseline).
Lab conversion: oracle/lab.py (skimage absent -> parity unpinned there); Gaussian / median:
scipy.ndimage exactly as the reference calls them (:210-214, :221-222).
"""
import types

import numpy as np
import torch
from scipy.ndimage import gaussian_filter, median_filter

from oracle import lab as olab


def rgb2lab_planes(rgb, dtype=torch.float32):
    """src/train/transform.py:6-14 on [B,3,H,W] -> [B,3,H,W] (L/100, (a,b+128)/255)."""
    lab = olab.rgb2lab(rgb.permute(0, 2, 3, 1).numpy())
    lab[..., 0] /= 100
    lab[..., 1:] += 128
    lab[..., 1:] /= 255
    return torch.tensor(lab).permute(0, 3, 1, 2).to(dtype)


def lab2rgb_planes(lab, dtype=torch.float32):
    """src/train/transform.py:28-37."""
    x = lab.clone().permute(0, 2, 3, 1).numpy().astype(np.float64)
    x[..., 0] *= 100
    x[..., 1:] *= 255
    x[..., 1:] -= 128
    return torch.tensor(olab.lab2rgb(x)).permute(0, 3, 1, 2).to(dtype)


def oracle_backend(state, kernel_size=5, dilation=1, threads=1, height=None, hw=None, precision="fp32"):
    """Building blocks from oracle/nets.py; ``state`` = dict(phase_net=..., fusion_net=..., adacof=...) state_dicts.
    ``precision="fp64"``: the same restatement with every stage in double (networks .double(), complex128 pyramid, the fp64
    instantiation of the C warp) -- NOT what the reference computes, but the arbiter that says how far the reference's own
    fp32 run is from the exact result of the same recipe (tests: |GPU - fp64| <= 2 |reference_fp32 - fp64| per stage)."""
    from oracle import nets
    H, W = hw
    pyr = nets.Pyramid(height or nets.calc_pyr_height(torch.empty(3, H, W)), 4, np.sqrt(2), precision=precision)
    pn = nets.PhaseNet(pyr).eval()
    fn = nets.FusionNet().eval()
    an = nets.AdaCoFNet(kernel_size, dilation, threads=threads).eval()
    pn.load_state_dict(state["phase_net"])
    fn.load_state_dict(state["fusion_net"])
    an.load_state_dict(state["adacof"])
    dtype = torch.float64 if precision == "fp64" else torch.float32
    if precision == "fp64":
        pn, fn, an = pn.double(), fn.double(), an.double()
    return types.SimpleNamespace(dtype=dtype, pyr=pyr, phase_net=pn, fusion_net=fn, adacof=an, separate_vals=nets.separate_vals,
                                 get_concat_layers_inf=nets.get_concat_layers_inf,
                                 get_last_value_levels=nets.get_last_value_levels,
                                 get_first_value_levels=nets.get_first_value_levels,
                                 subtract_values=nets.subtract_values)


@torch.no_grad()
def interp(backend, rgb1, rgb2, stages=None, decomps=None):
    """rgb1, rgb2: [B,3,H,W] in [0,1] (CPU).  Returns the fused frame [B,3,H,W]; ``stages`` (dict) receives
    the intermediate tensors named as in the reference script, ``decomps`` (dict) the two raw decompositions
    ("phasenet", "uncertainty") -- what oracle/wrap_align.py needs to know the reference's branch of phases at +-pi."""
    be = backend
    dt = getattr(be, "dtype", torch.float32)
    rgb1, rgb2 = rgb1.to(dt), rgb2.to(dt)
    B, _, H, W = rgb1.shape
    r_shape = (B, 3, H, W)
    lab1, lab2 = rgb2lab_planes(rgb1, dt), rgb2lab_planes(rgb2, dt)                           # :148-149
    ada_frame1, ada_frame2, ada_pred, flow_var_map = be.adacof(rgb1, rgb2)                    # :156
    flow_var_map = flow_var_map.squeeze(1)                                                     # :165
    # PhaseNet branch :168-192
    img_batch = torch.cat((lab1.reshape(-1, H, W), lab2.reshape(-1, H, W)), 0)
    vals_raw = be.pyr.filter(img_batch.to(dt))
    if decomps is not None:
        decomps["phasenet"] = vals_raw
    vals_list = be.separate_vals(vals_raw, 2)
    inp = be.phase_net.normalize_vals(be.get_concat_layers_inf(be.pyr, vals_list))
    vals_pred = be.phase_net(inp)
    lab_pred = be.pyr.inv_filter(vals_pred).reshape(r_shape).to(dt)
    rgb_pred = lab2rgb_planes(lab_pred, dt)
    phase_pred = rgb_pred.clone()
    # uncertainty maps :197-225
    img_batch = torch.cat((ada_pred.reshape(-1, H, W), rgb_pred.reshape(-1, H, W)), 0)
    vals_raw = be.pyr.filter(img_batch.to(dt))
    if decomps is not None:
        decomps["uncertainty"] = vals_raw
    vals_ada, vals_ph = be.separate_vals(vals_raw, 2)
    h_freq = be.pyr.inv_filter(be.get_last_value_levels(vals_ada, use_levels=1)).reshape(r_shape).mean(1)
    h_freq_ph = be.pyr.inv_filter(be.get_last_value_levels(vals_ph, use_levels=1)).reshape(r_shape).mean(1)
    h_freq_diff = (torch.abs(h_freq - h_freq_ph) * 100).clamp(min=0, max=1.0)
    phase_uncertainty = torch.stack([torch.as_tensor(gaussian_filter(h.numpy(), 5)) for h in h_freq_diff])
    vals_diff = be.get_first_value_levels(be.subtract_values(vals_ph, vals_ada), use_levels=6)
    freq_diff = be.pyr.inv_filter(vals_diff).reshape(r_shape).mean(1) * 30
    freq_med = torch.stack([torch.as_tensor(median_filter(f.numpy(), size=50)) for f in freq_diff])
    ada_uncertainty = (torch.abs(freq_diff - freq_med) * 5).clamp(0, 1)
    # baseline :228-238
    inb1 = be.adacof(rgb1, phase_pred)[2].to(dt)
    inb2 = be.adacof(phase_pred, rgb2)[2].to(dt)
    base = be.adacof(inb1, inb2)[2].to(dt)
    # fusion :324-330
    other = torch.cat([lab1, lab2], 1).to(dt)
    maps = torch.stack([ada_uncertainty, phase_uncertainty, flow_var_map], 1).to(dt)
    final = be.fusion_net(base, ada_pred.to(dt), phase_pred, other, maps)
    if stages is not None:
        stages.update(lab1=lab1, lab2=lab2, ada_pred=ada_pred, flow_var_map=flow_var_map, lab_pred=lab_pred,
                      phase_pred=phase_pred, phase_uncertainty=phase_uncertainty, ada_uncertainty=ada_uncertainty,
                      freq_diff=freq_diff, h_freq_diff=h_freq_diff, base=base, final=final)
    return final


@torch.no_grad()
def interp_phasenet(backend, rgb1, rgb2, stages=None, decomps=None):
    """BASELINE.json configs[0]: PhaseNet decompose -> phase/amplitude prediction -> reconstruct on a frame pair
    (src/phase_net/interpolate_twoframe.py:52-107: rgb2lab -> pyr.filter -> concat layers -> normalize_vals -> phase_net ->
    inv_filter -> lab2rgb; the three colour planes are a batch here instead of that script's per-channel loop "to save
    memory" (:83), which is the form the fusion recipe uses, src/fusion_net/interpolate_twoframe.py:168-192).
    Returns the interpolated RGB frame; ``stages`` receives lab_pred, phase_pred and the predicted pyramid values."""
    be = backend
    dt = getattr(be, "dtype", torch.float32)
    rgb1, rgb2 = rgb1.to(dt), rgb2.to(dt)
    B, _, H, W = rgb1.shape
    lab1, lab2 = rgb2lab_planes(rgb1, dt), rgb2lab_planes(rgb2, dt)
    img_batch = torch.cat((lab1.reshape(-1, H, W), lab2.reshape(-1, H, W)), 0)
    vals_raw = be.pyr.filter(img_batch.to(dt))
    if decomps is not None:
        decomps["phasenet"] = vals_raw
    vals_list = be.separate_vals(vals_raw, 2)
    vals_pred = be.phase_net(be.phase_net.normalize_vals(be.get_concat_layers_inf(be.pyr, vals_list)))
    lab_pred = be.pyr.inv_filter(vals_pred).reshape(B, 3, H, W).to(dt)
    phase_pred = lab2rgb_planes(lab_pred, dt)
    if stages is not None:
        stages.update(lab_pred=lab_pred, phase_pred=phase_pred, low_level=vals_pred.low_level)
        for l, (p, a) in enumerate(zip(vals_pred.phase, vals_pred.amplitude)):
            stages["phase%d" % l], stages["amp%d" % l] = p, a
    return phase_pred


def seeded_state(seed, kernel_size=5):
    """Random-init state_dicts (nn.Module default init under torch.manual_seed) shared by reference,
    oracle and product; BatchNorm running stats stay at 0/1 (eval)."""
    from oracle import nets
    torch.manual_seed(seed)
    pyr = types.SimpleNamespace(height=8, nbands=4)
    st = {"phase_net": nets.PhaseNet(pyr).state_dict(), "fusion_net": nets.FusionNet().state_dict(),
          "adacof": nets.AdaCoFNet(kernel_size).state_dict()}
    return st


def seeded_frames(B, H, W, seed):
    """Smooth synthetic frames in [0,1] (low-pass noise + a shifted copy) so Lab stays in gamut."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand((B, 3, H // 4 + 2, W // 4 + 2), generator=g)
    up = torch.nn.functional.interpolate(base, size=(H + 8, W + 8), mode='bicubic', align_corners=False).clamp(0.02, 0.98)
    noise = 0.03 * torch.rand((B, 3, H + 8, W + 8), generator=g)
    full = (up + noise).clamp(0, 1)
    return full[:, :, 2:2 + H, 1:1 + W].contiguous(), full[:, :, 5:5 + H, 6:6 + W].contiguous()
