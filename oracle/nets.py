"""oracle/nets.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

CPU (torch, fp32) restatements of the three networks on the fusion path and of the DecompValues
plumbing around them, each citing the reference lines it follows.  They are checked against the
REAL reference modules (imported from /root/reference by oracle/ref_import.py) in
tests/test_models_oracle.py and against the committed golden fixtures produced by
tests/golden/make_golden_models.py, and they are what bench.py times as the CPU baseline on the
GPU box (where /root/reference does not exist).  State-dict keys equal the reference's, so one
seeded state_dict drives reference, oracle and product.
"""
import math
from collections import namedtuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import adacof as oa

DecompValues = namedtuple('values', 'high_level, phase, amplitude, low_level')   # src/train/pyramid.py:12-18


# ---------------------------------------------------------------------------- pyramid wrapper
class Pyramid:
    """src/train/pyramid.py:20-112 over the steerable shim (vectorised loops, same layouts)."""

    def __init__(self, height, nbands, scale_factor, device=None, precision="fp32"):
        from oracle.steerable_shim import SCFpyr_PyTorch
        self.height, self.nbands, self.scale_factor, self.device = height, nbands, scale_factor, torch.device("cpu")
        self.pyr = SCFpyr_PyTorch(height=height, nbands=nbands, scale_factor=scale_factor, precision=precision)

    def filter(self, img):                                                   # :35-39
        return self.coeff_to_values(self.pyr.build(img.unsqueeze(1)))

    def inv_filter(self, vals):                                              # :41-46
        return self.pyr.reconstruct(self.values_to_coeff(vals))

    def coeff_to_values(self, coeff):                                        # :48-78
        phase, amplitude = [], []
        for lv in coeff[1:-1]:
            z = torch.stack([torch.view_as_complex(b) for b in lv], 1)       # [N, nb, h, w]: channel = plane*nb + band (:64-66)
            z = z.reshape(-1, 1, z.shape[-2], z.shape[-1])
            phase.append(torch.imag(torch.log(z)))                           # :63
            amplitude.append(torch.abs(z))                                   # :67
        return DecompValues(high_level=coeff[0].unsqueeze(1), low_level=coeff[-1].unsqueeze(1), phase=phase,
                            amplitude=amplitude)

    def values_to_coeff(self, values):                                       # :85-112
        ndims = values.high_level.shape[0]
        coeff = [values.high_level.squeeze(1)]
        for ph, am in zip(values.phase, values.amplitude):
            if isinstance(ph, (int, float)):
                coeff.append(0)
                continue
            nb = ph.shape[0] // ndims
            z = torch.stack((torch.cos(ph) * am, torch.sin(ph) * am), -1).reshape(ndims, nb, ph.shape[2], ph.shape[3], 2)
            coeff.append([z[:, b].contiguous() for b in range(nb)])
        coeff.append(values.low_level.squeeze(1))
        return coeff


def separate_vals(vals, num_input):                                          # src/train/utils.py:83-127
    sp = lambda t: t.reshape(num_input, -1, t.shape[2], t.shape[3])
    return [DecompValues(high_level=sp(vals.high_level)[i].unsqueeze(1), low_level=sp(vals.low_level)[i].unsqueeze(1),
                         phase=[sp(p)[i].unsqueeze(1) for p in vals.phase],
                         amplitude=[sp(a)[i].unsqueeze(1) for a in vals.amplitude]) for i in range(num_input)]


def get_concat_layers_inf(pyr, vals_list):                                   # src/train/utils.py:47-80
    nb = pyr.nbands
    rs = lambda x: x.reshape(x.shape[0] // nb, nb, x.shape[2], x.shape[3])
    n = pyr.height - 2
    return DecompValues(high_level=torch.cat([e.high_level for e in vals_list], 1),
                        low_level=torch.cat([e.low_level for e in vals_list], 1),
                        phase=[torch.cat([rs(e.phase[i]) for e in vals_list], 1) for i in range(n)][::-1],
                        amplitude=[torch.cat([rs(e.amplitude[i]) for e in vals_list], 1) for i in range(n)][::-1])


def get_last_value_levels(vals, use_levels=1):                               # src/train/utils.py:242-280
    z = torch.zeros_like
    return DecompValues(high_level=vals.high_level.clone(), low_level=z(vals.low_level),
                        phase=[p.clone() if i < use_levels else z(p) for i, p in enumerate(vals.phase)],
                        amplitude=[a.clone() if i < use_levels else z(a) for i, a in enumerate(vals.amplitude)])


def get_first_value_levels(vals, use_levels=1):                              # src/train/utils.py:282-320
    z = torch.zeros_like
    n = len(vals.phase)
    return DecompValues(high_level=z(vals.high_level), low_level=vals.low_level.clone(),
                        phase=[z(p) if i < n - use_levels else p.clone() for i, p in enumerate(vals.phase)],
                        amplitude=[z(a) if i < n - use_levels else a.clone() for i, a in enumerate(vals.amplitude)])


def subtract_values(v1, v2):                                                 # src/train/utils.py:322-346
    return DecompValues(high_level=(v1.high_level - v2.high_level).abs(), low_level=(v1.low_level - v2.low_level).abs(),
                        phase=[(a - b).abs() for a, b in zip(v1.phase, v2.phase)],
                        amplitude=[(a - b).abs() for a, b in zip(v1.amplitude, v2.amplitude)])


def calc_pyr_height(img):                                                    # src/train/utils.py:168-171
    return int(np.ceil((np.log2(min(img.shape[1:])) - 3) * 2) + 2)


# ---------------------------------------------------------------------------- PhaseNet
class PhaseNetBlock(nn.Module):                                              # src/phase_net/phase_net.py:179-207
    def __init__(self, c_in, c_out, pred_out, kernel_size):
        super().__init__()
        pad = 1 if kernel_size == (3, 3) else 0
        self.feature_map = nn.Sequential(nn.Conv2d(c_in, c_out, kernel_size, padding=pad, padding_mode='reflect'),
                                         nn.BatchNorm2d(c_out), nn.ELU(),
                                         nn.Conv2d(c_out, c_out, kernel_size, padding=pad, padding_mode='reflect'), nn.ELU())
        self.prediction_map = nn.Sequential(nn.Conv2d(c_out, pred_out, (1, 1), padding_mode='reflect'), nn.Tanh())

    def forward(self, x):
        f = self.feature_map(x)
        return f, self.prediction_map(f)


class PhaseNet(nn.Module):                                                   # src/phase_net/phase_net.py:7-177 (num_img = 2)
    def __init__(self, pyr):
        super().__init__()
        self.pyr, self.eps = pyr, 1e-8
        self.layers = nn.ModuleList([PhaseNetBlock(2, 64, 1, (1, 1)), PhaseNetBlock(64 + 1 + 16, 64, 8, (1, 1)),
                                     PhaseNetBlock(64 + 8 + 16, 64, 8, (1, 1)),
                                     *[PhaseNetBlock(64 + 8 + 16, 64, 8, (3, 3)) for _ in range(5)]])

    def normalize_vals(self, vals):                                          # :42-78
        bs = vals.amplitude[0].shape[0]
        self.max_amplitudes = [a.reshape(bs, -1).max(1)[0] + self.eps for a in vals.amplitude]
        amps = [a / m.view(-1, 1, 1, 1) for a, m in zip(vals.amplitude, self.max_amplitudes)]
        self.max_low_level = vals.low_level.reshape(bs, -1).max(1)[0] + self.eps
        return DecompValues(high_level=vals.high_level, low_level=vals.low_level / self.max_low_level.view(-1, 1, 1, 1),
                            amplitude=amps, phase=[x / math.pi for x in vals.phase])

    def forward(self, vals):                                                 # :107-177
        m = self.pyr.height - 2
        feature, prediction = self.layers[0](vals.low_level)
        alpha = (prediction[:, 0] + 1) / 2
        low = (alpha * vals.low_level[:, 0] + (1 - alpha) * vals.low_level[:, 1]).unsqueeze(1)
        phases, amps = [], []
        for idx in range(m):
            res = tuple(vals.phase[idx].shape[2:])
            fr = nn.Upsample(res, mode='bilinear')(feature)
            pr = nn.Upsample(res, mode='bilinear')(prediction)
            i = idx + 1 if idx + 1 < len(self.layers) - 1 else len(self.layers) - 1
            feature, prediction = self.layers[i](torch.cat((fr, vals.phase[idx], vals.amplitude[idx], pr), 1))
            beta = (prediction[:, 4:8] + 1) / 2
            amp = beta * vals.amplitude[idx][:, 4:8] + (1 - beta) * vals.amplitude[idx][:, :4]
            phases.append(prediction[:, :4].reshape(-1, 1, *res))
            amps.append(amp.reshape(-1, 1, *res))
        # reverse_normalize :80-105
        phases = [x * math.pi for x in phases]
        amps = [(a.reshape(a.shape[0] // self.pyr.nbands, -1) * mx.view(-1, 1)).reshape(a.shape)
                for a, mx in zip(amps, self.max_amplitudes)]
        hl = vals.high_level.shape
        return DecompValues(high_level=torch.zeros((hl[0], 1, hl[2], hl[3]), dtype=low.dtype), low_level=low * self.max_low_level.view(-1, 1, 1, 1),
                            amplitude=amps[::-1], phase=phases[::-1])


# ---------------------------------------------------------------------------- FusionNet
class FusionNet(nn.Module):                                                  # src/fusion_net/fusion_net.py:6-77
    def __init__(self, num_imgs=5, uncertainty_maps=3, kernel=3, pad=3, dil=3):
        super().__init__()
        cin = 3 * num_imgs + uncertainty_maps
        self.net = nn.Sequential(nn.Conv2d(cin, 64, kernel, 1, pad, dil), nn.ReLU(), nn.Conv2d(64, 64, kernel, 1, pad, dil),
                                 nn.ReLU(), nn.Conv2d(64, 64, kernel, 1, pad, dil), nn.ReLU(),
                                 nn.Conv2d(64, 3, kernel, 1, pad, dil), nn.Tanh())           # dead weights, :11-20
        r = dict(stride=1, padding_mode='reflect')
        self.encoder_layers = nn.ModuleList([nn.Conv2d(cin, 32, 5, padding=2, **r), nn.Conv2d(32, 64, 5, padding=2, **r),
                                             nn.Conv2d(64, 128, 3, padding=1, **r)])
        self.bottleneck_layer = nn.Conv2d(128, 128, 3, padding=1, **r)
        self.decoder_layers = nn.ModuleList([nn.Conv2d(128, 64, 5, padding=2, **r), nn.Conv2d(64, 32, 5, padding=2, **r),
                                             nn.Conv2d(32, 3, 1)])

    def forward(self, base, adacof, phase, other, maps):                     # :46-77 (variant 0)
        x = torch.cat([base, adacof, phase, other, maps], 1)
        skip = []
        for layer in self.encoder_layers:
            x = F.relu(layer(x))
            skip.append(x)
            x = F.max_pool2d(x, 2, 2)
        x = self.bottleneck_layer(x)
        for layer, s in zip(self.decoder_layers, skip[::-1]):
            x = F.interpolate(F.relu(x), scale_factor=2, mode='bilinear') + s
            x = layer(x)
        return (base + torch.tanh(x)).clamp(0, 1)


# ---------------------------------------------------------------------------- AdaCoFNet
class KernelEstimation(nn.Module):                                           # src/fusion_net/fusion_adacofnet.py:14-155
    def __init__(self, kernel_size):
        super().__init__()
        c3 = lambda i, o: nn.Conv2d(i, o, 3, 1, 1)
        basic = lambda i, o: nn.Sequential(c3(i, o), nn.ReLU(), c3(o, o), nn.ReLU(), c3(o, o), nn.ReLU())
        up = lambda c: nn.Sequential(nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True), c3(c, c), nn.ReLU())

        def subnet(ks, tail=None, mid=None):
            mid = ks if mid is None else mid
            layers = [c3(64, 64), nn.ReLU(), c3(64, 64), nn.ReLU(), c3(64, mid), nn.ReLU(),
                      nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True), c3(mid, ks)]
            return nn.Sequential(*(layers + ([tail] if tail is not None else [])))
        ks = kernel_size ** 2
        self.moduleConv1, self.modulePool1 = basic(6, 32), nn.AvgPool2d(2, 2)
        self.moduleConv2, self.modulePool2 = basic(32, 64), nn.AvgPool2d(2, 2)
        self.moduleConv3, self.modulePool3 = basic(64, 128), nn.AvgPool2d(2, 2)
        self.moduleConv4, self.modulePool4 = basic(128, 256), nn.AvgPool2d(2, 2)
        self.moduleConv5, self.modulePool5 = basic(256, 512), nn.AvgPool2d(2, 2)
        self.moduleDeconv5, self.moduleUpsample5 = basic(512, 512), up(512)
        self.moduleDeconv4, self.moduleUpsample4 = basic(512, 256), up(256)
        self.moduleDeconv3, self.moduleUpsample3 = basic(256, 128), up(128)
        self.moduleDeconv2, self.moduleUpsample2 = basic(128, 64), up(64)
        self.moduleWeight1, self.moduleAlpha1, self.moduleBeta1 = subnet(ks, nn.Softmax(dim=1)), subnet(ks), subnet(ks)
        self.moduleWeight2, self.moduleAlpha2, self.moduleBeta2 = subnet(ks, nn.Softmax(dim=1)), subnet(ks), subnet(ks)
        self.moduleOcclusion = subnet(1, nn.Sigmoid(), mid=64)

    def forward(self, r0, r2):                                               # :109-155
        c1 = self.moduleConv1(torch.cat([r0, r2], 1))
        c2 = self.moduleConv2(self.modulePool1(c1))
        c3 = self.moduleConv3(self.modulePool2(c2))
        c4 = self.moduleConv4(self.modulePool3(c3))
        c5 = self.moduleConv5(self.modulePool4(c4))
        x = self.moduleUpsample5(self.moduleDeconv5(self.modulePool5(c5))) + c5
        x = self.moduleUpsample4(self.moduleDeconv4(x)) + c4
        x = self.moduleUpsample3(self.moduleDeconv3(x)) + c3
        x = self.moduleUpsample2(self.moduleDeconv2(x)) + c2
        return (self.moduleWeight1(x), self.moduleAlpha1(x), self.moduleBeta1(x), self.moduleWeight2(x),
                self.moduleAlpha2(x), self.moduleBeta2(x), self.moduleOcclusion(x))


class AdaCoFNet(nn.Module):                                                  # src/fusion_net/fusion_adacofnet.py:158-240
    def __init__(self, kernel_size=5, dilation=1, threads=1):
        super().__init__()
        self.kernel_size, self.dilation, self.threads = kernel_size, dilation, threads
        self.kernel_pad = int(((kernel_size - 1) * dilation) / 2.0)
        self.get_kernel = KernelEstimation(kernel_size)

    def forward(self, frame0, frame2):
        h0, w0 = frame0.shape[2:]
        if h0 % 32:                                                           # :182-186
            frame0, frame2 = (F.pad(f, (0, 0, 0, 32 - h0 % 32), mode='reflect') for f in (frame0, frame2))
        if w0 % 32:                                                           # :188-192
            frame0, frame2 = (F.pad(f, (0, 32 - w0 % 32, 0, 0), mode='reflect') for f in (frame0, frame2))
        mean = torch.tensor([0.4631, 0.4352, 0.3990], dtype=frame0.dtype).view(1, 3, 1, 1)       # src/adacof/utility.py:86-87
        maps = [m.contiguous() for m in self.get_kernel(frame0 - mean, frame2 - mean)]
        W1, A1, B1, W2, A2, B2, Occ = [m.numpy() for m in maps]
        p = self.kernel_pad
        pad = lambda f: F.pad(f, [p, p, p, p], mode='replicate').numpy()     # :168
        t1 = oa.forward(pad(frame0), W1, A1, B1, self.dilation, threads=self.threads)   # :195
        t2 = oa.forward(pad(frame2), W2, A2, B2, self.dilation, threads=self.threads)   # :196
        frame1, mask = oa.adacofnet_tail(t1, t2, Occ, W1, A1, B1, W2, A2, B2, threads=self.threads)   # :198-213
        crop = lambda a: torch.from_numpy(a)[:, :, :h0, :w0]
        return crop(t1), crop(t2), crop(frame1), crop(mask)
