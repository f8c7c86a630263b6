"""Drop-in for the reference's ``src/train/pyramid.py`` (Pyramid, DecompValues).

``Pyramid(height, nbands, scale_factor, device)`` with ``filter(img[N,H,W]) -> DecompValues`` and
``inv_filter(DecompValues) -> img[N,H,W]`` (pyramid.py:23-46).  ``filter`` is the fused path: one
forward FFT2 per plane, then per level band-mask -> inverse FFT -> amplitude/phase epilogue, so
the complex coefficients never reach HBM and ``coeff_to_values``'s Python loops (pyramid.py:48-78)
disappear.  ``coeff_to_values`` / ``values_to_coeff`` are kept for API compatibility (vectorised).
"""
from collections import namedtuple

import torch

from . import _lib
from .pyr_plan import PyrPlan, ptr_array
from .steerable import SCFpyr_PyTorch

DecompValues = namedtuple(
    'values',
    'high_level, '
    'phase, '
    'amplitude, '
    'low_level'
)


class _InvFilterFn(torch.autograd.Function):
    """Pyramid.inv_filter with a backward (fvfi_pyr_reconstruct_backward): what PhaseNet training needs to push the L1 image
    loss through the reconstruction (src/train/trainer.py:139-147).  Inputs: plan, L, high-or-None, low-or-None, then the L
    phase tensors and the L amplitude tensors (None = level absent)."""

    @staticmethod
    def forward(ctx, plan, L, high, low, *levels):
        phase, amp = list(levels[:L]), list(levels[L:])
        ref = next(t for t in [high, low] + phase + amp if t is not None)
        N = (high if high is not None else low).shape[0]
        H, W = plan.H, plan.W
        out = torch.empty((N, H, W), dtype=torch.float32, device=ref.device)
        with torch.cuda.device(ref.device):
            _lib.check(_lib.lib().fvfi_pyr_reconstruct(plan.handle, _lib.ptr(high), ptr_array(phase), ptr_array(amp),
                                                       _lib.ptr(low), N, out.data_ptr(), plan.workspace(N).data_ptr(),
                                                       _lib.stream_ptr()))
        ctx.plan, ctx.L, ctx.N = plan, L, N
        ctx.has = (high is not None, low is not None)
        ctx.save_for_backward(*[t for t in phase + amp if t is not None])
        ctx.present = [t is not None for t in phase]
        return out

    @staticmethod
    def backward(ctx, gout):
        plan, L, N = ctx.plan, ctx.L, ctx.N
        saved = list(ctx.saved_tensors)
        npres = sum(ctx.present)
        ph_s, am_s = saved[:npres], saved[npres:]
        phase, amp, gph, gam = [None] * L, [None] * L, [None] * L, [None] * L
        it = 0
        for l in range(L):
            if ctx.present[l]:
                phase[l], amp[l] = ph_s[it], am_s[it]
                gph[l], gam[l] = torch.empty_like(ph_s[it]), torch.empty_like(am_s[it])
                it += 1
        g = gout.contiguous().float()
        new = lambda *s: torch.empty(s, dtype=torch.float32, device=g.device)
        ghigh = new(N, 1, plan.H, plan.W) if ctx.has[0] else None
        glow = new(N, 1, *plan.shapes[-1]) if ctx.has[1] else None
        with torch.cuda.device(g.device):
            _lib.check(_lib.lib().fvfi_pyr_reconstruct_backward(plan.handle, g.data_ptr(), N, ptr_array(phase), ptr_array(amp),
                                                                _lib.ptr(ghigh), ptr_array(gph), ptr_array(gam),
                                                                _lib.ptr(glow), plan.workspace(N).data_ptr(),
                                                                _lib.stream_ptr()))
        return (None, None, ghigh, glow) + tuple(gph) + tuple(gam)


class Pyramid:
    """ Steerable Pyramid Decomposition (B200-native). """

    def __init__(self, height, nbands, scale_factor, device):
        self.height = height
        self.nbands = nbands
        self.scale_factor = scale_factor
        self.device = device
        self.pyr = SCFpyr_PyTorch(height=self.height, nbands=self.nbands, scale_factor=self.scale_factor,
                                  device=self.device)
        self.last_amp_max = None  # [L, N] per-level per-plane max amplitude of the last filter() call

    def _plan(self, H, W, device):
        return PyrPlan.get(H, W, self.height, self.nbands, self.scale_factor, device)

    def filter(self, img, want_high=True, levels=None):
        """ Psi filter: img [N,H,W] -> DecompValues (layouts of src/train/pyramid.py:48-78).
        ``levels`` (optional): only these band levels are computed, the others are ``None`` in the result (the
        uncertainty branch of the recipe reads level 0 and the six coarsest levels only). """
        if not img.is_cuda:
            raise NotImplementedError("fvfi pyramid: CUDA tensors only")
        img = img.contiguous().float()
        N, H, W = img.shape
        plan = self._plan(H, W, img.device)
        nb = self.nbands
        new = lambda *s: torch.empty(s, dtype=torch.float32, device=img.device)
        high = new(N, 1, H, W) if want_high else None
        keep = set(range(plan.L)) if levels is None else set(levels)
        phase = [new(N * nb, 1, h, w) if l in keep else None for l, (h, w) in enumerate(plan.shapes[:-1])]
        amp = [new(N * nb, 1, h, w) if l in keep else None for l, (h, w) in enumerate(plan.shapes[:-1])]
        low = new(N, 1, *plan.shapes[-1])
        amp_max = new(plan.L, N)
        with torch.cuda.device(img.device):
            _lib.check(_lib.lib().fvfi_pyr_decompose(plan.handle, img.data_ptr(), N, _lib.ptr(high), ptr_array(phase),
                                                     ptr_array(amp), low.data_ptr(), amp_max.data_ptr(),
                                                     plan.workspace(N).data_ptr(), _lib.stream_ptr()))
        self.last_amp_max = amp_max
        if high is None:     # not computed: a stride-0 view of one zero (shape-compatible with the reference's layout, no storage)
            high = torch.zeros((1,), dtype=torch.float32, device=img.device).expand(N, 1, H, W)
        return DecompValues(high_level=high, phase=phase, amplitude=amp, low_level=low)

    def inv_filter(self, vals):
        """ Psi^{-1} filter: DecompValues -> img [N,H,W].  Levels given as the int 0 are skipped
        (src/phase_net/phase_net.py:91-93). """
        high = vals.high_level
        N, _, H, W = high.shape
        plan = self._plan(H, W, high.device)
        keep = []

        def prep(t):
            if isinstance(t, (int, float)):
                return None
            t = t.contiguous().float()
            keep.append(t)
            return t
        phase = [prep(t) for t in vals.phase]
        amp = [prep(t) for t in vals.amplitude]
        high_c, low_c = prep(high), prep(vals.low_level)
        if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in phase + amp + [high_c, low_c]):
            for l in range(len(phase)):            # a level is used only if both of its tensors are present
                if phase[l] is None or amp[l] is None:
                    phase[l] = amp[l] = None
            return _InvFilterFn.apply(plan, plan.L, high_c, low_c, *(phase + amp))
        out = torch.empty((N, H, W), dtype=torch.float32, device=high.device)
        with torch.cuda.device(high.device):
            _lib.check(_lib.lib().fvfi_pyr_reconstruct(plan.handle, high_c.data_ptr(), ptr_array(phase), ptr_array(amp),
                                                       low_c.data_ptr(), N, out.data_ptr(),
                                                       plan.workspace(N).data_ptr(), _lib.stream_ptr()))
        return out

    def inv_filter_sparse(self, vals, use_high=True, use_low=True, levels=None):
        """inv_filter with whole components dropped instead of zero-filled copies -- the fused form
        of get_last_value_levels / get_first_value_levels (src/train/utils.py:242-320)."""
        high = vals.high_level
        N, _, H, W = high.shape
        plan = self._plan(H, W, high.device)
        L = plan.L
        levels = set(range(L)) if levels is None else set(levels)
        ok = lambda t: torch.is_tensor(t)          # levels given as the int 0 are absent (phase_net.py:91-93)
        phase = [vals.phase[l].contiguous() if l in levels and ok(vals.phase[l]) and ok(vals.amplitude[l]) else None for l in range(L)]
        amp = [vals.amplitude[l].contiguous() if phase[l] is not None else None for l in range(L)]
        high_c = high.contiguous() if use_high else None
        low_c = vals.low_level.contiguous() if use_low else None
        out = torch.empty((N, H, W), dtype=torch.float32, device=high.device)
        with torch.cuda.device(high.device):
            _lib.check(_lib.lib().fvfi_pyr_reconstruct(plan.handle, _lib.ptr(high_c), ptr_array(phase), ptr_array(amp),
                                                       _lib.ptr(low_c), N, out.data_ptr(),
                                                       plan.workspace(N).data_ptr(), _lib.stream_ptr()))
        return out

    def highband_filter(self, img):
        """``inv_filter(get_last_value_levels(filter(img), 1))`` (src/train/utils.py:242-280; the h_freq maps of
        src/fusion_net/interpolate_twoframe.py:205-209): the high residual plus the finest band level of ``img`` [N,H,W], put
        back together.  One spectral multiplication with a plan table instead of a decomposition and a reconstruction."""
        if not img.is_cuda:
            raise NotImplementedError("fvfi pyramid: CUDA tensors only")
        img = img.contiguous().float()
        N, H, W = img.shape
        plan = self._plan(H, W, img.device)
        out = torch.empty_like(img)
        with torch.cuda.device(img.device):
            _lib.check(_lib.lib().fvfi_pyr_highband_filter(plan.handle, img.data_ptr(), N, out.data_ptr(),
                                                           plan.workspace(N).data_ptr(), _lib.stream_ptr()))
        return out

    def inv_filter_bands(self, bands, N, H, W, high=None):
        """Reconstruction from COMPLEX band coefficients of some levels only: ``bands`` = {level: [nb tensors [N,h,w,2]]}, optionally
        the high-pass residual ``high`` [N,H,W]; the low residual and the other levels are absent (contribute nothing).  -> [N,H,W]."""
        dev = next(iter(bands.values()))[0].device
        plan = self._plan(H, W, dev)
        flat = [None] * (plan.L * self.nbands)
        keep = []
        for l, lst in bands.items():
            for b, t in enumerate(lst):
                assert tuple(t.shape) == (N,) + tuple(plan.shapes[l]) + (2,) and t.is_cuda
                t = t.contiguous().float()
                keep.append(t)
                flat[l * self.nbands + b] = t
        out = torch.empty((N, H, W), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            if high is not None:
                high = high.reshape(N, H, W).contiguous().float()
            _lib.check(_lib.lib().fvfi_pyr_reconstruct_complex(plan.handle, _lib.ptr(high), ptr_array(flat), None, N, out.data_ptr(),
                                                               plan.workspace(N).data_ptr(), _lib.stream_ptr()))
        return out

    # ---- API-compat helpers (the reference calls these from filter / inv_filter) -------------
    def coeff_to_values(self, coeff):
        """pyramid.py:48-78, vectorised: band list -> (phase, amplitude) in [N*nb,1,h,w] layout."""
        nlevels = len(coeff) - 2
        phase, amplitude = [], []
        for level in range(nlevels):
            z = torch.stack([torch.view_as_complex(b.contiguous()) for b in coeff[level + 1]], 1)  # [N,nb,h,w]
            z = z.reshape(-1, 1, z.shape[-2], z.shape[-1])
            phase.append(torch.angle(z))        # imag(log z)
            amplitude.append(torch.abs(z))
        return DecompValues(high_level=coeff[0].unsqueeze(1), low_level=coeff[-1].unsqueeze(1), phase=phase,
                            amplitude=amplitude)

    def reorder(self, input, ndims):
        """pyramid.py:80-83."""
        nbands = int(input[0].shape[0] / ndims)
        return [[x.reshape(ndims, nbands, x.shape[2], x.shape[3])[:, j] for j in range(nbands)] for x in input]

    def values_to_coeff(self, values):
        """pyramid.py:85-112, vectorised: polar -> complex band list."""
        ndims = values.high_level.shape[0]
        coeff = [values.high_level.squeeze(1)]
        for ph, am in zip(values.phase, values.amplitude):
            nb = ph.shape[0] // ndims
            z = torch.stack((torch.cos(ph) * am, torch.sin(ph) * am), -1)           # [N*nb,1,h,w,2]
            z = z.reshape(ndims, nb, ph.shape[2], ph.shape[3], 2)
            coeff.append([z[:, b].contiguous() for b in range(nb)])
        coeff.append(values.low_level.squeeze(1))
        return coeff
