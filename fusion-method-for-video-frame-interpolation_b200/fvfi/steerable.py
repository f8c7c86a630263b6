"""Drop-in for the third-party ``steerable.SCFpyr_PyTorch`` module the reference imports
(src/train/pyramid.py:7-8; constructor :28-33, ``build`` :37, ``reconstruct`` :44).

Same contract: ``build(im[N,1,H,W]) -> [hi0 [N,H,W], [band_b [N,h_l,w_l,2]]*nbands per level..., lo [N,h_L,w_L]]``
and ``reconstruct(coeff) -> [N,H,W]``; band tensors accept ``torch.view_as_complex``
(pyramid.py:58).  The arithmetic runs in libfvfi's shared-memory FFT kernels.
"""
import torch

from . import _lib
from .pyr_plan import PyrPlan, ptr_array


class SCFpyr_PyTorch(object):
    def __init__(self, height=5, nbands=4, scale_factor=2, device=None):
        self.height = height
        self.nbands = nbands
        self.scale_factor = scale_factor
        self.device = torch.device("cuda") if device is None else torch.device(device)

    def _plan(self, H, W, device):
        return PyrPlan.get(H, W, self.height, self.nbands, self.scale_factor, device)

    def build(self, im_batch):
        assert im_batch.dim() == 4 and im_batch.shape[1] == 1, "Image batch must be of shape [N,1,H,W]"
        if not im_batch.is_cuda:
            raise NotImplementedError("fvfi pyramid: CUDA tensors only")
        im = im_batch.squeeze(1).contiguous().float()
        N, H, W = im.shape
        plan = self._plan(H, W, im.device)
        new = lambda *s: torch.empty(s, dtype=torch.float32, device=im.device)
        hi = new(N, H, W)
        bands = [[new(N, h, w, 2) for _ in range(self.nbands)] for (h, w) in plan.shapes[:-1]]
        lo = new(N, *plan.shapes[-1])
        flat = [b for lv in bands for b in lv]
        with torch.cuda.device(im.device):
            _lib.check(_lib.lib().fvfi_pyr_build_complex(plan.handle, im.data_ptr(), N, hi.data_ptr(), ptr_array(flat),
                                                         lo.data_ptr(), plan.workspace(N).data_ptr(),
                                                         _lib.stream_ptr()))
        return [hi] + bands + [lo]

    def reconstruct(self, coeff):
        if self.nbands != len(coeff[1]):
            raise Exception("Unmatched number of orientations")
        hi, lo = coeff[0], coeff[-1]
        N, H, W = hi.shape
        plan = self._plan(H, W, hi.device)
        flat = []
        keep = [hi, lo]
        for lv in coeff[1:-1]:
            for b in range(self.nbands):
                t = None
                if not isinstance(lv, (int, float)) and not isinstance(lv[b], (int, float)):
                    t = lv[b].contiguous().float()
                    keep.append(t)
                flat.append(t)
        # a level is either fully present or skipped (the reference passes 0 for whole levels)
        for l in range(plan.L):
            grp = flat[l * self.nbands:(l + 1) * self.nbands]
            if any(g is None for g in grp):
                for b in range(self.nbands):
                    flat[l * self.nbands + b] = None
        out = torch.empty((N, H, W), dtype=torch.float32, device=hi.device)
        hi_c, lo_c = hi.contiguous().float(), lo.contiguous().float()
        with torch.cuda.device(hi.device):
            _lib.check(_lib.lib().fvfi_pyr_reconstruct_complex(plan.handle, hi_c.data_ptr(), ptr_array(flat),
                                                               lo_c.data_ptr(), N, out.data_ptr(),
                                                               plan.workspace(N).data_ptr(), _lib.stream_ptr()))
        return out
