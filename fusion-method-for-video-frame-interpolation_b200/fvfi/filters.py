"""GPU replacements of the two scipy.ndimage calls in the reference's uncertainty-map stage
(src/fusion_net/interpolate_twoframe.py:210-214, :221-222): ``gaussian_filter(h, 5)`` and
``median_filter(f, size=50)``; same boundary mode ('reflect'), truncation (4 sigma) and rank."""
import torch

from . import _lib


def gaussian_filter(maps, sigma):
    """maps [N,H,W] -> [N,H,W]."""
    x = maps.contiguous().float()
    N, H, W = x.shape
    out, tmp = torch.empty_like(x), torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().fvfi_gaussian_filter(x.data_ptr(), out.data_ptr(), tmp.data_ptr(), N, H, W, float(sigma),
                                                   _lib.stream_ptr()))
    return out


def median_filter(maps, size):
    """maps [N,H,W] -> [N,H,W]; exact rank filter (rank size*size//2)."""
    x = maps.contiguous().float()
    N, H, W = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().fvfi_median_filter(x.data_ptr(), out.data_ptr(), N, H, W, int(size), _lib.stream_ptr()))
    return out
