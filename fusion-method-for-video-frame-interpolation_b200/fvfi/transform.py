"""Drop-in for the reference's ``src/train/transform.py`` (rgb2lab, rgb2lab_single, lab2rgb,
lab2rgb_single) -- device-resident: the reference round-trips through skimage on the host CPU
(transform.py:8,19,35,46); here one CUDA kernel per direction, no host copy, same L/100 and
(a,b+128)/255 scaling."""
import torch

from . import _lib


def _run(fn_name, img):
    if not img.is_cuda:
        raise NotImplementedError("fvfi colour transforms run on CUDA tensors only")
    x = img.contiguous().float()
    B, C, H, W = x.shape
    assert C == 3
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(getattr(_lib.lib(), fn_name)(x.data_ptr(), out.data_ptr(), B, H, W, _lib.stream_ptr()))
    return out


def rgb2lab(img, light=100, ab_mul=255, ab_max=128):
    """[B,3,H,W] RGB in [0,1] -> scaled Lab (transform.py:6-14)."""
    assert (light, ab_mul, ab_max) == (100, 255, 128)
    return _run("fvfi_rgb2lab", img)


def rgb2lab_single(img, light=100, ab_mul=255, ab_max=128):
    """[3,H,W] (transform.py:17-25)."""
    return rgb2lab(img.unsqueeze(0), light, ab_mul, ab_max)[0]


def lab2rgb(img, light=100, ab_mul=255, ab_max=128):
    """[B,3,H,W] scaled Lab -> RGB in [0,1] (transform.py:28-37)."""
    assert (light, ab_mul, ab_max) == (100, 255, 128)
    return _run("fvfi_lab2rgb", img)


def lab2rgb_single(img, light=100, ab_mul=255, ab_max=128):
    """[3,H,W] (transform.py:40-49)."""
    return lab2rgb(img.unsqueeze(0), light, ab_mul, ab_max)[0]
