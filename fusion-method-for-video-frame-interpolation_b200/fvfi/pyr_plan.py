"""Plan + workspace cache for the steerable-pyramid kernels (C-ABI: fvfi_pyr_* in include/fvfi.h).

A plan owns the immutable device tables for one (H, W, height, nbands, scale_factor) on one
device; the scratch workspace is a torch uint8 tensor (so the caching allocator stays in charge,
SURVEY.md 8(b) "Ownership"), grown on demand and reused across calls.
"""
import ctypes
import math

import torch

from . import _lib


def level_sizes(H, W, height, scale_factor):
    """[(h_l, w_l) for the height-2 band levels] + [(h_low, w_low)]; rule n' = ceil((n-0.5)/s)."""
    L = _lib.lib()
    out = [(H, W)]
    for _ in range(height - 2):
        h, w = out[-1]
        out.append((L.fvfi_pyr_next_size(h, float(scale_factor)), L.fvfi_pyr_next_size(w, float(scale_factor))))
    return out


class PyrPlan:
    _cache = {}

    def __init__(self, H, W, height, nbands, scale_factor, device):
        self.H, self.W, self.height, self.nbands = int(H), int(W), int(height), int(nbands)
        self.scale_factor = float(scale_factor)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise NotImplementedError("the steerable pyramid runs on CUDA only (no CPU fallback)")
        L = _lib.lib()
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.fvfi_pyr_plan_create(self.H, self.W, self.height, self.nbands, self.scale_factor,
                                              ctypes.byref(handle)))
        self.handle = handle
        self.L = L.fvfi_pyr_num_levels(handle)
        self.shapes = []
        for l in range(self.L + 1):
            h, w = ctypes.c_int(), ctypes.c_int()
            _lib.check(L.fvfi_pyr_level_shape(handle, l, ctypes.byref(h), ctypes.byref(w)))
            self.shapes.append((h.value, w.value))
        self._ws = {}          # one scratch buffer per CUDA stream: calls on different streams / threads never share scratch

    @classmethod
    def get(cls, H, W, height, nbands, scale_factor, device):
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        key = (int(H), int(W), int(height), int(nbands), round(float(scale_factor), 12), str(device))
        p = cls._cache.get(key)
        if p is None:
            p = cls(H, W, height, nbands, scale_factor, device)
            cls._cache[key] = p
        return p

    def workspace(self, N):
        """Scratch for a call with N planes on the CURRENT stream (grown on demand, reused by later calls on that stream; calls
        on one stream are ordered, so forward / backward / filter / inv_filter can share it)."""
        need = _lib.lib().fvfi_pyr_workspace_bytes(self.handle, int(N))
        key = torch.cuda.current_stream(self.device).cuda_stream
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            self._ws.pop(key, None)
            ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().fvfi_pyr_plan_destroy(self.handle)
        except Exception:
            pass


def ptr_array(tensors):
    """Host array of device pointers (float* const*); None -> NULL."""
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def calc_pyr_height(img):
    """src/train/utils.py:168-171."""
    size = img.shape[1:]
    return int(math.ceil((math.log2(min(size)) - 3) * 2) + 2)
