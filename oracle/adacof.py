"""oracle/adacof.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

numpy/ctypes front end of oracle/adacof_oracle.c, the CPU restatement of the
reference's AdaCoF kernels (src/adacof/cupy_module/adacof.py:6-258) and of the
AdaCoFNet blend/uncertainty tail (src/fusion_net/fusion_adacofnet.py:198-213).
Host threads split the output rows; ctypes releases the GIL during each call.
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile liboracle.so with gcc (recipe: oracle/Makefile)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _prep(arrays):
    """Contiguous operands in ONE dtype: float64 if the first operand is float64 (the fp64 arbiter instantiation
    oracle_*_f64 of the same C source), else float32 (the reference's arithmetic).  -> (arrays, dtype, symbol suffix)"""
    dt = np.float64 if np.asarray(arrays[0]).dtype == np.float64 else np.float32
    return [np.ascontiguousarray(np.asarray(a, dtype=dt)) for a in arrays], dt, ("_f64" if dt == np.float64 else "")


def _slabs(H, threads):
    threads = max(1, min(int(threads), H))
    edges = np.linspace(0, H, threads + 1).astype(int)
    return [(int(edges[t]), int(edges[t + 1])) for t in range(threads) if edges[t + 1] > edges[t]]


def _run(fn, H, threads):
    slabs = _slabs(H, threads)
    if len(slabs) == 1:
        fn(*slabs[0])
        return
    with ThreadPoolExecutor(len(slabs)) as ex:
        list(ex.map(lambda s: fn(*s), slabs))


def _dims(inp, weight, dilation):
    B, C, Hin, Win = inp.shape
    _, FF, H, W = weight.shape
    F = int(np.sqrt(FF))
    assert Hin - ((F - 1) * dilation + 1) == H - 1 and Win - ((F - 1) * dilation + 1) == W - 1
    return B, C, Hin, Win, H, W, F


def forward(inp, weight, off_i, off_j, dilation=1, threads=1):
    (inp, weight, off_i, off_j), dt, suf = _prep((inp, weight, off_i, off_j))
    B, C, Hin, Win, H, W, F = _dims(inp, weight, dilation)
    out = np.empty((B, C, H, W), dt)
    L = _lib()
    fwd = getattr(L, "oracle_adacof_forward" + suf)
    _run(lambda i0, i1: fwd(_p(inp), _p(weight), _p(off_i), _p(off_j), _p(out),
                                                B, C, Hin, Win, H, W, F, dilation, i0, i1), H, threads)
    return out


def backward(gout, inp, weight, off_i, off_j, dilation=1, threads=1):
    """Returns (gradWeight, gradOffset_i, gradOffset_j); gradInput is zeros in the reference."""
    gout, inp, weight, off_i, off_j = map(_f32, (gout, inp, weight, off_i, off_j))
    B, C, Hin, Win, H, W, F = _dims(inp, weight, dilation)
    gw, gi, gj = (np.empty_like(weight) for _ in range(3))
    L = _lib()
    _run(lambda i0, i1: L.oracle_adacof_backward(_p(gout), _p(inp), _p(weight), _p(off_i), _p(off_j),
                                                 _p(gw), _p(gi), _p(gj), B, C, Hin, Win, H, W, F,
                                                 dilation, i0, i1), H, threads)
    return gw, gi, gj


def grad_input(gout, inp_shape, weight, off_i, off_j, dilation=1):
    gout, weight, off_i, off_j = map(_f32, (gout, weight, off_i, off_j))
    B, C, Hin, Win = inp_shape
    _, FF, H, W = weight.shape
    F = int(np.sqrt(FF))
    gin = np.empty((B, C, Hin, Win), np.float32)
    _lib().oracle_adacof_grad_input(_p(gout), _p(weight), _p(off_i), _p(off_j), _p(gin),
                                    B, C, Hin, Win, H, W, F, dilation)
    return gin


def adacofnet_tail(t1, t2, occ, w1, a1, b1, w2, a2, b2, threads=1):
    (t1, t2, occ, w1, a1, b1, w2, a2, b2), dt, suf = _prep((t1, t2, occ, w1, a1, b1, w2, a2, b2))
    B, C, H, W = t1.shape
    FF = w1.shape[1]
    frame = np.empty_like(t1)
    mask = np.empty((B, 1, H, W), dt)
    L = _lib()
    tail = getattr(L, "oracle_adacofnet_tail" + suf)
    _run(lambda i0, i1: tail(_p(t1), _p(t2), _p(occ), _p(w1), _p(a1), _p(b1), _p(w2),
                                                _p(a2), _p(b2), _p(frame), _p(mask), B, C, H, W, FF,
                                                i0, i1), H, threads)
    return frame, mask


def synth(B, C, H, W, F=5, dilation=1, seed=0):
    """SURVEY.md 8(d) synthetic operands: negative fractional offsets and
    out-of-range taps included on purpose."""
    rng = np.random.default_rng(seed)
    pad = (F - 1) * dilation
    inp = rng.random((B, C, H + pad, W + pad), dtype=np.float32)
    logits = rng.standard_normal((B, F * F, H, W), dtype=np.float32)
    e = np.exp(logits - logits.max(1, keepdims=True))
    weight = (e / e.sum(1, keepdims=True)).astype(np.float32)
    off_i = np.clip(3.0 * rng.standard_normal((B, F * F, H, W), dtype=np.float32), -16, 16).astype(np.float32)
    off_j = np.clip(3.0 * rng.standard_normal((B, F * F, H, W), dtype=np.float32), -16, 16).astype(np.float32)
    gout = rng.standard_normal((B, C, H, W), dtype=np.float32)
    return inp, weight, off_i, off_j, gout
