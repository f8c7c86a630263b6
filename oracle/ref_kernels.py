"""oracle/ref_kernels.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Loads the reference's own AdaCoF CUDA kernels (cubins under oracle/_ref/, built
by oracle/build_ref_kernels.py from /root/reference/src/adacof/cupy_module/adacof.py:6-258)
with cuda-python and launches them exactly like FunctionAdaCoF does
(adacof.py:334-354 forward, :382-438 backward): outputs pre-zeroed with
new_zeros, grid ceil(n/512) x block 512, on torch's current stream.
This is the GPU oracle ("kind": "reference") and the same-box bar for the warp.
"""
import json
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")
_cache = {}


def tag(B, C, H, W, F, d):
    return "B%d_C%d_H%d_W%d_F%d_D%d" % (B, C, H, W, F, d)


def have(B, C, H, W, F, d):
    mf = os.path.join(REFDIR, "manifest.json")
    if not os.path.exists(mf):
        return False
    files = json.load(open(mf))["files"]
    return ("kernel_AdaCoF_updateOutput/" + tag(B, C, H, W, F, d)) in files


def _fn(name, t):
    from cuda.bindings import driver
    key = (name, t)
    if key not in _cache:
        torch.cuda.init()
        torch.zeros(1, device="cuda")  # make sure the primary context is current
        data = open(os.path.join(REFDIR, "%s_%s.cubin" % (name, t)), "rb").read()
        err, mod = driver.cuModuleLoadData(data)
        assert err == driver.CUresult.CUDA_SUCCESS, err
        err, fn = driver.cuModuleGetFunction(mod, name.encode())
        assert err == driver.CUresult.CUDA_SUCCESS, err
        _cache[key] = (mod, fn)
    return _cache[key][1]


def _launch(fn, n, ptrs):
    from cuda.bindings import driver
    stream = torch.cuda.current_stream().cuda_stream
    args = [np.array([n], dtype=np.int32)] + [np.array([p], dtype=np.uint64) for p in ptrs]
    argv = np.array([a.ctypes.data for a in args], dtype=np.uint64)
    (err,) = driver.cuLaunchKernel(fn, (n + 511) // 512, 1, 1, 512, 1, 1, 0, stream, argv.ctypes.data, 0)
    assert err == driver.CUresult.CUDA_SUCCESS, err


def forward(inp, weight, off_i, off_j, dilation):
    B, C, Hin, Win = inp.shape
    _, FF, H, W = weight.shape
    F = int(FF ** 0.5)
    t = tag(B, C, H, W, F, dilation)
    out = inp.new_zeros(B, C, H, W)                                   # adacof.py:334
    _launch(_fn("kernel_AdaCoF_updateOutput", t), out.numel(),
            [x.data_ptr() for x in (inp, weight, off_i, off_j, out)])
    return out


def backward(gout, inp, weight, off_i, off_j, dilation):
    B, C, Hin, Win = inp.shape
    _, FF, H, W = weight.shape
    F = int(FF ** 0.5)
    t = tag(B, C, H, W, F, dilation)
    gin = inp.new_zeros(inp.shape)                                    # adacof.py:382
    gw, gi, gj = (inp.new_zeros(weight.shape) for _ in range(3))     # :383-385
    n = gw.numel()
    _launch(_fn("kernel_AdaCoF_updateGradWeight", t), n,
            [x.data_ptr() for x in (gout, inp, off_i, off_j, gw)])
    _launch(_fn("kernel_AdaCoF_updateGradAlpha", t), n,
            [x.data_ptr() for x in (gout, inp, weight, off_i, off_j, gi)])
    _launch(_fn("kernel_AdaCoF_updateGradBeta", t), n,
            [x.data_ptr() for x in (gout, inp, weight, off_i, off_j, gj)])
    return gin, gw, gi, gj
