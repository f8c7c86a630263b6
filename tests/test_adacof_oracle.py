"""CPU tests: the AdaCoF oracle against (a) the golden outputs of the reference's own CUDA
kernels (tests/golden/adacof_ref_*.npz, generated on a B200 by tests/golden/make_adacof_golden.py)
and (b) an independent numpy restatement / finite differences."""
import glob
import os

import numpy as np
import pytest

from oracle import adacof as oa

# fp32, [0,1] frames, softmax weights: reference kernels use FMA contraction (NVRTC default),
# the oracle does not -> a few ulp per term.
TOL = 2e-6


def _golden_files(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "adacof_ref_*.npz")))


def test_golden_present(golden_dir):
    assert len(_golden_files(golden_dir)) >= 3, "golden vectors of the reference kernels are missing"


def test_oracle_matches_reference_kernels(golden_dir):
    for f in _golden_files(golden_dir):
        z = np.load(f)
        B, C, H, W, F, d, seed = [int(v) for v in z["shape"]]
        inp, w, oi, oj, g = oa.synth(B, C, H, W, F, d, seed)
        out = oa.forward(inp, w, oi, oj, d, threads=2)
        assert np.abs(out - z["out"]).max() <= TOL, f
        gw, gi, gj = oa.backward(g, inp, w, oi, oj, d, threads=2)
        # gradients scale with |gout| ~ N(0,1) x 3 channels
        assert np.abs(gw - z["gw"]).max() <= 2e-5, f
        assert np.abs(gi - z["gi"]).max() <= 2e-5, f
        assert np.abs(gj - z["gj"]).max() <= 2e-5, f


def _np_forward(inp, w, oi, oj, d):
    B, C, Hin, Win = inp.shape
    _, FF, H, W = w.shape
    F = int(np.sqrt(FF))
    out = np.zeros((B, C, H, W), np.float64)
    ii, jj = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    for k in range(F):
        for l in range(F):
            al, be = oi[:, k * F + l].astype(np.float64), oj[:, k * F + l].astype(np.float64)
            A, Bq = np.trunc(al).astype(int), np.trunc(be).astype(int)
            a, b = al - A, be - Bq
            r0, r1 = np.clip(ii + k * d + A, 0, Hin - 1), np.clip(ii + k * d + A + 1, 0, Hin - 1)
            c0, c1 = np.clip(jj + l * d + Bq, 0, Win - 1), np.clip(jj + l * d + Bq + 1, 0, Win - 1)
            for n in range(B):
                for c in range(C):
                    I = inp[n, c].astype(np.float64)
                    out[n, c] += w[n, k * F + l] * (I[r0[n], c0[n]] * (1 - a[n]) * (1 - b[n]) + I[r1[n], c0[n]] * a[n] * (1 - b[n]) +
                                                    I[r0[n], c1[n]] * (1 - a[n]) * b[n] + I[r1[n], c1[n]] * a[n] * b[n])
    return out


@pytest.mark.parametrize("shape", [(2, 3, 17, 23, 5, 1), (1, 3, 9, 31, 3, 2), (1, 2, 8, 8, 1, 1)])
def test_forward_vs_numpy(shape):
    B, C, H, W, F, d = shape
    inp, w, oi, oj, _ = oa.synth(B, C, H, W, F, d, seed=3)
    assert (oi < 0).any() and (np.abs(oi) > 6).any()  # negative fractional + out-of-range taps are covered
    assert np.abs(oa.forward(inp, w, oi, oj, d) - _np_forward(inp, w, oi, oj, d)).max() < 2e-6


def test_trunc_not_floor():
    """(int) truncates toward zero, so alpha=-0.5 samples rows i, i+1 with weights 1.5/-0.5."""
    inp = np.zeros((1, 1, 4, 4), np.float32)
    inp[0, 0] = np.arange(16).reshape(4, 4)
    w = np.ones((1, 1, 4, 4), np.float32)
    oi = np.full((1, 1, 4, 4), -0.5, np.float32)
    oj = np.zeros((1, 1, 4, 4), np.float32)
    out = oa.forward(inp, w, oi, oj, 1)
    # row 1: 1.5*I[1] - 0.5*I[2]  (floor would give 0.5*I[0] + 0.5*I[1])
    assert np.allclose(out[0, 0, 1], 1.5 * inp[0, 0, 1] - 0.5 * inp[0, 0, 2])


def test_backward_finite_differences():
    B, C, H, W, F, d = 1, 3, 10, 12, 3, 1
    inp, w, oi, oj, g = oa.synth(B, C, H, W, F, d, seed=5)
    # keep offsets away from integer boundaries where the trunc makes the op non-differentiable
    frac = np.abs(oi - np.trunc(oi))
    oi = np.where((frac < 0.05) | (frac > 0.95), oi + 0.3, oi).astype(np.float32)
    frac = np.abs(oj - np.trunc(oj))
    oj = np.where((frac < 0.05) | (frac > 0.95), oj + 0.3, oj).astype(np.float32)
    gw, gi, gj = oa.backward(g, inp, w, oi, oj, d)
    base = oa.forward(inp, w, oi, oj, d).astype(np.float64)
    rng = np.random.default_rng(0)
    for _ in range(20):
        idx = tuple(rng.integers(0, s) for s in w.shape)
        eps = 1e-2
        for arr, grad in ((w, gw), (oi, gi), (oj, gj)):
            pert = arr.copy()
            pert[idx] += eps
            args = [pert if x is arr else x for x in (w, oi, oj)]
            fd = ((oa.forward(inp, *args, d).astype(np.float64) - base) * g).sum() / eps
            assert abs(fd - grad[idx]) < 5e-3 * max(1.0, abs(fd)), (idx, fd, grad[idx])


def test_grad_input_is_adjoint():
    B, C, H, W, F, d = 1, 3, 9, 11, 3, 1
    inp, w, oi, oj, g = oa.synth(B, C, H, W, F, d, seed=7)
    gin = oa.grad_input(g, inp.shape, w, oi, oj, d)
    # forward is linear in input: <forward(x), g> == <x, gin>
    lhs = (oa.forward(inp, w, oi, oj, d).astype(np.float64) * g).sum()
    rhs = (inp.astype(np.float64) * gin).sum()
    assert abs(lhs - rhs) < 1e-3 * abs(lhs)


def test_tail_matches_torch_restatement():
    import torch
    B, C, H, W, FF = 2, 3, 6, 7, 25
    rng = np.random.default_rng(1)
    t1, t2 = rng.random((B, C, H, W), np.float32), rng.random((B, C, H, W), np.float32)
    occ = rng.random((B, 1, H, W), np.float32)
    maps = []
    for _ in range(2):
        lg = rng.standard_normal((B, FF, H, W), np.float32)
        wt = np.exp(lg) / np.exp(lg).sum(1, keepdims=True)
        maps += [wt.astype(np.float32), (2 * rng.standard_normal((B, FF, H, W))).astype(np.float32),
                 (2 * rng.standard_normal((B, FF, H, W))).astype(np.float32)]
    frame, mask = oa.adacofnet_tail(t1, t2, occ, *maps)
    # torch restatement with the reference's own tensor expressions (fusion_adacofnet.py:198-213)
    T = [torch.from_numpy(x) for x in maps]
    W1, A1, B1, W2, A2, B2 = T
    D1, D2 = torch.stack([A1, B1], 0), torch.stack([A2, B2], 0)
    M1, M2 = (W1 * D1).sum(-3), (W2 * D2).sum(-3)
    V1 = (W1 * ((M1 - D1.permute(2, 0, 1, 3, 4)) ** 2).permute(1, 2, 0, 3, 4)).sum(-3)
    V2 = (W2 * ((M2 - D2.permute(2, 0, 1, 3, 4)) ** 2).permute(1, 2, 0, 3, 4)).sum(-3)
    U = (torch.clip(torch.max(V1.sum(0), V2.sum(0)), 0, 20) / 20).unsqueeze(1)
    fr = torch.from_numpy(occ) * torch.from_numpy(t1) + (1 - torch.from_numpy(occ)) * torch.from_numpy(t2)
    assert np.abs(frame - fr.numpy()).max() < 1e-6
    assert np.abs(mask - U.numpy()).max() < 1e-5
