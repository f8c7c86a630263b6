"""CPU tests of the steerable-pyramid oracle (oracle/steerable_shim.py; parity UNPINNED for the
FFT/mask arithmetic -- the third-party package is absent): invariants of SURVEY 8(c)
plus the golden fixtures produced by the reference's own Pyramid wrapper on top of the shim."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import steerable_shim as ss

S2 = np.sqrt(2)


def test_level_size_table():
    # SURVEY.md Appendix A.7
    assert [h for h, _ in ss.level_sizes(256, 256, 12, S2)] == [256, 181, 128, 91, 64, 45, 32, 23, 16, 11, 8]
    sz = ss.level_sizes(1080, 1920, 17, S2)
    assert [h for h, _ in sz] == [1080, 764, 540, 382, 270, 191, 135, 96, 68, 48, 34, 24, 17, 12, 9, 7]
    assert [w for _, w in sz] == [1920, 1358, 960, 679, 480, 340, 241, 171, 121, 86, 61, 43, 31, 22, 16, 11]
    # s = 2 reduces to upstream's octave rule
    assert [h for h, _ in ss.level_sizes(128, 128, 5, 2)] == [128, 64, 32, 16]


@pytest.mark.parametrize("H,W,height", [(256, 256, 12), (90, 150, 8), (135, 241, 9), (48, 64, 6)])
def test_perfect_reconstruction_and_layout(H, W, height):
    pyr = ss.SCFpyr_PyTorch(height=height, nbands=4, scale_factor=S2)
    x = torch.rand(2, 1, H, W, generator=torch.Generator().manual_seed(0))
    c = pyr.build(x)
    assert len(c) == height and len(c[1]) == 4
    assert c[0].shape == (2, H, W) and c[0].dtype == torch.float32
    for l, (h, w) in enumerate(ss.level_sizes(H, W, height, S2)[:-1]):
        for b in c[1 + l]:
            assert b.shape == (2, h, w, 2)
            torch.view_as_complex(b)  # must be accepted (pyramid.py:58)
    assert c[-1].shape[1:] == ss.level_sizes(H, W, height, S2)[-1]
    r = pyr.reconstruct(c)
    # the LUT-interpolated masks are power complementary only to ~1e-5 (inherent to the algorithm)
    assert float((r - x.squeeze(1)).abs().max()) <= 3e-5


def test_skipped_levels_are_zero_contribution():
    pyr = ss.SCFpyr_PyTorch(height=7, nbands=4, scale_factor=S2)
    x = torch.rand(1, 1, 64, 80, generator=torch.Generator().manual_seed(1))
    c = pyr.build(x)
    c0 = list(c)
    c0[2] = [torch.zeros_like(b) for b in c[2]]
    c1 = list(c)
    c1[2] = 0  # the reference passes the int 0 (phase_net.py:91-93 -> values_to_coeff)
    assert float((pyr.reconstruct(c0) - pyr.reconstruct(c1)).abs().max()) < 1e-6


def test_linearity_and_analytic_bands():
    pyr = ss.SCFpyr_PyTorch(height=6, nbands=4, scale_factor=S2, precision="fp64")
    g = torch.Generator().manual_seed(2)
    x, y = torch.rand(1, 1, 40, 56, generator=g), torch.rand(1, 1, 40, 56, generator=g)
    cx, cy, cxy = pyr.build(x), pyr.build(y), pyr.build(x + 2 * y)
    for l in range(1, 5):
        for b in range(4):
            assert float((cxy[l][b] - (cx[l][b] + 2 * cy[l][b])).abs().max()) < 1e-4
    # the real part of the oriented bands + residuals rebuilds the image (analytic-signal property)
    z = torch.view_as_complex(cx[1][0])
    spec = torch.fft.fft2(z)
    # one-sided: at most half of the spectrum carries energy
    assert int((spec.abs() > 1e-3 * spec.abs().max()).sum()) <= spec.numel() // 2 + spec.shape[-1]


def test_golden_reference_wrapper(golden_dir):
    """Fixtures were produced by the REFERENCE's Pyramid.filter/inv_filter (real coeff_to_values /
    values_to_coeff, src/train/pyramid.py:48-112) over the shim; the oracle restatement of those two
    functions (torch.angle / abs, cos/sin*amp) must reproduce them."""
    files = sorted(glob.glob(os.path.join(golden_dir, "pyramid_ref_*.npz")))
    assert files
    for f in files:
        z = np.load(f)
        N, H, W, height, seed = [int(v) for v in z["meta"]]
        img = torch.rand((N, H, W), generator=torch.Generator().manual_seed(seed))
        pyr = ss.SCFpyr_PyTorch(height=height, nbands=4, scale_factor=S2)
        c = pyr.build(img.unsqueeze(1))
        assert np.abs(c[0].numpy() - z["high"][:, 0]).max() < 1e-6
        assert np.abs(c[-1].numpy() - z["low"][:, 0]).max() < 1e-5
        for l in range(height - 2):
            zc = torch.stack([torch.view_as_complex(b) for b in c[1 + l]], 1).reshape(N * 4, 1, *c[1 + l][0].shape[1:3])
            assert np.abs(zc.abs().numpy() - z["amp%d" % l]).max() < 1e-5
            ph = z["phase%d" % l]
            d = np.angle(np.exp(1j * (torch.angle(zc).numpy() - ph)))
            assert np.abs(d[z["amp%d" % l] > 1e-3]).max() < 1e-3


@pytest.mark.parametrize("H,W,height", [(256, 256, 12), (1080, 1920, 17), (2160, 3840, 19), (2048, 2048, 18), (64, 64, 8)])
def test_calc_pyr_height_is_buildable(H, W, height):
    """SURVEY 8(c) invariant (iv): the height the recipe derives from the frame size (src/train/utils.py:168-171) leaves a low-pass
    residual of at least 7 px per side at every size of BASELINE.json's configs, and the level table has `height - 1` entries."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "fusion-method-for-video-frame-interpolation_b200"))
    from fvfi import pyr_plan
    assert pyr_plan.calc_pyr_height(torch.empty((1, H, W))) == height
    sizes = ss.level_sizes(H, W, height, S2)
    assert len(sizes) == height - 1
    assert min(sizes[-1]) >= 7
    assert all(a[0] >= b[0] and a[1] >= b[1] for a, b in zip(sizes, sizes[1:]))


def test_power_complementary_masks_fp64():
    """SURVEY 8(c) invariant (ii): analysis followed by synthesis is the identity because the radial masks are power complementary
    (lo_l^2 + hi_l^2 = 1) and the four angular masks tile the half plane -- in fp64 the round trip is limited only by the 1024-step
    lookup table of the published algorithm (~1e-5), and it is the same for an image and for 3x that image (linearity)."""
    pyr = ss.SCFpyr_PyTorch(height=9, nbands=4, scale_factor=S2, precision="fp64")
    x = torch.rand(1, 1, 96, 160, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    r1 = pyr.reconstruct(pyr.build(x))
    r3 = pyr.reconstruct(pyr.build(3 * x))
    e1 = float((r1 - x.squeeze(1)).abs().max())
    assert e1 <= 3e-5
    assert float((r3 - 3 * r1).abs().max()) <= 1e-12
    # energy bookkeeping of one level: removing a band level removes exactly that level's contribution
    c = pyr.build(x)
    c_wo = list(c)
    c_wo[3] = 0
    only = [torch.zeros_like(c[0])] + [0] * (len(c) - 2) + [torch.zeros_like(c[-1])]
    only[1] = [torch.zeros_like(b) for b in c[1]]       # reconstruct() reads the band count from level 1
    only[3] = c[3]
    assert float((pyr.reconstruct(c_wo) + pyr.reconstruct(only) - r1).abs().max()) <= 1e-12
