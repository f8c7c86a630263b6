"""oracle/lab.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED (skimage absent).

CPU restatement of ``skimage.color.rgb2lab`` / ``lab2rgb`` (the reference calls them in
src/train/transform.py:6-49; scikit-image is not installed here and not vendored).  Published
algorithm: sRGB companding (IEC 61966-2-1), linear RGB -> XYZ (D65, 2 degree observer), CIE 1976 L*a*b*.
Coarse independent anchor: OpenCV's float RGB<->Lab (same standard, table-interpolated) agrees to <= 0.5 Lab units / 2e-3 RGB
(tests/test_models_oracle.py::test_lab_oracle_agrees_with_opencv) -- it catches a wrong constant, it does not pin the last digits.
"""
import numpy as np

_M = np.array([[0.412453, 0.357580, 0.180423],
               [0.212671, 0.715160, 0.072169],
               [0.019334, 0.119193, 0.950227]])
_MINV = np.linalg.inv(_M)
_WHITE = np.array([0.95047, 1.0, 1.08883])  # D65, observer "2"


def rgb2lab(rgb):
    rgb = np.asarray(rgb, dtype=np.float64)
    lin = np.where(rgb > 0.04045, ((rgb + 0.055) / 1.055) ** 2.4, rgb / 12.92)
    xyz = lin @ _M.T
    xyz = xyz / _WHITE
    f = np.where(xyz > 0.008856, np.cbrt(xyz), 7.787 * xyz + 16.0 / 116.0)
    L = 116.0 * f[..., 1] - 16.0
    a = 500.0 * (f[..., 0] - f[..., 1])
    b = 200.0 * (f[..., 1] - f[..., 2])
    return np.stack([L, a, b], -1)


def lab2rgb(lab):
    lab = np.asarray(lab, dtype=np.float64)
    L, a, b = lab[..., 0], lab[..., 1], lab[..., 2]
    y = (L + 16.0) / 116.0
    x = a / 500.0 + y
    z = y - b / 200.0
    z = np.maximum(z, 0)  # skimage clips negative z
    f = np.stack([x, y, z], -1)
    xyz = np.where(f > 0.2068966, f ** 3, (f - 16.0 / 116.0) / 7.787)
    xyz = xyz * _WHITE
    lin = xyz @ _MINV.T
    rgb = np.where(lin > 0.0031308, 1.055 * np.power(np.maximum(lin, 0), 1 / 2.4) - 0.055, lin * 12.92)
    return np.clip(rgb, 0, 1)
