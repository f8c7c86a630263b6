"""GPU parity of the tcgen05 3xTF32 convolution against torch's fp32 convolution (TF32 off)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # B, Cin, Cout, K, H, W, mode, act
    (1, 64, 64, 3, 32, 40, "zeros", "relu"),
    (2, 6, 32, 3, 48, 72, "zeros", "relu"),
    (1, 16, 16, 1, 16, 8, "zeros", None),
    (1, 88, 64, 3, 23, 31, "reflect", "elu"),
    (1, 18, 32, 5, 40, 56, "reflect", "relu"),
    (1, 64, 8, 1, 45, 45, "zeros", "tanh"),
    (1, 64, 25, 3, 64, 96, "zeros", None),
    (1, 64, 1, 3, 33, 47, "zeros", "sigmoid"),
    (1, 128, 256, 3, 20, 28, "zeros", "relu"),
    (1, 512, 512, 3, 17, 30, "zeros", "relu"),
    (2, 32, 64, 5, 70, 130, "reflect", "relu"),
]


@pytest.mark.parametrize("B,Cin,Cout,K,H,W,mode,act", CASES)
def test_conv_matches_fp32(B, Cin, Cout, K, H, W, mode, act):
    from fvfi import conv
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((B, Cin, H, W), device="cuda", generator=g)
    w = torch.randn((Cout, Cin, K, K), device="cuda", generator=g) / (Cin * K * K) ** 0.5
    b = torch.randn((Cout,), device="cuda", generator=g)
    y = conv.conv2d(x, w, b, mode, act)
    torch.cuda.synchronize()
    p = K // 2
    xr = F.pad(x.double(), (p, p, p, p), mode="reflect") if (mode == "reflect" and p) else x.double()
    ref = F.conv2d(xr, w.double(), b.double(), padding=0 if (mode == "reflect" and p) else p)
    ref = {None: lambda t: t, "relu": F.relu, "elu": F.elu, "tanh": torch.tanh, "sigmoid": torch.sigmoid}[act](ref)
    err = float((y.double() - ref).abs().max())
    # fp32 cuDNN itself is ~1e-6 from the fp64 result here; 3xTF32 must be in the same class (plain TF32: ~1e-3)
    assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
    print('max abs err %.2e (max |ref| %.2f)' % (err, float(ref.abs().max())))
    assert err <= 2e-5 * max(1.0, float(ref.abs().max())), err
