"""GPU parity of the tcgen05 split-operand convolution (3xFP16 default, 3xTF32) against an fp64 convolution."""
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
    (2, 32, 3, 1, 40, 56, "zeros", None),
    (1, 24, 5, 3, 31, 45, "reflect", "elu"),
    (1, 6, 2, 3, 20, 20, "zeros", "relu"),
]


@pytest.fixture(params=["f16x3", "tf32x3"])
def prec(request, monkeypatch):
    from fvfi import conv
    monkeypatch.setattr(conv, "precision", conv.PRECISIONS[request.param])
    return request.param


@pytest.mark.parametrize("B,Cin,Cout,K,H,W,mode,act", CASES)
def test_conv_matches_fp32(B, Cin, Cout, K, H, W, mode, act, prec):
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
    conv.check_overflow()


def test_conv_f16x3_range_and_small_values(monkeypatch):
    """3xFP16 keeps fp32-grade accuracy for small activations / weights (power-of-two scaling keeps both halves in
    fp16's normal range) and reports -- instead of hiding -- activations beyond its range."""
    from fvfi import conv
    monkeypatch.setattr(conv, "precision", conv.PRECISIONS["f16x3"])
    g = torch.Generator(device="cuda").manual_seed(3)
    for xs, ws in ((1e-3, 1.0), (1.0, 1e-4), (100.0, 30.0), (3e-4, 2e-3)):
        x = xs * torch.randn((1, 48, 24, 40), device="cuda", generator=g)
        w = ws * torch.randn((40, 48, 3, 3), device="cuda", generator=g) / 20
        y = conv.conv2d(x, w, None, "zeros", None)
        ref = F.conv2d(x.double(), w.double(), padding=1)
        assert float((y.double() - ref).abs().max()) <= 3e-6 * float(ref.abs().max()) + 1e-9, (xs, ws)
    conv.check_overflow()
    x = torch.full((1, 16, 16, 16), 5000.0, device="cuda")
    conv.conv2d(x, torch.ones((16, 16, 1, 1), device="cuda"), None, "zeros", None)
    with pytest.raises(FloatingPointError):
        conv.check_overflow()
    conv.check_overflow()   # the flag is cleared by the read


@pytest.mark.parametrize("align", [True, False])
def test_resize_bilinear_nhwc_matches_torch(align):
    from fvfi import conv
    g = torch.Generator(device="cuda").manual_seed(1)
    for (B, C, Hi, Wi, Ho, Wo) in [(2, 64, 17, 30, 34, 60), (1, 8, 23, 23, 32, 32), (1, 1, 8, 11, 12, 16), (2, 25, 20, 28, 40, 56)]:
        x = torch.randn((B, C, Hi, Wi), device="cuda", generator=g)
        y = conv.resize_bilinear(x, (Ho, Wo), align)
        ref = F.interpolate(x, size=(Ho, Wo), mode="bilinear", align_corners=align)
        assert float((y - ref).abs().max()) <= 2e-6
    # channel-slice destination (PhaseNet concat assembly)
    x = torch.randn((1, 64, 11, 11), device="cuda", generator=g)
    buf = torch.zeros((1, 88, 16, 16), device="cuda").contiguous(memory_format=torch.channels_last)
    conv.resize_bilinear(x, (16, 16), False, out=buf, out_channel_offset=8)
    assert float((buf[:, 8:72] - F.interpolate(x, size=(16, 16), mode="bilinear", align_corners=False)).abs().max()) <= 2e-6
    assert float(buf[:, :8].abs().max()) == 0.0 and float(buf[:, 72:].abs().max()) == 0.0
    # fused decoder step of FusionNet: Upsample(ReLU(x)) + skip (fusion_net.py:60-62), every vector width / channel-group count
    for (B, C, Hi, Wi) in [(2, 64, 9, 13), (1, 88, 7, 5), (1, 12, 6, 10), (2, 3, 5, 4), (1, 32, 33, 17)]:
        x = torch.randn((B, C, Hi, Wi), device="cuda", generator=g)
        sk = torch.randn((B, C, 2 * Hi, 2 * Wi), device="cuda", generator=g)
        ref = F.interpolate(torch.relu(x), scale_factor=2, mode="bilinear", align_corners=align) + sk
        got = conv.resize_bilinear(x, (2 * Hi, 2 * Wi), align, relu_input=True, add=sk)
        assert float((got - ref).abs().max()) <= 2e-6, (B, C, Hi, Wi)
        got = conv.resize_bilinear(x, (2 * Hi, 2 * Wi), align, add=sk.contiguous(memory_format=torch.channels_last))
        assert float((got - (F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=align) + sk)).abs().max()) <= 2e-6


def test_conv_softmax_and_nchw_epilogues(prec):
    from fvfi import conv
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn((2, 25, 40, 72), device="cuda", generator=g)
    w = torch.randn((25, 25, 3, 3), device="cuda", generator=g) / 15
    b = torch.randn((25,), device="cuda", generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    y = conv.conv2d(x, w, b, "zeros", "softmax", nchw_out=True)
    assert y.is_contiguous() and float((y.double() - torch.softmax(ref, 1)).abs().max()) <= 5e-6
    y = conv.conv2d(x, w, b, "zeros", None, nchw_out=True)
    assert y.is_contiguous() and float((y.double() - ref).abs().max()) <= 2e-5
    w1 = torch.randn((1, 25, 3, 3), device="cuda", generator=g) / 15
    y = conv.conv2d(x, w1, b[:1], "zeros", "sigmoid", nchw_out=True)
    assert float((y.double() - torch.sigmoid(F.conv2d(x.double(), w1.double(), b[:1].double(), padding=1))).abs().max()) <= 2e-6


def test_conv_padded_channel_chain(prec):
    """64 -> 25 conv with zero-padded NHWC output (32 channels), bilinear x2, 25 -> 25 conv reading the padded tensor
    with 16-byte loads: the KernelEstimation head tail (fusion_adacofnet.py:36-59)."""
    from fvfi import conv
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn((2, 64, 20, 36), device="cuda", generator=g)
    w1 = torch.randn((25, 64, 3, 3), device="cuda", generator=g) / 24
    b1 = torch.randn((25,), device="cuda", generator=g)
    w2 = torch.randn((25, 25, 3, 3), device="cuda", generator=g) / 15
    b2 = torch.randn((25,), device="cuda", generator=g)
    y1 = conv.conv2d(x, w1, b1, "zeros", "relu", pad_out=True)
    assert y1.shape == (2, 32, 20, 36) and float(y1[:, 25:].abs().max()) == 0.0
    up = conv.resize_bilinear(y1, (40, 72), True)
    assert float(up[:, 25:].abs().max()) == 0.0
    y2 = conv.conv2d(up, w2, b2, "zeros", "softmax", nchw_out=True)
    r1 = F.relu(F.conv2d(x.double(), w1.double(), b1.double(), padding=1))
    r2 = F.conv2d(F.interpolate(r1, size=(40, 72), mode="bilinear", align_corners=True), w2.double(), b2.double(), padding=1)
    assert y2.shape == (2, 25, 40, 72) and float((y2.double() - torch.softmax(r2, 1)).abs().max()) <= 5e-6


def test_conv_full_size_layer_vs_cudnn_fp32():
    """One KernelEstimation layer at its real size (64 -> 64, 3x3, 544x960, batch 2) against cuDNN's fp32 path
    (TF32 off): the two fp32-grade implementations agree to 2e-5 of the output range."""
    from fvfi import conv
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randn((2, 64, 544, 960), device="cuda", generator=g).relu_()
    w = torch.randn((64, 64, 3, 3), device="cuda", generator=g) / 24
    b = torch.randn((64,), device="cuda", generator=g)
    y = conv.conv2d(x, w, b, "zeros", "relu")
    ref = F.relu(F.conv2d(x, w, b, padding=1))
    assert float((y - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    conv.check_overflow()


def test_avg_pool2_nhwc_matches_torch():
    from fvfi import conv
    g = torch.Generator(device="cuda").manual_seed(8)
    for (B, C, H, W) in [(2, 32, 32, 64), (1, 64, 34, 60), (1, 5, 9, 7), (2, 512, 68, 120)]:
        x = torch.randn((B, C, H, W), device="cuda", generator=g)
        y = conv.avg_pool2(x)
        ref = F.avg_pool2d(x, 2, 2)
        assert y.shape == ref.shape and float((y - ref).abs().max()) <= 1e-6
        ym = conv.max_pool2(x)                                      # nn.MaxPool2d(2, stride=2) (fusion_net.py:39): exact
        assert torch.equal(ym, F.max_pool2d(x, 2, 2))


def test_put_planar_slice():
    from fvfi import conv
    g = torch.Generator(device="cuda").manual_seed(9)
    for C, off in ((8, 64), (8, 72), (5, 3)):
        x = torch.randn((3, C, 23, 31), device="cuda", generator=g)
        buf = torch.zeros((3, 88, 23, 31), device="cuda").contiguous(memory_format=torch.channels_last)
        conv.put_planar(x, buf, off)
        assert torch.equal(buf[:, off:off + C], x) and float(buf[:, :off].abs().max()) == 0 and float(buf[:, off + C:].abs().max()) == 0


@pytest.mark.parametrize("Cin,Cout,act,shape", [(64, 8, "tanh", (3, 37, 53)), (128, 8, None, (1, 9, 301)), (32, 3, None, (2, 40, 56)),
                                                (8, 1, "sigmoid", (1, 5, 7)), (24, 5, "elu", (2, 31, 17))])
def test_conv1x1_direct_kernel(Cin, Cout, act, shape):
    """Cout <= 8 1x1 convolutions take the direct fp32 kernel (fvfi_conv1x1_nhwc): PhaseNet prediction 64 -> 8 + tanh
    (phase_net.py:197-200), FusionNet 32 -> 3 (fusion_net.py:36); also on a channel slice of a wider NHWC tensor."""
    from fvfi import conv
    B, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(1)
    wide = torch.randn((B, Cin + 8, H, W), device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    w = torch.randn((Cout, Cin, 1, 1), device="cuda", generator=g) / Cin ** 0.5
    b = torch.randn((Cout,), device="cuda", generator=g)
    fn = {None: lambda t: t, "elu": F.elu, "tanh": torch.tanh, "sigmoid": torch.sigmoid}[act]
    for x in (wide[:, :Cin].contiguous(memory_format=torch.channels_last), wide[:, :Cin]):
        n0 = conv._lib.lib().fvfi_launch_count()
        y = conv.conv2d(x, w, b, "zeros", act)
        assert conv._lib.lib().fvfi_launch_count() == n0 + 1          # one direct kernel, no tensor-core launch
        ref = fn(F.conv2d(x.double(), w.double(), b.double()))
        assert y.shape == ref.shape
        assert float((y.double() - ref).abs().max()) <= 2e-6 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 17, 23), (1, 64, 34, 60), (1, 16, 1, 9)])
def test_upsample2_conv_single_channel(B, C, H, W, prec):
    """Occlusion-head tail (fusion_adacofnet.py:50-59,103-104): Upsample(x2, align_corners=True) -> Conv2d(C, 1, 3) -> Sigmoid with
    the channels contracted at half resolution (fvfi_upsample2_tapsum) == the reference's order of operations."""
    from fvfi import conv
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn((B, C, H, W), device="cuda", generator=g)
    torch.manual_seed(2)                                   # the module's default init draws from the global generator
    m = torch.nn.Conv2d(C, 1, 3, 1, 1).cuda()
    with torch.no_grad():
        y = conv.upsample2_conv3x3_single(m, x, "sigmoid")
        up = F.interpolate(x.double(), scale_factor=2, mode="bilinear", align_corners=True)
        ref = torch.sigmoid(F.conv2d(up, m.weight.double(), m.bias.double(), padding=1))
    assert y.shape == ref.shape and y.is_contiguous()
    # pre-activation error <= 2e-5 * max|z| (the convolution's own bound, both operand splits); sigmoid' <= 1/4
    assert float((y.double() - ref).abs().max()) <= 8e-6
    conv.check_overflow()


@pytest.mark.parametrize("B,Cin,Cout,H,W,act", [(3, 25, 25, 50, 100, None), (2, 32, 25, 37, 52, "softmax"), (1, 64, 25, 19, 44, "relu"),
                                                 (2, 25, 25, 33, 51, None), (1, 16, 40, 70, 36, None), (5, 8, 3, 64, 260, "sigmoid")])
def test_conv_planar_output_tiles(B, Cin, Cout, H, W, act, prec):
    """Planar [B,Cout,H,W] outputs (the coefficient maps the AdaCoF warp streams) at ragged sizes: several tiles per persistent CTA,
    image edges inside a tile, nothing written outside the planes."""
    from fvfi import conv
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn((B, Cin, H, W), device="cuda", generator=g)
    w = torch.randn((Cout, Cin, 3, 3), device="cuda", generator=g) / (3 * Cin ** 0.5)
    b = torch.randn((Cout,), device="cuda", generator=g)
    canary = torch.full((B * Cout * H * W + 64,), 7.5, device="cuda")
    out = canary[:B * Cout * H * W].view(B, Cout, H, W)
    y = conv.conv2d(x, w, b, "zeros", act, out=out, nchw_out=True)
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    ref = {None: lambda t: t, "relu": F.relu, "sigmoid": torch.sigmoid, "softmax": lambda t: torch.softmax(t, 1)}[act](ref)
    assert float((y.double() - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))
    assert float((canary[B * Cout * H * W:] - 7.5).abs().max()) == 0.0           # nothing written past the last plane
    conv.check_overflow()


@pytest.mark.parametrize("B,Cin,Cout,H,W", [(2, 64, 64, 40, 56), (1, 128, 512, 17, 30), (1, 32, 25, 33, 47)])
def test_conv_residual_epilogue(B, Cin, Cout, H, W, prec):
    """y = relu(conv(x)) + skip in one launch (fvfi_conv2d_nhwc_residual): KernelEstimation's decoder additions d_k + c_k
    (fusion_adacofnet.py:128-138); vector and scalar residual loads, Cout > 256 split."""
    from fvfi import conv
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn((B, Cin, H, W), device="cuda", generator=g)
    w = torch.randn((Cout, Cin, 3, 3), device="cuda", generator=g) / (3 * Cin ** 0.5)
    b = torch.randn((Cout,), device="cuda", generator=g)
    skip = torch.randn((B, Cout, H, W), device="cuda", generator=g)
    y = conv.conv2d(x, w, b, "zeros", "relu", residual=skip)
    ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), padding=1)) + skip.double()
    assert y.shape == ref.shape
    assert float((y.double() - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))
    conv.check_overflow()


@pytest.mark.parametrize("B,Cin,Cout,Hs,Ws,mode,act,align,extra", [
    (2, 25, 25, 37, 53, "zeros", None, True, "nchw"),        # head tail: 25 channels stored as 32, planar output
    (1, 25, 25, 20, 31, "zeros", "softmax", True, "nchw"),
    (2, 64, 64, 17, 30, "zeros", "relu", True, "residual"),  # moduleUpsample: Upsample -> Conv -> ReLU, + skip in the epilogue
    (1, 128, 128, 9, 15, "zeros", "relu", True, None),       # several K chunks per tile
    (1, 512, 512, 5, 8, "zeros", "relu", True, "residual"),  # Cout > 256: two launches
    (1, 40, 16, 23, 19, "reflect", "elu", False, None),      # align_corners=False, reflect padding of the upsampled image
])
def test_conv_fused_bilinear_upsample(B, Cin, Cout, Hs, Ws, mode, act, align, extra, prec):
    """Upsample(x2, bilinear) -> Conv2d as ONE kernel (fvfi_conv2d_nhwc_upsampled, fusion_adacofnet.py:29-36,41-48): the loaders
    evaluate the resampling.  Bit-identical to resize kernel + convolution (same arithmetic), and within the convolution's own
    tolerance of an fp64 F.interpolate -> conv2d."""
    from fvfi import conv
    g = torch.Generator(device="cuda").manual_seed(11)
    Cs = (Cin + 15) // 16 * 16
    src = torch.zeros((B, Cs, Hs, Ws), device="cuda").contiguous(memory_format=torch.channels_last)
    src[:, :Cin] = torch.randn((B, Cin, Hs, Ws), device="cuda", generator=g)
    w = torch.randn((Cout, Cin, 3, 3), device="cuda", generator=g) / (3 * Cin ** 0.5)
    b = torch.randn((Cout,), device="cuda", generator=g)
    H, W = 2 * Hs, 2 * Ws
    skip = torch.randn((B, Cout, H, W), device="cuda", generator=g) if extra == "residual" else None
    kw = dict(nchw_out=(extra == "nchw"), residual=skip)
    y = conv.conv2d(src, w, b, mode, act, upsample=((H, W), align), **kw)
    up = conv.resize_bilinear(src, (H, W), align)
    y2 = conv.conv2d(up, w, b, mode, act, **kw)
    assert y.shape == (B, Cout, H, W) and torch.equal(y, y2)
    xr = F.interpolate(src[:, :Cin].double(), size=(H, W), mode="bilinear", align_corners=align)
    if mode == "reflect":
        xr = F.pad(xr, (1, 1, 1, 1), mode="reflect")
    ref = F.conv2d(xr, w.double(), b.double(), padding=0 if mode == "reflect" else 1)
    ref = {None: lambda t: t, "relu": F.relu, "elu": F.elu, "softmax": lambda t: torch.softmax(t, 1)}[act](ref)
    if skip is not None:
        ref = ref + skip.double()
    assert float((y.double() - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))
    conv.check_overflow()


@pytest.mark.parametrize("B,Cup,Cd,Cout,K,hs,ws,H,W", [
    (2, 64, 24, 64, 3, 27, 38, 38, 54),      # PhaseNet level: 64 resampled feature channels + 16 value + 8 prediction channels, ratio sqrt(2)
    (1, 64, 17, 64, 1, 9, 13, 13, 18),       # the first level: 1x1 convolution, 81 input channels (direct record padded to 24)
    (1, 32, 40, 48, 3, 20, 20, 40, 40),      # several direct chunks
])
def test_conv_two_source_resampled_concat(B, Cup, Cd, Cout, K, hs, ws, H, W, prec):
    """conv(cat(interpolate(x, size), x_direct)) without the concat (fvfi_conv2d_nhwc_upsampled, two-source form; phase_net.py:138-148):
    bit-identical to resize kernel -> torch.cat -> convolution, and within the convolution's tolerance of the fp64 composition."""
    from fvfi import conv
    g = torch.Generator(device="cuda").manual_seed(13)
    x = torch.randn((B, Cup, hs, ws), device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    Cs = (Cd + 7) // 8 * 8
    xd = torch.zeros((B, Cs, H, W), device="cuda").contiguous(memory_format=torch.channels_last)
    xd[:, :Cd] = torch.randn((B, Cd, H, W), device="cuda", generator=g)
    Cin = Cup + Cd
    w = torch.randn((Cout, Cin, K, K), device="cuda", generator=g) / (K * Cin ** 0.5)
    b = torch.randn((Cout,), device="cuda", generator=g)
    mode = "reflect" if K == 3 else "zeros"
    y = conv.conv2d(x, w, b, mode, "elu", upsample=((H, W), False), x_direct=xd)
    cat = torch.cat((conv.resize_bilinear(x, (H, W), False), xd[:, :Cd]), 1)
    y2 = conv.conv2d(cat, w, b, mode, "elu")
    assert y.shape == (B, Cout, H, W) and torch.equal(y, y2)
    xr = torch.cat((F.interpolate(x.double(), size=(H, W), mode="bilinear", align_corners=False), xd[:, :Cd].double()), 1)
    if K == 3:
        xr = F.pad(xr, (1, 1, 1, 1), mode="reflect")
    ref = F.elu(F.conv2d(xr, w.double(), b.double()))
    assert float((y.double() - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))
    conv.check_overflow()


BWD_CASES = [
    # B, Cin, Cout, K, H, W, mode, act, scale of gy  -- the seven layers of FusionNet (fusion_net.py:24-36) at reduced size + edge cases
    (2, 18, 32, 5, 40, 56, "reflect", "relu", 1.0),
    (2, 32, 64, 5, 20, 28, "reflect", "relu", 1.0),
    (2, 64, 128, 3, 10, 14, "reflect", "relu", 1.0),
    (2, 128, 128, 3, 5, 7, "reflect", "relu", 1.0),
    (1, 128, 64, 5, 10, 14, "reflect", None, 1.0),
    (1, 64, 32, 5, 20, 28, "reflect", None, 1e-7),       # gradients far below fp16's range: the data gradient runs in 3xTF32
    (2, 32, 3, 1, 40, 56, "zeros", None, 1.0),
    (1, 24, 5, 3, 31, 45, "zeros", "elu", 1.0),
    (1, 16, 40, 3, 17, 19, "reflect", "tanh", 1e3),
    (3, 7, 33, 3, 9, 11, "zeros", "sigmoid", 1.0),
]


@pytest.mark.parametrize("B,Cin,Cout,K,H,W,mode,act,gscale", BWD_CASES)
def test_conv_backward_matches_fp64(B, Cin, Cout, K, H, W, mode, act, gscale):
    """dL/dx, dL/dw, dL/db of conv2d's autograd Function (csrc/conv_bwd.cu + the tcgen05 kernel as dgrad) against torch's fp64
    autograd of the same expression; no ATen convolution kernel may run in the backward."""
    from fvfi import conv, _lib
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((B, Cin, H, W), device="cuda", generator=g)
    w = torch.randn((Cout, Cin, K, K), device="cuda", generator=g) / (Cin * K * K) ** 0.5
    b = torch.randn((Cout,), device="cuda", generator=g)
    gy = torch.randn((B, Cout, H, W), device="cuda", generator=g) * gscale
    xg, wg, bg = (t.clone().requires_grad_(True) for t in (x, w, b))
    with torch.enable_grad():
        y = conv.conv2d(xg, wg, bg, mode, act)
    n0 = _lib.lib().fvfi_launch_count()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        y.backward(gy)
        torch.cuda.synchronize()
    assert _lib.lib().fvfi_launch_count() - n0 >= 5
    foreign = [e.key for e in prof.key_averages() if "cudnn" in e.key.lower() or "convolution" in e.key.lower() or "gemm" in e.key.lower()]
    assert not foreign, foreign
    # fp64 reference built on the GPU kernel's own forward output for the activation mask (a ReLU mask is discontinuous)
    xd, wd, bd = (t.double().clone().requires_grad_(True) for t in (x, w, b))
    p = K // 2
    xr = F.pad(xd, (p, p, p, p), mode="reflect") if (mode == "reflect" and p) else xd
    pre = F.conv2d(xr, wd, bd, padding=0 if (mode == "reflect" and p) else p)
    yd = y.detach().double()
    dact = {None: torch.ones_like(yd), "relu": (yd > 0).double(), "elu": torch.where(yd > 0, torch.ones_like(yd), yd + 1),
            "tanh": 1 - yd * yd, "sigmoid": yd * (1 - yd)}[act]
    pre.backward(gy.double() * dact)
    for name, got, ref in (("gx", xg.grad, xd.grad), ("gw", wg.grad, wd.grad), ("gb", bg.grad, bd.grad)):
        err = float((got.double() - ref).abs().max())
        scale = float(ref.abs().max())
        print("%s max abs err %.2e (max |ref| %.2e)" % (name, err, scale))
        assert got.shape == ref.shape
        assert err <= 2e-5 * scale + 1e-30, (name, err, scale)


def test_conv_backward_is_reproducible_and_handles_padded_input_channels():
    """Two backward passes give identical bits (fixed summation order: the data-parallel parity check relies on it); an input
    with zero padding channels beyond the weight's Cin gets zero gradient there."""
    from fvfi import conv
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.zeros((2, 24, 24, 40), device="cuda")
    x[:, :18] = torch.randn((2, 18, 24, 40), device="cuda", generator=g)
    w = torch.randn((32, 18, 5, 5), device="cuda", generator=g) * 0.05
    b = torch.randn((32,), device="cuda", generator=g)
    gy = torch.randn((2, 32, 24, 40), device="cuda", generator=g)
    outs = []
    for _ in range(2):
        xg, wg, bg = (t.clone().requires_grad_(True) for t in (x, w, b))
        with torch.enable_grad():
            conv.conv2d(xg, wg, bg, "reflect", "relu").backward(gy)
        outs.append((xg.grad.clone(), wg.grad.clone(), bg.grad.clone()))
    for a, c in zip(*outs):
        assert torch.equal(a, c)
    assert outs[0][0].shape == x.shape and float(outs[0][0][:, 18:].abs().max()) == 0.0
    xd, wd, bd = x[:, :18].double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    F.relu(F.conv2d(F.pad(xd, (2, 2, 2, 2), mode="reflect"), wd, bd)).backward(gy.double())
    assert float((outs[0][0][:, :18].double() - xd.grad).abs().max()) <= 2e-5 * float(xd.grad.abs().max())
    assert float((outs[0][1].double() - wd.grad).abs().max()) <= 2e-5 * float(wd.grad.abs().max())


@pytest.mark.parametrize("B,C,H,W", [(2, 32, 16, 24), (1, 3, 9, 7), (1, 64, 8, 8)])
def test_max_pool2_backward_matches_torch(B, C, H, W):
    """fvfi_max_pool2_backward_nhwc vs torch autograd of F.max_pool2d (incl. ties after a ReLU and odd sizes)."""
    from fvfi import conv
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.relu(torch.randn((B, C, H, W), device="cuda", generator=g))        # ~half zeros: tied windows
    gy = torch.randn((B, C, H // 2, W // 2), device="cuda", generator=g)
    xa = x.clone().requires_grad_(True)
    with torch.enable_grad():
        ya = conv.max_pool2(xa)
    ya.backward(gy)
    xb = x.clone().requires_grad_(True)
    yb = F.max_pool2d(xb, 2, 2)
    yb.backward(gy)
    assert torch.equal(ya.detach(), yb.detach())
    assert torch.equal(xa.grad, xb.grad)


@pytest.mark.parametrize("B,C,Hi,Wi,Ho,Wo,align,relu_in", [
    (2, 32, 8, 12, 16, 24, False, True), (1, 128, 4, 4, 8, 8, False, False), (1, 5, 7, 9, 14, 18, True, False),
    (1, 8, 5, 6, 13, 17, False, True), (1, 4, 6, 6, 6, 6, False, False)])
def test_resize_bilinear_backward_matches_torch(B, C, Hi, Wi, Ho, Wo, align, relu_in):
    """fvfi_resize_bilinear_backward_nhwc (adjoint of the fused resize: ReLU on the input, skip added on the way out) vs torch autograd."""
    from fvfi import conv
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn((B, C, Hi, Wi), device="cuda", generator=g)
    add = torch.randn((B, C, Ho, Wo), device="cuda", generator=g)
    gy = torch.randn((B, C, Ho, Wo), device="cuda", generator=g)
    xa, aa = x.clone().requires_grad_(True), add.clone().requires_grad_(True)
    with torch.enable_grad():
        ya = conv.resize_bilinear(xa, (Ho, Wo), align, relu_input=relu_in, add=aa)
    ya.backward(gy)
    xb, ab = x.double().requires_grad_(True), add.double().requires_grad_(True)
    yb = F.interpolate(torch.relu(xb) if relu_in else xb, size=(Ho, Wo), mode="bilinear", align_corners=align) + ab
    yb.backward(gy.double())
    assert float((ya.detach().double() - yb.detach()).abs().max()) <= 1e-5
    assert float((xa.grad.double() - xb.grad).abs().max()) <= 1e-5 * max(1.0, float(xb.grad.abs().max()))
    assert torch.equal(aa.grad, gy)


def test_fusion_blend_backward_matches_torch():
    from fvfi.fusion_net import fusion_blend
    g = torch.Generator(device="cuda").manual_seed(5)
    base = torch.rand((2, 3, 20, 28), device="cuda", generator=g)
    x = torch.randn((2, 3, 20, 28), device="cuda", generator=g)
    gout = torch.randn((2, 3, 20, 28), device="cuda", generator=g)
    ba, xa = base.clone().requires_grad_(True), x.clone().requires_grad_(True)
    with torch.enable_grad():
        oa = fusion_blend(ba, xa)
    oa.backward(gout)
    bb, xb = base.double().requires_grad_(True), x.double().requires_grad_(True)
    ob = (bb + torch.tanh(xb)).clamp(0, 1)
    ob.backward(gout.double())
    assert float((oa.detach().double() - ob.detach()).abs().max()) <= 2e-6
    assert float((xa.grad.double() - xb.grad).abs().max()) <= 1e-5 and float((ba.grad.double() - bb.grad).abs().max()) <= 1e-6


def test_planar_concat_nhwc_matches_cat():
    """fvfi_planar_concat_nhwc == torch.cat(..., 1) padded with zero channels to a multiple of 4, stored channels_last (FusionNet's input,
    fusion_net.py:47)."""
    from fvfi import conv
    g = torch.Generator(device="cuda").manual_seed(7)
    for chans, (B, H, W) in (((3, 3, 3, 6, 3), (2, 30, 44)), ((1,), (1, 5, 7)), ((8, 8, 16), (3, 16, 16)), ((5, 2), (1, 33, 19))):
        parts = [torch.randn((B, c, H, W), device="cuda", generator=g) for c in chans]
        out = conv.planar_concat_nhwc(parts)
        C = sum(chans)
        ref = torch.cat(parts, 1)
        assert out.shape == (B, (C + 3) // 4 * 4, H, W) and out.is_contiguous(memory_format=torch.channels_last)
        assert torch.equal(out[:, :C], ref) and float(out[:, C:].abs().sum()) == 0.0
