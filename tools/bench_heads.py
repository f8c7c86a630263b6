"""Times the KernelEstimation head tails and the direct kernels at 1088x1920 (CUDA events): python tools/bench_heads.py [B]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi import conv as tc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H, W = 1088, 1920
torch.manual_seed(0)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    x32 = torch.randn((B, 32, H, W), device="cuda").contiguous(memory_format=torch.channels_last)
    x32[:, 25:] = 0
    w25 = torch.randn((25, 25, 3, 3), device="cuda") / 15
    b25 = torch.randn(25, device="cuda")
    for act in (None, "softmax"):
        ms = timeit(lambda: tc.conv2d(x32, w25, b25, "zeros", act, nchw_out=True))
        gb = B * H * W * (32 + 25) * 4 / 1e9
        print("head 25->25 3x3 %-8s B=%d  %.3f ms  %.0f GB/s  %.1f TF/s" % (act, B, ms, gb / ms * 1e3, 2 * B * H * W * 625 * 9 / ms / 1e9))
    xh = torch.randn((B, 64, H // 2, W // 2), device="cuda").contiguous(memory_format=torch.channels_last)
    m = torch.nn.Conv2d(64, 1, 3, 1, 1).cuda()
    ms = timeit(lambda: tc.upsample2_conv3x3_single(m, xh, "sigmoid"))
    print("occlusion tail (contract at half res + tapsum) B=%d  %.3f ms" % (B, ms))
    ms = timeit(lambda: tc.conv_module(m, tc.resize_bilinear(xh, (H, W), True), "sigmoid", nchw_out=True))
    print("occlusion tail (upsample + tensor-core 64->1)   B=%d  %.3f ms" % (B, ms))
    x64 = torch.randn((6, 64, 1080, 1920), device="cuda").contiguous(memory_format=torch.channels_last)
    w8 = torch.randn((8, 64, 1, 1), device="cuda") / 8
    b8 = torch.randn(8, device="cuda")
    ms = timeit(lambda: tc.conv2d(x64, w8, b8, "zeros", "tanh"))
    gb = 6 * 1080 * 1920 * (64 + 8) * 4 / 1e9
    print("PhaseNet 64->8 1x1 tanh direct  B=6  %.3f ms  %.0f GB/s" % (ms, gb / ms * 1e3))
