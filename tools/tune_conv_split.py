"""Output-channel split / tile height of the wide convolutions: python tools/tune_conv_split.py
(run once per FVFI_CONV_MT value: the tile override is read once per process)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi import conv


def t(fn, reps=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


B = 8
SHAPES = [("heads 64->448 @544x960", 64, 448, 544, 960, (256, 224, 128, 112, 64)),
          ("512->512 @68x120", 512, 512, 68, 120, (256, 128, 64)),
          ("256->512 @68x120", 256, 512, 68, 120, (256, 128)),
          ("512->256 @136x240", 512, 256, 136, 240, (256, 128, 64)),
          ("256->256 @136x240", 256, 256, 136, 240, (256, 128, 64)),
          ("128->128 @272x480", 128, 128, 272, 480, (128, 64)),
          ("256->128 @272x480", 256, 128, 272, 480, (128, 64))]
print("FVFI_CONV_MT =", os.environ.get("FVFI_CONV_MT", "auto"))
for name, ci, co, h, w, splits in SHAPES:
    x = torch.randn((B, ci, h, w), device="cuda").contiguous(memory_format=torch.channels_last)
    wt = torch.randn((co, ci, 3, 3), device="cuda") / (ci * 9) ** 0.5
    bias = torch.randn((co,), device="cuda")
    flops = 2.0 * B * h * w * ci * co * 9
    ref = None
    for sp in splits:
        conv.n_split = sp if sp < co else None
        ms = t(lambda: conv.conv2d(x, wt, bias, "zeros", "relu"))
        y = conv.conv2d(x, wt, bias, "zeros", "relu")
        if ref is None: ref = y
        print("%-24s split %3d: %.3f ms  %.0f TF/s  (max diff vs first split %.2g)" % (name, sp, ms, flops / ms / 1e9, float((y - ref).abs().max())))
    conv.n_split = None
