"""Narrow layers (<= 128 output channels): regular vs role-swapped kernel."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from discogan_modernized_b200 import ops  # noqa: E402


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


for (B, H, Cb, Cs) in [(32, 256, 64, 128), (32, 128, 128, 256), (64, 32, 64, 128)]:
    big = torch.randn(B, H, H, Cb, device="cuda").to(torch.bfloat16)
    small = torch.randn(B, H // 2, H // 2, Cs, device="cuda").to(torch.bfloat16)
    w = torch.randn(Cs, Cb, 4, 4, device="cuda") * 0.01
    wd, wu = ops.pack_weights(w)
    fl = 2.0 * B * (H // 2) ** 2 * Cs * Cb * 16
    for name, bn in (("auto-noswap", 0), ("swap", 1)):
        ops.set_conv_tiling(bn, -1)
        md = timeit(lambda: ops.conv_down(big, wd))
        mu = timeit(lambda: ops.conv_up(small, wu))
        mm = timeit(lambda: ops.conv_up(small, wu, mask=big, slope=0.2))
        print(f"{name} B{B} H{H} {Cb}->{Cs}: down {md * 1e3:.1f} us {fl / md / 1e9:.0f} TF | "
              f"up {mu * 1e3:.1f} us {fl / mu / 1e9:.0f} TF | masked up {mm * 1e3:.1f} us {fl / mm / 1e9:.0f} TF")
    ops.set_conv_tiling(0, -1)
