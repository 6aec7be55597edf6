"""Time conv_wgrad on the 512^2 layer shapes (B=32)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from discogan_modernized_b200 import ops  # noqa: E402

shapes = [(32, 128, 128, 64), (32, 32, 512, 256), (32, 8, 2048, 1024), (32, 4, 2048, 2048)]
for B, Hs, Cs, Cb in shapes:
    small = torch.randn(B, Hs, Hs, Cs, device="cuda").to(torch.bfloat16)
    big = torch.randn(B, 2 * Hs, 2 * Hs, Cb, device="cuda").to(torch.bfloat16)
    dw = torch.zeros(Cs, Cb, 4, 4, device="cuda")
    for _ in range(3):
        ops.conv_wgrad(small, big, dw, 0.0)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ops.conv_wgrad(small, big, dw, 0.0)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    fl = 2.0 * B * Hs * Hs * Cs * Cb * 16
    print(f"wgrad B{B} {Hs} {Cs}x{Cb}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.0f} TFLOP/s")
