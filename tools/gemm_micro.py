"""Time conv_down / conv_up on the 512^2 layer shapes (B=32), CTA pairs on/off."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from discogan_modernized_b200 import ops  # noqa: E402

shapes = [(32, 256, 64, 128), (32, 128, 128, 256), (32, 64, 256, 512), (32, 32, 512, 1024), (32, 16, 1024, 2048)]
for pair in (0, 1):
    ops.set_conv_tiling(0, pair)
    for B, H, Cb, Cs in shapes:
        big = torch.randn(B, H, H, Cb, device="cuda").to(torch.bfloat16)
        small = torch.randn(B, H // 2, H // 2, Cs, device="cuda").to(torch.bfloat16)
        w = torch.randn(Cs, Cb, 4, 4, device="cuda") * 0.01
        wd, wu = ops.pack_weights(w)
        for name, fn in (("down", lambda: ops.conv_down(big, wd)), ("up", lambda: ops.conv_up(small, wu))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                fn()
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 10
            fl = 2.0 * B * (H // 2) ** 2 * Cs * Cb * 16
            print(f"pair={pair} {name} B{B} {H}->{H // 2} {Cb}->{Cs}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.0f} TFLOP/s")
