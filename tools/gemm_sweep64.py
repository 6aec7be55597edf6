"""Tile / pair / split-K sweep over the 64x64 (B=64) layer shapes, graph-replayed back-to-back launches."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from discogan_modernized_b200 import ops  # noqa: E402


def timed(fn, reps=20, iters=10):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / (reps * iters)


B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for (H, Cb, Cs) in [(32, 64, 128), (16, 128, 256), (8, 256, 512)]:
    big = torch.randn(B, H, H, Cb, device="cuda").to(torch.bfloat16)
    small = torch.randn(B, H // 2, H // 2, Cs, device="cuda").to(torch.bfloat16)
    w = torch.randn(Cs, Cb, 4, 4, device="cuda") * 0.01
    wd, wu = ops.pack_weights(w)
    fl = 2.0 * B * (H // 2) ** 2 * Cs * Cb * 16
    for sk in (0, 1):
        ctx = ops.OpsContext()
        with ops.use_context(ctx):
            if sk:
                ops.enable_splitk(big.device)
            for bn in (0, 1, 64, 128, 256):
                for pair in ((-1,) if bn == 0 else (0, 1)):
                    try:
                        ops.set_conv_tiling(bn, pair)
                        md = timed(lambda: ops.conv_down(big, wd))
                        mu = timed(lambda: ops.conv_up(small, wu))
                        print(f"B{B} H{H} {Cb}->{Cs} splitk={sk} bn={bn:3d} pair={pair:2d}: down {md:6.1f} us {fl / md / 1e6:5.0f} TF | "
                              f"up {mu:6.1f} us {fl / mu / 1e6:5.0f} TF")
                    except Exception as ex:  # noqa: BLE001
                        print(f"B{B} H{H} {Cb}->{Cs} splitk={sk} bn={bn} pair={pair}: {str(ex)[:80]}")
