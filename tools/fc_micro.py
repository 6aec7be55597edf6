"""In-graph cost of the 4x4 valid heads (fc_down / fc_up / fc_wgrad) at the 64x64 (B=64, C=512) and 512x512 (B=32, C=2048) shapes."""
import sys
import torch
sys.path.insert(0, ".")
from discogan_modernized_b200 import ops


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


for B, C in ((64, 512), (32, 2048)):
    K = 16 * C
    big = torch.randn(B, K, device="cuda").bfloat16()
    small = torch.randn(B, 100, device="cuda").bfloat16()
    wd = torch.randn(100, K, device="cuda").bfloat16()
    dw = torch.zeros(100, C, 4, 4, device="cuda")
    wd1 = torch.randn(1, K, device="cuda").bfloat16()
    dl = torch.randn(B, 1, device="cuda")
    dw1 = torch.zeros(1, C, 4, 4, device="cuda")
    print(f"B={B} C={C}: fc_down {timed(lambda: ops.fc_down(big, wd)):6.1f} us | fc_up {timed(lambda: ops.fc_up(small, wd)):6.1f} us | "
          f"fc_wgrad {timed(lambda: ops.fc_wgrad(small, big, dw, 0.0)):6.1f} us | D head: fc_down {timed(lambda: ops.fc_down(big, wd1, out_f32=True)):6.1f} us "
          f"fc_up {timed(lambda: ops.fc_up(dl, wd1)):6.1f} us fc_wgrad {timed(lambda: ops.fc_wgrad(dl, big, dw1, 0.0)):6.1f} us")
