"""Per-layer cost of the folded eval BatchNorm epilogue vs conv + separate bn_act pass (512x512 generator shapes)."""
import sys
import torch
sys.path.insert(0, ".")
from discogan_modernized_b200 import ops


def timed(fn, reps=10, iters=5):
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
for (H, Cb, Cs) in [(256, 64, 128), (128, 128, 256), (64, 256, 512), (32, 512, 1024), (16, 1024, 2048), (8, 2048, 2048)]:
    big = torch.randn(B, H, H, Cb, device="cuda").to(torch.bfloat16)
    small = torch.randn(B, H // 2, H // 2, Cs, device="cuda").to(torch.bfloat16)
    w = torch.randn(Cs, Cb, 4, 4, device="cuda") * 0.01
    wd, wu = ops.pack_weights(w)
    for mode, C in (("down", Cs), ("up", Cb)):
        stats = ops.bn_eval_stats(torch.ones(C, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda"),
                                  torch.ones(C, device="cuda"))
        if mode == "down":
            plain = lambda: ops.conv_down(big, wd)
            fold = lambda: ops.conv_down(big, wd, affine=(stats, 1, 0.2))
            two = lambda: ops.bn_act_fwd(ops.conv_down(big, wd).view(-1, C), stats, 1, 0.2)
        else:
            plain = lambda: ops.conv_up(small, wu)
            fold = lambda: ops.conv_up(small, wu, affine=(stats, 2, 0.2))
            two = lambda: ops.bn_act_fwd(ops.conv_up(small, wu).view(-1, C), stats, 2, 0.2)
        tp, tf, tt = timed(plain), timed(fold), timed(two)
        print(f"B{B} {mode:4s} H{H} {Cb}->{Cs}: conv {tp:8.1f} us | folded {tf:8.1f} us | conv + bn_act {tt:8.1f} us")
