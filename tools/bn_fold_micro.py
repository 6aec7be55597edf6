"""Micro-benchmark of the folded BatchNorm finalize against the three-launch path at the 64x64 / 512x512 layer shapes
(CUDA-graph replay of 20 dependent repetitions, so launch gaps count the way they do inside the step graph)."""
import sys
import torch
sys.path.insert(0, ".")
from discogan_modernized_b200 import ops


def timed(fn, reps=20, iters=20):
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


ctx = ops.OpsContext()
with ops.use_context(ctx):
    for P, C in ((65536, 64), (16384, 128), (4096, 256), (1024, 512), (64, 100), (2097152, 64), (131072, 256)):
        z = torch.randn(P, C, device="cuda").bfloat16()
        dy = torch.randn(P, C, device="cuda").bfloat16()
        gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        part = torch.randn(2, 128, C, device="cuda").abs()
        acc = part.sum(1).contiguous()
        st = ops.bn_stats(z, gamma, beta)
        y = ops.bn_act_fwd(z, st, 1, 0.2)

        def fwd3():
            s = ops.bn_stats_finalize(part, P, gamma, beta)
            ops.bn_act_fwd(z, s, 1, 0.2)

        def fwd2():
            ops.bn_act_fwd_acc(z, acc, gamma, beta, 1, 0.2)

        def bwd(fold):
            def f():
                ctx.fold_stats = fold
                ctx.arena.cursor = 0
                ops.bn_act_bwd(dy, y, z, st, gamma, 1, 0.2, dg, db, 0.0)
            return f
        t3, t2 = timed(fwd3), timed(fwd2)
        b3, b2 = timed(bwd(False)), timed(bwd(True))
        print(f"P={P:8d} C={C:4d}: fwd finalize+act {t3:7.2f} us  folded {t2:7.2f} us | bwd 3-launch {b3:7.2f} us  folded {b2:7.2f} us")
