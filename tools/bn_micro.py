"""BN kernels on one large activation ([P, C] = [B*H*W, C]) for ncu / quick timing."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from discogan_modernized_b200 import ops  # noqa: E402

P, C = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (524288, 128)
z = torch.randn(P, C, device="cuda").to(torch.bfloat16)
dy = torch.randn(P, C, device="cuda").to(torch.bfloat16)
gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
for it in range(3):
    stats = ops.bn_stats(z, gamma, beta)
    y = ops.bn_act_fwd(z, stats, 1)
    dz = ops.bn_act_bwd(dy, y, z, stats, gamma, 1, 0.2, dg, db, 0.0)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
ev[0].record(); stats = ops.bn_stats(z, gamma, beta)
ev[1].record(); y = ops.bn_act_fwd(z, stats, 1)
ev[2].record(); dz = ops.bn_act_bwd(dy, y, z, stats, gamma, 1, 0.2, dg, db, 0.0)
ev[3].record(); torch.cuda.synchronize()
E = P * C
for name, i, b in (("bn_stats", 0, 2), ("bn_act_fwd", 1, 4), ("bn_act_bwd", 2, 10)):
    ms = ev[i].elapsed_time(ev[i + 1])
    print(f"{name}: {ms * 1e3:.1f} us, {b * E / ms / 1e6:.0f} GB/s")
