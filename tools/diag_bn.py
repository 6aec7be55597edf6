"""Element-level diagnosis of dg_bn_act_bwd against torch autograd (rows whose dz deviates by more than a bf16 ulp)."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from discogan_modernized_b200 import ops

torch.backends.cudnn.allow_tf32 = False
for P, C, act in ((65536, 64, 2), (65536, 64, 1), (65536, 128, 2), (16384, 64, 2), (131072, 64, 2), (2097152, 64, 2)):
    g = torch.Generator(device="cuda").manual_seed(1)
    z = (torch.randn(P, C, device="cuda", generator=g) * 1.3 + 0.4).bfloat16()
    dy = (torch.randn(P, C, device="cuda", generator=g) * 1e-3).bfloat16()
    gamma = torch.rand(C, device="cuda", generator=g) + 0.5
    beta = torch.randn(C, device="cuda", generator=g) * 0.1
    zf = z.float().requires_grad_(True)
    gp, bp = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yb = F.batch_norm(zf, None, None, gp, bp, True, 0.1, 1e-5)
    yr = F.relu(yb) if act == 2 else F.leaky_relu(yb, 0.2)
    yr.backward(dy.float())
    stats = ops.bn_stats(z, gamma, beta)
    y = ops.bn_act_fwd(z, stats, act, 0.2)
    dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(beta)
    dz = ops.bn_act_bwd(dy, y, z, stats, gamma, act, 0.2, dgamma, dbeta, 0.0)
    torch.cuda.synchronize()
    ref = zf.grad
    err = (dz.float() - ref).abs()
    tol = ref.abs() * 2 ** -7 + 1e-9
    bad = (err > tol)
    rows = bad.any(1).nonzero().flatten()
    rel = float((dz.float() - ref).norm() / ref.norm())
    print(f"P={P} C={C} act={act}: rel-L2 {rel:.3e}  bad elements {int(bad.sum())} in {rows.numel()} rows; first rows {rows[:12].tolist()} "
          f"dgamma {float((dgamma - gp.grad).norm() / gp.grad.norm()):.2e} dbeta {float((dbeta - bp.grad).norm() / bp.grad.norm()):.2e}")
    if rows.numel():
        r = int(rows[0])
        cols = bad[r].nonzero().flatten()[:6].tolist()
        print("   row", r, "cols", cols, "got", dz[r, cols].float().tolist(), "ref", ref[r, cols].tolist(), "y", y[r, cols].float().tolist(),
              "yb", yb[r, cols].tolist())
