"""On-device diagnosis of the tcgen05 convolution kernels: tap-isolated weights make each output a pure shifted
copy of the input, so a wrong TMA coordinate shows up as a match against another tap's reference while a wrong
smem/UMMA descriptor shows up as garbage.  Usage: python tools/diag_tc.py > gpurun_out/diag_tc.txt"""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from discogan_modernized_b200 import ops  # noqa: E402

BF16 = torch.bfloat16
torch.backends.cudnn.allow_tf32 = False


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12))


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(BF16)


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


def probe(name, fn):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"[{name}] EXCEPTION: {e}")
        return False
    return True


def diag_down(B=2, H=16, Cb=64, Cs=128):
    print(f"== DOWN B={B} H={H} Cb={Cb} Cs={Cs}")
    x = torch.randn(B, Cb, H, H, device="cuda").to(BF16)
    refs = []
    for t in range(16):
        w = torch.zeros(Cs, Cb, 4, 4, device="cuda")
        for c in range(min(Cs, Cb)):
            w[c, c, t // 4, t % 4] = 1.0
        refs.append(F.conv2d(x.float(), w, stride=2, padding=1))
    for t in range(16):
        w = torch.zeros(Cs, Cb, 4, 4, device="cuda")
        for c in range(min(Cs, Cb)):
            w[c, c, t // 4, t % 4] = 1.0
        wd, _ = ops.pack_weights(w)
        out = nchw(ops.conv_down(nhwc(x.float()), wd))
        errs = [rel(out, r) for r in refs]
        best = min(range(16), key=lambda i: errs[i])
        print(f"  tap {t:2d}: err vs own ref {errs[t]:.3e}; best match tap {best} ({errs[best]:.3e}); "
              f"out norm {float(out.norm()):.3f} ref norm {float(refs[t].norm()):.3f}")
    w = (torch.randn(Cs, Cb, 4, 4, device="cuda") / 32).to(BF16).float()
    wd, _ = ops.pack_weights(w)
    out = nchw(ops.conv_down(nhwc(x.float()), wd))
    ref = F.conv2d(x.float(), w, stride=2, padding=1)
    print(f"  random w: rel {rel(out, ref):.3e}")
    e = (out - ref).abs()
    print("  err by channel block of 32:", [f"{float(e[:, i:i + 32].mean()):.2e}" for i in range(0, Cs, 32)])
    print("  err by out row:", [f"{float(e[:, :, i].mean()):.2e}" for i in range(e.shape[2])])
    print("  err by batch:", [f"{float(e[i].mean()):.2e}" for i in range(B)])


def diag_up(B=2, Hs=8, Cb=64, Cs=128):
    print(f"== UP B={B} Hs={Hs} Cb={Cb} Cs={Cs}")
    s = torch.randn(B, Cs, Hs, Hs, device="cuda").to(BF16)
    refs = []
    for t in range(16):
        w = torch.zeros(Cs, Cb, 4, 4, device="cuda")
        for c in range(min(Cs, Cb)):
            w[c, c, t // 4, t % 4] = 1.0
        refs.append(F.conv_transpose2d(s.float(), w, stride=2, padding=1))
    for t in range(16):
        w = torch.zeros(Cs, Cb, 4, 4, device="cuda")
        for c in range(min(Cs, Cb)):
            w[c, c, t // 4, t % 4] = 1.0
        _, wu = ops.pack_weights(w)
        out = nchw(ops.conv_up(nhwc(s.float()), wu))
        errs = [rel(out, r) for r in refs]
        best = min(range(16), key=lambda i: errs[i])
        print(f"  tap {t:2d}: err vs own ref {errs[t]:.3e}; best match tap {best} ({errs[best]:.3e})")
    w = (torch.randn(Cs, Cb, 4, 4, device="cuda") / 32).to(BF16).float()
    _, wu = ops.pack_weights(w)
    out = nchw(ops.conv_up(nhwc(s.float()), wu))
    ref = F.conv_transpose2d(s.float(), w, stride=2, padding=1)
    print(f"  random w: rel {rel(out, ref):.3e}")
    e = (out - ref).abs()
    print("  err by out row:", [f"{float(e[:, :, i].mean()):.2e}" for i in range(e.shape[2])])


def diag_wgrad(B=2, H=16, Cb=64, Cs=128):
    print(f"== WGRAD B={B} H={H} Cb={Cb} Cs={Cs}")
    big = torch.randn(B, Cb, H, H, device="cuda").to(BF16)
    small = torch.randn(B, Cs, H // 2, H // 2, device="cuda").to(BF16)
    ref = torch.nn.grad.conv2d_weight(big.float(), (Cs, Cb, 4, 4), small.float(), stride=2, padding=1)
    dw = torch.zeros(Cs, Cb, 4, 4, device="cuda")
    ops.conv_wgrad(nhwc(small.float()), nhwc(big.float()), dw, beta=0.0)
    print(f"  rel {rel(dw, ref):.3e}")
    for t in range(16):
        errs = [rel(dw[:, :, t // 4, t % 4], ref[:, :, u // 4, u % 4]) for u in range(16)]
        best = min(range(16), key=lambda i: errs[i])
        print(f"  tap {t:2d}: err vs own {errs[t]:.3e}; best match tap {best} ({errs[best]:.3e}); "
              f"transposed-match {rel(dw[:, :, t // 4, t % 4][:Cb, :], ref[:, :, t // 4, t % 4][:Cb, :Cb].t()) if Cs >= Cb else -1:.3e}")
    e = (dw - ref).abs()
    print("  err by cs block of 32:", [f"{float(e[i:i + 32].mean()):.2e}" for i in range(0, Cs, 32)])
    print("  err by cb block of 16:", [f"{float(e[:, i:i + 16].mean()):.2e}" for i in range(0, Cb, 16)])


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    ops.device_check()
    if which in ("all", "down"):
        probe("down", diag_down)
    if which in ("all", "up"):
        probe("up", diag_up)
    if which in ("all", "wgrad"):
        probe("wgrad", diag_wgrad)
