"""Parity of the B200 train step at the BENCHMARKED configurations (64x64 B=64, 512x512 B=32) against oracles run on
the same device with TF32 off:

* ``fp32``   -- the reference restated in stock fp32 PyTorch (oracle/step.py; image_translation.py:335-390);
* ``bf16e``  -- the same oracle with bf16 rounding at the points where the kernels store bf16 (oracle/bf16_emul.py);
* ``bf16e2`` -- a twin of ``bf16e`` whose pre-rounding values carry 1e-6 relative noise (a different fp32 summation
                order).  Rounding noise cascades through the BatchNorm stack (oracle/bf16_emul.py docstring), so two
                correct bf16 computations agree end to end only to the distance(bf16e2, bf16e) = the NOISE FLOOR.

What is asserted, and with which stated tolerance:

1. teacher-forced per-layer parity (``test_layers_teacher_forced``): every tensor-core conv / convT layer and its
   BatchNorm, fed the emulation's own bf16 inputs at the benchmarked batch, reproduces that layer's output, data
   gradient, weight gradient, BatchNorm output / input gradient / gamma-beta gradients:
       bf16 outputs  rel-L2 <= 4e-3 (one bf16 rounding of an fp32-accurate value measures 1.66e-3),  fp32 outputs <= 1e-3,
       fused BatchNorm statistics: mean within 5e-5 sigma, invstd within 1e-4 of the fp64 statistics.
   No cascade is involved, so these bounds are tight.
2. end-to-end gradients after one backward from identical weights (``test_per_layer_gradient_table``), every parameter:
       rel-L2(kernel, bf16e) <= 1.3 * floor + 0.02         (as close to bf16e as its own twin is)
       rel-L2(kernel, fp32)  <= 1.25 * rel-L2(bf16e, fp32) + 0.02   (no further from fp32 than bf16 storage itself is)
3. losses over the first steps (``test_step_parity_at_benchmarked_batch``): 64x64: |d| <= 2 % + 0.01 vs bf16e and
   5 % + 0.02 vs fp32 over the first D,G,G cycle (measured <= 0.8 % / 2.4 %), 5 % + 0.02 / 8 % + 0.03 over the second;
   512x512: the first band at steps 0-1; at step 2 (two Adam steps driven by gradients whose
   measured bf16 noise floor is 20-45 % at this size, and a discriminator that saturated at step 1: gen loss 29 -> 1)
   reconstruction / FM losses 5 % + 0.02, GAN losses 35 % + 0.10 (run-to-run spread of the kernel itself: 0.89 / 1.07).
4. accumulated update after those steps: rel-L2(w - w0) vs bf16e <= 1.3 * floor + 0.03.
5. 100-step loss curves at 64x64 B=64: 50-step means within 3 % (recon) / 5 % (dis) / 8 % (fm) / 15 % (gen) of fp32;
   worst single step bounded by what bf16e itself shows against fp32.
6. data-parallel semantics on one GPU (two in-process ranks) against the oracle's R = 2 emulation of the DDP step.

Every test appends its measured numbers to gpurun_out/parity_report.jsonl (summarised in profiles/r02_parity.md).
"""
import json
import os
from pathlib import Path

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
REPORT = Path(os.environ.get("DG_PARITY_REPORT", Path(__file__).resolve().parent.parent / "gpurun_out" / "parity_report.jsonl"))
LOSSES = ("dis_loss_A", "gen_loss_A", "dis_loss_B", "gen_loss_B", "fm_loss_A", "fm_loss_B", "recon_loss_A", "recon_loss_B")
NAMES = ("G_A", "G_B", "D_A", "D_B")


@pytest.fixture(scope="module", autouse=True)
def _fp32_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def report(name, payload):
    try:
        REPORT.parent.mkdir(parents=True, exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(json.dumps({"test": name, **payload}) + "\n")
    except OSError:
        pass


def rel_l2(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cos(a, b):
    a, b = a.detach().float().flatten(), b.detach().float().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))


def make_all(S, arch="discogan", variant="image_translation", first_iter=0, want=("fp32", "bf16e", "bf16e2"), **trainer_kw):
    """Trainer + oracles from identical seed-1234 weights (construction order distributed_image_translation.py:372-376)."""
    from discogan_modernized_b200 import DiscoGANTrainer, model
    from oracle.bf16_emul import emulate
    from oracle.step import OracleStep, build_nets
    src = build_nets(S, seed=1234, device="cuda")
    nets = []
    with torch.device("meta"):
        shells = [model.Generator(True, S), model.Generator(True, S), model.Discriminator(S), model.Discriminator(S)]
    for shell, r in zip(shells, src):
        n = shell.to_empty(device="cuda")
        n.load_state_dict(r.state_dict())
        nets.append(n)
    tr = DiscoGANTrainer(image_size=S, nets=nets, model_arch=arch, variant=variant, data_parallel=False, **trainer_kw)
    tr.iters = first_iter
    oracles = {}
    for kind in want:
        onets = src if kind == "fp32" else build_nets(S, seed=1234, device="cuda")
        wrapped = onets if kind == "fp32" else emulate(onets, perturb=1e-6 if kind == "bf16e2" else 0.0)
        st = OracleStep(wrapped, model_arch=arch, variant=variant, device="cuda")
        st.iters = first_iter
        st.raw_nets = onets
        oracles[kind] = st
    return tr, oracles


# ---------------------------------------------------------------------------------------------------------------------
# 1. teacher-forced per-layer parity
# ---------------------------------------------------------------------------------------------------------------------
def _layer_rows(tape, which):
    from discogan_modernized_b200 import ops
    nhwc, nchw = ops.nchw_f32_to_nhwc, ops.nhwc_to_nchw_f32
    rows = []
    for i, r in enumerate(tape):
        conv, bn = r["conv"], r["bn"]
        if conv.stride[0] != 2:
            continue                        # the 4x4 valid heads are SIMT kernels with B-independent plans (test_fc_heads)
        is_t = isinstance(conv, nn.ConvTranspose2d)
        x, gz, w = r["x_in"].detach(), r["z32"].grad, conv.weight.detach()
        wd, wu = ops.pack_weights(w.contiguous())
        xk, gzk = nhwc(x.contiguous()), nhwc(gz.contiguous())
        row = {"net": which, "layer": i, "kind": "convT" if is_t else "conv", "x": list(x.shape), "z": list(gz.shape)}
        if is_t:                                  # the product path: statistics accumulated by the conv epilogue /
            z, acc = ops.conv_up_acc(xk, wu)       # the split-K finish kernel (dg_conv_opts.stat_accumulate)
            dx = ops.conv_down(gzk, wd)
            dw = torch.empty_like(w)
            ops.conv_wgrad(xk, gzk, dw, beta=0.0)
        else:
            z, acc = ops.conv_down_acc(xk, wd)
            dx = ops.conv_up(gzk, wu)
            dw = torch.empty_like(w)
            ops.conv_wgrad(gzk, xk, dw, beta=0.0)
        row["fprop"] = rel_l2(nchw(z), r["z32"])
        row["dgrad"] = rel_l2(nchw(dx), r["x_in"].grad)
        row["wgrad"] = rel_l2(dw, conv.weight.grad)
        # BatchNorm (+activation) on the emulation's own bf16 z.  torch normalises with statistics of that bf16 z, and so
        # does dg_bn_stats_acc: the BN checks below use it, so they see identical data.  The product path takes the sums
        # from the conv kernel's fp32 accumulators instead (closer to the fp32 reference; the mean moves by ~1e-5 sigma):
        # checked separately against the fp64 statistics of the unrounded conv output.
        C = z.shape[-1]
        zk = nhwc(r["z"].detach().contiguous())
        z2 = zk.view(-1, C)
        g, b = bn.weight.detach(), bn.bias.detach()
        act = ops.ACT_LRELU if r["act"] == "lrelu" else ops.ACT_RELU
        _, fused = ops.bn_act_fwd_acc(z2, acc, g, b, act, 0.2)
        z64 = r["z32"].detach().double()
        mean, var = z64.mean((0, 2, 3)), z64.var((0, 2, 3), unbiased=False)
        row["fused_mean_err_sigma"] = float(((fused[0].double() - mean).abs() / var.sqrt()).max())
        row["fused_invstd_rel"] = float(((fused[1].double() * (var + bn.eps).sqrt()) - 1).abs().max())
        row["stats_fused"] = True
        y, stats = ops.bn_act_fwd_acc(z2, ops.bn_stats_acc(z2), g, b, act, 0.2)
        row["bn_fwd"] = rel_l2(nchw(y.view(zk.shape)), r["y"])
        dyk = nhwc(r["y"].grad.contiguous())
        dgamma, dbeta = torch.empty_like(g), torch.empty_like(b)
        dz = ops.bn_act_bwd(dyk.view(-1, C), y, z2, stats, g, act, 0.2, dgamma, dbeta, 0.0)
        row["bn_bwd_dz"] = rel_l2(nchw(dz.view(zk.shape)), r["zr"].grad)
        row["bn_dgamma"] = rel_l2(dgamma, bn.weight.grad)
        row["bn_dbeta"] = rel_l2(dbeta, bn.bias.grad)
        rows.append(row)
    return rows


@pytest.mark.parametrize("S,B", [(64, 64), (512, 32)])
def test_layers_teacher_forced(S, B):
    """Per-layer outputs and gradients at the benchmarked batch, each layer fed the emulation's own tensors."""
    from discogan_modernized_b200 import ops
    from oracle import bf16_emul
    from oracle.step import build_nets, synthetic_batch
    G, _, D, _ = build_nets(S, seed=1234, device="cuda")
    A, Bt = synthetic_batch(B, S, step=0, device="cuda")
    ctx = ops.OpsContext()
    ctx.fold_stats = True                                # also covers the accumulator-mode kernels (opt-in in the product)
    rows = []
    with ops.use_context(ctx):
        ops.enable_splitk(A.device)                      # the trainer's launch plan: split-K available
        tape = []
        y = bf16_emul.generator_forward(G, A, tape)
        ((y - Bt) ** 2).mean().backward()
        rows += _layer_rows(tape, "G")
        del tape, y
        tape = []
        p, feats = bf16_emul.discriminator_forward(D, Bt, tape)
        p.log().mean().backward()       # (no feature-map term: its fp32 gradient would not be a bf16-exact BN-backward input)
        rows += _layer_rows(tape, "D")
    report("layers_teacher_forced", {"S": S, "B": B, "rows": rows})
    bad = []
    for r in rows:
        for k in ("fprop", "dgrad", "bn_fwd"):
            if r[k] > 4e-3:
                bad.append((k, r))
        # BatchNorm backward: elements whose pre-activation lies within fp32 rounding of zero (channels with
        # |mean| / std >> 1) may take the other ReLU branch than torch's two-step normalisation does; on random data the
        # kernel measures 1.66e-3 like every other bf16 output (tools/diag_bn.py), on the real 64x64 decoder tail 4.2e-3
        if r["bn_bwd_dz"] > 6e-3:
            bad.append(("bn_bwd_dz", r))
        if r["wgrad"] > 1e-3:
            bad.append(("wgrad", r))
        for k in ("bn_dgamma", "bn_dbeta"):      # sums with cancellation: the few near-zero ReLU-branch flips show (<= 1.1e-3 seen)
            if r[k] > 5e-3:
                bad.append((k, r))
        if r["fused_mean_err_sigma"] > 5e-5 or r["fused_invstd_rel"] > 1e-4:
            bad.append(("fused statistics", r))
    assert not bad, bad[:6]


# ---------------------------------------------------------------------------------------------------------------------
# 2. end-to-end gradient table
# ---------------------------------------------------------------------------------------------------------------------
def grad_rows(tr, oracles, nets_idx):
    rows = []
    for idx in nets_idx:
        mine = dict(tr.nets()[idx].named_parameters())
        others = {k: dict(st.raw_nets[idx].named_parameters()) for k, st in oracles.items()}
        for pname, p in mine.items():
            row = {"net": NAMES[idx], "param": pname}
            for k in ("fp32", "bf16e"):
                row[f"rel_{k}"] = rel_l2(p.grad, others[k][pname].grad)
                row[f"cos_{k}"] = cos(p.grad, others[k][pname].grad)
            row["rel_bf16e_vs_fp32"] = rel_l2(others["bf16e"][pname].grad, others["fp32"][pname].grad)
            row["floor"] = rel_l2(others["bf16e2"][pname].grad, others["bf16e"][pname].grad)
            rows.append(row)
    return rows


@pytest.mark.parametrize("S,B", [(64, 64), (512, 32)])
@pytest.mark.parametrize("kind", ["D", "G"])
def test_per_layer_gradient_table(S, B, kind):
    """Every conv / convT / BatchNorm parameter gradient of the stepped networks after ONE backward from identical
    weights, at the benchmarked batch (the launch plan -- tile widths, split-K, CTA pairs, role swap -- depends on B)."""
    from oracle.step import synthetic_batch
    tr, oracles = make_all(S, first_iter=0 if kind == "D" else 1)
    A, Bt = synthetic_batch(B, S, step=0, device="cuda")
    assert tr.step(A, Bt) == (kind == "D")
    got = tr.losses()
    logs = {k: st.backward(A, Bt) for k, st in oracles.items()}
    rows = grad_rows(tr, oracles, (2, 3) if kind == "D" else (0, 1))
    report("per_layer_gradient_table", {"S": S, "B": B, "kind": kind, "rows": rows, "losses": got,
                                        "losses_fp32": logs["fp32"], "losses_bf16e": logs["bf16e"],
                                        "losses_bf16e2": logs["bf16e2"]})
    for k in LOSSES:     # forward pass from identical weights
        assert abs(got[k] - logs["bf16e"][k]) <= 0.01 * abs(logs["bf16e"][k]) + 0.003, (k, got[k], logs["bf16e"][k])
        assert abs(got[k] - logs["fp32"][k]) <= 0.05 * abs(logs["fp32"][k]) + 0.02, (k, got[k], logs["fp32"][k])
    bad = [r for r in rows if r["rel_bf16e"] > 1.3 * r["floor"] + 0.02 or r["rel_fp32"] > 1.25 * r["rel_bf16e_vs_fp32"] + 0.02]
    assert not bad, bad[:6]
    tr.close()


# ---------------------------------------------------------------------------------------------------------------------
# 3/4. first steps: losses and accumulated update
# ---------------------------------------------------------------------------------------------------------------------
def run_steps(tr, oracles, S, B, steps, tag, meta):
    from oracle.step import synthetic_batch
    curves = {k: [] for k in ("kernel", *oracles)}
    for it in range(steps):
        A, Bt = synthetic_batch(B, S, step=it, device="cuda")
        was_dis = tr.step(A, Bt)
        curves["kernel"].append(tr.losses())
        for k, st in oracles.items():
            want = st.step(A, Bt)
            assert was_dis == want.pop("is_dis_step")
            curves[k].append(want)
            if S >= 256:                       # tens of GB of cached autograd buffers per oracle: hand them back before the
                torch.cuda.empty_cache()       # trainer captures its next graph (private pool)
    report(tag, {**meta, "curves": curves})
    return curves


def update_rows(tr, oracles, w0):
    rows = []
    for idx, net in enumerate(tr.nets()):
        for pname, p in net.named_parameters():
            if p.dim() != 4:
                continue
            base = w0[idx][pname]
            if float((p.detach() - base).abs().max()) == 0.0:
                continue                                     # never stepped (gan / recongan leave some nets alone)
            q = {k: dict(st.raw_nets[idx].named_parameters())[pname] for k, st in oracles.items()}
            rows.append({"net": NAMES[idx], "param": pname,
                         "rel_bf16e": rel_l2(p - base, q["bf16e"] - base), "cos_bf16e": cos(p - base, q["bf16e"] - base),
                         "rel_fp32": rel_l2(p - base, q["fp32"] - base), "cos_fp32": cos(p - base, q["fp32"] - base),
                         "floor": rel_l2(q["bf16e2"] - base, q["bf16e"] - base) if "bf16e2" in q else None})
    return rows


@pytest.mark.parametrize("S,B,steps,variant,arch", [
    (64, 64, 6, "image_translation", "discogan"),      # BASELINE configs 1/2: two D,G,G cycles
    (64, 64, 3, "angle_pairing", "discogan"),          # config 3
    (64, 64, 3, "image_translation", "recongan"),      # config 5
    (64, 64, 3, "image_translation", "gan"),
    (512, 32, 3, "image_translation", "discogan"),     # config 4: one D,G,G cycle of the reference topology
])
def test_step_parity_at_benchmarked_batch(S, B, steps, variant, arch):
    from oracle.step import build_nets
    # (512x512: trainer + three oracle trainings would need ~170 of the 180 GB; the noise-floor twin is left to the
    # gradient-table test there and the accumulated update gets a direction check instead of the floor comparison)
    tr, oracles = make_all(S, arch=arch, variant=variant, want=("fp32", "bf16e", "bf16e2") if S < 256 else ("fp32", "bf16e"))
    w0 = [{k: v.detach().clone() for k, v in n.named_parameters()} for n in build_nets(S, seed=1234, device="cuda")]
    curves = run_steps(tr, oracles, S, B, steps, "step_parity", {"S": S, "B": B, "variant": variant, "arch": arch})
    for it in range(steps):
        got = curves["kernel"][it]
        loose = S == 512 and it >= 2
        for k in LOSSES:
            e, f = curves["bf16e"][it][k], curves["fp32"][it][k]
            if loose:      # GAN losses one step after a saturated discriminator (gen loss 29 -> 1) are exponentially sensitive
                rel, ab = (0.35, 0.10) if k.startswith(("gen_", "dis_")) else (0.05, 0.02)
                assert abs(got[k] - f) <= rel * abs(f) + ab, (it, k, got[k], f)
            else:      # first D,G,G cycle: tight; afterwards the three trainings are independent runs drifting apart
                re_, ae, rf, af = (0.02, 0.01, 0.05, 0.02) if it < 3 else (0.05, 0.02, 0.08, 0.03)
                assert abs(got[k] - e) <= re_ * abs(e) + ae, (it, k, got[k], e)
                assert abs(got[k] - f) <= rf * abs(f) + af, (it, k, got[k], f)
    rows = update_rows(tr, oracles, w0)
    report("accumulated_update", {"S": S, "B": B, "variant": variant, "arch": arch, "steps": steps, "rows": rows})
    bad = [r for r in rows if (r["rel_bf16e"] > 1.3 * r["floor"] + 0.03 if r["floor"] is not None else r["cos_bf16e"] < 0.5)]
    assert not bad, bad[:6]
    tr.close()
    del tr, oracles, w0
    torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------------------------------------
# 5. loss curves
# ---------------------------------------------------------------------------------------------------------------------
def test_loss_curves_track_over_100_steps():
    """north_star: 'loss curves that track over N steps' -- 100 iterations at 64x64, B=64.  Independent trainings from
    the same weights and batches (kernels, fp32 oracle, bf16-storage oracle).  Stated drift bounds, per loss, on the mean
    over the last 50 steps and on the worst single step relative to what bf16 storage alone shows against fp32 (the GAN
    losses oscillate by 2-5x from step to step, so single steps of independent runs differ by O(1))."""
    S, B, steps = 64, 64, 100
    tr, oracles = make_all(S, want=("fp32", "bf16e"))
    curves = run_steps(tr, oracles, S, B, steps, "loss_curve_100", {"S": S, "B": B})
    summary = {}
    for k in LOSSES:
        ker = torch.tensor([c[k] for c in curves["kernel"]])
        f32 = torch.tensor([c[k] for c in curves["fp32"]])
        emu = torch.tensor([c[k] for c in curves["bf16e"]])
        scale = f32.abs().mean()
        summary[k] = dict(mean_fp32=float(scale),
                          max_kernel_vs_fp32=float((ker - f32).abs().max() / scale),
                          max_bf16e_vs_fp32=float((emu - f32).abs().max() / scale),
                          max_kernel_vs_bf16e=float((ker - emu).abs().max() / scale),
                          tail_kernel_vs_fp32=float((ker[50:].mean() - f32[50:].mean()).abs() / scale),
                          tail_bf16e_vs_fp32=float((emu[50:].mean() - f32[50:].mean()).abs() / scale))
    report("loss_curve_100_summary", {"summary": summary})
    # 50-step means: reconstruction 3 %, discriminator 5 %, feature matching 8 %, generator GAN loss 15 % (it swings between
    # 0.1 and 5 from step to step; measured over runs: <= 1.3 %, 1.9 %, 4.9 %, 8.0 % -- bf16e itself: up to 5.0 %)
    tail_bound = {"recon": 0.03, "dis_l": 0.05, "fm_lo": 0.08, "gen_l": 0.15}
    for k, s in summary.items():
        assert s["tail_kernel_vs_fp32"] <= tail_bound[k[:5]], (k, s)
        # single steps of independent GAN trainings differ by O(1) of the loss's mean (bf16e vs fp32 itself: up to 1.07)
        assert s["max_kernel_vs_fp32"] <= 3.0 * s["max_bf16e_vs_fp32"] + 0.25, (k, s)
    tr.close()


# ---------------------------------------------------------------------------------------------------------------------
# 6. data parallel semantics, one GPU
# ---------------------------------------------------------------------------------------------------------------------
def test_data_parallel_semantics_on_one_gpu():
    """a14 on a single GPU: two trainers act as the two ranks of a data-parallel group (own shard, own BatchNorm
    statistics and FM means), their flat gradients are summed, Adam applies the mean -- compared against the oracle's
    R=2 emulation of the reference DDP step (distributed_image_translation.py:396-404,465-518), fp32 and bf16e."""
    from discogan_modernized_b200 import DiscoGANTrainer, model
    from oracle.bf16_emul import emulate
    from oracle.step import OracleDataParallel, OracleStep, build_nets, synthetic_batch
    S, B, R, steps = 64, 32, 2, 4
    src = build_nets(S, seed=1234, device="cuda")
    ranks = []
    for r in range(R):
        nets = [model.Generator(True, S), model.Generator(True, S), model.Discriminator(S), model.Discriminator(S)]
        for n, o in zip(nets, src):
            n.load_state_dict(o.state_dict())
        ranks.append(DiscoGANTrainer(image_size=S, nets=nets, data_parallel=False, use_graphs=False))
    dp = {"fp32": OracleDataParallel(lambda r: OracleStep(build_nets(S, seed=1234, device="cuda"), device="cuda"), R),
          "bf16e": OracleDataParallel(lambda r: OracleStep(emulate(build_nets(S, seed=1234, device="cuda")), device="cuda"), R)}
    rows, grad_rows_ = [], []
    for it in range(steps):
        shards = [synthetic_batch(B, S, step=it, rank=r, device="cuda") for r in range(R)]
        for tr, (A, Bt) in zip(ranks, shards):
            tr.step(A, Bt, defer_update=True)
        for n0, n1 in zip(ranks[0].stepped_nets(), ranks[1].stepped_nets()):     # the all-reduce(sum)
            g = ranks[0].flat[n0].flat_g + ranks[1].flat[n1].flat_g
            ranks[0].flat[n0].flat_g.copy_(g)
            ranks[1].flat[n1].flat_g.copy_(g)
        stepped_idx = [ranks[0].nets().index(n) for n in ranks[0].stepped_nets()]
        for tr in ranks:
            tr.apply_update(grad_scale=1.0 / R)
        logs = {k: d.step(shards) for k, d in dp.items()}
        if it == 0:     # identical weights: the summed gradient / R equals the oracle's averaged gradient
            for idx in stepped_idx:
                for (pn, p), (_, q), (_, f) in zip(ranks[0].nets()[idx].named_parameters(),
                                                   dp["bf16e"].replicas[0].nets()[idx].inner.named_parameters(),
                                                   dp["fp32"].replicas[0].nets()[idx].named_parameters()):
                    mean_grad = p.grad / R
                    grad_rows_.append({"net": NAMES[idx], "param": pn, "rel_bf16e": rel_l2(mean_grad, q.grad),
                                       "rel_fp32": rel_l2(mean_grad, f.grad), "rel_bf16e_vs_fp32": rel_l2(q.grad, f.grad),
                                       "norm_ratio": float(mean_grad.norm() / q.grad.norm())})
        for r, tr in enumerate(ranks):
            got = tr.losses()
            for k in LOSSES:
                e, f = logs["bf16e"][r][k], logs["fp32"][r][k]
                rows.append({"it": it, "rank": r, "loss": k, "kernel": got[k], "bf16e": e, "fp32": f})
                re_, ae, rf, af = (0.02, 0.01, 0.05, 0.02) if it < 3 else (0.05, 0.02, 0.08, 0.03)
                assert abs(got[k] - e) <= re_ * abs(e) + ae, (it, r, k, got[k], e)
                assert abs(got[k] - f) <= rf * abs(f) + af, (it, r, k, got[k], f)
    report("dp_semantics_one_gpu", {"rows": rows, "grad_rows": grad_rows_})
    for g in grad_rows_:
        assert 0.95 < g["norm_ratio"] < 1.05, g                                  # a sum instead of a mean would read 2.0
        assert g["rel_fp32"] <= 1.25 * g["rel_bf16e_vs_fp32"] + 0.02, g
    # ranks stay weight-synchronised bit for bit, BatchNorm running statistics stay per rank and follow their own replica
    for a, b in zip(ranks[0].nets(), ranks[1].nets()):
        assert torch.equal(ranks[0].flat[a].flat_p, ranks[1].flat[b].flat_p)
    assert not torch.equal(ranks[0].G_A.encoder[3].running_mean, ranks[1].G_A.encoder[3].running_mean)
    for r in range(R):
        a = ranks[r].G_B.encoder[3].running_mean
        b = dp["bf16e"].replicas[r].G_B.inner.encoder[3].running_mean
        assert torch.allclose(a, b, rtol=2e-2, atol=3e-3), float((a - b).abs().max())   # 4 independent noisy steps apart
    for tr in ranks:
        tr.close()


def test_two_trainers_interleaved_in_one_process():
    """De-globalised launch state: two trainers (own contexts, scratch, split-K workspaces, CUDA graphs) stepping
    alternately in one process each follow the oracle exactly as a lone trainer does."""
    from oracle.step import synthetic_batch
    S, B = 64, 16
    tr1, o1 = make_all(S, want=("fp32",))
    tr2, o2 = make_all(S, want=("fp32",), arch="gan")
    assert tr1.ctx is not tr2.ctx
    for it in range(6):
        A, Bt = synthetic_batch(B, S, step=it, device="cuda")
        tr1.step(A, Bt)
        tr2.step(Bt, A)
        w1, w2 = o1["fp32"].step(A, Bt), o2["fp32"].step(Bt, A)
        g1, g2 = tr1.losses(), tr2.losses()
        rel, ab = (0.05, 0.02) if it < 2 else ((0.08, 0.03) if it == 2 else (0.12, 0.04))   # small batch: noisier than B=64
        for k in LOSSES:
            assert abs(g1[k] - w1[k]) <= rel * abs(w1[k]) + ab, (it, k, g1[k], w1[k])
            assert abs(g2[k] - w2[k]) <= rel * abs(w2[k]) + ab, (it, k, g2[k], w2[k])
    tr1.close()
    tr2.close()
