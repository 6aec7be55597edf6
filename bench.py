#!/usr/bin/env python
"""bench.py -- DiscoGAN train-step throughput (image-pairs/s) on B200, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--image-size 64] [--batch 64]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

A step is one iteration of the reference loop (image_translation.py:335-390) on one synthetic A/B batch pair:
D step when iter % 3 == 0, else G step; exactly K steps are timed (the D:G:G schedule simply continues across them)
after exactly W warm-up steps.  Before the warm-up the trainer is primed for two D:G:G cycles (`config.prime_steps`):
the first cycle runs eagerly and sizes the scratch buffers, the second captures the step graphs -- set-up, like building
the model, not warm-up.  `value` is whole-job image-pairs/s with the batches already resident in HBM; `e2e` is the same
loop fed from pinned host buffers with the H2D copy of both batches and a D2H read of the eight losses inside the timed
region every step.  `also` repeats the measurement at 512x512, B=32 per GPU (BASELINE config 4) at every N;
`inference` is the eval-mode AtoB generator sweep over batch sizes (config 5, N=1).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--image-size", type=int, default=64)
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default 64 at 64^2, 32 at 512^2)")
    ap.add_argument("--model-arch", default="discogan")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--also-512", type=int, default=1, help="also report a 512^2 B=32-per-GPU measurement (every N)")
    ap.add_argument("--also-steps", type=int, default=12)
    ap.add_argument("--no-inference", action="store_true", help="skip the eval-mode AtoB batch sweep")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the image-file input-pipeline measurement")
    return ap.parse_args()


def run_config(S, B, world, arch):
    """The `config` object -- identical keys for the B200 arm and the reference arm."""
    return {"workload": workload_name(S, B), "image_size": S, "batch_per_gpu": B, "global_batch": B * world,
            "model_arch": arch, "parallelism": f"dp{world}"}


def workload_name(S, B):
    names = {64: "celebA Male->Smiling discogan 64x64 (synthetic faces)", 512: "tops2hanbok discogan 512x512 (synthetic)"}
    return f"{names.get(S, f'discogan {S}x{S}')} batch {B} per GPU, D:G:G schedule, Adam"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ---------------------------------------------------------------------------------------------------
# work accounting (SURVEY.md 8-A0 / BASELINE.md section 2)
# ---------------------------------------------------------------------------------------------------
def conv_flops(B, Hs, Cs, Cb):
    return 2.0 * B * Hs * Hs * Cs * Cb * 16


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.monotonic()

    def mark_end(self):
        self.t1 = time.monotonic()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # the sampler runs from before the warm-up; only samples that arrived inside the marked window (the timed
        # steps and the host-fed timed steps that follow them) are reported
        lines = [ln for t, ln in self.lines if self.t0 is None or (self.t0 <= t <= (self.t1 or t))]
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# reference arm: the restated reference step (oracle port) on the host CPU cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference_pairs_per_s(S, B, steps, warmup, arch, budget_s=200.0):
    import torch
    from oracle.step import OracleStep, build_nets, synthetic_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    st = OracleStep(build_nets(S, seed=1234), model_arch=arch)
    A, Bt = synthetic_batch(B, S, step=0)
    t0 = time.perf_counter()
    st.step(A, Bt)                                   # first (untimed) step, also calibrates the sample size
    t1 = time.perf_counter() - t0
    Bs = B
    if t1 * (steps + warmup) > budget_s:              # bound the sample: shrink the per-step batch
        Bs = max(2, int(B * budget_s / (t1 * (steps + warmup))))
        Bs = min(B, max(2, Bs))
    for i in range(max(0, warmup - 1)):
        A, Bt = synthetic_batch(Bs, S, step=i + 1)
        st.step(A, Bt)
    batches = [synthetic_batch(Bs, S, step=100 + i) for i in range(min(steps, 3))]
    t0 = time.perf_counter()
    for i in range(steps):
        A, Bt = batches[i % len(batches)]
        st.step(A, Bt)
    dt = time.perf_counter() - t0
    sample = f"{steps} steps (D:G:G) of the oracle port at {S}x{S}, batch {Bs} per step, fp32, {cores} threads"
    return Bs * steps / dt, dt / steps * 1e3, cores, sample, Bs


def run_reference(args, S, B):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)                 # exactly K steps (the D:G:G schedule simply continues across them)
    v, ms, cores, sample, Bs = cpu_reference_pairs_per_s(S, B, steps, max(1, args.warmup), args.model_arch)
    line = {
        "impl": "reference", "metric": "train image-pairs/sec", "value": v, "unit": "image-pairs/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": max(1, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": run_config(S, B, max(1, args.gpus), args.model_arch),
        "details": {"sample_batch": Bs, "note": "CPU arm: rank 0 only, one process on all host cores"},
        "cpu_baseline": {"value": v, "unit": "image-pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "image-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
def timed_steps(tr, batches, steps, torch, dist, world):
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    start.record()
    for i in range(steps):
        A, B = batches[i % len(batches)]
        tr.step(A, B)
    end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = start.elapsed_time(end)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    return ms


def timed_steps_e2e(tr, host_batches, steps, torch, dist, world):
    """Public-API loop fed from pinned host memory.  Every step's two batches are copied host->device inside the timed
    region (double-buffered on a copy stream, so the copy of step i+1 overlaps the kernels of step i) and every step's
    eight losses are read back device->host (asynchronously, consumed one step later, as a logging loop would)."""
    dev = [(torch.empty_like(host_batches[0][0], device="cuda"), torch.empty_like(host_batches[0][1], device="cuda"))
           for _ in range(2)]
    host_loss = [torch.empty(8, dtype=torch.float32).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]
    main = torch.cuda.current_stream()
    seen = 0.0

    def upload(i):
        slot = i % 2
        hA, hB = host_batches[i % len(host_batches)]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])           # the step that last read this buffer is done with it
            dev[slot][0].copy_(hA, non_blocking=True)
            dev[slot][1].copy_(hB, non_blocking=True)
            copied[slot].record(copy_stream)

    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for ev in consumed:
        ev.record(main)
    start.record()
    upload(0)
    for i in range(steps):
        slot = i % 2
        if i + 1 < steps:
            upload(i + 1)
        main.wait_event(copied[slot])
        tr.step(dev[slot][0], dev[slot][1])
        consumed[slot].record(main)
        host_loss[slot].copy_(tr.loss_buf, non_blocking=True)
        loss_ready[slot].record(main)
        if i > 0:                                            # consume the previous step's losses on the host
            loss_ready[1 - slot].synchronize()
            seen += float(host_loss[1 - slot][0])
    loss_ready[(steps - 1) % 2].synchronize()
    seen += float(host_loss[(steps - 1) % 2][0])
    end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = start.elapsed_time(end)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    assert seen == seen, "losses are NaN"
    h2d = 2 * host_batches[0][0].numel() * 4
    return ms, h2d, 32


def measure_roofline(tr, batches, torch, pk):
    """Per-launch CUDA-event timing over one D,G,G cycle, eager launches on the trainer's stream (separate from the
    throughput region so the events do not perturb it).  Tensor-bound: the tcgen05 convolution kernels, algorithmic
    FLOPs 2*B*Ho*Wo*Co*Ci*16 per launch.  HBM-bound: BatchNorm(+activation) forward/backward and Adam, algorithmic
    bytes per DESIGN.md section 3.4."""
    from discogan_modernized_b200 import ops
    recs, calls = [], []

    def conv_work(name, a):
        if name in ("conv_down", "conv_down_stats", "conv_down_acc"):
            big, wd = a[0], a[1]
            return "conv_gemm", conv_flops(big.shape[0], big.shape[1] // 2, wd.shape[0], big.shape[3]), \
                f"down B{big.shape[0]} {big.shape[1]}->{big.shape[1] // 2} {big.shape[3]}->{wd.shape[0]}"
        if name in ("conv_up", "conv_up_stats", "conv_up_acc"):
            small, wu = a[0], a[1]
            return "conv_gemm", conv_flops(small.shape[0], small.shape[1], small.shape[3], wu.shape[0]), \
                f"up B{small.shape[0]} {small.shape[1]}->{2 * small.shape[1]} {small.shape[3]}->{wu.shape[0]}"
        small, big = a[0], a[1]
        return "conv_wgrad", conv_flops(small.shape[0], small.shape[1], small.shape[3], big.shape[3]), \
            f"wgrad B{small.shape[0]} {small.shape[1]} {small.shape[3]}x{big.shape[3]}"

    def hbm_work(name, a):
        if name in ("bn_act_fwd", "bn_act_fwd_acc"):
            return "bn_act_fwd", 4.0 * a[0].numel(), f"bn_fwd {a[0].shape[0]}x{a[0].shape[1]}"   # read z, write y (bf16)
        if name == "bn_act_bwd":                                   # reduce: dy,z ; dx: dy,z + write dz (bf16)
            return "bn_act_bwd", 10.0 * a[0].numel(), f"bn_bwd {a[0].shape[0]}x{a[0].shape[1]}"
        return "adam", 28.0 * a[0].numel(), ""                     # p,g,m,v read + p,m,v write (fp32); +repack excluded

    names = {n: conv_work for n in ("conv_down", "conv_up", "conv_wgrad", "conv_down_stats", "conv_up_stats",
                                    "conv_down_acc", "conv_up_acc")}
    names.update({n: hbm_work for n in ("bn_act_fwd", "bn_act_fwd_acc", "bn_act_bwd", "adam_step")})
    orig = {n: getattr(ops, n) for n in names}

    def wrap(name, fn, work):
        def inner(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = fn(*a, **k)
            e.record()
            cls, amount, label = work(name, a)
            recs.append((cls, amount, label, s, e))
            if cls.startswith("conv"):
                calls.append((cls, amount, fn, a, k))      # arguments kept alive for the back-to-back replay below
            return out
        return inner

    for n, f in orig.items():
        setattr(ops, n, wrap(n, f, names[n]))
    saved_graphs, saved_side = tr.use_graphs, tr._side  # (_streams() is empty when _side is None)
    saved_buckets, saved_cb = tr.adam_buckets, tr.reducer.on_bucket
    tr.use_graphs = False                      # per-launch events need eager launches ...
    tr._side = None                            # ... on one stream (no concurrent lane sharing the SMs)
    tr.adam_buckets, tr.reducer.on_bucket = False, None    # ... and no per-bucket Adam running beside the timed kernels
    try:
        for cycle in range(2):                 # first cycle: untimed, lets the caching allocator grow its eager pool
            recs.clear()                       # (cudaMalloc stalls between the events would be charged to kernels)
            calls.clear()
            for i in range(3):
                A, B = batches[i % len(batches)]
                tr.step(A, B)
            torch.cuda.synchronize()
    finally:
        tr.use_graphs, tr._side = saved_graphs, saved_side
        tr.adam_buckets, tr.reducer.on_bucket = saved_buckets, saved_cb
        for n, f in orig.items():
            setattr(ops, n, f)
    # Eager per-launch events include host launch gaps whenever the GPU is faster than Python (always at 64x64).  So the
    # tensor-core kernels are timed a second time: exactly the cycle's launches of one kernel, same arguments, captured
    # into a CUDA graph and replayed back to back between two events (no host gaps, nothing else on the GPU).
    replay = {}
    for cls in ("conv_gemm", "conv_wgrad"):
        mine = [c for c in calls if c[0] == cls]
        if not mine:
            continue
        try:
            g = torch.cuda.CUDAGraph()
            tr.ctx.arena.reset()                  # the replayed launches draw their statistics accumulators afresh
            torch.cuda.synchronize()
            with ops.use_context(tr.ctx), torch.cuda.graph(g):      # the trainer's launch plan (split-K workspaces, lanes)
                keep = [fn(*a, **k) for _, _, fn, a, k in mine]
            g.replay()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(3):
                g.replay()
            e.record()
            torch.cuda.synchronize()
            replay[cls] = (sum(c[1] for c in mine), s.elapsed_time(e) * 1e-3 / 3, len(mine))
            del g, keep
        except Exception as ex:  # noqa: BLE001
            replay[cls] = None
            print(f"bench: graph replay timing of {cls} failed: {ex}", file=sys.stderr)
    calls.clear()
    by, layers = {}, {}
    for cls, amount, label, s, e in recs:
        sec = s.elapsed_time(e) * 1e-3
        d = by.setdefault(cls, [0.0, 0.0, 0])
        d[0] += amount; d[1] += sec; d[2] += 1
        if label:
            l = layers.setdefault(label, [0.0, 0.0, 0])
            l[0] += amount; l[1] += sec; l[2] += 1
    out = {}
    for cls, (amount, sec, n) in by.items():
        if cls.startswith("conv"):
            out[cls] = {"launches": n, "tflops": amount / sec / 1e12, "avg_us": sec / n * 1e6, "ms_per_cycle": sec * 1e3,
                        "frac_of_tensor_peak": amount / sec / 1e12 / pk["tf_sustained"]}
        else:
            out[cls] = {"launches": n, "gbs": amount / sec / 1e9, "avg_us": sec / n * 1e6, "ms_per_cycle": sec * 1e3,
                        "frac_of_hbm_peak": amount / sec / 1e9 / pk["hbm"]}
    top = sorted(layers.items(), key=lambda kv: -kv[1][1])[:16]
    out["layers"] = {k: ({"launches": n, "gbs": round(a / sec / 1e9, 1), "ms_per_cycle": round(sec * 1e3, 3)}
                         if k.startswith("bn_") else
                         {"launches": n, "tflops": round(a / sec / 1e12, 1), "ms_per_cycle": round(sec * 1e3, 3)})
                     for k, (a, sec, n) in top}
    for cls, r in replay.items():
        if r:
            out[cls]["replay_tflops"] = r[0] / r[1] / 1e12
            out[cls]["replay_avg_us"] = r[1] / r[2] * 1e6
    g = by.get("conv_gemm", [0.0, 1.0, 1])
    timing = "per-launch CUDA events, eager launches"
    peak, peak_src = pk["tf_sustained"], f"{pk['src']} bf16 sustained (kernel timed inside the step)"
    if replay.get("conv_gemm"):
        g = list(replay["conv_gemm"])
        timing = "CUDA events around a back-to-back graph replay of the cycle's launches of this kernel"
        if 3 * g[1] < 0.02:      # a burst of a few milliseconds: compare with the burst figure
            peak, peak_src = pk["tf_burst"], f"{pk['src']} bf16 burst (kernel timed alone, {3 * g[1] * 1e3:.0f} ms of launches)"
    ach = g[0] / g[1] / 1e12
    # DRAM traffic per launch from the committed `ncu --set full` captures of this kernel (profiles/README.md)
    traffic = None
    tfile = ROOT / "profiles" / "ncu_traffic.json"
    big = batches[0][0]
    if tfile.exists():
        key = {(512, 32): "conv_gemm_kernel", (64, 64): "conv_gemm_kernel_64x64_b64"}.get((big.shape[-1], big.shape[0]))
        ent = json.loads(tfile.read_text()).get(key) if key else None
        traffic = ent["dram_bytes_per_launch"] if ent else None
    roof = {"bound": "tensor", "kernel": "conv_gemm_kernel<1|2> + conv_gemm_swap_kernel (tcgen05 implicit GEMM: conv fprop/dgrad, convT fprop/dgrad)",
            "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "peak_source": peak_src, "traffic": traffic,
            "launches_per_cycle": g[2], "avg_launch_us": g[1] / max(g[2], 1) * 1e6,
            "algorithmic": "2*B*Ho*Wo*Co*Ci*16 FLOP per launch", "timing": timing}
    return roof, out


def run_b200(args, S, B):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from discogan_modernized_b200 import DiscoGANTrainer, _lib

    def synthetic_batch(batch, image_size, step=0, rank=0, device="cpu"):
        """A, B = uniform [0,1) fp32 images from seed 1000*rank+step; same recipe as the oracle's
        helper, restated here so that nothing under oracle/ is imported on this arm."""
        g = torch.Generator().manual_seed(1000 * rank + step)
        A = torch.rand(batch, 3, image_size, image_size, generator=g)
        Bt = torch.rand(batch, 3, image_size, image_size, generator=g)
        return A.to(device), Bt.to(device)

    PRIME = 6      # two D:G:G cycles of set-up: eager (buffer sizing, kernel attributes), then graph capture + first replay
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    pk = peaks()

    def measure(size, bsz, n_steps, n_warmup, with_clocks):
        """prime -> W warm-up steps -> K timed steps (device-resident batches) -> K timed steps fed from the host."""
        tr = DiscoGANTrainer(image_size=size, device=f"cuda:{local}", model_arch=args.model_arch, seed=1234)
        host = [tuple(t.pin_memory() for t in synthetic_batch(bsz, size, step=i, rank=rank)) for i in range(3)]
        batches = [(a.cuda(non_blocking=True), b.cuda(non_blocking=True)) for a, b in host]
        sampler = ClockSampler(local) if (with_clocks and rank == 0) else None
        if sampler:
            sampler.start()                     # nvidia-smi needs ~0.1 s to deliver its first sample: start it early
        for i in range(PRIME + n_warmup):
            tr.step(*batches[i % 3])
        torch.cuda.synchronize()
        L0 = tr.kernel_launches
        if sampler:
            sampler.mark_begin()
        ms = timed_steps(tr, batches, n_steps, torch, dist, world)
        launches = tr.kernel_launches - L0      # kernels of libdiscogan_b200.so (eager launches + graph-replayed nodes)
        ms_e2e, h2d, d2h = timed_steps_e2e(tr, host, n_steps, torch, dist, world)
        clocks = None
        if sampler:
            sampler.mark_end()
            clocks = sampler.stop()
        pairs = bsz * world * n_steps
        res = {"value": pairs / (ms * 1e-3), "unit": "image-pairs/s", "ms_per_step": ms / n_steps, "steps": n_steps,
               "warmup": n_warmup,
               "e2e": {"value": pairs / (ms_e2e * 1e-3), "unit": "image-pairs/s", "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / n_steps},
               "gpu_launches": int(launches)}
        lanes = 1 + (1 + len(tr._more) if tr._side is not None else 0)
        return tr, batches, res, clocks, lanes

    tr, batches, res, clocks, lanes = measure(S, B, steps, warmup, True)
    line = {
        "metric": "train image-pairs/sec", "value": res["value"], "unit": "image-pairs/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": run_config(S, B, world, args.model_arch),
        "details": {"prime_steps": PRIME, "cuda_graphs": bool(tr.use_graphs), "lanes": lanes,
                    "grad_buckets": {n: len(getattr(tr.flat[net], "_buckets", []) or []) for n, net in
                                     zip(("G_A", "G_B", "D_A", "D_B"), tr.nets())} if world > 1 else None,
                    "l2": "per-step working set (weights+grads+Adam moments+activations) exceeds the 126 MB L2; no explicit flush",
                    "library": {"version": _lib.lib().dg_version(), "sources": _lib.lib().dg_source_hash().decode()}},
        "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_roofline:
        roof, per_kernel = measure_roofline(tr, batches, torch, pk)
        line["roofline"] = roof
        line["kernels"] = per_kernel
    tr.close()
    del tr, batches
    torch.cuda.empty_cache()
    if args.also_512 and S != 512:
        # BASELINE config 4 (512x512, B=32 per GPU) at the same N: every rank takes part (the gradient exchange is
        # 1.8 GB per G step there), rank 0 reports
        try:
            tr5, b5, res5, _, _ = measure(512, 32, max(1, args.also_steps), 6, False)
            also = {"workload": workload_name(512, 32), **res5}
            if rank == 0 and world == 1 and not args.no_roofline:
                also["roofline"], also["kernels"] = measure_roofline(tr5, b5, torch, pk)
            tr5.close()
            del tr5, b5
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            also = {"workload": workload_name(512, 32), "error": str(e)[:300]}
        line["also"] = also
    if rank == 0 and world == 1:
        if not args.no_inference:
            try:
                line["inference"] = inference_sweep(torch, local)
            except Exception as e:  # noqa: BLE001
                line["inference"] = {"error": str(e)[:300]}
        if not args.no_pipeline:
            try:
                line["pipeline"] = pipeline_e2e(torch, S, B, min(steps, 150), local, args.model_arch, DiscoGANTrainer)
            except Exception as e:  # noqa: BLE001
                line["pipeline"] = {"error": str(e)[:300]}
        if not args.no_cpu_baseline:
            v, msc, cores, sample, _ = cpu_reference_pairs_per_s(S, B, 3, 1, args.model_arch, budget_s=60.0)
            line["cpu_baseline"] = {"value": v, "unit": "image-pairs/s", "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # skip destroy_process_group (an NCCL communicator that was captured into CUDA graphs can block in teardown) --
        # exiting the process is the documented alternative
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


def pipeline_e2e(torch, S, B, steps, local, arch, DiscoGANTrainer):
    """The train loop fed from IMAGE FILES through the package's own loader (dataset.DiscoGANDataset): synthetic
    CelebA-sized JPEGs are written to a temp folder, decoded once into HBM, and every step's batch pair is produced by the
    preprocessing kernel (read_images arithmetic) right before the step.  `value` = image-pairs/s over `steps` steps with
    a loss readback per step, resident mode; `streaming` = the same loop when the decoded images are NOT cached (host
    decode per batch), for contrast."""
    import tempfile
    import numpy as np
    from PIL import Image
    from discogan_modernized_b200 import dataset
    n_img = 1024
    with tempfile.TemporaryDirectory() as tmp:
        rng = np.random.default_rng(0)
        base = np.linspace(0, 255, 218 * 178 * 3, dtype=np.float32).reshape(218, 178, 3)
        for d in ("A", "B"):
            os.makedirs(os.path.join(tmp, d))
            for i in range(n_img):
                img = np.clip(np.roll(base, int(rng.integers(0, 178)), 1) + rng.normal(0, 12, base.shape), 0, 255).astype(np.uint8)
                Image.fromarray(img).save(os.path.join(tmp, d, f"{i:05d}.jpg"), quality=90)
        fa, fb = dataset.list_images(os.path.join(tmp, "A")), dataset.list_images(os.path.join(tmp, "B"))
        out = {"unit": "image-pairs/s", "images_per_domain": n_img, "image_hw": [218, 178], "format": "jpeg"}
        for mode in ("resident", "streaming"):
            ds = dataset.DiscoGANDataset(fa, fb, None, None, S, device=f"cuda:{local}", cache_bytes=None if mode == "resident" else 0)
            t0 = time.perf_counter()
            ds.load()
            torch.cuda.synchronize()
            load_s = time.perf_counter() - t0
            tr = DiscoGANTrainer(image_size=S, device=f"cuda:{local}", model_arch=arch, seed=1234, data_parallel=False)

            def stream():
                epoch = 0
                while True:
                    yield from ds.batches(B, epoch=epoch, drop_last=True)
                    epoch += 1
            it = stream()
            n_steps = steps if mode == "resident" else min(steps, 24)
            for _ in range(12):
                tr.step(*next(it))
            host_loss = torch.empty(8, dtype=torch.float32).pin_memory()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            s.record()
            for _ in range(n_steps):
                tr.step(*next(it))
                host_loss.copy_(tr.loss_buf, non_blocking=True)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e)
            if mode == "resident":
                out.update({"value": B * n_steps / (ms * 1e-3), "ms_per_step": ms / n_steps, "steps": n_steps,
                            "mode": "decoded images resident in HBM; per step: index gather + dg_preprocess_u8 x2",
                            "one_time_decode_and_upload_s": round(load_s, 3),
                            "resident_bytes": int(ds._stores[0].bytes + ds._stores[1].bytes)})
            else:
                out["streaming"] = {"value": B * n_steps / (ms * 1e-3), "ms_per_step": ms / n_steps, "steps": n_steps,
                                    "mode": f"host decode per batch on {ds.workers} threads, prefetch 3"}
            tr.close()
            del tr, ds
            torch.cuda.empty_cache()
    return out


def inference_sweep(torch, local):
    """BASELINE config 5: eval-mode AtoB generator forward (inference.py:149,168-172 batched) over batch sizes, images
    already on the device; CUDA events around `reps` back-to-back forwards after 3 warm-up calls."""
    from discogan_modernized_b200.inference import GraphedGenerator
    from discogan_modernized_b200.model import Generator
    out = {"unit": "images/s", "direction": "AtoB", "sizes": {},
           "mode": "eval (BatchNorm running statistics), CUDA-graph replay per batch size (inference.GraphedGenerator)"}
    for size, batches in ((64, (1, 2, 4, 8, 16, 32, 64, 128, 256)), (512, (1, 2, 4, 8, 16, 32, 64, 128, 256))):
        torch.manual_seed(1234)
        g = GraphedGenerator(Generator(extra_layers=True, image_size=size).to(f"cuda:{local}").eval())
        rows = {}
        with torch.no_grad():
            for bsz in batches:
                x = torch.rand(bsz, 3, size, size, device=f"cuda:{local}")
                for _ in range(3):
                    g(x)
                reps = max(3, min(50, int(2e4 / (bsz * (size / 64) ** 2))))
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                s.record()
                for _ in range(reps):
                    g(x)
                e.record()
                torch.cuda.synchronize()
                ms = s.elapsed_time(e) / reps
                rows[str(bsz)] = {"images_per_s": round(bsz / (ms * 1e-3), 1), "ms_per_batch": round(ms, 4)}
        out["sizes"][str(size)] = rows
        del g
        torch.cuda.empty_cache()
    return out


def main():
    args = parse_args()
    S = args.image_size
    B = args.batch or (32 if S == 512 else 64)
    if args.impl == "reference":
        run_reference(args, S, B)
    else:
        run_b200(args, S, B)


if __name__ == "__main__":
    main()
