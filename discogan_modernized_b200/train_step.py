"""The DiscoGAN train step on the B200 kernels (single GPU and data parallel).

Implements the body of the reference loop -- ``image_translation.py:335-390`` (single device),
``distributed_image_translation.py:465-518`` (DDP) and the ``angle_pairing.py:291-346`` variant --
as one hand-scheduled forward/backward over the kernel engines in ``model.py``:

* all eight forwards run every iteration (BatchNorm running statistics advance exactly as in the
  reference), but only the gradients that reach an optimiser step are computed: a D step back-props
  the two discriminators only, a G step back-props the generators and the *data* path of the
  discriminators' fake passes (the reference computes and discards the rest);
* GAN-BCE, reconstruction-MSE and feature-matching losses are fused kernels; loss weights are host
  constants so no autograd graph is built;
* parameters, gradients and Adam moments of each network live in flat fp32 buffers: Adam is one
  kernel per network; the data-parallel exchange is bucketed like the reference's DDP reducer
  (``distributed_image_translation.py:401-404,513-518``): the flat gradient of a stepped network is cut
  into contiguous buckets in backward-completion order and each bucket's NCCL all-reduce is launched on
  a side stream the moment the kernels writing its last gradient are enqueued, overlapping the rest of
  the backward pass (``GradReducer``); gradients are averaged over ranks, BatchNorm statistics stay per
  rank, the learning rate is not scaled (``distributed_image_translation.py:396-427`` semantics with
  the ``broadcast_buffers`` crash of SURVEY.md F4 avoided);
* the whole iteration (several hundred kernel launches, micro-seconds each at 64x64) is captured once
  per (step kind, loss weights) into a CUDA graph and replayed: inputs are copied into static
  buffers, the Adam step counter lives on the device.
"""
import contextlib
import os

import torch
import torch.distributed as dist

from . import ops
from .model import (Discriminator, Generator, discriminator_backward, discriminator_forward, generator_backward,
                    generator_forward)

LOSS_NAMES = ("dis_loss_A", "gen_loss_A", "dis_loss_B", "gen_loss_B", "fm_loss_A", "fm_loss_B", "recon_loss_A",
              "recon_loss_B")
_ALIGN = 64  # floats; keeps every parameter view 256-byte aligned


class FlatNet:
    """Flat fp32 parameter / gradient / Adam-moment buffers of one network; ``p.data`` and ``p.grad``
    of every parameter become views into them (registration order, ``image_translation.py:272-273``)."""

    def __init__(self, net):
        self.net = net
        params = list(net.parameters())
        dev = params[0].device
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = total
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.adam_state = torch.zeros(4, dtype=torch.float32, device=dev)   # {steps, 1-b1^t, sqrt(1-b2^t), -}
        for p, o in zip(params, offs):
            view = self.flat_p[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_g[o:o + p.numel()].view(p.shape)
        self.params, self.offsets = params, offs
        # one int64 buffer for every BatchNorm counter of the network (one add per forward instead of one per layer)
        bns = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d) and m.track_running_stats]
        if bns:
            flat = torch.stack([m.num_batches_tracked.to(dev) for m in bns]).contiguous()
            for i, m in enumerate(bns):
                m._buffers["num_batches_tracked"] = flat[i]
            net._nbt_flat = flat
        net._packed.invalidate()

    def zero_grad(self):
        """No memset: every parameter's gradient is marked fresh, so the first kernel that writes it this iteration
        overwrites (beta = 0) and later passes accumulate.  Every parameter of a stepped network is written by each
        of its backward passes; the alignment gaps between parameters stay zero for ever."""
        for p, o in zip(self.params, self.offsets):  # re-attach in case someone set .grad = None
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * o:
                p.grad = self.flat_g[o:o + p.numel()].view(p.shape)
            p._dg_fresh = True

    def adam(self, lr, beta1, beta2, eps, weight_decay, grad_scale):
        """Adam over the flat buffers, then refresh the bf16 GEMM copies of the conv weights in place."""
        ops.adam_step(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, lr, beta1, beta2, eps, weight_decay,
                      self.adam_state, grad_scale)
        self.net._packed.refresh()

    def adam_tick(self, beta1, beta2):
        """Advance the device-side step state once; ``adam_range`` calls then update pieces of the buffers."""
        ops.adam_tick(self.adam_state, beta1, beta2)

    def adam_range(self, lo, hi, lr, beta1, beta2, eps, weight_decay, grad_scale):
        ops.adam_apply(self.flat_p[lo:hi], self.flat_g[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi], lr, beta1, beta2,
                       eps, weight_decay, self.adam_state, grad_scale)

    @property
    def steps(self):
        return int(self.adam_state[0].item())


def plan_buckets(sizes, cap_elems):
    """Partition parameters (given by their padded element counts in registration order) into contiguous buckets in
    BACKWARD-completion order, i.e. walking the registration order from the end (the last-registered layers finish
    their backward first).  A bucket closes once it holds >= cap_elems.  Returns [(first_param, last_param)] index
    pairs (inclusive), first-finishing bucket first."""
    buckets, hi, acc = [], len(sizes) - 1, 0
    for i in range(len(sizes) - 1, -1, -1):
        acc += sizes[i]
        if acc >= cap_elems or i == 0:
            buckets.append((i, hi))
            hi, acc = i - 1, 0
    return buckets


class GradReducer:
    """Data-parallel gradient exchange (reference: the DDP reducer, distributed_image_translation.py:396-404,513-518).

    Each stepped network's flat fp32 gradient is cut into contiguous buckets in backward-completion order
    (``plan_buckets``; cap = min(25 MiB -- DDP's default --, a third of the network)).  The backward pass reports every
    layer's parameters as soon as the kernels writing their final gradients are enqueued (``hook``); when a bucket is
    complete its NCCL all-reduce (sum; the 1/world factor is folded into Adam) is launched on a side stream ordered
    after exactly those kernels, so the exchange of the deep layers overlaps the backward of the shallow ones -- inside
    the captured CUDA graph as well.  ``join`` makes the optimiser wait for all buckets.  The two lanes of the step use
    two communicators, so one network's last bucket never queues behind the other's first.
    CPU tensors (gloo, tests): blocking all-reduce per bucket."""

    DDP_BUCKET_BYTES = 25 << 20

    def __init__(self, group=None, enabled=True, bucket_bytes=None):
        self.enabled = enabled and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.group = group
        self.world = dist.get_world_size(group) if self.enabled else 1
        env = os.environ.get("DISCOGAN_B200_BUCKET_MB")
        self.bucket_bytes = bucket_bytes if bucket_bytes is not None else (int(float(env) * (1 << 20)) if env else None)
        self._streams = {}       # comm slot -> side stream
        self._groups = {0: group}
        self._pending = []
        self._state = {}         # FlatNet -> dict(buckets, left, passes, slot)
        self.launched = []       # (net numel, lo, hi) of every all-reduce issued (tests / accounting)
        # on_bucket(flat, lo, hi): called on the bucket's stream right behind its all-reduce (the trainer hangs the Adam
        # update of that range here); with it set, buckets are tracked even without a process group (world 1)
        self.on_bucket = None

    @property
    def active(self):
        return self.enabled or self.on_bucket is not None

    # -- planning -------------------------------------------------------------------------------------
    def _plan(self, flat):
        plan = getattr(flat, "_buckets", None)
        if plan is None:
            sizes = [(p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN for p in flat.params]
            cap = self.bucket_bytes if self.bucket_bytes else min(self.DDP_BUCKET_BYTES, -(-flat.numel * 4 // 3))
            plan = []
            for lo, hi in plan_buckets(sizes, max(1, cap // 4)):
                plan.append((flat.offsets[lo], flat.offsets[hi] + sizes[hi], [id(p) for p in flat.params[lo:hi + 1]]))
            flat._buckets = plan
        return plan

    def _group_for(self, slot, device_is_cuda):
        """Communicator of a lane slot: slot 0 = the trainer's group, slot 1 = a second communicator over the same ranks
        (created collectively on first use; every rank reaches this point in the same order)."""
        if slot not in self._groups:
            if not device_is_cuda or os.environ.get("DISCOGAN_B200_DP_COMMS", "2") == "1":
                self._groups[slot] = self.group
            else:
                ranks = dist.get_process_group_ranks(self.group) if self.group is not None else None
                self._groups[slot] = dist.new_group(ranks=ranks)
        return self._groups[slot]

    def prepare(self, flat_nets):
        """Create the communicators and bucket plans outside the step (never during CUDA-graph capture)."""
        if self.active and not self.enabled:
            for fn in flat_nets:
                self._plan(fn)
        if not self.enabled:
            return
        for i, fn in enumerate(flat_nets):
            self._plan(fn)
        if any(fn.flat_g.is_cuda for fn in flat_nets):
            for slot in (0, 1):
                self._group_for(slot, True)

    # -- per-step protocol ----------------------------------------------------------------------------
    def begin(self, flat, passes, slot=0, first=None):
        """Start tracking one stepped network: every parameter will be reported ``passes`` times (once per backward
        pass of this iteration); its buckets go to communicator ``slot``.  ``first()`` runs on the slot's stream ahead of
        every bucket of this network (the trainer advances the Adam step state there)."""
        if not self.active:
            return
        if first is not None:
            if flat.flat_g.is_cuda:
                stream = self._streams.get(slot)
                if stream is None:
                    stream = self._streams[slot] = torch.cuda.Stream()
                stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(stream):
                    first()
            else:
                first()
        plan = self._plan(flat)
        self._state[flat] = dict(left=[len(ids) * passes for _, _, ids in plan],
                                 where={pid: b for b, (_, _, ids) in enumerate(plan) for pid in ids}, slot=slot)

    def hook(self, flat):
        """The ``grad_ready`` callback for the backward passes of this network (None when not reducing)."""
        if not self.active or flat not in self._state:
            return None
        st = self._state[flat]
        plan = flat._buckets

        def grad_ready(params):
            for p in params:
                b = st["where"][id(p)]
                st["left"][b] -= 1
                if st["left"][b] == 0:
                    lo, hi, _ = plan[b]
                    self._launch(flat, lo, hi, st["slot"])
        return grad_ready

    def finish(self, flat):
        """Launch whatever has not been reported complete (a safety net: every bucket is normally launched by the
        hook) and stop tracking."""
        st = self._state.pop(flat, None)
        if st is None:
            return
        for b, left in enumerate(st["left"]):
            if left > 0:
                lo, hi, _ = flat._buckets[b]
                self._launch(flat, lo, hi, st["slot"])

    def _launch(self, flat, lo, hi, slot):
        buf = flat.flat_g[lo:hi]
        self.launched.append((flat.numel, lo, hi))
        if not buf.is_cuda:
            if self.enabled:
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self._group_for(slot, False))
            if self.on_bucket is not None:
                self.on_bucket(flat, lo, hi)
            return
        from . import ops
        stream = self._streams.get(slot)
        if stream is None:
            stream = self._streams[slot] = torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        ready = [torch.cuda.Event()]
        ready[0].record(cur)                                   # dgamma / dbeta (and inline weight gradients)
        ctx = ops.current()
        side = ctx.wgrad_streams.get(ctx.lane)
        if side is not None:                                   # weight-gradient kernels of this lane run on a side stream
            ev = torch.cuda.Event()
            ev.record(side[0])
            ready.append(ev)
        with torch.cuda.stream(stream):
            for ev in ready:
                stream.wait_event(ev)
            if self.enabled:
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self._group_for(slot, True))
            if self.on_bucket is not None:
                self.on_bucket(flat, lo, hi)
            done = torch.cuda.Event()
            done.record(stream)
        self._pending.append(done)

    def launch(self, flat_g):
        """Whole-buffer all-reduce (kept for callers that exchange a buffer in one piece)."""
        if not self.enabled:
            return
        class _Whole:  # noqa: N801 -- minimal FlatNet stand-in
            pass
        w = _Whole()
        w.flat_g, w.numel = flat_g, flat_g.numel()
        self._launch(w, 0, flat_g.numel(), 0)

    def join(self):
        for ev in self._pending:
            torch.cuda.current_stream().wait_event(ev)
        self._pending = []

    @property
    def grad_scale(self):
        return 1.0 / self.world

    def broadcast_params(self, flat_nets):
        """Rank 0's weights to everyone once at start (DDP constructor semantics, C2)."""
        if not self.enabled:
            return
        for fn in flat_nets:
            dist.broadcast(fn.flat_p, src=0, group=self.group)
            fn.net._packed.invalidate()


def format_log_line(iteration, total_iterations, l):
    """The reference's log line (image_translation.py:394-398), parsed by hyperparameter_search.py:269-271 with
    ``GEN: (\\d+\\.\\d+)/(\\d+\\.\\d+)`` etc."""
    return (f"Iter [{iteration}/{total_iterations}] "
            f"GEN: {l['gen_loss_A']:.4f}/{l['gen_loss_B']:.4f}, "
            f"FM: {l['fm_loss_A']:.4f}/{l['fm_loss_B']:.4f}, "
            f"RECON: {l['recon_loss_A']:.4f}/{l['recon_loss_B']:.4f}, "
            f"DIS: {l['dis_loss_A']:.4f}/{l['dis_loss_B']:.4f}")


class LossReadback:
    """Non-blocking readback of the eight logged losses: ``request(tag)`` enqueues a device->pinned-host copy behind the
    step that produced them and returns immediately; ``poll()`` yields (tag, losses) for every copy that has landed
    (``drain()`` waits for the rest).  The training loop never stalls the GPU to print a log line (the reference calls
    ``.item()`` eight times per logged iteration, image_translation.py:394-398)."""

    def __init__(self, loss_buf, slots=4):
        self.buf = loss_buf
        self.host = [torch.empty(len(LOSS_NAMES), dtype=torch.float32).pin_memory() for _ in range(slots)]
        self.events = [torch.cuda.Event() for _ in range(slots)]
        self.pending = []          # (slot, tag), oldest first
        self.free = list(range(slots))

    def request(self, tag):
        out = []
        if not self.free:          # ring full: wait for the oldest (never happens with log intervals >> ring size)
            out = [self._take(wait=True)]
        slot = self.free.pop()
        self.host[slot].copy_(self.buf, non_blocking=True)
        self.events[slot].record(torch.cuda.current_stream())
        self.pending.append((slot, tag))
        return out

    def _take(self, wait):
        slot, tag = self.pending[0]
        if wait:
            self.events[slot].synchronize()
        elif not self.events[slot].query():
            return None
        self.pending.pop(0)
        vals = dict(zip(LOSS_NAMES, self.host[slot].tolist()))
        self.free.append(slot)
        return tag, vals

    def poll(self):
        out = []
        while self.pending:
            r = self._take(wait=False)
            if r is None:
                break
            out.append(r)
        return out

    def drain(self):
        out = []
        while self.pending:
            out.append(self._take(wait=True))
        return out


def loss_coefficients(model_arch, rate):
    """d(gen_loss)/d(component) for image_translation.py:370-382.  Returns dict with keys
    gen_A/fm_A (through D_A's fake pass), gen_B/fm_B (through D_B's fake pass), recon_A, recon_B,
    and the D-step flags dis_A/dis_B."""
    if model_arch == "discogan":
        return dict(gen_A=0.1 * (1 - rate), fm_A=0.9 * (1 - rate), gen_B=0.1 * (1 - rate), fm_B=0.9 * (1 - rate),
                    recon_A=rate, recon_B=rate, dis_A=1.0, dis_B=1.0)
    if model_arch == "recongan":  # gen_loss = gen_loss_A_total = (fm_B*.9 + gen_B*.1)(1-rate) + recon_A*rate
        return dict(gen_A=0.0, fm_A=0.0, gen_B=0.1 * (1 - rate), fm_B=0.9 * (1 - rate), recon_A=rate, recon_B=0.0,
                    dis_A=0.0, dis_B=1.0)
    if model_arch == "gan":       # gen_loss = gen_B*.1 + fm_B*.9
        return dict(gen_A=0.0, fm_A=0.0, gen_B=0.1, fm_B=0.9, recon_A=0.0, recon_B=0.0, dis_A=0.0, dis_B=1.0)
    raise ValueError(f"unknown model_arch {model_arch!r}")


class DiscoGANTrainer:
    """Owns G_A, G_B, D_A, D_B and performs reference-equivalent iterations with ``step(A, B)``.

    variant='angle_pairing' skips the first feature map in the FM loss and defaults both rates to 0.9
    (``angle_pairing.py:55-57,115``).  ``use_graphs``: replay captured CUDA graphs (default on; set
    DISCOGAN_B200_GRAPHS=0 or pass False to launch every kernel eagerly).  ``deterministic``: see below."""

    def __init__(self, image_size=512, device="cuda", model_arch="discogan", learning_rate=2e-4, beta1=0.5,
                 beta2=0.999, weight_decay=1e-5, update_interval=3, gan_curriculum=10000, starting_rate=None,
                 default_rate=None, variant="image_translation", seed=None, nets=None, process_group=None,
                 use_graphs=None, data_parallel=True, deterministic=False):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DiscoGANTrainer runs on CUDA (sm_100a) only")
        if self.device.index is not None:
            torch.cuda.set_device(self.device)
        ops.device_check()
        # launch context of this trainer: lanes, scratch buffers and split-K workspaces are its own, so several trainers
        # (or host threads) in one process never share mutable state
        self.ctx = ops.OpsContext()
        # deterministic=True: bit-reproducible iterations (no split-K atomics, BatchNorm statistics in a separate ordered
        # pass instead of the conv epilogue's shared-memory atomics); a few percent slower at 64x64
        self.deterministic = deterministic
        if deterministic:
            self.ctx.fuse_stats = self.ctx.fold_stats = False
        if os.environ.get("DISCOGAN_B200_SPLITK", "1") != "0" and not deterministic:
            with ops.use_context(self.ctx):
                ops.enable_splitk(self.device)
        self.image_size = image_size
        self.model_arch = model_arch
        loss_coefficients(model_arch, 0.5)  # validates
        angle = variant == "angle_pairing"
        self.fm_skip_first = angle
        self.starting_rate = (0.9 if angle else 0.01) if starting_rate is None else starting_rate
        self.default_rate = (0.9 if angle else 0.5) if default_rate is None else default_rate
        self.update_interval, self.gan_curriculum = update_interval, gan_curriculum
        self.lr, self.beta1, self.beta2, self.eps, self.weight_decay = learning_rate, beta1, beta2, 1e-8, weight_decay
        if nets is None:
            if seed is not None:
                torch.manual_seed(seed)
            nets = [Generator(extra_layers=True, image_size=image_size), Generator(extra_layers=True, image_size=image_size),
                    Discriminator(image_size=image_size), Discriminator(image_size=image_size)]
        self.G_A, self.G_B, self.D_A, self.D_B = [n.to(self.device).train() for n in nets]
        self.flat = {n: FlatNet(n) for n in (self.G_A, self.G_B, self.D_A, self.D_B)}
        self.reducer = GradReducer(process_group, enabled=data_parallel)
        # Adam per gradient bucket, right behind the bucket's all-reduce (or, on one GPU, behind the kernels that complete
        # it): the optimiser overlaps the rest of the backward pass instead of forming a serial tail
        self.adam_buckets = os.environ.get("DISCOGAN_B200_ADAM_BUCKETS", "1") != "0"
        if self.adam_buckets:
            self.reducer.on_bucket = self._adam_bucket
        self.reducer.broadcast_params(self.flat.values())
        self.reducer.prepare(list(self.flat.values()))
        self.loss_buf = torch.zeros(len(LOSS_NAMES), dtype=torch.float32, device=self.device)
        self.iters = 0
        if use_graphs is None:
            use_graphs = os.environ.get("DISCOGAN_B200_GRAPHS", "1") != "0"
        self.use_graphs = use_graphs
        self._graphs = {}        # (is_dis, rate, batch) -> CUDAGraph
        self._eager_done = set()
        self._static = {}        # batch -> (A, B) static input buffers
        self._pool = None
        self._scratch_gen = -1
        use_lanes = os.environ.get("DISCOGAN_B200_LANES", "1") != "0"
        self._side = torch.cuda.Stream(device=self.device) if use_lanes else None          # lane 1
        self._more = [torch.cuda.Stream(device=self.device) for _ in range(4)] if use_lanes else []   # lanes 2..5
        self._deferred = None       # stepped networks of a step(..., defer_update=True) awaiting apply_update()
        self._graph_launches = {}   # kernels inside each captured graph
        self.kernel_launches = 0    # kernels of this library launched (eagerly or by graph replay) by step()

    # ------------------------------------------------------------------------------------------
    def _disc_fake_and_losses(self, D, real, fake_img, slot):
        """The fake pass of one discriminator + its BCE and FM losses against the real pass
        (image_translation.py:353-364)."""
        lr_, feats_r, ctx_r = real
        lf_, feats_f, ctx_f = discriminator_forward(D, fake_img, save=True)
        buf = self.loss_buf
        p_real, p_fake = ops.gan_bce_fwd(lr_, lf_, buf[2 * slot:2 * slot + 2])
        fm_out = buf[4 + slot:5 + slot]
        diffs = []
        first = True
        for i, (fr, ff) in enumerate(zip(feats_r, feats_f)):
            if self.fm_skip_first and i == 0:
                diffs.append(None)
                continue
            diffs.append(ops.fm_fwd(fr, ff, fm_out, accumulate=not first))
            first = False
        if first:
            fm_out.zero_()
        return dict(ctx_r=ctx_r, ctx_f=ctx_f, p_real=p_real, p_fake=p_fake, diffs=diffs, feats_f=feats_f)

    def step(self, A, B, defer_update=False):
        """One iteration on a batch pair (fp32 NCHW on the trainer's device).  Returns True if it was a
        discriminator step.  Losses of the iteration are in ``self.loss_buf`` (see ``losses()``).

        ``defer_update=True`` stops after the backward pass (eager launches, no gradient exchange, no Adam): the
        stepped networks' gradients are in their flat buffers and ``apply_update()`` finishes the iteration.  This is
        the seam an external reducer -- or a test emulating R data-parallel ranks in one process -- plugs into."""
        with ops.use_context(self.ctx):
            return self._step(A, B, defer_update)

    def _step(self, A, B, defer_update):
        is_dis = self.iters % self.update_interval == 0
        rate = self.starting_rate if self.iters < self.gan_curriculum else self.default_rate
        from ._lib import lib
        count0 = lib().dg_launch_count()
        if defer_update:
            if self._deferred is not None:
                raise RuntimeError("apply_update() has not been called for the previous deferred step")
            self._deferred = self._forward_backward(A, B, is_dis, rate)
            self.kernel_launches += lib().dg_launch_count() - count0
            return is_dis
        if not self.use_graphs:
            self._step_impl(A, B, is_dis, rate)
            self.kernel_launches += lib().dg_launch_count() - count0
        else:
            if self._scratch_gen != self.ctx.generation and self._graphs:
                self._graphs.clear()            # a scratch buffer moved: captured pointers are stale
                self._eager_done.clear()
            key = (is_dis, rate, A.shape[0])
            g = self._graphs.get(key)
            if g is None and key not in self._eager_done:
                # first encounter: run eagerly (sizes scratch buffers, sets kernel attributes), capture next time
                self._step_impl(A, B, is_dis, rate)
                self._eager_done.add(key)
                self.kernel_launches += lib().dg_launch_count() - count0
            else:
                st = self._static.get(A.shape[0])
                if st is None:
                    st = (torch.empty_like(A), torch.empty_like(B))
                    self._static[A.shape[0]] = st
                st[0].copy_(A, non_blocking=True)
                st[1].copy_(B, non_blocking=True)
                if g is None:
                    g = torch.cuda.CUDAGraph()
                    if self._pool is None:
                        self._pool = torch.cuda.graph_pool_handle()
                    torch.cuda.synchronize()
                    with torch.cuda.graph(g, pool=self._pool):
                        self._step_impl(st[0], st[1], is_dis, rate)
                    self._graphs[key] = g
                    self._graph_launches[key] = lib().dg_launch_count() - count0
                    self._scratch_gen = self.ctx.generation
                g.replay()
                self.kernel_launches += self._graph_launches[key]
        self.iters += 1
        return is_dis

    def stepped_nets(self):
        """Networks whose gradients the pending deferred step produced (in Adam order)."""
        if self._deferred is None:
            raise RuntimeError("no deferred step pending")
        return list(self._deferred)

    def apply_update(self, grad_scale=1.0):
        """Finish a ``step(..., defer_update=True)``: Adam on the stepped networks with gradients scaled by
        ``grad_scale`` (1/R when the flat gradient buffers hold a sum over R shards)."""
        if self._deferred is None:
            raise RuntimeError("no deferred step pending")
        stepped, self._deferred = self._deferred, None
        with ops.use_context(self.ctx):
            self._update(stepped, grad_scale)
        self.iters += 1

    # ------------------------------------------------------------------------------------------
    # two "lanes" (streams): the A->B->A and B->A->B halves of the step are independent between a few join points,
    # and at 64x64 most kernels are far too small to fill 148 SMs, so running the halves concurrently hides their
    # latency.  Inside a phase the two lanes never touch the same network, gradient buffer or scratch buffer.
    def _streams(self):
        return ([self._side] + self._more) if self._side is not None else []

    def _fork(self, n=2):
        for st in self._streams()[:n - 1]:
            st.wait_stream(torch.cuda.current_stream())

    def _join(self, n=2):
        for st in self._streams()[:n - 1]:
            torch.cuda.current_stream().wait_stream(st)

    @contextlib.contextmanager
    def _lane(self, i):
        if i == 0 or self._side is None:
            yield
            return
        self.ctx.lane = i
        try:
            with torch.cuda.stream(self._streams()[i - 1]):
                yield
        finally:
            self.ctx.lane = 0

    def _step_impl(self, A, B, is_dis, rate):
        stepped = self._forward_backward(A, B, is_dis, rate, reduce=True)
        self.reducer.join()
        self._update(stepped, self.reducer.grad_scale, done_in_buckets=self.reducer.active and self.adam_buckets)

    def _tick(self, net):
        """The Adam step-state advance of one network, to run ahead of its buckets (None when Adam is not bucketed)."""
        if not self.adam_buckets:
            return None
        return lambda: self.flat[net].adam_tick(self.beta1, self.beta2)

    def _adam_bucket(self, flat, lo, hi):
        flat.adam_range(lo, hi, self.lr, self.beta1, self.beta2, self.eps, self.weight_decay, self.reducer.grad_scale)

    def _update(self, stepped, grad_scale, done_in_buckets=False):
        self._fork()
        for i, n in enumerate(stepped):
            with self._lane(i):
                if done_in_buckets:          # every range was updated behind its bucket: only the bf16 GEMM copies are left
                    n._packed.refresh()
                else:
                    self.flat[n].adam(self.lr, self.beta1, self.beta2, self.eps, self.weight_decay, grad_scale)
        self._join()

    def _forward_backward(self, A, B, is_dis, rate, reduce=False):
        """All eight forwards, the losses and the backward passes that reach an optimiser step.  Returns the stepped
        networks; with ``reduce`` their flat gradients are handed to the data-parallel reducer as they complete."""
        co = loss_coefficients(self.model_arch, rate)
        G_A, G_B, D_A, D_B = self.G_A, self.G_B, self.D_A, self.D_B
        self.ctx.arena.reset()        # one memset re-zeroes every BatchNorm accumulator of the iteration (before the lanes fork)
        save_g = not is_dis
        lane, fork, join = self._lane, self._fork, self._join
        # forward, phase 1 (4 lanes): the two first generator passes and the two real discriminator passes
        # (big images fill the GPU with single kernels: the discriminators then share lanes 0/1 with the generators)
        small = self._side is not None and self.image_size <= 128
        l2, l3 = (2, 3) if small else (0, 1)
        nf = 4
        if small and is_dis:     # the discriminators' early backward: weight gradients on lanes 4/5
            self.ctx.wgrad_streams = {2: (self._more[2], 4), 3: (self._more[3], 5)}
            nf = 6
        fork(nf)
        with lane(0):
            AB, c_gb1 = generator_forward(G_B, A, save=save_g)       # A -> B
        with lane(1):
            BA, c_ga1 = generator_forward(G_A, B, save=save_g)       # B -> A
        # On a D step the discriminators' backward passes need nothing from the generators' second passes (the BCE
        # gradient of the real logit depends on the real pass only), so each pass is back-propagated on its lane right
        # after its forward, in the reference's accumulation order (real, then fake), beside the generator forwards.
        d_coef = {0: co["dis_A"], 1: co["dis_B"]}
        stepped = []
        red = self.reducer
        if is_dis:
            for slot, (D, c) in enumerate(((D_A, co["dis_A"]), (D_B, co["dis_B"]))):
                if c != 0.0:
                    self.flat[D].zero_grad()
                    stepped.append(D)
                    if reduce:                   # two backward passes (real, fake) write every parameter
                        red.begin(self.flat[D], passes=2, slot=slot, first=self._tick(D))

        def real_pass(D, img, slot):
            real = discriminator_forward(D, img, save=is_dis)
            if is_dis and d_coef[slot] != 0.0:
                p_real = ops.sigmoid_fwd(real[0])
                dlr, _ = ops.gan_bce_bwd(p_real, p_real, d_coef[slot], 0.0, want_fake=False)
                discriminator_backward(D, real[2], dlr, need_dx=False, need_wgrad=True,
                                       grad_ready=red.hook(self.flat[D]) if reduce else None)
            return real

        def fake_pass(D, real, img, slot):
            d = self._disc_fake_and_losses(D, real, img, slot)
            if is_dis and d_coef[slot] != 0.0:
                _, dlf = ops.gan_bce_bwd(d["p_real"], d["p_fake"], d_coef[slot], 0.0, want_real=False)
                discriminator_backward(D, d["ctx_f"], dlf, need_dx=False, need_wgrad=True,
                                       grad_ready=red.hook(self.flat[D]) if reduce else None)
            return d

        with lane(l2):
            real_a = real_pass(D_A, A, 0)
        with lane(l3):
            real_b = real_pass(D_B, B, 1)
        join(nf); fork(nf)
        # phase 2: second generator passes + reconstruction losses, fake discriminator passes + GAN / FM losses
        # (every network's second pass follows its first: BatchNorm running statistics advance in reference order)
        with lane(0):
            ABA, c_ga2 = generator_forward(G_A, AB, save=save_g)     # A -> B -> A
            ops.mse_fwd(ABA, A, self.loss_buf[6:7])
        with lane(1):
            BAB, c_gb2 = generator_forward(G_B, BA, save=save_g)     # B -> A -> B
            ops.mse_fwd(BAB, B, self.loss_buf[7:8])
        with lane(l2):
            da = fake_pass(D_A, real_a, BA, 0)
        with lane(l3):
            db = fake_pass(D_B, real_b, AB, 1)
        join(nf)
        self.ctx.wgrad_streams = {}

        # backward: lanes 0/1 carry the two chains; for small images each chain's weight-gradient kernels go to its
        # own side stream (lanes 2/3) and overlap the dgrad chain
        nb = 2
        wl = os.environ.get("DISCOGAN_B200_WGRAD_LANES", "auto")
        if self._side is not None and (wl == "1" or (wl == "auto" and self.image_size <= 128)):
            # (lanes 2/3 carry the discriminators' fake-pass backward at the same time, hence lanes 4/5 here)
            self.ctx.wgrad_streams = {0: (self._more[2], 4), 1: (self._more[3], 5)}
            nb = 6
        if is_dis:
            if reduce:                          # backward already done beside the forward passes; buckets went out as each
                for D in stepped:               # layer's fake-pass gradients were enqueued
                    red.finish(self.flat[D])
        else:
            use_a = co["gen_A"] != 0.0 or co["fm_A"] != 0.0      # losses through D_A(BA): reach G_A pass 1
            use_b = co["gen_B"] != 0.0 or co["fm_B"] != 0.0      # losses through D_B(AB): reach G_B pass 1
            if use_b or co["recon_A"] != 0.0 or co["recon_B"] != 0.0:
                stepped.append(G_B)
            if use_a or co["recon_B"] != 0.0 or co["recon_A"] != 0.0:
                stepped.append(G_A)
            for G in stepped:
                self.flat[G].zero_grad()
            if reduce:                          # G_B's last pass runs on lane 0, G_A's on lane 1: one communicator each
                n_b = int(co["recon_B"] != 0.0) + int(co["recon_A"] != 0.0 or use_b)
                n_a = int(co["recon_A"] != 0.0) + int(co["recon_B"] != 0.0 or use_a)
                if G_B in stepped:
                    red.begin(self.flat[G_B], passes=n_b, slot=0, first=self._tick(G_B))
                if G_A in stepped:
                    red.begin(self.flat[G_A], passes=n_a, slot=1, first=self._tick(G_A))
            hook = (lambda G: red.hook(self.flat[G])) if reduce else (lambda G: None)
            dAB = dBA = dAB_d = dBA_d = None
            fork(6 if small else nb)
            # phase A (lanes 2/3 for small images): the discriminators' fake passes, data gradients only -- independent of
            # phase B (lanes 0/1): the generators' second passes.  Both produce a gradient for AB (resp. BA); the two
            # are summed on the fly by the first kernel of phase C.
            la, lb = (2, 3) if small else (0, 1)
            with lane(la):
                if use_b:
                    dAB_d = self._disc_fake_backward(D_B, db, co["gen_B"], co["fm_B"], AB.shape[0])
            with lane(lb):
                if use_a:
                    dBA_d = self._disc_fake_backward(D_A, da, co["gen_A"], co["fm_A"], BA.shape[0])
            with lane(0):
                if co["recon_A"] != 0.0:
                    dABA = ops.mse_bwd(ABA, A, co["recon_A"])
                    dAB = generator_backward(G_A, c_ga2, dABA, need_dx=True, need_wgrad=True, grad_ready=hook(G_A))
            with lane(1):
                if co["recon_B"] != 0.0:
                    dBAB = ops.mse_bwd(BAB, B, co["recon_B"])
                    dBA = generator_backward(G_B, c_gb2, dBAB, need_dx=True, need_wgrad=True, grad_ready=hook(G_B))
            join(6 if small else nb); fork(nb)
            # phase C: the generators' first passes (each accumulates into the gradients its second pass just wrote)
            with lane(0):
                g1, g2 = (dAB, dAB_d) if dAB is not None else (dAB_d, None)
                if g1 is not None:
                    generator_backward(G_B, c_gb1, g1, need_dx=False, need_wgrad=True, dout2=g2, grad_ready=hook(G_B))
            with lane(1):
                g1, g2 = (dBA, dBA_d) if dBA is not None else (dBA_d, None)
                if g1 is not None:
                    generator_backward(G_A, c_ga1, g1, need_dx=False, need_wgrad=True, dout2=g2, grad_ready=hook(G_A))
            join(nb)
            if reduce:
                for G in stepped:
                    red.finish(self.flat[G])
        self.ctx.wgrad_streams = {}
        return stepped

    def _disc_fake_backward(self, D, d, c_gen, c_fm, B):
        """Back-prop c_gen*gen_loss + c_fm*fm_loss through the fake pass of D down to its input image."""
        _, dlf = ops.gan_bce_bwd(d["p_real"], d["p_fake"], 0.0, c_gen, want_real=False)
        bcast = []
        for diff, f in zip(d["diffs"], d["feats_f"]):
            if diff is None or c_fm == 0.0:
                bcast.append(None)
            else:
                n = diff.numel()
                bcast.append((diff, -c_fm * 2.0 / (float(n) * float(B))))
        return discriminator_backward(D, d["ctx_f"], dlf, None, bcast, need_dx=True, need_wgrad=False)

    # ------------------------------------------------------------------------------------------
    def losses(self):
        """Host copy of the last iteration's eight logged losses (synchronises)."""
        return dict(zip(LOSS_NAMES, self.loss_buf.tolist()))

    def log_line(self, total_iterations):
        """The reference's log format for the iteration just done (blocking; the CLI uses ``LossReadback`` instead)."""
        return format_log_line(self.iters - 1, total_iterations, self.losses())

    # ------------------------------------------------------------------------------------------
    # training state beyond the four state-dicts: the reference saves weights only
    # (image_translation.py:420-432) and restarts Adam and the GAN curriculum on --load_*
    # (distributed_image_translation.py:379-393); this is the part that makes a restart continue instead.
    def training_state(self):
        """Iteration counter + per-network Adam moments and step state, CPU tensors keyed like the checkpoints
        (gen_A / gen_B / dis_A / dis_B)."""
        st = {"iters": self.iters, "format": 1, "image_size": self.image_size, "model_arch": self.model_arch}
        for name, net in zip(("gen_A", "gen_B", "dis_A", "dis_B"), self.nets()):
            f = self.flat[net]
            st[name] = {"exp_avg": f.exp_avg.detach().cpu().clone(), "exp_avg_sq": f.exp_avg_sq.detach().cpu().clone(),
                        "adam_state": f.adam_state.detach().cpu().clone()}
        return st

    def load_training_state(self, st):
        if st.get("image_size", self.image_size) != self.image_size:
            raise ValueError(f"training state is for image_size {st.get('image_size')}, trainer is {self.image_size}")
        for name, net in zip(("gen_A", "gen_B", "dis_A", "dis_B"), self.nets()):
            f, src = self.flat[net], st[name]
            if src["exp_avg"].numel() != f.exp_avg.numel():
                raise ValueError(f"training state of {name} has {src['exp_avg'].numel()} elements, expected {f.exp_avg.numel()}")
            f.exp_avg.copy_(src["exp_avg"])
            f.exp_avg_sq.copy_(src["exp_avg_sq"])
            f.adam_state.copy_(src["adam_state"])
        self.iters = int(st["iters"])

    def load_weights(self, state_dicts):
        """Load the four reference-format state-dicts (gen_A, gen_B, dis_A, dis_B order or a dict by those names) into the
        flat buffers and refresh the bf16 GEMM copies."""
        if isinstance(state_dicts, dict):
            state_dicts = [state_dicts.get(k) for k in ("gen_A", "gen_B", "dis_A", "dis_B")]
        for net, sd in zip(self.nets(), state_dicts):
            if sd is not None:
                net.load_state_dict(sd)
                net._packed.invalidate()

    @torch.no_grad()
    def sample_images(self, test_A, test_B, mode="reference"):
        """The four generator passes of ``save_sample_images`` (image_translation.py:170-176): AB, BA, ABA, BAB.
        mode='reference': like the reference, the generators stay in train mode under no_grad, so the passes use the test
        batch's statistics AND advance the BatchNorm running statistics / counters.
        mode='eval': clean eval-mode passes that leave the training state untouched."""
        if mode not in ("reference", "eval"):
            raise ValueError(mode)
        with ops.use_context(self.ctx):
            if mode == "eval":
                self.G_A.eval(); self.G_B.eval()
            try:
                AB = self.G_B(test_A)
                BA = self.G_A(test_B)
                ABA = self.G_A(AB)
                BAB = self.G_B(BA)
            finally:
                self.G_A.train(); self.G_B.train()
        return AB, BA, ABA, BAB

    def nets(self):
        return self.G_A, self.G_B, self.D_A, self.D_B

    def close(self):
        """Drop the captured CUDA graphs and their memory pool.  Call before ``dist.destroy_process_group()``: NCCL
        cannot tear down a communicator while instantiated graphs still reference its kernels."""
        torch.cuda.synchronize()
        self._graphs.clear()
        self._graph_launches.clear()
        self._eager_done.clear()
        self._static.clear()
        self._pool = None
        torch.cuda.synchronize()
