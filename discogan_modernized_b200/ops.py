"""Tensor-level wrappers over the C ABI (include/discogan_b200.h).

Every function launches on ``torch.cuda.current_stream()`` and never synchronises.  Inputs are
checked (device, dtype, contiguity) and a wrong argument raises ``ValueError``/``KernelError``
synchronously -- there is no CPU fallback.
"""
import contextlib
import ctypes
import os
import threading

import torch

from ._lib import ConvOpts, KernelError, check, lib

ACT_NONE, ACT_LRELU, ACT_RELU = 0, 1, 2
BF16, F32 = torch.bfloat16, torch.float32


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t, dtype=None, name="tensor"):
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (discogan_modernized_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t.data_ptr()


class StatArena:
    """Zero-initialised fp32 accumulators for the folded BatchNorm reductions (``dg_conv_opts.stat_accumulate``,
    ``dg_bn_act_fwd_acc`` / ``dg_bn_act_bwd_acc``).  ``take(n)`` hands out the next n floats; a trainer calls ``reset()``
    once per iteration before any lane starts, which rewinds the cursor and re-zeroes what the previous iteration used
    with ONE memset (captured as the first node of the step graph).  Without resets (eager module calls) the arena simply
    keeps allocating fresh zeroed chunks."""

    def __init__(self, ctx, floats=1 << 20):
        self.ctx, self.floats = ctx, floats
        self.buf, self.cursor, self.high = None, 0, 0

    def take(self, n, device):
        n = (n + 63) // 64 * 64
        if self.buf is None or self.buf.device != torch.device(device) or self.cursor + n > self.buf.numel():
            if torch.cuda.is_current_stream_capturing():
                raise KernelError("statistics arena would be (re)allocated during CUDA-graph capture; run one eager step first")
            if self.buf is not None:
                self.floats = 2 * max(self.floats, n)        # an iteration must fit one buffer (graph capture needs it)
            self.buf = torch.zeros(max(self.floats, n), dtype=F32, device=device)
            self.cursor = 0
            self.ctx.generation += 1
        out = self.buf[self.cursor:self.cursor + n]
        self.cursor += n
        self.high = max(self.high, self.cursor)
        return out

    def reset(self):
        if self.buf is not None and self.high:
            self.buf[:self.high].zero_()
        self.cursor = 0


class OpsContext:
    """Everything the wrappers need beyond their arguments: which lane (stream slot) the caller is on, where weight-
    gradient kernels go, the grow-only scratch buffers and split-K workspaces (per device and lane), and the test-only
    tiling overrides.  A trainer owns one context and activates it around its step (``with ops.use_context(ctx)``);
    code that calls the wrappers directly uses a per-thread default.  Nothing here is process-global: two trainers, or
    two host threads, never share a context unless they are given the same one."""

    def __init__(self):
        self.lane = 0
        self.wgrad_streams = {}      # lane -> (side stream, its lane id) for weight-gradient kernels
        self.scratch = {}            # (kind, device, lane) -> uint8 buffer
        self.generation = 0          # bumped whenever a scratch buffer is (re)allocated: CUDA graphs that captured the
                                     # old pointer must be re-captured (DiscoGANTrainer checks before every replay)
        self.splitk_bytes = {}       # device -> workspace size (split-K enabled on that device)
        self.splitk_ws = {}          # (device, lane) -> uint8 buffer
        self.block_n, self.pair, self.wgrad_pair = 0, -1, -1    # test hooks (dg_conv_opts)
        # BatchNorm statistics from the conv epilogue's fp32 accumulators (shared-memory float atomics: the summation
        # order, hence the last bit, varies from run to run); off = a separate deterministic statistics pass
        self.fuse_stats = os.environ.get("DISCOGAN_B200_FUSE_STATS", "1") != "0"
        # folded finalize (opt-in, DISCOGAN_B200_FOLD_STATS=1): reductions add into arena accumulators, consumers derive
        # their coefficients -- two launches fewer per BatchNorm layer and pass.  Measured on B200 it does NOT pay: inside a
        # CUDA graph a finalize kernel costs only ~1.5 us of chain time, while ~148-way same-address red.global.add in the
        # reductions costs 4-13 us per launch (tools/bn_fold_micro.py; step 1.997 -> 2.086 ms at 64x64).  Default: per-CTA
        # partial rows + finalize kernels.
        self.fold_stats = self.fuse_stats and os.environ.get("DISCOGAN_B200_FOLD_STATS", "0") == "1"
        self.arena = StatArena(self)
        # eval-mode forwards: BatchNorm (running statistics) + activation folded into the producing conv's epilogue
        self.fold_eval_bn = os.environ.get("DISCOGAN_B200_FOLD_EVAL_BN", "1") != "0"
        self._opts = {}

    def conv_opts(self, device, splitk=True, accumulate=False, affine=None):
        """ctypes ``dg_conv_opts`` for a launch on the current lane (cached per (device, lane): the struct must stay
        alive until the call returns, and its address is stable for repeated launches).  affine = (stats [4,C], act,
        slope): fold the eval-mode BatchNorm scale / shift and the activation into the epilogue."""
        device = torch.device(device)
        nbytes = self.splitk_bytes.get(device, 0) if splitk else 0
        ptr = 0
        if nbytes:
            key = (device, self.lane)
            buf = self.splitk_ws.get(key)
            if buf is None:
                if torch.cuda.is_current_stream_capturing():
                    raise KernelError("split-K workspace would be allocated during CUDA-graph capture; run one eager step first")
                buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
                self.splitk_ws[key] = buf
                self.generation += 1
            ptr = buf.data_ptr()
        key = (device, self.lane, bool(nbytes), accumulate)
        o = self._opts.get(key)
        if o is None:
            o = self._opts[key] = ConvOpts()
        o.splitk_ws, o.splitk_ws_bytes = ptr or None, nbytes
        o.block_n, o.pair, o.wgrad_pair = self.block_n, self.pair, self.wgrad_pair
        o.stat_accumulate = int(accumulate)
        if affine is not None:
            stats, act, slope = affine
            C = stats.shape[1]
            o.affine_scale, o.affine_shift = stats.data_ptr() + 8 * C, stats.data_ptr() + 12 * C
            o.affine_act, o.affine_slope = int(act), float(slope)
        else:
            o.affine_scale = o.affine_shift = None
        return ctypes.byref(o)

    def plan_opts(self, device):
        """Options for a launch-plan query (dg_conv_stats_rows): the workspace size matters, the pointer does not."""
        o = ConvOpts()
        o.splitk_ws, o.splitk_ws_bytes = None, self.splitk_bytes.get(torch.device(device), 0)
        o.block_n, o.pair, o.wgrad_pair = self.block_n, self.pair, self.wgrad_pair
        return o


_tls = threading.local()


def current():
    """The active context of this thread (a per-thread default unless a trainer has activated its own)."""
    ctx = getattr(_tls, "ctx", None)
    if ctx is None:
        ctx = _tls.ctx = OpsContext()
    return ctx


@contextlib.contextmanager
def use_context(ctx):
    prev = getattr(_tls, "ctx", None)
    _tls.ctx = ctx
    try:
        yield ctx
    finally:
        _tls.ctx = prev


def set_lane(i):
    current().lane = i


@contextlib.contextmanager
def wgrad_side():
    """Run the enclosed weight-gradient kernels on the current lane's side stream, ordered after everything already
    enqueued on the lane: wgrads only feed the optimiser, so they overlap the dgrad chain that continues on the lane.
    The caller keeps the operand tensors alive until the streams are joined."""
    ctx = current()
    ent = ctx.wgrad_streams.get(ctx.lane)
    if ent is None:
        yield
        return
    st, side_lane = ent                      # (stream, scratch lane id of that stream)
    st.wait_stream(torch.cuda.current_stream())
    prev = ctx.lane
    ctx.lane = side_lane
    try:
        with torch.cuda.stream(st):
            yield
    finally:
        ctx.lane = prev


def scratch(kind, nbytes, device):
    """Grow-only per-device, per-lane scratch buffers (wgrad workspace, BN partials, reduction partials)."""
    ctx = current()
    key = (kind, device, ctx.lane)
    buf = ctx.scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        if torch.cuda.is_current_stream_capturing():
            raise KernelError(f"scratch buffer '{kind}' would be allocated during CUDA-graph capture; run one eager "
                              "step first")
        buf = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        ctx.scratch[key] = buf
        ctx.generation += 1
    return buf


def device_check():
    check(lib().dg_device_check(), "dg_device_check")


def set_conv_tiling(block_n=0, pair=-1, wgrad_pair=-1):
    """Test hook: force the GEMM N tile / CTA pairing of the conv kernels and the wgrad kernel variant in the current
    context (0, -1 = heuristics).  Travels with each call as ``dg_conv_opts``; no library state is touched."""
    ctx = current()
    o = ConvOpts()
    o.block_n, o.pair, o.wgrad_pair = int(block_n), int(pair), int(wgrad_pair)
    check(lib().dg_conv_opts_check(ctypes.byref(o)), "dg_conv_opts_check")
    ctx.block_n, ctx.pair, ctx.wgrad_pair = int(block_n), int(pair), int(wgrad_pair)


def enable_splitk(device, nbytes=64 << 20):
    """Allow split-K for the SM-starved deep layers on this device in the current context (the trainer turns it on).
    Each lane gets its own fp32 workspace, handed to the library with every launch."""
    current().splitk_bytes[torch.device(device)] = int(nbytes)


def _opts(device, splitk=True, accumulate=False, affine=None):
    return current().conv_opts(device, splitk, accumulate, affine)


# ---- weights / layout ---------------------------------------------------------------------
def pack_weights(w, want_wd=True, want_wu=True, out=None):
    """fp32 [Cs,Cb,4,4] -> (bf16 Wd [Cs,16,Cb], bf16 Wu [Cb,16,Cs]); `out=(wd, wu)` re-packs in place."""
    Cs, Cb = w.shape[0], w.shape[1]
    if out is not None:
        wd, wu = out
    else:
        wd = torch.empty(Cs, 16, Cb, dtype=BF16, device=w.device) if want_wd else None
        wu = torch.empty(Cb, 16, Cs, dtype=BF16, device=w.device) if want_wu else None
    check(lib().dg_pack_weights(_ptr(w, F32, "w"), _ptr(wd), _ptr(wu), Cs, Cb, _stream()), "dg_pack_weights")
    return wd, wu


def pack_weights_multi(table, n, total_items):
    """Re-pack every (w -> wd, wu) listed in the device int64 table [n,6] with one launch."""
    check(lib().dg_pack_weights_multi(_ptr(table, torch.int64, "table"), n, total_items, _stream()), "dg_pack_weights_multi")


def nhwc_to_nchw_f32(x):
    B, H, W, C = x.shape
    y = torch.empty(B, C, H, W, dtype=F32, device=x.device)
    check(lib().dg_nhwc_bf16_to_nchw_f32(_ptr(x, BF16, "x"), _ptr(y), B, H * W, C, _stream()), "dg_nhwc_bf16_to_nchw_f32")
    return y


def nchw_f32_to_nhwc(x):
    B, C, H, W = x.shape
    y = torch.empty(B, H, W, C, dtype=BF16, device=x.device)
    check(lib().dg_nchw_f32_to_nhwc_bf16(_ptr(x, F32, "x"), _ptr(y), B, H * W, C, _stream()), "dg_nchw_f32_to_nhwc_bf16")
    return y


# ---- tensor-core convolutions ---------------------------------------------------------------
_conv_impl = "tc"  # "simt" only for on-device debugging (DISCOGAN_B200_CONV=simt)


def set_conv_impl(name):
    global _conv_impl
    if name not in ("tc", "simt"):
        raise ValueError(name)
    _conv_impl = name


def _check_affine(affine, C):
    stats, act, _ = affine
    if stats.dtype != F32 or tuple(stats.shape) != (4, C) or not stats.is_contiguous() or C % 4:
        raise ValueError(f"affine statistics must be contiguous fp32 [4, {C}] with C a multiple of 4")
    if act not in (ACT_NONE, ACT_LRELU, ACT_RELU):
        raise ValueError(f"unknown activation code {act}")


def conv_down(big, wd, affine=None):
    """[B,H,W,Cb] x Wd[Cs,16,Cb] -> [B,H/2,W/2,Cs]  (Conv2d 4x4 s2 p1 forward / ConvTranspose2d dgrad).
    affine = (stats [4,Cs] from bn_eval_stats, act, slope): out = act(conv * scale + shift) in the epilogue (eval-mode
    BatchNorm + activation folded into the convolution)."""
    B, H, W, Cb = big.shape
    Cs = wd.shape[0]
    out = torch.empty(B, H // 2, W // 2, Cs, dtype=BF16, device=big.device)
    if affine is not None:
        _check_affine(affine, Cs)
        if _conv_impl != "tc":
            raise KernelError("the folded affine epilogue exists on the tensor-core path only")
    if _conv_impl == "tc":
        check(lib().dg_conv4x4s2_fprop(_ptr(big, BF16, "big"), _ptr(wd, BF16, "wd"), _ptr(out), B, H, W, Cb, Cs,
                                       _opts(big.device, affine=affine), _stream()), "dg_conv4x4s2_fprop")
    else:
        check(lib().dg_simt_conv4x4s2_fprop(_ptr(big, BF16, "big"), _ptr(wd, BF16, "wd"), _ptr(out), B, H, W, Cb, Cs,
                                            _stream()), "dg_simt_conv4x4s2_fprop")
    return out


def conv_up(small, wu, mask=None, slope=0.2, affine=None):
    """[B,Hs,Ws,Cs] x Wu[Cb,16,Cs] -> [B,2Hs,2Ws,Cb]  (Conv2d dgrad / ConvTranspose2d 4x4 s2 p1 forward).
    mask (bf16, output shape): multiply by the LeakyReLU derivative (mask > 0 ? 1 : slope) in the epilogue.
    affine: see conv_down (eval-mode BatchNorm + activation folded into the ConvTranspose2d)."""
    B, Hs, Ws, Cs = small.shape
    Cb = wu.shape[0]
    out = torch.empty(B, 2 * Hs, 2 * Ws, Cb, dtype=BF16, device=small.device)
    if affine is not None:
        _check_affine(affine, Cb)
        if _conv_impl != "tc" or mask is not None:
            raise KernelError("the folded affine epilogue exists on the tensor-core path only and excludes the mask")
        check(lib().dg_conv4x4s2_dgrad(_ptr(small, BF16, "small"), _ptr(wu, BF16, "wu"), _ptr(out), B, Hs, Ws, Cs, Cb,
                                       _opts(small.device, affine=affine), _stream()), "dg_conv4x4s2_dgrad")
        return out
    if _conv_impl == "tc" and mask is not None:
        check(lib().dg_conv4x4s2_dgrad_masked(_ptr(small, BF16, "small"), _ptr(wu, BF16, "wu"), _ptr(out),
                                              _ptr(mask, BF16, "mask"), slope, B, Hs, Ws, Cs, Cb, _opts(small.device),
                                              _stream()), "dg_conv4x4s2_dgrad_masked")
        return out
    if _conv_impl == "tc":
        check(lib().dg_conv4x4s2_dgrad(_ptr(small, BF16, "small"), _ptr(wu, BF16, "wu"), _ptr(out), B, Hs, Ws, Cs, Cb,
                                       _opts(small.device), _stream()), "dg_conv4x4s2_dgrad")
    else:
        check(lib().dg_simt_conv4x4s2_dgrad(_ptr(small, BF16, "small"), _ptr(wu, BF16, "wu"), _ptr(out), B, Hs, Ws, Cs,
                                            Cb, _stream()), "dg_simt_conv4x4s2_dgrad")
    if mask is not None:   # simt debug path: apply the mask with stock ops
        out = torch.where(mask > 0, out, out * slope)
    return out


# the *_stats wrappers fall back to these for split-K shapes; private aliases so that instrumentation which replaces the
# public names (bench.py's per-launch timing) never sees one launch twice
_conv_down_impl, _conv_up_impl = conv_down, conv_up


def conv_down_stats(big, wd):
    """conv_down that also returns per-CTA partial BatchNorm sums of its output: (out, part fp32 [2, rows, Cs])."""
    B, H, W, Cb = big.shape
    Cs = wd.shape[0]
    rows = lib().dg_conv_stats_rows(0, B, H // 2, W // 2, Cs, Cb, ctypes.byref(current().plan_opts(big.device)))
    if rows <= 0:        # split-K shape: no fused statistics
        return _conv_down_impl(big, wd), None
    out = torch.empty(B, H // 2, W // 2, Cs, dtype=BF16, device=big.device)
    part = torch.empty(2, rows, Cs, dtype=F32, device=big.device)
    check(lib().dg_conv4x4s2_fprop_stats(_ptr(big, BF16, "big"), _ptr(wd, BF16, "wd"), _ptr(out), _ptr(part), B, H, W, Cb,
                                         Cs, _opts(big.device), _stream()), "dg_conv4x4s2_fprop_stats")
    return out, part


def conv_down_acc(big, wd):
    """conv_down whose epilogue adds the per-channel sums / sums of squares of its fp32 output into a fresh zeroed arena
    accumulator: (out, acc fp32 [2, Cs]).  Works for every shape (split-K layers: the finish kernel sums)."""
    B, H, W, Cb = big.shape
    Cs = wd.shape[0]
    out = torch.empty(B, H // 2, W // 2, Cs, dtype=BF16, device=big.device)
    acc = current().arena.take(2 * Cs, big.device)[:2 * Cs].view(2, Cs)
    check(lib().dg_conv4x4s2_fprop_stats(_ptr(big, BF16, "big"), _ptr(wd, BF16, "wd"), _ptr(out), acc.data_ptr(), B, H, W, Cb,
                                         Cs, _opts(big.device, accumulate=True), _stream()), "dg_conv4x4s2_fprop_stats")
    return out, acc


def conv_up_acc(small, wu):
    """conv_up (ConvTranspose2d forward) with accumulated statistics: (out, acc fp32 [2, Cb])."""
    B, Hs, Ws, Cs = small.shape
    Cb = wu.shape[0]
    out = torch.empty(B, 2 * Hs, 2 * Ws, Cb, dtype=BF16, device=small.device)
    acc = current().arena.take(2 * Cb, small.device)[:2 * Cb].view(2, Cb)
    check(lib().dg_convT4x4s2_fprop_stats(_ptr(small, BF16, "small"), _ptr(wu, BF16, "wu"), _ptr(out), acc.data_ptr(), B, Hs,
                                          Ws, Cs, Cb, _opts(small.device, accumulate=True), _stream()),
          "dg_convT4x4s2_fprop_stats")
    return out, acc


def conv_up_stats(small, wu):
    """conv_up (ConvTranspose2d forward) with fused partial BatchNorm sums: (out, part fp32 [2, rows, Cb])."""
    B, Hs, Ws, Cs = small.shape
    Cb = wu.shape[0]
    rows = lib().dg_conv_stats_rows(1, B, Hs, Ws, Cs, Cb, ctypes.byref(current().plan_opts(small.device)))
    if rows <= 0:        # split-K shape: no fused statistics
        return _conv_up_impl(small, wu), None
    out = torch.empty(B, 2 * Hs, 2 * Ws, Cb, dtype=BF16, device=small.device)
    part = torch.empty(2, rows, Cb, dtype=F32, device=small.device)
    check(lib().dg_convT4x4s2_fprop_stats(_ptr(small, BF16, "small"), _ptr(wu, BF16, "wu"), _ptr(out), _ptr(part), B, Hs,
                                          Ws, Cs, Cb, _opts(small.device), _stream()), "dg_convT4x4s2_fprop_stats")
    return out, part


def conv_wgrad(small, big, dw, beta=1.0):
    """dw[Cs,Cb,4,4] = beta*dw + sum_pixels small (x) shifted big."""
    B, Hs, Ws, Cs = small.shape
    Cb = big.shape[3]
    if tuple(dw.shape) != (Cs, Cb, 4, 4):
        raise ValueError(f"dw shape {tuple(dw.shape)} != {(Cs, Cb, 4, 4)}")
    if _conv_impl == "simt":
        check(lib().dg_simt_conv4x4s2_wgrad(_ptr(small, BF16, "small"), _ptr(big, BF16, "big"), _ptr(dw, F32, "dw"),
                                            beta, B, Hs, Ws, Cs, Cb, _stream()), "dg_simt_conv4x4s2_wgrad")
        return
    need = lib().dg_conv4x4s2_wgrad_workspace(B, Hs, Ws, Cs, Cb)
    if need == 0:
        raise KernelError(f"wgrad: unsupported shape B={B} Hs={Hs} Ws={Ws} Cs={Cs} Cb={Cb}")
    ws = scratch("wgrad", need, small.device)
    check(lib().dg_conv4x4s2_wgrad(_ptr(small, BF16, "small"), _ptr(big, BF16, "big"), _ptr(dw, F32, "dw"), beta, B, Hs,
                                   Ws, Cs, Cb, ws.data_ptr(), ws.numel(), _opts(small.device, splitk=False), _stream()),
          "dg_conv4x4s2_wgrad")


# ---- image-side 3-channel layers, tensor-core path -----------------------------------------------
def c3_pack_weights(w, out=None):
    """fp32 [64,3,4,4] -> (wc bf16 [64,64] for the down GEMM, wu3 bf16 [16,16,64] for the up GEMM)."""
    if out is not None:
        wc, wu3 = out
    else:
        wc = torch.empty(64, 64, dtype=BF16, device=w.device)
        wu3 = torch.empty(16, 16, 64, dtype=BF16, device=w.device)
    check(lib().dg_c3_pack_weights(_ptr(w, F32, "w"), _ptr(wc), _ptr(wu3), _stream()), "dg_c3_pack_weights")
    return wc, wu3


def img_pad_nhwc4(img, yimg=None, img2=None):
    """fp32 NCHW [B,3,S,S] (+ img2) (times yimg*(1-yimg) if given) -> zero-padded bf16 NHWC4 [B,S+2,S+2,4]."""
    B, _, S, _ = img.shape
    out = torch.empty(B, S + 2, S + 2, 4, dtype=BF16, device=img.device)
    check(lib().dg_img_pad_nhwc4(_ptr(img, F32, "img"), _ptr(img2, F32, "img2"), _ptr(yimg, F32, "yimg"), _ptr(out), B, S,
                                 _stream()), "dg_img_pad_nhwc4")
    return out


def c3_down_tc(xp, wc, act, slope=0.2):
    """padded NHWC4 image -> bf16 [B,S/2,S/2,64] = act(conv 4x4 s2 p1)."""
    B, S = xp.shape[0], xp.shape[1] - 2
    y = torch.empty(B, S // 2, S // 2, 64, dtype=BF16, device=xp.device)
    check(lib().dg_c3_down_tc(_ptr(xp, BF16, "xp"), _ptr(wc, BF16, "wc"), _ptr(y), B, S, act, slope, _stream()),
          "dg_c3_down_tc")
    return y


def c3_up_tc(x64, wu3, sigmoid, out=None, accumulate=False):
    """bf16 [B,S/2,S/2,64] -> fp32 NCHW [B,3,S,S] (+)= [sigmoid](convT 4x4 s2 p1)."""
    B, Hs = x64.shape[0], x64.shape[1]
    S = 2 * Hs
    if out is None:
        out = torch.empty(B, 3, S, S, dtype=F32, device=x64.device)
        accumulate = False
    check(lib().dg_c3_up_tc(_ptr(x64, BF16, "x64"), _ptr(wu3, BF16, "wu3"), _ptr(out, F32, "img"), B, S, int(sigmoid),
                            int(accumulate), _stream()), "dg_c3_up_tc")
    return out


def c3_wgrad_tc(v64, xp, dw, beta=1.0):
    """dw[64,3,4,4] = beta*dw + sum_pixels v64 (x) patches(xp)."""
    B, S = xp.shape[0], xp.shape[1] - 2
    need = lib().dg_c3_wgrad_workspace(B, S)
    ws = scratch("c3wgrad", need, v64.device)
    check(lib().dg_c3_wgrad_tc(_ptr(v64, BF16, "v64"), _ptr(xp, BF16, "xp"), _ptr(dw, F32, "dw"), beta, B, S,
                               ws.data_ptr(), ws.numel(), _stream()), "dg_c3_wgrad_tc")


# ---- image-side 3-channel layers, direct SIMT kernels (debug reference) ---------------------------
def conv_c3_in_fwd(x, w, slope=0.2):
    B, _, S, _ = x.shape
    y = torch.empty(B, S // 2, S // 2, 64, dtype=BF16, device=x.device)
    check(lib().dg_conv_c3_in_fwd(_ptr(x, F32, "x"), _ptr(w, F32, "w"), _ptr(y), B, S, slope, _stream()), "dg_conv_c3_in_fwd")
    return y


def conv_c3_in_bwd(x, w, y, dy, dx=None, dx_accumulate=False, dw=None, slope=0.2):
    B, _, S, _ = x.shape
    check(lib().dg_conv_c3_in_bwd(_ptr(x, F32, "x"), _ptr(w, F32, "w"), _ptr(y, BF16, "y"), _ptr(dy, BF16, "dy"),
                                  _ptr(dx, F32, "dx"), int(dx_accumulate), _ptr(dw, F32, "dw"), B, S, slope, _stream()),
          "dg_conv_c3_in_bwd")


def convT_c3_out_fwd(x, w):
    B, Hs, _, _ = x.shape
    S = 2 * Hs
    y = torch.empty(B, 3, S, S, dtype=F32, device=x.device)
    check(lib().dg_convT_c3_out_fwd(_ptr(x, BF16, "x"), _ptr(w, F32, "w"), _ptr(y), B, S, _stream()), "dg_convT_c3_out_fwd")
    return y


def convT_c3_out_bwd(x, w, y, dy, want_dx=True, dw=None):
    B, _, S, _ = y.shape
    dx = torch.empty_like(x) if want_dx else None
    check(lib().dg_convT_c3_out_bwd(_ptr(x, BF16, "x"), _ptr(w, F32, "w"), _ptr(y, F32, "y"), _ptr(dy, F32, "dy"),
                                    _ptr(dx), _ptr(dw, F32, "dw"), B, S, _stream()), "dg_convT_c3_out_bwd")
    return dx


# ---- FC heads -------------------------------------------------------------------------------
def fc_down(big2d, wd2d, out_f32=False):
    """[B,K] x Wd[Ns,K] -> [B,Ns]."""
    B, K = big2d.shape
    Ns = wd2d.shape[0]
    out = torch.empty(B, Ns, dtype=F32 if out_f32 else BF16, device=big2d.device)
    check(lib().dg_fc_down(_ptr(big2d, BF16, "big"), _ptr(wd2d, BF16, "wd"), _ptr(out), int(out_f32), B, Ns, K, _stream()),
          "dg_fc_down")
    return out


def fc_up(small2d, wd2d):
    """[B,Ns] x Wd[Ns,K] -> [B,K] bf16."""
    B, Ns = small2d.shape
    K = wd2d.shape[1]
    f32 = small2d.dtype == F32
    out = torch.empty(B, K, dtype=BF16, device=small2d.device)
    check(lib().dg_fc_up(_ptr(small2d, F32 if f32 else BF16, "small"), int(f32), _ptr(wd2d, BF16, "wd"), _ptr(out), B, Ns, K,
                         _stream()), "dg_fc_up")
    return out


def fc_wgrad(small2d, big2d, dw, beta=1.0):
    """dw[Ns,C,4,4] = beta*dw + small^T . big, big = [B,16*C]."""
    B, Ns = small2d.shape
    C = big2d.shape[1] // 16
    if tuple(dw.shape) != (Ns, C, 4, 4):
        raise ValueError(f"dw shape {tuple(dw.shape)} != {(Ns, C, 4, 4)}")
    f32 = small2d.dtype == F32
    check(lib().dg_fc_wgrad(_ptr(small2d, F32 if f32 else BF16, "small"), int(f32), _ptr(big2d, BF16, "big"),
                            _ptr(dw, F32, "dw"), beta, B, Ns, C, _stream()), "dg_fc_wgrad")


# ---- BatchNorm + activation -----------------------------------------------------------------
def _bn_scratch(P, C, device):
    n = lib().dg_bn_scratch_floats(P, C)
    return scratch("bn", 4 * n, device)


def bn_stats(z2d, gamma, beta, running_mean=None, running_var=None, eps=1e-5, momentum=0.1):
    """Training statistics of z[P,C] -> stats fp32 [4,C] = (mean, invstd, scale, shift); updates running stats."""
    P, C = z2d.shape
    stats = torch.empty(4, C, dtype=F32, device=z2d.device)
    sc = _bn_scratch(P, C, z2d.device)
    check(lib().dg_bn_stats(_ptr(z2d, BF16, "z"), P, C, _ptr(gamma, F32, "gamma"), _ptr(beta, F32, "beta"), eps, momentum,
                            _ptr(stats), _ptr(running_mean, F32, "running_mean"), _ptr(running_var, F32, "running_var"),
                            sc.data_ptr(), _stream()), "dg_bn_stats")
    return stats


def bn_stats_finalize(part, P, gamma, beta, running_mean=None, running_var=None, eps=1e-5, momentum=0.1):
    """Statistics from partial sums produced by conv_down_stats / conv_up_stats -> stats fp32 [4,C]."""
    _, rows, C = part.shape
    stats = torch.empty(4, C, dtype=F32, device=part.device)
    check(lib().dg_bn_stats_finalize(_ptr(part, F32, "part"), rows, P, C, _ptr(gamma, F32, "gamma"), _ptr(beta, F32, "beta"),
                                     eps, momentum, _ptr(stats), _ptr(running_mean, F32, "running_mean"),
                                     _ptr(running_var, F32, "running_var"), _stream()), "dg_bn_stats_finalize")
    return stats


def bn_stats_acc(z2d):
    """Sums / sums of squares of z[P,C] added into a fresh zeroed arena accumulator -> acc fp32 [2, C]."""
    P, C = z2d.shape
    acc = current().arena.take(2 * C, z2d.device)[:2 * C].view(2, C)
    check(lib().dg_bn_stats_acc(_ptr(z2d, BF16, "z"), P, C, acc.data_ptr(), _stream()), "dg_bn_stats_acc")
    return acc


def bn_act_fwd_acc(z2d, acc, gamma, beta, act, slope=0.2, running_mean=None, running_var=None, eps=1e-5, momentum=0.1):
    """y = act(BN_train(z)) with the statistics derived from acc inside the kernel; returns (y, stats fp32 [4, C])."""
    P, C = z2d.shape
    y = torch.empty_like(z2d)
    stats = torch.empty(4, C, dtype=F32, device=z2d.device)
    check(lib().dg_bn_act_fwd_acc(_ptr(z2d, BF16, "z"), _ptr(y), P, C, _ptr(acc, F32, "acc"), _ptr(gamma, F32, "gamma"),
                                  _ptr(beta, F32, "beta"), eps, momentum, _ptr(stats), _ptr(running_mean, F32, "running_mean"),
                                  _ptr(running_var, F32, "running_var"), act, slope, _stream()), "dg_bn_act_fwd_acc")
    return y, stats


def bn_eval_stats(gamma, beta, running_mean, running_var, eps=1e-5):
    C = gamma.numel()
    stats = torch.zeros(4, C, dtype=F32, device=gamma.device)
    check(lib().dg_bn_eval_coeffs(_ptr(gamma, F32), _ptr(beta, F32), _ptr(running_mean, F32), _ptr(running_var, F32), eps,
                                  C, _ptr(stats), _stream()), "dg_bn_eval_coeffs")
    return stats


def bn_act_fwd(z2d, stats, act, slope=0.2, out=None):
    P, C = z2d.shape
    y = torch.empty_like(z2d) if out is None else out
    check(lib().dg_bn_act_fwd(_ptr(z2d, BF16, "z"), _ptr(y, BF16, "y"), P, C, _ptr(stats, F32, "stats"), act, slope, _stream()),
          "dg_bn_act_fwd")
    return y


def bn_act_bwd(dy2d, y2d, z2d, stats, gamma, act, slope=0.2, dgamma=None, dbeta=None, grad_beta=1.0, dy2=None,
               bcast=None, bcast_coef=0.0):
    """Backward of y = act(BN_train(z)); returns dz.  dgamma/dbeta (fp32 [C]) are updated as
    grad_beta*old + new when given."""
    P, C = z2d.shape
    dz = torch.empty_like(z2d)
    bcast_rows = 0
    if bcast is not None:
        bcast_rows = bcast.numel() // C
    if current().fold_stats:      # two launches: reduction into an arena accumulator, dx derives its coefficients
        acc2 = current().arena.take(2 * C, z2d.device)
        check(lib().dg_bn_act_bwd_acc(_ptr(dy2d, BF16, "dy"), _ptr(dy2, BF16, "dy2"), _ptr(bcast, F32, "bcast"), bcast_coef,
                                      bcast_rows, _ptr(z2d, BF16, "z"), _ptr(stats, F32, "stats"), _ptr(gamma, F32, "gamma"),
                                      P, C, act, slope, _ptr(dgamma, F32, "dgamma"), _ptr(dbeta, F32, "dbeta"), grad_beta,
                                      _ptr(dz), acc2.data_ptr(), _stream()), "dg_bn_act_bwd_acc")
        return dz
    coefs = torch.empty(3, C, dtype=F32, device=z2d.device)
    sc = _bn_scratch(P, C, z2d.device)
    check(lib().dg_bn_act_bwd(_ptr(dy2d, BF16, "dy"), _ptr(dy2, BF16, "dy2"), _ptr(bcast, F32, "bcast"), bcast_coef, bcast_rows,
                              _ptr(y2d, BF16, "y"), _ptr(z2d, BF16, "z"), _ptr(stats, F32, "stats"), _ptr(gamma, F32, "gamma"),
                              P, C, act, slope, _ptr(dgamma, F32, "dgamma"), _ptr(dbeta, F32, "dbeta"), grad_beta, _ptr(dz),
                              _ptr(coefs), sc.data_ptr(), _stream()), "dg_bn_act_bwd")
    return dz


# ---- losses -----------------------------------------------------------------------------------
def _red_scratch(device):
    return scratch("reduce", 4 * lib().dg_reduce_scratch_floats(), device)


def gan_bce_fwd(logit_real, logit_fake, out2):
    """out2[0] = dis_loss, out2[1] = gen_loss; returns (p_real, p_fake)."""
    B = logit_real.numel()
    p_real, p_fake = torch.empty_like(logit_real), torch.empty_like(logit_fake)
    check(lib().dg_gan_bce_fwd(_ptr(logit_real, F32), _ptr(logit_fake, F32), B, _ptr(p_real), _ptr(p_fake), _ptr(out2, F32),
                               _stream()), "dg_gan_bce_fwd")
    return p_real, p_fake


def gan_bce_bwd(p_real, p_fake, g_dis, g_gen, want_real=True, want_fake=True):
    B = p_real.numel()
    dr = torch.empty_like(p_real) if want_real else None
    df = torch.empty_like(p_fake) if want_fake else None
    check(lib().dg_gan_bce_bwd(_ptr(p_real, F32), _ptr(p_fake, F32), B, g_dis, g_gen, _ptr(dr), _ptr(df), _stream()),
          "dg_gan_bce_bwd")
    return dr, df


def sigmoid_fwd(x):
    y = torch.empty_like(x)
    check(lib().dg_sigmoid_fwd(_ptr(x, F32), _ptr(y), x.numel(), _stream()), "dg_sigmoid_fwd")
    return y


def sigmoid_bwd(y, dy):
    dx = torch.empty_like(y)
    check(lib().dg_sigmoid_bwd(_ptr(y, F32), _ptr(dy, F32), _ptr(dx), y.numel(), _stream()), "dg_sigmoid_bwd")
    return dx


def mse_fwd(a, b, out):
    """out[0] = mean((a-b)^2); out is a 1-element fp32 view."""
    check(lib().dg_mse_fwd(_ptr(a, F32, "a"), _ptr(b, F32, "b"), a.numel(), _ptr(out, F32), _red_scratch(a.device).data_ptr(),
                           _stream()), "dg_mse_fwd")


def mse_bwd(a, b, g, da=None, accumulate=False):
    """da (+)= g * d mean((a-b)^2) / da."""
    if da is None:
        da = torch.empty_like(a)
        accumulate = False
    check(lib().dg_mse_bwd(_ptr(a, F32, "a"), _ptr(b, F32, "b"), a.numel(), g, _ptr(da, F32, "da"), int(accumulate), _stream()),
          "dg_mse_bwd")
    return da


def fm_fwd(real, fake, out, accumulate, want_diff=True):
    """One feature map (NHWC bf16, same shape): out[0] (+)= mean((mean_b real - mean_b fake)^2); returns diff."""
    B = real.shape[0]
    n = real.numel() // B
    diff = torch.empty(n, dtype=F32, device=real.device) if want_diff else None
    check(lib().dg_fm_fwd(_ptr(real, BF16, "real"), _ptr(fake, BF16, "fake"), B, n, _ptr(diff), _ptr(out, F32),
                          int(accumulate), _red_scratch(real.device).data_ptr(), _stream()), "dg_fm_fwd")
    return diff


def fm_bwd(diff, B, shape, g):
    """d(g*fm)/d fake as a bf16 tensor of `shape` (negate g for the real branch)."""
    n = diff.numel()
    dfeat = torch.empty(shape, dtype=BF16, device=diff.device)
    check(lib().dg_fm_bwd(_ptr(diff, F32), B, n, g, _ptr(dfeat), _stream()), "dg_fm_bwd")
    return dfeat


# ---- input pipeline ---------------------------------------------------------------------------
def preprocess_u8(table, image_size, out=None):
    """table: device int64 [n,8] rows {src ptr, H, W, x0, crop_w, mode, 0, 0} (decoded uint8 HWC images in device memory)
    -> fp32 [n,3,S,S] in [0,1]: crop + (mode 1: 3x3 edge thickening) + bilinear resize + /255 + CHW, cv2-exact."""
    n = table.shape[0]
    if table.dim() != 2 or table.shape[1] != 8:
        raise ValueError(f"table must be [n,8] int64, got {tuple(table.shape)}")
    if out is None:
        out = torch.empty(n, 3, image_size, image_size, dtype=F32, device=table.device)
    elif tuple(out.shape) != (n, 3, image_size, image_size):
        raise ValueError(f"out shape {tuple(out.shape)} != {(n, 3, image_size, image_size)}")
    if n == 0:
        return out
    check(lib().dg_preprocess_u8(_ptr(table, torch.int64, "table"), n, image_size, _ptr(out, F32, "out"), _stream()),
          "dg_preprocess_u8")
    return out


# ---- optimiser --------------------------------------------------------------------------------
def adam_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, state, grad_scale=1.0):
    """One torch.optim.Adam-equivalent step on flat buffers; `state` = device fp32[4] (zero-initialised once; holds
    the step count and bias corrections so the call is CUDA-graph replayable)."""
    check(lib().dg_adam_step(_ptr(p, F32, "p"), _ptr(g, F32, "g"), _ptr(m, F32, "m"), _ptr(v, F32, "v"), p.numel(), lr, beta1,
                             beta2, eps, weight_decay, _ptr(state, F32, "state"), grad_scale, _stream()), "dg_adam_step")


def adam_tick(state, beta1, beta2):
    """Advance the device-side Adam step state once (the first half of adam_step)."""
    check(lib().dg_adam_tick(_ptr(state, F32, "state"), beta1, beta2, _stream()), "dg_adam_tick")


def adam_apply(p, g, m, v, lr, beta1, beta2, eps, weight_decay, state, grad_scale=1.0):
    """Adam update of a (contiguous, 16-byte aligned) range with the current step state (the second half of adam_step)."""
    check(lib().dg_adam_apply(_ptr(p, F32, "p"), _ptr(g, F32, "g"), _ptr(m, F32, "m"), _ptr(v, F32, "v"), p.numel(), lr, beta1,
                              beta2, eps, weight_decay, _ptr(state, F32, "state"), grad_scale, _stream()), "dg_adam_apply")
