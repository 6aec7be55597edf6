"""Generator / Discriminator with the reference's nn.Module API, running on the sm_100a kernels.

Drop-in boundary (reference ``model.py``): same class names, constructor arguments, forward
outputs, ``state_dict`` keys/shapes and parameter order -- ``Discriminator`` mirrors
``model.py:5-69`` and ``Generator`` ``model.py:72-225`` -- so the reference entry points
(``image_translation.py:260-263``, ``inference.py:126-136``) can import these instead.  The
``torch.nn`` layer objects below are parameter containers only (they give the reference's default
initialisation and state-dict layout); the arithmetic is done by the hand-written kernels through
one ``torch.autograd.Function`` per network.  Internally activations are NHWC bf16 with fp32
accumulation and fp32 BatchNorm statistics.  There is no CPU or cuDNN path: a non-CUDA input raises.

``image_size`` (default 512, the only size the reference supports -- SURVEY.md F1) selects a member
of the depth-parametrised family: ``n_down = log2(S) - 2`` stride-2 layers with channels
``64 * 2**min(i, 5)``.
"""
import math
import os

import torch
import torch.nn as nn

from . import ops
from .ops import ACT_LRELU, ACT_RELU

LRELU_SLOPE = 0.2


def family_channels(image_size: int):
    n_down = int(round(math.log2(image_size))) - 2
    if image_size < 16 or 2 ** (n_down + 2) != image_size:
        raise ValueError(f"image_size must be a power of two >= 16, got {image_size}")
    return [64 * 2 ** min(i, 5) for i in range(n_down)]


# ---------------------------------------------------------------------------------------------
# packed bf16 weights, refreshed when the fp32 parameter changes
# ---------------------------------------------------------------------------------------------
class _PackedWeights:
    """bf16 GEMM-layout copies of the fp32 conv weights, keyed by parameter; re-packed (in place, so captured CUDA
    graphs keep valid pointers) when the parameter's version counter or storage changes.  Code that rewrites
    parameters behind autograd's back (the fused Adam kernel) calls ``refresh()`` right after the update."""

    def __init__(self):
        self._cache = {}

    def get(self, p, want_wd, want_wu):
        """GEMM layers: (Wd [Cs,16,Cb], Wu [Cb,16,Cs])."""
        return self._get(p, "gemm", want_wd, want_wu)

    def get_c3(self, p):
        """Image-side 3-channel layers: (wc [64,64], wu3 [16,16,64])."""
        return self._get(p, "c3", True, True)

    @staticmethod
    def _pack(p, kind, want_a, want_b, out=None):
        if kind == "c3":
            return ops.c3_pack_weights(p.detach(), out=out)
        return ops.pack_weights(p.detach(), want_a, want_b, out=out)

    def _get(self, p, kind, want_a, want_b):
        key = id(p)
        tag = (p._version, p.data_ptr())
        ent = self._cache.get(key)
        if ent is None:
            a, b = self._pack(p, kind, want_a, want_b)
            ent = [tag, a, b, p, kind]
            self._cache[key] = ent
        elif ent[0] != tag:
            self._pack(p, kind, True, True, out=(ent[1], ent[2]))
            ent[0] = tag
        return ent[1], ent[2]

    def refresh(self):
        """Re-pack every cached weight in place from the current fp32 values: one multi-tensor launch for the GEMM
        layouts plus one per image-side layer."""
        gemm = [e for e in self._cache.values() if e[4] == "gemm"]
        key = tuple((e[3].data_ptr(), 0 if e[1] is None else e[1].data_ptr(), 0 if e[2] is None else e[2].data_ptr())
                    for e in gemm)
        if gemm and getattr(self, "_table_key", None) != key:
            rows, end = [], 0
            for e, k in zip(gemm, key):
                Cs, Cb = e[3].shape[0], e[3].shape[1]
                end += 2 * Cs * Cb
                rows.append([k[0], k[1], k[2], Cs, Cb, end])
            self._table = torch.tensor(rows, dtype=torch.int64, device=gemm[0][3].device)
            self._table_key, self._table_total = key, end
        if gemm:
            ops.pack_weights_multi(self._table, len(gemm), self._table_total)
        for ent in self._cache.values():
            p = ent[3]
            if ent[4] != "gemm":
                self._pack(p, ent[4], True, True, out=(ent[1], ent[2]))
            ent[0] = (p._version, p.data_ptr())

    def invalidate(self):
        self._cache.clear()


def _new_packed():
    return _PackedWeights()


def _grad_buf(p):
    """(fp32 gradient buffer of a parameter, beta): the kernels compute grad = beta*grad + new.
    * inside an ``autograd.Function`` backward (module-level API) the parameter carries ``_dg_out``, a private buffer the
      Function returns to autograd -- so ``torch.autograd.grad``, parameter hooks and DDP see ordinary gradients;
    * the fused trainer marks its flat gradient views ``_dg_fresh`` instead of zero-filling them: the first write of an
      iteration then overwrites (beta 0);
    * otherwise a missing ``p.grad`` is allocated zeroed (beta 1, accumulate semantics)."""
    out = getattr(p, "_dg_out", None)
    if out is not None:
        if out[1]:
            out[1] = False
            return out[0], 0.0
        return out[0], 1.0
    if getattr(p, "_dg_fresh", False) and p.grad is not None:
        p._dg_fresh = False
        return p.grad, 0.0
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return p.grad, 1.0


def _with_param_outputs(mod, needs, need_wgrad, run):
    """Run a backward engine with every parameter's gradient directed into a private buffer and hand those buffers back
    as the ``autograd.Function``'s parameter gradients (None where autograd did not ask).  The engines write all
    parameters of a network whenever they write any."""
    params = list(mod.parameters())
    if not need_wgrad:
        return run(), (None,) * len(params)
    for p in params:
        p._dg_out = [torch.empty_like(p, memory_format=torch.contiguous_format), True]
    try:
        dx = run()
        if any(p._dg_out[1] for p in params):      # never written: would hand uninitialised memory to autograd
            raise RuntimeError("internal error: a parameter gradient was not produced by the backward engine")
        grads = tuple(p._dg_out[0] if need else None for p, need in zip(params, needs))
    finally:
        for p in params:
            p._dg_out = None
    return dx, grads


def _check_input(x, image_size, training):
    if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != image_size or x.shape[3] != image_size:
        raise RuntimeError(f"expected input of shape [B, 3, {image_size}, {image_size}], got {tuple(x.shape)}")
    if not x.is_cuda:
        raise RuntimeError("discogan_modernized_b200 runs on CUDA (sm_100a) only; move the module and input to the GPU")
    if x.dtype != torch.float32:
        raise RuntimeError(f"expected a float32 image, got {x.dtype}")
    if training and x.shape[0] < 2:
        raise ValueError("Expected more than 1 value per channel when training (BatchNorm over a 1x1 map needs B >= 2)")


class _BnSave:
    __slots__ = ("z", "y", "stats")

    def __init__(self, z, y, stats):
        self.z, self.y, self.stats = z, y, stats


def _bump_counters(mod, training):
    """BatchNorm's num_batches_tracked: when the trainer has flattened the counters of a network into one int64
    buffer (views keep the state-dict layout) they advance with a single add per forward; otherwise per layer.
    Returns True if _bn_act must bump each layer's own counter."""
    flat = getattr(mod, "_nbt_flat", None)
    if flat is None:
        return True
    if training:
        flat.add_(1)
    return False


def _conv_bn(conv_fn, conv_stats_fn, x, w, training):
    """Run a GEMM convolution; in training mode its epilogue also produces the BatchNorm sums: one accumulator pair
    [2, C] per layer (folded finalize, the default) or per-CTA partial rows [2, rows, C]."""
    ctx = ops.current()
    if training and ops._conv_impl == "tc" and ctx.fuse_stats:
        if ctx.fold_stats:
            return (ops.conv_down_acc if conv_fn is ops.conv_down else ops.conv_up_acc)(x, w)
        return conv_stats_fn(x, w)
    return conv_fn(x, w), None


def _conv_bn_act(conv_fn, conv_stats_fn, x, w, bn, act, training, bump):
    """conv -> BatchNorm -> activation.  Returns (z, y, stats); in eval mode with running statistics the BatchNorm affine
    and the activation are applied inside the conv epilogue (one bf16 rounding, no separate pass): z is None then."""
    if not training and bn.track_running_stats and ops._conv_impl == "tc" and ops.current().fold_eval_bn:
        stats = ops.bn_eval_stats(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.eps)
        return None, conv_fn(x, w, affine=(stats, act, LRELU_SLOPE)), stats
    z, part = _conv_bn(conv_fn, conv_stats_fn, x, w, training)
    y, stats = _bn_act(z, bn, act, training, part, bump)
    return z, y, stats


def _bn_act(z, bn, act, training, part=None, bump=True):
    """z: NHWC bf16 (any leading dims, channels last); part: partial sums from the producing conv's epilogue.
    Returns (y, stats)."""
    C = z.shape[-1]
    z2 = z.view(-1, C)
    if training or not bn.track_running_stats:
        if z2.shape[0] < 2:
            raise ValueError("Expected more than 1 value per channel when training")
        rm, rv = (bn.running_mean, bn.running_var) if (training and bn.track_running_stats) else (None, None)
        mom = bn.momentum if bn.momentum is not None else 0.1
        if ops.current().fold_stats and (part is None or part.dim() == 2):
            # folded finalize: the consumer derives the coefficients from the accumulated sums (of the conv epilogue, or of
            # a reduction pass for the 4x4 valid heads)
            acc = part if part is not None else ops.bn_stats_acc(z2)
            y, stats = ops.bn_act_fwd_acc(z2, acc, bn.weight.detach(), bn.bias.detach(), act, LRELU_SLOPE, rm, rv, bn.eps, mom)
            if rm is not None and bump:
                bn.num_batches_tracked.add_(1)
            return y.view(z.shape), stats
        if part is not None:
            stats = ops.bn_stats_finalize(part, z2.shape[0], bn.weight.detach(), bn.bias.detach(), rm, rv, bn.eps, mom)
        else:
            stats = ops.bn_stats(z2, bn.weight.detach(), bn.bias.detach(), rm, rv, bn.eps, mom)
        if rm is not None and bump:
            bn.num_batches_tracked.add_(1)
    else:
        stats = ops.bn_eval_stats(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.eps)
    y = ops.bn_act_fwd(z2, stats, act, LRELU_SLOPE).view(z.shape)
    return y, stats


def _bn_act_bwd(dy, sv, bn, act, need_wgrad, dy2=None, bcast=None, bcast_coef=0.0):
    C = sv.z.shape[-1]
    dgamma, gb = _grad_buf(bn.weight) if need_wgrad else (None, 1.0)
    dbeta, gb2 = _grad_buf(bn.bias) if need_wgrad else (None, 1.0)
    assert gb == gb2
    dz = ops.bn_act_bwd(dy.view(-1, C), sv.y.view(-1, C), sv.z.view(-1, C), sv.stats, bn.weight.detach(), act,
                        LRELU_SLOPE, dgamma, dbeta, gb, None if dy2 is None else dy2.view(-1, C), bcast, bcast_coef)
    return dz.view(sv.z.shape)


def _ready(grad_ready, *params):
    """Tell the data-parallel reducer that the kernels producing these parameters' gradients (this pass) are enqueued."""
    if grad_ready is not None:
        grad_ready(params)


def _conv1_backward(pk, weight, ctx, dz1, need_dx, need_wgrad, dx_out, dx_accumulate, grad_ready=None):
    """Backward of the image-side Conv2d(3,64,4,2,1)+LeakyReLU given dz1 = d(loss)/d(pre-activation)."""
    if need_wgrad:
        ctx.keep.append(dz1)
        with ops.wgrad_side():
            ops.c3_wgrad_tc(dz1, ctx.xp, *_grad_buf(weight))
        _ready(grad_ready, weight)
    if not need_dx:
        return None
    _, wu3 = pk.get_c3(weight)
    return ops.c3_up_tc(dz1, wu3, sigmoid=False, out=dx_out, accumulate=dx_accumulate and dx_out is not None)


# ---------------------------------------------------------------------------------------------
# Discriminator
# ---------------------------------------------------------------------------------------------
class _DiscCtx:
    __slots__ = ("xp", "y1", "bn", "B", "shape", "keep")


def discriminator_forward(mod, x, save=True):
    """-> (logit fp32 [B], feats: list of NHWC bf16 tensors, ctx)."""
    training = mod.training
    _check_input(x, mod.image_size, training)
    x = x.contiguous()
    pk = mod._packed
    bump = _bump_counters(mod, training)
    xp = ops.img_pad_nhwc4(x)
    wc, _ = pk.get_c3(mod.conv1.weight)
    y = ops.c3_down_tc(xp, wc, ACT_LRELU, LRELU_SLOPE)
    ctx = _DiscCtx()
    ctx.keep = []   # tensors used by kernels on the wgrad side stream: kept alive until the context is dropped
    ctx.xp, ctx.y1, ctx.bn, ctx.B, ctx.shape = (xp if save else None), y, [], x.shape[0], x.shape
    feats = []
    for k in range(2, mod.n_down + 1):
        conv, bn = getattr(mod, f"conv{k}"), getattr(mod, f"bn{k}")
        wd, _ = pk.get(conv.weight, True, True)
        z, y, stats = _conv_bn_act(ops.conv_down, ops.conv_down_stats, y, wd, bn, ACT_LRELU, training, bump)
        ctx.bn.append(_BnSave(z if save else None, y, stats))
        feats.append(y)
    head = getattr(mod, f"conv{mod.n_down + 1}")
    wd, _ = pk.get(head.weight, True, False)
    logit = ops.fc_down(y.view(x.shape[0], -1), wd.view(1, -1), out_f32=True).view(-1)
    return logit, feats, ctx


def discriminator_backward(mod, ctx, dlogit, dfeats=None, fm_bcast=None, need_dx=True, need_wgrad=True,
                           dx_out=None, dx_accumulate=False, grad_ready=None):
    """dlogit fp32 [B]; dfeats[i] optional bf16 NHWC grads on feats; fm_bcast[i] optional (diff, coef).
    Returns d(loss)/d(input image) (fp32 NCHW) or None.  ``grad_ready(params)`` is called, deepest layer first, as soon
    as the kernels writing those parameters' gradients are enqueued (the bucketed gradient exchange hangs off it)."""
    pk = mod._packed
    B = ctx.B
    head = getattr(mod, f"conv{mod.n_down + 1}")
    wd, _ = pk.get(head.weight, True, False)
    y_last = ctx.bn[-1].y
    dl = dlogit.contiguous().view(B, 1)
    if need_wgrad:
        ctx.keep.append(dl)
        with ops.wgrad_side():
            ops.fc_wgrad(dl, y_last.view(B, -1), *_grad_buf(head.weight))
        _ready(grad_ready, head.weight)
    dy = ops.fc_up(dl, wd.view(1, -1)).view(y_last.shape)
    for k in range(mod.n_down, 1, -1):
        i = k - 2
        conv, bn = getattr(mod, f"conv{k}"), getattr(mod, f"bn{k}")
        sv = ctx.bn[i]
        dy2 = dfeats[i] if dfeats is not None else None
        bc, coef = fm_bcast[i] if (fm_bcast is not None and fm_bcast[i] is not None) else (None, 0.0)
        dz = _bn_act_bwd(dy, sv, bn, ACT_LRELU, need_wgrad, dy2, bc, coef)
        y_prev = ctx.bn[i - 1].y if i > 0 else ctx.y1
        if need_wgrad:
            ctx.keep.append(dz)
            with ops.wgrad_side():
                ops.conv_wgrad(dz, y_prev, *_grad_buf(conv.weight))
            _ready(grad_ready, conv.weight, bn.weight, bn.bias)
        _, wu = pk.get(conv.weight, True, True)
        # the gradient reaching conv1's output also takes conv1's LeakyReLU derivative (fused in the epilogue)
        dy = ops.conv_up(dz, wu, mask=ctx.y1 if i == 0 else None, slope=LRELU_SLOPE)
    return _conv1_backward(pk, mod.conv1.weight, ctx, dy, need_dx, need_wgrad, dx_out, dx_accumulate, grad_ready)


class _DiscFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, *params):
        logit, feats, sv = discriminator_forward(mod, x, save=True)
        prob = ops.sigmoid_fwd(logit)
        ctx.mod, ctx.sv, ctx.prob = mod, sv, prob
        ctx.feat_dtype = mod.feature_dtype
        if mod.feature_dtype == torch.bfloat16:
            outs = [f.permute(0, 3, 1, 2) for f in feats]
        else:
            outs = [ops.nhwc_to_nchw_f32(f) for f in feats]
        return (prob.view(-1, 1, 1, 1),) + tuple(outs)

    @staticmethod
    def backward(ctx, dprob, *dfeats):
        mod, sv = ctx.mod, ctx.sv
        if not mod.training:
            raise RuntimeError("backward through eval-mode BatchNorm is not supported by the B200 kernels")
        B = sv.B
        if dprob is None:
            dlogit = torch.zeros(B, dtype=torch.float32, device=ctx.prob.device)
        else:
            dlogit = ops.sigmoid_bwd(ctx.prob, dprob.reshape(-1).float().contiguous())
        dfs = []
        for g in dfeats:
            if g is None:
                dfs.append(None)
            elif g.dtype == torch.bfloat16:
                dfs.append(g.permute(0, 2, 3, 1).contiguous())
            else:
                dfs.append(ops.nchw_f32_to_nhwc(g.float().contiguous()))
        need_dx = ctx.needs_input_grad[1]
        need_wgrad = any(ctx.needs_input_grad[2:])
        dx, pgrads = _with_param_outputs(mod, ctx.needs_input_grad[2:], need_wgrad,
                                         lambda: discriminator_backward(mod, sv, dlogit, dfs, None, need_dx, need_wgrad))
        return (None, dx) + pgrads


class Discriminator(nn.Module):
    """Reference ``model.py:5-69``.  forward(x[B,3,S,S]) -> (prob[B,1,1,1], [feat_2 .. feat_n_down])."""

    def __init__(self, image_size: int = 512):
        super().__init__()
        ch = family_channels(image_size)
        self.image_size = image_size
        self.n_down = len(ch)
        self.conv1 = nn.Conv2d(3, ch[0], 4, 2, 1, bias=False)
        self.relu1 = nn.LeakyReLU(LRELU_SLOPE, inplace=True)
        for i in range(1, self.n_down):
            k = i + 1
            setattr(self, f"conv{k}", nn.Conv2d(ch[i - 1], ch[i], 4, 2, 1, bias=False))
            setattr(self, f"bn{k}", nn.BatchNorm2d(ch[i]))
            setattr(self, f"relu{k}", nn.LeakyReLU(LRELU_SLOPE, inplace=True))
        setattr(self, f"conv{self.n_down + 1}", nn.Conv2d(ch[-1], 1, 4, 1, 0, bias=False))
        self.sigmoid = nn.Sigmoid()
        self.feature_dtype = torch.float32  # torch.bfloat16: zero-copy NHWC-strided views (used by the fused step)
        self._packed = _new_packed()

    def forward(self, input_tensor):
        if torch.is_grad_enabled() and (input_tensor.requires_grad or any(p.requires_grad for p in self.parameters())):
            out = _DiscFn.apply(self, input_tensor, *list(self.parameters()))
            return out[0], list(out[1:])
        logit, feats, _ = discriminator_forward(self, input_tensor, save=False)
        prob = ops.sigmoid_fwd(logit).view(-1, 1, 1, 1)
        if self.feature_dtype == torch.bfloat16:
            return prob, [f.permute(0, 3, 1, 2) for f in feats]
        return prob, [ops.nhwc_to_nchw_f32(f) for f in feats]


# ---------------------------------------------------------------------------------------------
# Generator
# ---------------------------------------------------------------------------------------------
class _GenCtx:
    __slots__ = ("xp", "y1", "enc", "head", "dec0", "dec", "out", "B", "shape", "keep")


def _gen_layers(mod):
    """Index helpers into the reference's Sequential numbering (model.py:79-143)."""
    n = mod.n_down
    enc_convs = [mod.encoder[0]] + [mod.encoder[2 + 3 * (i - 1)] for i in range(1, n)]
    enc_bns = [None] + [mod.encoder[3 + 3 * (i - 1)] for i in range(1, n)]
    head_conv, head_bn = mod.encoder[2 + 3 * (n - 1)], mod.encoder[3 + 3 * (n - 1)]
    dec_convs = [mod.decoder[3 * j] for j in range(n + 1)]
    dec_bns = [mod.decoder[3 * j + 1] for j in range(n)]
    return enc_convs, enc_bns, head_conv, head_bn, dec_convs, dec_bns


def generator_forward(mod, x, save=True):
    """-> (image fp32 NCHW [B,3,S,S] in (0,1), ctx)."""
    training = mod.training
    _check_input(x, mod.image_size, training)
    x = x.contiguous()
    B = x.shape[0]
    pk = mod._packed
    enc_convs, enc_bns, head_conv, head_bn, dec_convs, dec_bns = _gen_layers(mod)
    ctx = _GenCtx()
    ctx.keep = []
    bump = _bump_counters(mod, training)
    xp = ops.img_pad_nhwc4(x)
    ctx.B, ctx.xp, ctx.shape = B, (xp if save else None), x.shape
    wc, _ = pk.get_c3(enc_convs[0].weight)
    y = ops.c3_down_tc(xp, wc, ACT_LRELU, LRELU_SLOPE)
    ctx.y1, ctx.enc = y, []
    for conv, bn in zip(enc_convs[1:], enc_bns[1:]):
        wd, _ = pk.get(conv.weight, True, True)
        z, y, stats = _conv_bn_act(ops.conv_down, ops.conv_down_stats, y, wd, bn, ACT_LRELU, training, bump)
        ctx.enc.append(_BnSave(z if save else None, y, stats))
    # 4x4 valid conv to the 100-d bottleneck (model.py:107-109)
    wd, _ = pk.get(head_conv.weight, True, False)
    z = ops.fc_down(y.view(B, -1), wd.view(wd.shape[0], -1))
    y, stats = _bn_act(z, head_bn, ACT_LRELU, training, None, bump)
    ctx.head = _BnSave(z if save else None, y, stats)
    # ConvTranspose2d(100, C, 4, 1, 0) from the 1x1 bottleneck (model.py:114-116)
    wd0, _ = pk.get(dec_convs[0].weight, True, False)
    C = wd0.shape[2]
    z = ops.fc_up(y, wd0.view(wd0.shape[0], -1)).view(B, 4, 4, C)
    y, stats = _bn_act(z, dec_bns[0], ACT_RELU, training, None, bump)
    ctx.dec0 = _BnSave(z if save else None, y, stats)
    ctx.dec = []
    for conv, bn in zip(dec_convs[1:-1], dec_bns[1:]):
        _, wu = pk.get(conv.weight, True, True)
        z, y, stats = _conv_bn_act(ops.conv_up, ops.conv_up_stats, y, wu, bn, ACT_RELU, training, bump)
        ctx.dec.append(_BnSave(z if save else None, y, stats))
    _, wu3 = pk.get_c3(dec_convs[-1].weight)
    out = ops.c3_up_tc(y, wu3, sigmoid=True)
    ctx.out = out
    return out, ctx


def generator_backward(mod, ctx, dout, need_dx=True, need_wgrad=True, dx_out=None, dx_accumulate=False, dout2=None,
                       grad_ready=None):
    """dout (+ dout2): d(loss)/d(output image), fp32 NCHW.  Returns d(loss)/d(input image) or None.
    ``grad_ready(params)``: see discriminator_backward."""
    pk = mod._packed
    B = ctx.B
    enc_convs, enc_bns, head_conv, head_bn, dec_convs, dec_bns = _gen_layers(mod)
    dec_in = [ctx.dec0] + ctx.dec          # dec_in[j].y is the input of dec_convs[j+1]
    last = dec_convs[-1]
    dpre = ops.img_pad_nhwc4(dout.contiguous(), yimg=ctx.out, img2=dout2)   # d(loss)/d(pre-sigmoid), padded NHWC4 bf16
    if need_wgrad:
        ctx.keep.append(dpre)
        with ops.wgrad_side():
            ops.c3_wgrad_tc(dec_in[-1].y, dpre, *_grad_buf(last.weight))
        _ready(grad_ready, last.weight)
    wc_last, _ = pk.get_c3(last.weight)
    dy = ops.c3_down_tc(dpre, wc_last, ops.ACT_NONE)
    for j in range(len(ctx.dec), 0, -1):
        conv, bn = dec_convs[j], dec_bns[j]
        sv = ctx.dec[j - 1]
        dz = _bn_act_bwd(dy, sv, bn, ACT_RELU, need_wgrad)
        x_in = dec_in[j - 1].y
        if need_wgrad:
            ctx.keep.append(dz)
            with ops.wgrad_side():
                ops.conv_wgrad(x_in, dz, *_grad_buf(conv.weight))      # convT wgrad: small = input, big = dz
            _ready(grad_ready, conv.weight, bn.weight, bn.bias)
        wd, _ = pk.get(conv.weight, True, True)
        dy = ops.conv_down(dz, wd)                                      # convT dgrad
    # decoder.0: ConvTranspose2d(100, C, 4, 1, 0)
    dz = _bn_act_bwd(dy, ctx.dec0, dec_bns[0], ACT_RELU, need_wgrad)
    wd0, _ = pk.get(dec_convs[0].weight, True, False)
    if need_wgrad:
        ctx.keep.append(dz)
        with ops.wgrad_side():
            ops.fc_wgrad(ctx.head.y, dz.view(B, -1), *_grad_buf(dec_convs[0].weight))
        _ready(grad_ready, dec_convs[0].weight, dec_bns[0].weight, dec_bns[0].bias)
    dy = ops.fc_down(dz.view(B, -1), wd0.view(wd0.shape[0], -1))
    # encoder head: Conv2d(C, 100, 4, 1, 0)
    dz = _bn_act_bwd(dy, ctx.head, head_bn, ACT_LRELU, need_wgrad)
    y_last = ctx.enc[-1].y
    if need_wgrad:
        ctx.keep.append(dz)
        with ops.wgrad_side():
            ops.fc_wgrad(dz, y_last.view(B, -1), *_grad_buf(head_conv.weight))
        _ready(grad_ready, head_conv.weight, head_bn.weight, head_bn.bias)
    wdh, _ = pk.get(head_conv.weight, True, False)
    dy = ops.fc_up(dz, wdh.view(wdh.shape[0], -1)).view(y_last.shape)
    for i in range(len(ctx.enc), 0, -1):
        conv, bn = enc_convs[i], enc_bns[i]
        sv = ctx.enc[i - 1]
        dz = _bn_act_bwd(dy, sv, bn, ACT_LRELU, need_wgrad)
        y_prev = ctx.enc[i - 2].y if i > 1 else ctx.y1
        if need_wgrad:
            ctx.keep.append(dz)
            with ops.wgrad_side():
                ops.conv_wgrad(dz, y_prev, *_grad_buf(conv.weight))
            _ready(grad_ready, conv.weight, bn.weight, bn.bias)
        _, wu = pk.get(conv.weight, True, True)
        dy = ops.conv_up(dz, wu, mask=ctx.y1 if i == 1 else None, slope=LRELU_SLOPE)
    return _conv1_backward(pk, enc_convs[0].weight, ctx, dy, need_dx, need_wgrad, dx_out, dx_accumulate, grad_ready)


class _GenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, *params):
        out, sv = generator_forward(mod, x, save=True)
        ctx.mod, ctx.sv = mod, sv
        return out

    @staticmethod
    def backward(ctx, dout):
        mod = ctx.mod
        if not mod.training:
            raise RuntimeError("backward through eval-mode BatchNorm is not supported by the B200 kernels")
        need_dx = ctx.needs_input_grad[1]
        need_wgrad = any(ctx.needs_input_grad[2:])
        dx, pgrads = _with_param_outputs(mod, ctx.needs_input_grad[2:], need_wgrad,
                                         lambda: generator_backward(mod, ctx.sv, dout.float(), need_dx, need_wgrad))
        return (None, dx) + pgrads


class Generator(nn.Module):
    """Reference ``model.py:72-225``: forward(x[B,3,S,S]) -> image [B,3,S,S].  ``extra_layers`` is accepted and
    has no effect, exactly as in the reference (both branches build the same layers)."""

    def __init__(self, extra_layers: bool = False, image_size: int = 512):
        super().__init__()
        ch = family_channels(image_size)
        self.image_size = image_size
        self.n_down = len(ch)
        self.main = None
        enc = [nn.Conv2d(3, ch[0], 4, 2, 1, bias=False), nn.LeakyReLU(LRELU_SLOPE, inplace=True)]
        for i in range(1, len(ch)):
            enc += [nn.Conv2d(ch[i - 1], ch[i], 4, 2, 1, bias=False), nn.BatchNorm2d(ch[i]),
                    nn.LeakyReLU(LRELU_SLOPE, inplace=True)]
        enc += [nn.Conv2d(ch[-1], 100, 4, 1, 0, bias=False), nn.BatchNorm2d(100), nn.LeakyReLU(LRELU_SLOPE, inplace=True)]
        dec = [nn.ConvTranspose2d(100, ch[-1], 4, 1, 0, bias=False), nn.BatchNorm2d(ch[-1]), nn.ReLU(True)]
        for i in range(len(ch) - 1, 0, -1):
            dec += [nn.ConvTranspose2d(ch[i], ch[i - 1], 4, 2, 1, bias=False), nn.BatchNorm2d(ch[i - 1]), nn.ReLU(True)]
        dec += [nn.ConvTranspose2d(ch[0], 3, 4, 2, 1, bias=False), nn.Sigmoid()]
        self.encoder = nn.Sequential(*enc)
        self.decoder = nn.Sequential(*dec)
        self._packed = _new_packed()

    def forward(self, input):
        if torch.is_grad_enabled() and (input.requires_grad or any(p.requires_grad for p in self.parameters())):
            return _GenFn.apply(self, input, *list(self.parameters()))
        out, _ = generator_forward(self, input, save=False)
        return out
