"""bf16-storage emulation of the oracle nets (oracle; TEST INFRASTRUCTURE ONLY).

The B200 kernels compute every contraction with bf16 operands and fp32 accumulation and keep
activations / activation gradients in HBM as bf16; the reference computes
in fp32 (``model.py:5-225`` on stock torch).  Comparing the kernels with the fp32 oracle therefore
mixes two things: implementation errors and the stated precision choice.  This module separates
them.  It runs the *oracle's own layers and parameters* (``oracle/family.py``) in fp32 torch but
rounds to bf16 at exactly the points where the kernels store bf16:

    forward :  input image, every conv / convT weight, every conv output z, every BN+activation
               output y (conv1: after LeakyReLU; the generator's sigmoid image and the
               discriminator's logit stay fp32)
    backward:  the gradient leaving every conv / convT towards its input (dgrad output), the
               gradient leaving every BatchNorm backward (dz), the generator's pre-sigmoid gradient

so a kernel fed the emulation's own layer inputs must reproduce that layer's outputs to fp32
summation-order accuracy (the teacher-forced per-layer test in tests/test_parity_gpu.py), and
emulation vs fp32 oracle is the precision gap of bf16 storage itself.

Rounding noise CASCADES: a 1e-6 relative perturbation anywhere in the forward pass flips ~2e-4 of the
bf16 roundings of that tensor, each flip perturbs every output of the next convolution by ~1e-4, which
flips ~1 % of the next layer's roundings, and after three or four layers the rounding noise of two
runs is statistically independent (measured: identical end-to-end gradient distance for perturbations
of 1e-7 ... 1e-4).  Two correct bf16 implementations that differ only in fp32 summation order therefore
agree end to end no better than each agrees with fp32.  ``perturb`` reproduces exactly that: a second
emulation whose pre-rounding values carry 1e-6 relative noise is the *noise floor* against which the
kernels' end-to-end distance is judged.  Nothing here is a kernel: it is stock torch ops plus two
rounding ``autograd.Function``s.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


_PERTURB = 0.0      # relative noise added before every forward rounding (see set_perturb)


def set_perturb(eps):
    """Emulate a different fp32 summation order: multiply every value by (1 + eps*N(0,1)) before the forward bf16
    rounding (eps ~ 1e-6, the size of fp32 reassociation differences).  0 switches it off."""
    global _PERTURB
    _PERTURB = float(eps)


class _RoundFwd(torch.autograd.Function):
    """bf16 rounding in the forward pass, identity gradient."""

    @staticmethod
    def forward(ctx, x):
        if _PERTURB:
            x = x * (1.0 + _PERTURB * torch.randn_like(x))
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBwd(torch.autograd.Function):
    """identity in the forward pass, bf16 rounding of the gradient."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


rf, rb = _RoundFwd.apply, _RoundBwd.apply


def _conv(layer, x):
    w = rf(layer.weight)
    if isinstance(layer, nn.ConvTranspose2d):
        return F.conv_transpose2d(x, w, None, layer.stride, layer.padding)
    return F.conv2d(x, w, None, layer.stride, layer.padding)


def _bn_block(conv, bn, act, y_prev, round_input_grad=True, tape=None, act_name=None):
    """conv -> z (bf16) -> BN (module call: running statistics advance as usual) -> act -> y (bf16).
    ``tape`` (a list) records the layer's tensors for the teacher-forced per-layer tests: after backward,
    ``x_in.grad`` is the unrounded dgrad output, ``z32.grad`` the bf16 gradient the conv backward consumes,
    ``zr.grad`` the unrounded BatchNorm-backward output and ``y.grad`` the gradient entering BatchNorm backward."""
    x_in = rb(y_prev) if round_input_grad else y_prev.view_as(y_prev)
    z32 = _conv(conv, x_in)
    zr = rb(z32)
    z = rf(zr)
    y = act(bn(z))
    if tape is not None:
        for t in (x_in, z32, zr, y):
            if t.requires_grad:
                t.retain_grad()
        tape.append(dict(conv=conv, bn=bn, act=act_name, x_in=x_in, z32=z32, zr=zr, z=z, y=y))
    return rf(y)


def discriminator_forward(net, x, tape=None):
    """oracle.family.Discriminator.forward (model.py:38-69) with bf16 storage points."""
    lrelu = lambda t: F.leaky_relu(t, 0.2)
    # conv1 + LeakyReLU: the kernels apply the LeakyReLU derivative inside conv2's dgrad epilogue before the single
    # bf16 rounding, i.e. the rounded quantity is d/d(pre-activation)
    h = rf(lrelu(rb(_conv(net.conv1, rf(x)))))
    feats = []
    for k in range(2, net.n_down + 1):
        h = _bn_block(getattr(net, f"conv{k}"), getattr(net, f"bn{k}"), lrelu, h, round_input_grad=(k > 2), tape=tape,
                      act_name="lrelu")
        feats.append(h)
    logit = _conv(getattr(net, f"conv{net.n_down + 1}"), rb(h))
    return torch.sigmoid(logit), feats


def generator_forward(net, x, tape=None):
    """oracle.family.Generator.forward (model.py:217-225) with bf16 storage points."""
    lrelu = lambda t: F.leaky_relu(t, 0.2)
    enc, dec = list(net.encoder), list(net.decoder)
    h = rf(lrelu(rb(_conv(enc[0], rf(x)))))
    i, first = 2, True
    while i < len(enc):
        h = _bn_block(enc[i], enc[i + 1], lrelu, h, round_input_grad=not first, tape=tape, act_name="lrelu")
        first = False
        i += 3
    j = 0
    while j + 2 < len(dec):                       # convT + BN + ReLU blocks
        h = _bn_block(dec[j], dec[j + 1], F.relu, h, tape=tape, act_name="relu")
        j += 3
    pre = rb(_conv(dec[j], rb(h)))               # last convT: pre-sigmoid gradient is stored as bf16
    return torch.sigmoid(pre)


class Bf16Emulated(nn.Module):
    """Wraps an oracle net: same parameters and buffers (so an optimiser over ``parameters()`` updates the wrapped
    net), forward through the emulation above.  ``perturb`` > 0 makes this instance a noise-floor twin (see the module
    docstring): its forward roundings see 1e-6-scale relative noise."""

    def __init__(self, net, perturb=0.0):
        super().__init__()
        self.inner = net
        self.perturb = perturb
        self._is_gen = hasattr(net, "encoder")

    def forward(self, x):
        prev = _PERTURB
        set_perturb(self.perturb)
        try:
            return generator_forward(self.inner, x) if self._is_gen else discriminator_forward(self.inner, x)
        finally:
            set_perturb(prev)


def emulate(nets, perturb=0.0):
    return [Bf16Emulated(n, perturb) for n in nets]
