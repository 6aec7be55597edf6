"""Per-kernel parity on the GPU: every C-ABI kernel against the same op in stock fp32 PyTorch
(cuDNN with TF32 disabled) on identical inputs.  bf16-operand / fp32-accumulate kernels are compared on
bf16-rounded inputs, so the only differences are accumulation order and the final bf16 rounding of the
output (rel-L2 tolerance 4e-3, SURVEY.md F9 measured 2.4e-3 for bf16 operands + fp32 outputs)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16 = torch.bfloat16


@pytest.fixture(scope="module", autouse=True)
def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from discogan_modernized_b200 import ops
    ops.device_check()
    yield


def ops_mod():
    from discogan_modernized_b200 import ops
    return ops


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, generator=g, device="cuda") * scale


def to_nhwc_bf16(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(BF16)


def to_nchw_f32(x_nhwc):
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous()


def test_pack_weights():
    ops = ops_mod()
    for Cs, Cb in ((128, 64), (100, 512), (1, 256)):
        w = rnd(Cs, Cb, 4, 4, seed=Cs)
        wd, wu = ops.pack_weights(w)
        assert torch.equal(wd, w.view(Cs, Cb, 16).permute(0, 2, 1).contiguous().to(BF16))
        assert torch.equal(wu, w.view(Cs, Cb, 16).permute(1, 2, 0).contiguous().to(BF16))


def test_layout_roundtrip():
    ops = ops_mod()
    x = rnd(3, 40, 6, 10, seed=1)               # NCHW fp32, ragged sizes
    y = ops.nchw_f32_to_nhwc(x)
    assert torch.equal(y, to_nhwc_bf16(x))
    z = ops.nhwc_to_nchw_f32(y)
    assert torch.equal(z, x.to(BF16).float())


CONV_SHAPES = [
    # B, H(big), Cb, Cs
    (2, 8, 64, 128),      # deepest 64^2 shape class: 8->4, batch padded inside a tile
    (4, 16, 128, 256),
    (3, 32, 64, 128),     # odd batch
    (2, 64, 64, 128),     # Wt < W rows
    (1, 256, 64, 128),    # the 512^2 conv2 geometry (Ws = 128 = one tile row)
    (9, 8, 256, 512),     # batch tiles with remainder, N=512 -> 2 n-tiles
    (2, 16, 512, 1024),
]


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("B,H,Cb,Cs", CONV_SHAPES)
def test_conv_down(impl, B, H, Cb, Cs):
    ops = ops_mod()
    if impl == "simt" and B * H * H * Cs * Cb > 2e10 / 16:
        pytest.skip("simt reference kernel too slow for this shape")
    x = rnd(B, Cb, H, H, seed=1).to(BF16)
    w = (rnd(Cs, Cb, 4, 4, seed=2) / (4 * Cb ** 0.5)).to(BF16).float()
    ref = F.conv2d(x.float(), w, stride=2, padding=1)
    wd, _ = ops.pack_weights(w)
    ops.set_conv_impl(impl)
    try:
        out = ops.conv_down(to_nhwc_bf16(x.float()), wd)
    finally:
        ops.set_conv_impl("tc")
    torch.cuda.synchronize()
    assert rel_l2(to_nchw_f32(out), ref) < 4e-3


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("B,H,Cb,Cs", CONV_SHAPES)
def test_conv_up(impl, B, H, Cb, Cs):
    ops = ops_mod()
    Hs = H // 2
    if impl == "simt" and B * H * H * Cs * Cb > 2e10 / 4:
        pytest.skip("simt reference kernel too slow for this shape")
    s = rnd(B, Cs, Hs, Hs, seed=3).to(BF16)
    w = (rnd(Cs, Cb, 4, 4, seed=4) / (2 * Cs ** 0.5)).to(BF16).float()
    ref = F.conv_transpose2d(s.float(), w, stride=2, padding=1)
    _, wu = ops.pack_weights(w)
    ops.set_conv_impl(impl)
    try:
        out = ops.conv_up(to_nhwc_bf16(s.float()), wu)
    finally:
        ops.set_conv_impl("tc")
    torch.cuda.synchronize()
    assert rel_l2(to_nchw_f32(out), ref) < 4e-3


@pytest.mark.parametrize("impl", ["simt", "tc", "tc_cluster"])
@pytest.mark.parametrize("B,H,Cb,Cs", CONV_SHAPES)
def test_conv_wgrad(impl, B, H, Cb, Cs):
    """tc = the default plan (cta_group::2 pair kernel where Cs % 256 == 0), tc_cluster = the multicast-cluster kernel
    (dg_conv_opts.wgrad_pair = 0)."""
    ops = ops_mod()
    Hs = H // 2
    if impl == "tc_cluster":
        ops.set_conv_tiling(0, -1, wgrad_pair=0)
        try:
            return test_conv_wgrad("tc", B, H, Cb, Cs)
        finally:
            ops.set_conv_tiling(0, -1, -1)
    if impl == "simt" and B * H * H * Cs * Cb > 2e10 / 16:
        pytest.skip("simt reference kernel too slow for this shape")
    big = rnd(B, Cb, H, H, seed=5).to(BF16)
    small = rnd(B, Cs, Hs, Hs, seed=6).to(BF16)
    ref = torch.nn.grad.conv2d_weight(big.float(), (Cs, Cb, 4, 4), small.float(), stride=2, padding=1)
    dw = torch.full((Cs, Cb, 4, 4), 0.5, device="cuda")
    ops.set_conv_impl(impl)
    try:
        ops.conv_wgrad(to_nhwc_bf16(small.float()), to_nhwc_bf16(big.float()), dw, beta=1.0)
    finally:
        ops.set_conv_impl("tc")
    torch.cuda.synchronize()
    assert rel_l2(dw - 0.5, ref) < 1e-3
    dw2 = torch.full((Cs, Cb, 4, 4), 7.0, device="cuda")
    ops.conv_wgrad(to_nhwc_bf16(small.float()), to_nhwc_bf16(big.float()), dw2, beta=0.0)
    assert rel_l2(dw2, ref) < 1e-3


@pytest.mark.parametrize("B,S", [(2, 16), (3, 64), (1, 128)])
def test_conv_c3_in(B, S):
    ops = ops_mod()
    x = torch.rand(B, 3, S, S, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    w = rnd(64, 3, 4, 4, seed=2, scale=0.2)
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    ref = F.leaky_relu(F.conv2d(xr, wr, stride=2, padding=1), 0.2)
    y = ops.conv_c3_in_fwd(x, w)
    assert rel_l2(to_nchw_f32(y), ref) < 4e-3
    dy = rnd(B, 64, S // 2, S // 2, seed=3).to(BF16)
    # backward reference evaluated with the kernel's own (bf16) activation signs
    mask = torch.where(to_nchw_f32(y) > 0, 1.0, 0.2)
    z = F.conv2d(xr, wr, stride=2, padding=1)
    z.backward(dy.float() * mask)
    dx = torch.empty_like(x)
    dw = torch.zeros_like(w)
    ops.conv_c3_in_bwd(x, w, y, to_nhwc_bf16(dy.float()), dx, False, dw)
    assert rel_l2(dx, xr.grad) < 1e-4
    assert rel_l2(dw, wr.grad) < 1e-4
    dx2 = dx.clone()
    ops.conv_c3_in_bwd(x, w, y, to_nhwc_bf16(dy.float()), dx2, True, None)
    assert rel_l2(dx2, 2 * xr.grad) < 1e-4


@pytest.mark.parametrize("B,S", [(2, 16), (3, 64), (1, 128)])
def test_convT_c3_out(B, S):
    ops = ops_mod()
    x = rnd(B, 64, S // 2, S // 2, seed=1).to(BF16)
    w = rnd(64, 3, 4, 4, seed=2, scale=0.1)
    xr = x.float().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    ref = torch.sigmoid(F.conv_transpose2d(xr, wr, stride=2, padding=1))
    y = ops.convT_c3_out_fwd(to_nhwc_bf16(x.float()), w)
    assert rel_l2(y, ref) < 1e-5
    dy = rnd(B, 3, S, S, seed=3)
    ref.backward(dy)
    dw = torch.zeros_like(w)
    dx = ops.convT_c3_out_bwd(to_nhwc_bf16(x.float()), w, y, dy, True, dw)
    assert rel_l2(to_nchw_f32(dx), xr.grad) < 4e-3
    assert rel_l2(dw, wr.grad) < 1e-4


@pytest.mark.parametrize("B,Ns,C", [(4, 100, 512), (64, 100, 512), (5, 1, 512), (32, 100, 2048), (2, 1, 2048)])
def test_fc_heads(B, Ns, C):
    ops = ops_mod()
    K = 16 * C
    w = (rnd(Ns, C, 4, 4, seed=1) / K ** 0.5)
    wd, _ = ops.pack_weights(w, True, False)
    wd2 = wd.view(Ns, K)
    big = rnd(B, 4, 4, C, seed=2).to(BF16)
    ref = F.conv2d(to_nchw_f32(big), w.to(BF16).float()).view(B, Ns)       # 4x4 valid conv
    out = ops.fc_down(big.view(B, K), wd2, out_f32=True)
    assert rel_l2(out, ref) < 1e-4
    outb = ops.fc_down(big.view(B, K), wd2, out_f32=False)
    assert rel_l2(outb, ref) < 4e-3
    small = rnd(B, Ns, seed=3)
    refu = F.conv_transpose2d(small.to(BF16).float().view(B, Ns, 1, 1), w.to(BF16).float())   # [B,C,4,4]
    for s in (small.to(BF16), small.to(BF16).float()):
        up = ops.fc_up(s.contiguous(), wd2).view(B, 4, 4, C)
        assert rel_l2(to_nchw_f32(up), refu) < 4e-3
    refw = torch.einsum("bn,bhwc->nchw", small.to(BF16).float(), big.float())
    dw = torch.full((Ns, C, 4, 4), 1.0, device="cuda")
    ops.fc_wgrad(small.to(BF16), big.view(B, K), dw, beta=1.0)
    assert rel_l2(dw - 1.0, refw) < 1e-4
    dw0 = torch.full((Ns, C, 4, 4), 3.0, device="cuda")
    ops.fc_wgrad(small.to(BF16).float(), big.view(B, K), dw0, beta=0.0)
    assert rel_l2(dw0, refw) < 1e-4


@pytest.mark.parametrize("P,C,act", [(64 * 16 * 16, 128, 1), (37, 256, 2), (8, 100, 1), (4 * 16, 2048, 2),
                                      (2 * 128 * 128, 64, 2)])
def test_bn_act_fwd_bwd(P, C, act):
    ops = ops_mod()
    z = (rnd(P, C, seed=1) * 1.5 + 0.3).to(BF16)
    gamma = 1.0 + 0.2 * rnd(C, seed=2)
    beta = 0.1 * rnd(C, seed=3)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    rm_ref, rv_ref = rm.clone(), rv.clone()
    zr = z.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    pre = F.batch_norm(zr, rm_ref, rv_ref, gr, br, training=True, momentum=0.1, eps=1e-5)
    ref = F.leaky_relu(pre, 0.2) if act == 1 else F.relu(pre)
    stats = ops.bn_stats(z, gamma, beta, rm, rv)
    y = ops.bn_act_fwd(z, stats, act)
    assert torch.allclose(stats[0], zr.detach().mean(0), atol=1e-5, rtol=1e-5)
    assert torch.allclose(rm, rm_ref, atol=1e-6, rtol=1e-5) and torch.allclose(rv, rv_ref, atol=1e-6, rtol=1e-4)
    assert rel_l2(y, ref) < 4e-3
    dy = rnd(P, C, seed=4).to(BF16)
    dy2 = rnd(P, C, seed=5).to(BF16)
    rows = 4 if P % 4 == 0 else 1
    bc = rnd(rows, C, seed=6)
    coef = 0.37
    g_total = dy.float() + dy2.float() + coef * bc.repeat(P // rows, 1)
    # reference backward with the kernel's own activation signs (sign of the bf16 output)
    slope = 0.2 if act == 1 else 0.0
    mask = torch.where(y.float() > 0, 1.0, slope)
    pre.backward(g_total * mask)
    dgamma = torch.full((C,), 2.0, device="cuda")
    dbeta = torch.full((C,), -1.0, device="cuda")
    dz = ops.bn_act_bwd(dy, y, z, stats, gamma, act, 0.2, dgamma, dbeta, 1.0, dy2, bc.contiguous(), coef)
    assert rel_l2(dgamma - 2.0, gr.grad) < 2e-3
    assert rel_l2(dbeta + 1.0, br.grad) < 2e-3
    assert rel_l2(dz, zr.grad) < 6e-3


def test_bn_eval():
    ops = ops_mod()
    P, C = 50, 128
    z = rnd(P, C, seed=1).to(BF16)
    gamma, beta = 1 + 0.1 * rnd(C, seed=2), 0.1 * rnd(C, seed=3)
    rm, rv = 0.2 * rnd(C, seed=4), 1 + 0.3 * torch.rand(C, device="cuda")
    ref = F.relu(F.batch_norm(z.float(), rm, rv, gamma, beta, training=False, eps=1e-5))
    y = ops.bn_act_fwd(z, ops.bn_eval_stats(gamma, beta, rm, rv), 2)
    assert rel_l2(y, ref) < 4e-3


def test_gan_bce():
    ops = ops_mod()
    B = 64
    lr_ = rnd(B, seed=1) * 3
    lf_ = rnd(B, seed=2) * 3
    lf_[0], lf_[1], lr_[2] = -200.0, 200.0, -150.0        # saturated sigmoid: the -100 clamp and zero-gradient cases
    a = lr_.clone().requires_grad_(True)
    b = lf_.clone().requires_grad_(True)
    bce = torch.nn.BCELoss()
    pr, pf = torch.sigmoid(a).view(B, 1), torch.sigmoid(b).view(B, 1)
    dis = 0.5 * (bce(pr, torch.ones_like(pr)) + bce(pf, torch.zeros_like(pf)))
    gen = bce(pf, torch.ones_like(pf))
    out = torch.zeros(2, device="cuda")
    p_real, p_fake = ops.gan_bce_fwd(lr_, lf_, out)
    assert torch.allclose(out, torch.stack([dis, gen]).detach(), rtol=1e-5, atol=1e-6)
    (0.7 * dis + 0.3 * gen).backward()
    dr, df = ops.gan_bce_bwd(p_real, p_fake, 0.7, 0.3)
    assert torch.allclose(dr, a.grad, rtol=1e-4, atol=1e-7)
    assert torch.allclose(df, b.grad, rtol=1e-4, atol=1e-7)


def test_mse_fm():
    ops = ops_mod()
    a, b = rnd(3, 3, 64, 64, seed=1), rnd(3, 3, 64, 64, seed=2)
    out = torch.zeros(1, device="cuda")
    ops.mse_fwd(a, b, out)
    assert torch.allclose(out[0], F.mse_loss(a, b), rtol=1e-5)
    ar = a.clone().requires_grad_(True)
    (0.3 * F.mse_loss(ar, b)).backward()
    da = ops.mse_bwd(a, b, 0.3)
    assert torch.allclose(da, ar.grad, rtol=1e-5, atol=1e-9)
    da2 = ops.mse_bwd(a, b, 0.3, da.clone(), accumulate=True)
    assert torch.allclose(da2, 2 * ar.grad, rtol=1e-5, atol=1e-9)

    Bn = 6
    real = rnd(Bn, 8, 8, 128, seed=3).to(BF16)
    fake = rnd(Bn, 8, 8, 128, seed=4).to(BF16)
    fr = fake.float().requires_grad_(True)
    d = real.float().mean(0) - fr.mean(0)
    fm_ref = (d * d).mean()
    out = torch.full((1,), 5.0, device="cuda")
    diff = ops.fm_fwd(real, fake, out, accumulate=True)
    assert torch.allclose(out[0] - 5.0, fm_ref.detach(), rtol=1e-4)
    (0.9 * fm_ref).backward()
    dfeat = ops.fm_bwd(diff, Bn, fake.shape, 0.9)
    assert rel_l2(dfeat, fr.grad) < 4e-3


def test_adam_matches_torch():
    ops = ops_mod()
    n = 10008
    p = rnd(n, seed=1)
    pt = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pt], lr=2e-4, betas=(0.5, 0.999), weight_decay=1e-5)
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    state = torch.zeros(4, device="cuda")
    for step in range(1, 4):
        g = rnd(n, seed=10 + step)
        pt.grad = g.clone()
        opt.step()
        ops.adam_step(p, g * 4.0, m, v, 2e-4, 0.5, 0.999, 1e-8, 1e-5, state, grad_scale=0.25)
        assert torch.allclose(p, pt.detach(), rtol=1e-5, atol=1e-7), step
        assert int(state[0]) == step


# ---- tensor-core path of the image-side 3-channel layers ---------------------------------------------------
@pytest.mark.parametrize("B,S", [(2, 16), (3, 64), (5, 8), (1, 256)])
def test_c3_tc_down_up_wgrad(B, S):
    ops = ops_mod()
    gen = torch.Generator(device="cuda").manual_seed(S)
    x = torch.rand(B, 3, S, S, device="cuda", generator=gen)
    w = rnd(64, 3, 4, 4, seed=2, scale=0.2)
    wc, wu3 = ops.c3_pack_weights(w)
    wq = w.to(BF16).float()
    xq = x.to(BF16).float()
    # padded NHWC4 repack
    xp = ops.img_pad_nhwc4(x)
    assert xp.shape == (B, S + 2, S + 2, 4)
    assert torch.equal(xp[:, 1:-1, 1:-1, :3].float(), xq.permute(0, 2, 3, 1))
    assert float(xp[:, 0].abs().sum() + xp[:, -1].abs().sum() + xp[:, :, 0].abs().sum() + xp[:, :, -1].abs().sum()
                 + xp[..., 3].abs().sum()) == 0.0
    # down: conv1 forward
    ref = F.leaky_relu(F.conv2d(xq, wq, stride=2, padding=1), 0.2)
    y = ops.c3_down_tc(xp, wc, 1, 0.2)
    assert rel_l2(to_nchw_f32(y), ref) < 4e-3
    # down with the sigmoid derivative folded into the repack (dgrad of the last ConvTranspose2d)
    yimg = torch.rand(B, 3, S, S, device="cuda", generator=gen)
    dpre = (x * yimg * (1 - yimg)).to(BF16).float()
    dp = ops.img_pad_nhwc4(x, yimg)
    half = 0.5 * x
    assert torch.equal(ops.img_pad_nhwc4(half, yimg, img2=half), dp)          # two summed gradient contributions
    ref2 = F.conv2d(dpre, wq, stride=2, padding=1)
    assert rel_l2(to_nchw_f32(ops.c3_down_tc(dp, wc, 0)), ref2) < 4e-3
    # up: last ConvTranspose2d forward (+sigmoid) and accumulate mode
    s64 = rnd(B, 64, S // 2, S // 2, seed=3).to(BF16)
    refu = F.conv_transpose2d(s64.float(), wq, stride=2, padding=1)
    img = ops.c3_up_tc(to_nhwc_bf16(s64.float()), wu3, sigmoid=True)
    assert torch.allclose(img, torch.sigmoid(refu), atol=2e-5, rtol=1e-4)
    acc = torch.full_like(img, 0.25)
    ops.c3_up_tc(to_nhwc_bf16(s64.float()), wu3, sigmoid=False, out=acc, accumulate=True)
    assert rel_l2(acc - 0.25, refu) < 1e-4
    # wgrad
    refw = torch.nn.grad.conv2d_weight(xq, (64, 3, 4, 4), s64.float(), stride=2, padding=1)
    dw = torch.full((64, 3, 4, 4), 0.5, device="cuda")
    ops.c3_wgrad_tc(to_nhwc_bf16(s64.float()), xp, dw, beta=1.0)
    assert rel_l2(dw - 0.5, refw) < 1e-3
    dw0 = torch.full((64, 3, 4, 4), 9.0, device="cuda")
    ops.c3_wgrad_tc(to_nhwc_bf16(s64.float()), xp, dw0, beta=0.0)
    assert rel_l2(dw0, refw) < 1e-3


def test_conv_up_masked():
    ops = ops_mod()
    B, Hs, Cs, Cb = 3, 8, 128, 64
    s = rnd(B, Cs, Hs, Hs, seed=3).to(BF16)
    w = (rnd(Cs, Cb, 4, 4, seed=4) / (2 * Cs ** 0.5)).to(BF16).float()
    mask = rnd(B, 2 * Hs, 2 * Hs, Cb, seed=5).to(BF16)
    ref = F.conv_transpose2d(s.float(), w, stride=2, padding=1) * torch.where(to_nchw_f32(mask) > 0, 1.0, 0.2)
    _, wu = ops.pack_weights(w)
    out = ops.conv_up(to_nhwc_bf16(s.float()), wu, mask=mask, slope=0.2)
    assert rel_l2(to_nchw_f32(out), ref) < 4e-3


@pytest.mark.parametrize("B,H,Cb,Cs", [(2, 8, 64, 128), (3, 32, 64, 128), (9, 8, 256, 512), (2, 16, 512, 1024)])
def test_conv_fused_bn_stats(B, H, Cb, Cs):
    """Partial sums from the GEMM epilogue -> same statistics as the stand-alone reduction over the stored output."""
    ops = ops_mod()
    x = rnd(B, H, H, Cb, seed=1).to(BF16)
    w = rnd(Cs, Cb, 4, 4, seed=2) / (4 * Cb ** 0.5)
    wd, wu = ops.pack_weights(w)
    gamma, beta = 1 + 0.1 * rnd(Cs, seed=3), 0.1 * rnd(Cs, seed=4)
    z, part = ops.conv_down_stats(x, wd)
    assert torch.equal(z, ops.conv_down(x, wd))
    P = z.numel() // Cs
    rm, rv = torch.zeros(Cs, device="cuda"), torch.ones(Cs, device="cuda")
    rm2, rv2 = rm.clone(), rv.clone()
    st = ops.bn_stats_finalize(part, P, gamma, beta, rm, rv)
    ref = ops.bn_stats(z.view(P, Cs), gamma, beta, rm2, rv2)
    # fp32 accumulators vs the bf16-rounded stored output: the rounding noise (2^-9 relative per element) averages
    # down only as 1/sqrt(P), and P is as small as 32 here
    assert torch.allclose(st[0], ref[0], atol=3e-3, rtol=1e-2)           # mean
    assert torch.allclose(st[1], ref[1], rtol=1e-2)                      # invstd
    assert torch.allclose(rv, rv2, rtol=1e-2) and torch.allclose(rm, rm2, atol=3e-4, rtol=1e-2)
    # transposed conv: statistics over all four output parities
    s = rnd(B, H // 2, H // 2, Cs, seed=5).to(BF16)
    gb, bb = 1 + 0.1 * rnd(Cb, seed=6), 0.1 * rnd(Cb, seed=7)
    zu, partu = ops.conv_up_stats(s, wu)
    assert torch.equal(zu, ops.conv_up(s, wu))
    Pu = zu.numel() // Cb
    stu = ops.bn_stats_finalize(partu, Pu, gb, bb)
    refu = ops.bn_stats(zu.view(Pu, Cb), gb, bb)
    assert torch.allclose(stu[0], refu[0], atol=3e-3, rtol=1e-2) and torch.allclose(stu[1], refu[1], rtol=1e-2)


def test_conv_splitk_deep_layer():
    """The SM-starved deep shapes (M = B*16 pixels, K = 16*2048) run split-K: fp32 partial tiles reduced in a
    workspace, then converted -- same numbers as the direct kernel."""
    ops = ops_mod()
    B, H, Cb, Cs = 32, 8, 2048, 2048
    x = rnd(B, Cb, H, H, seed=1).to(BF16)
    w = (rnd(Cs, Cb, 4, 4, seed=2) / (4 * Cb ** 0.5)).to(BF16).float()
    wd, wu = ops.pack_weights(w)
    xn = to_nhwc_bf16(x.float())
    ref = F.conv2d(x.float(), w, stride=2, padding=1)
    plain = ops.conv_down(xn, wd)                      # split-K is off until enabled
    assert rel_l2(to_nchw_f32(plain), ref) < 4e-3
    ops.enable_splitk(xn.device)
    try:
        out = ops.conv_down(xn, wd)
        assert rel_l2(to_nchw_f32(out), ref) < 4e-3
        z, part = ops.conv_down_stats(xn, wd)          # no fused statistics for a split-K shape
        assert part is None and rel_l2(z, out) < 2e-3   # (atomic accumulation order: not bit-reproducible)
        s = rnd(B, Cs, H // 2, H // 2, seed=3).to(BF16)
        refu = F.conv_transpose2d(s.float(), w, stride=2, padding=1)
        up = ops.conv_up(to_nhwc_bf16(s.float()), wu)
        assert rel_l2(to_nchw_f32(up), refu) < 4e-3
    finally:
        ops.current().splitk_bytes.clear()


@pytest.mark.parametrize("bn", [128, 256])
@pytest.mark.parametrize("B,H,Cb,Cs", [(4, 16, 256, 256),     # 2 tiles per plane: one CTA pair per N tile
                                       (6, 32, 256, 256),     # batch remainder inside the second tile of a pair
                                       (2, 128, 256, 256),    # Wt < W: pairs of row segments
                                       (32, 64, 128, 256)])   # what the heuristic itself pairs (bn = 256 / 128)
def test_conv_cta_pairs(bn, B, H, Cb, Cs):
    """CTA pairs (tcgen05 cta_group::2: two M tiles per MMA, each CTA loads half of the weight tile) give the same
    result as single-CTA tiles, for every N tile, with and without fused statistics."""
    ops = ops_mod()
    x = rnd(B, Cb, H, H, seed=1).to(BF16)
    w = (rnd(Cs, Cb, 4, 4, seed=2) / (4 * Cb ** 0.5)).to(BF16).float()
    wd, wu = ops.pack_weights(w)
    xn = to_nhwc_bf16(x.float())
    s = rnd(B, Cs, H // 2, H // 2, seed=3).to(BF16)
    sn = to_nhwc_bf16(s.float())
    ref_d = F.conv2d(x.float(), w, stride=2, padding=1)
    ref_u = F.conv_transpose2d(s.float(), w, stride=2, padding=1)
    try:
        ops.set_conv_tiling(bn, 0)
        d1, u1 = ops.conv_down(xn, wd), ops.conv_up(sn, wu)
        ops.set_conv_tiling(bn, 1)
        d2, u2 = ops.conv_down(xn, wd), ops.conv_up(sn, wu)
        z, part = ops.conv_down_stats(xn, wd)
        zu, partu = ops.conv_up_stats(sn, wu)
        torch.cuda.synchronize()
    finally:
        ops.set_conv_tiling(0, -1)
    assert rel_l2(to_nchw_f32(d2), ref_d) < 4e-3 and rel_l2(to_nchw_f32(u2), ref_u) < 4e-3
    assert torch.equal(d1, d2) and torch.equal(u1, u2)          # same K order, same fp32 accumulation
    assert torch.equal(z, d2) and torch.equal(zu, u2)
    P = z.numel() // Cs
    g, b = torch.ones(Cs, device="cuda"), torch.zeros(Cs, device="cuda")
    st = ops.bn_stats_finalize(part, P, g, b)
    ref = ops.bn_stats(z.view(P, Cs), g, b)
    assert torch.allclose(st[0], ref[0], atol=3e-3, rtol=1e-2) and torch.allclose(st[1], ref[1], rtol=1e-2)
    Pu = zu.numel() // Cb
    gu, bu = torch.ones(Cb, device="cuda"), torch.zeros(Cb, device="cuda")
    stu = ops.bn_stats_finalize(partu, Pu, gu, bu)
    refu = ops.bn_stats(zu.view(Pu, Cb), gu, bu)
    assert torch.allclose(stu[0], refu[0], atol=3e-3, rtol=1e-2) and torch.allclose(stu[1], refu[1], rtol=1e-2)


@pytest.mark.parametrize("B,H,Cb,Cs", [(4, 16, 64, 128),      # DOWN: 128 output channels; UP: 64 (M = 64 accumulator)
                                       (3, 32, 64, 128),      # tiles of half an image, odd batch
                                       (11, 16, 128, 256),    # UP: 128 output channels, batch remainder in the last tile
                                       (1, 256, 64, 128)])    # the 512^2 geometry: one image row per tile
def test_conv_role_swapped(B, H, Cb, Cs):
    """Layers with <= 128 output channels run with the weights as the MMA M operand and 256 pixels as N (transposed
    accumulator): same results as the regular kernel, incl. fused statistics and the masked dgrad epilogue."""
    ops = ops_mod()
    x = rnd(B, Cb, H, H, seed=1).to(BF16)
    w = (rnd(Cs, Cb, 4, 4, seed=2) / (4 * Cb ** 0.5)).to(BF16).float()
    wd, wu = ops.pack_weights(w)
    xn = to_nhwc_bf16(x.float())
    s = rnd(B, Cs, H // 2, H // 2, seed=3).to(BF16)
    sn = to_nhwc_bf16(s.float())
    mask = to_nhwc_bf16(rnd(B, Cb, H, H, seed=4))
    ref_d = F.conv2d(x.float(), w, stride=2, padding=1)
    ref_u = F.conv_transpose2d(s.float(), w, stride=2, padding=1)
    try:
        ops.set_conv_tiling(64, 0)
        d1, u1, um1 = ops.conv_down(xn, wd), ops.conv_up(sn, wu), ops.conv_up(sn, wu, mask=mask, slope=0.2)
        ops.set_conv_tiling(1, 0)
        d2, u2, um2 = ops.conv_down(xn, wd), ops.conv_up(sn, wu), ops.conv_up(sn, wu, mask=mask, slope=0.2)
        z, part = ops.conv_down_stats(xn, wd)
        zu, partu = ops.conv_up_stats(sn, wu)
        torch.cuda.synchronize()
    finally:
        ops.set_conv_tiling(0, -1)
    assert rel_l2(to_nchw_f32(d2), ref_d) < 4e-3 and rel_l2(to_nchw_f32(u2), ref_u) < 4e-3
    assert torch.equal(d1, d2) and torch.equal(u1, u2) and torch.equal(um1, um2)   # same K order, fp32 accumulation
    assert torch.equal(z, d2) and torch.equal(zu, u2)
    for zz, pp, C in ((z, part, Cs), (zu, partu, Cb)):
        P = zz.numel() // C
        g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        st = ops.bn_stats_finalize(pp, P, g, b)
        ref = ops.bn_stats(zz.view(P, C), g, b)
        assert torch.allclose(st[0], ref[0], atol=3e-3, rtol=1e-2) and torch.allclose(st[1], ref[1], rtol=1e-2)


@pytest.mark.parametrize("P,C,act", [(64 * 16 * 16, 128, 1), (37, 256, 2), (8, 100, 1), (64 * 32 * 32, 64, 2), (4 * 16, 2048, 2)])
def test_bn_folded_finalize_matches_three_launch_path(P, C, act):
    """dg_bn_stats_acc + dg_bn_act_fwd_acc / dg_bn_act_bwd_acc (accumulators + coefficients derived in the consumer)
    against the partial-rows + finalize path: same statistics, outputs, running statistics and gradients."""
    ops = ops_mod()
    g = torch.Generator(device="cuda").manual_seed(P + C)
    z = (torch.randn(P, C, device="cuda", generator=g) * 1.5 + 0.3).to(BF16)
    dy = (torch.randn(P, C, device="cuda", generator=g) * 0.01).to(BF16)
    gamma = torch.rand(C, device="cuda", generator=g) + 0.5
    beta = torch.randn(C, device="cuda", generator=g) * 0.1
    rm1, rv1 = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    rm2, rv2 = rm1.clone(), rv1.clone()
    ctx = ops.OpsContext()
    with ops.use_context(ctx):
        ctx.fold_stats = False
        st_ref = ops.bn_stats(z, gamma, beta, rm1, rv1)
        y_ref = ops.bn_act_fwd(z, st_ref, act, 0.2)
        dg_ref, db_ref = torch.full((C,), 2.0, device="cuda"), torch.full((C,), -1.0, device="cuda")
        dz_ref = ops.bn_act_bwd(dy, y_ref, z, st_ref, gamma, act, 0.2, dg_ref, db_ref, 1.0)
        ctx.fold_stats = True
        acc = ops.bn_stats_acc(z)
        y, st = ops.bn_act_fwd_acc(z, acc, gamma, beta, act, 0.2, rm2, rv2)
        dg, db = torch.full((C,), 2.0, device="cuda"), torch.full((C,), -1.0, device="cuda")
        dz = ops.bn_act_bwd(dy, y, z, st, gamma, act, 0.2, dg, db, 1.0)
        # a second use after the per-iteration reset sees zeroed accumulators again
        ctx.arena.reset()
        acc_b = ops.bn_stats_acc(z)
        assert acc_b.data_ptr() == acc.data_ptr()
        y_b, _ = ops.bn_act_fwd_acc(z, acc_b, gamma, beta, act, 0.2)
    torch.cuda.synchronize()
    assert torch.allclose(st, st_ref, rtol=2e-5, atol=2e-6)
    assert torch.allclose(rm2, rm1, rtol=1e-5, atol=1e-7) and torch.allclose(rv2, rv1, rtol=1e-5, atol=1e-7)
    assert rel_l2(y.float(), y_ref.float()) < 1e-3 and rel_l2(y_b.float(), y_ref.float()) < 1e-3
    assert rel_l2(dz.float(), dz_ref.float()) < 2e-3
    assert torch.allclose(dg, dg_ref, rtol=1e-4, atol=1e-5) and torch.allclose(db, db_ref, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("B,H,Cb,Cs,splitk", [(2, 8, 64, 128, False), (9, 8, 256, 512, False), (64, 8, 256, 512, True),
                                              (32, 8, 2048, 2048, True)])
def test_conv_accumulated_stats(B, H, Cb, Cs, splitk):
    """Convolutions in accumulator mode (dg_conv_opts.stat_accumulate): the [2, C] sums equal the sums over the per-CTA
    partial rows -- also for split-K shapes, where the finish kernel produces them."""
    ops = ops_mod()
    x = rnd(B, Cb, H, H, seed=1).to(BF16)
    w = (rnd(Cs, Cb, 4, 4, seed=2) / (4 * Cb ** 0.5)).to(BF16).float()
    wd, wu = ops.pack_weights(w)
    xn = to_nhwc_bf16(x.float())
    s = to_nhwc_bf16(rnd(B, Cs, H // 2, H // 2, seed=3))
    ctx = ops.OpsContext()
    with ops.use_context(ctx):
        if splitk:
            ops.enable_splitk(xn.device)
        z, acc = ops.conv_down_acc(xn, wd)
        zu, accu = ops.conv_up_acc(s, wu)
        zr, part = ops.conv_down_stats(xn, wd)
    torch.cuda.synchronize()
    assert (part is None) == splitk
    zf = z.float().view(-1, Cs)
    ref = F.conv2d(x.float(), w, stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, Cs)
    assert rel_l2(zf, ref) < 4e-3
    assert torch.allclose(acc[0], ref.sum(0), rtol=2e-3, atol=2e-3 * float(ref.abs().sum(0).max()))
    assert torch.allclose(acc[1], (ref * ref).sum(0), rtol=2e-3)
    refu = F.conv_transpose2d(to_nchw_f32(s), w, stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, Cb)
    assert rel_l2(zu.float().view(-1, Cb), refu) < 4e-3
    assert torch.allclose(accu[1], (refu * refu).sum(0), rtol=2e-3)
    if part is not None:
        assert torch.allclose(acc, part.sum(1), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("B,H,Cb,Cs,splitk", [(4, 16, 64, 128, False),     # DOWN N=128 / UP N=64: role-swapped kernel when eligible
                                              (8, 64, 64, 128, False),     # many tiles: swap kernel for both directions
                                              (3, 32, 128, 256, False), (9, 8, 256, 512, False),
                                              (32, 8, 2048, 2048, True)])  # split-K: the finish kernel applies the affine
def test_conv_folded_eval_batchnorm(B, H, Cb, Cs, splitk):
    """Eval-mode BatchNorm + activation folded into the conv / convT epilogue (dg_conv_opts.affine_*): against
    act(BN_eval(conv)) in fp32 torch, and against the unfolded kernel pair."""
    ops = ops_mod()
    x = rnd(B, Cb, H, H, seed=1).to(BF16)
    s_in = rnd(B, Cs, H // 2, H // 2, seed=3).to(BF16)
    w = (rnd(Cs, Cb, 4, 4, seed=2) / (4 * Cb ** 0.5)).to(BF16).float()
    wd, wu = ops.pack_weights(w)
    ctx = ops.OpsContext()
    with ops.use_context(ctx):
        if splitk:
            ops.enable_splitk(x.device)
        for mode, C, act in (("down", Cs, ops.ACT_LRELU), ("up", Cb, ops.ACT_RELU)):
            g = torch.Generator(device="cuda").manual_seed(C)
            gamma, beta = torch.rand(C, device="cuda", generator=g) + 0.5, torch.randn(C, device="cuda", generator=g) * 0.2
            rm, rv = torch.randn(C, device="cuda", generator=g) * 0.1, torch.rand(C, device="cuda", generator=g) + 0.3
            stats = ops.bn_eval_stats(gamma, beta, rm, rv)
            if mode == "down":
                ref = F.conv2d(x.float(), w, stride=2, padding=1)
                folded = ops.conv_down(to_nhwc_bf16(x.float()), wd, affine=(stats, act, 0.2))
                plain = ops.conv_down(to_nhwc_bf16(x.float()), wd)
            else:
                ref = F.conv_transpose2d(s_in.float(), w, stride=2, padding=1)
                folded = ops.conv_up(to_nhwc_bf16(s_in.float()), wu, affine=(stats, act, 0.2))
                plain = ops.conv_up(to_nhwc_bf16(s_in.float()), wu)
            ref = F.batch_norm(ref, rm, rv, gamma, beta, False, 0.0, 1e-5)
            ref = F.leaky_relu(ref, 0.2) if act == ops.ACT_LRELU else F.relu(ref)
            two_pass = ops.bn_act_fwd(plain.view(-1, C), stats, act, 0.2).view(plain.shape)
            e_fold, e_two = rel_l2(to_nchw_f32(folded), ref), rel_l2(to_nchw_f32(two_pass), ref)
            assert e_fold < 4e-3, (mode, e_fold)
            assert e_fold <= e_two + 1e-4, (mode, e_fold, e_two)      # one rounding instead of two
    with pytest.raises(ValueError):
        ops.conv_down(to_nhwc_bf16(x.float()), wd, affine=(torch.zeros(4, Cs + 4, device="cuda"), 1, 0.2))
