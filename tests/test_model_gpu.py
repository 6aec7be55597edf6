"""Module- and step-level parity on the GPU against the oracle (the reference restated in stock fp32 PyTorch,
run here on the same device with TF32 disabled), identical seeded weights and synthetic batches.

Tolerances follow SURVEY.md F9: bf16-operand kernels chained through the BatchNorm stack deviate a few percent
on activations and up to tens of percent (rel-L2) on early-layer weight gradients, while losses agree to ~1e-3;
so activations are held to 5e-2 rel-L2, gradients to cosine similarity >= 0.9 / rel-L2 <= 0.5, losses to
5 % + 0.02 absolute."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))


def make_pair(kind, S, seed=1234):
    import oracle
    from discogan_modernized_b200 import model
    torch.manual_seed(seed)
    ref = (oracle.Generator(True, S) if kind == "G" else oracle.Discriminator(S)).cuda()
    new = (model.Generator(True, S) if kind == "G" else model.Discriminator(S))
    new.load_state_dict(ref.state_dict(), strict=True)          # state-dict compatibility is part of the boundary
    return ref, new.cuda()


@pytest.mark.parametrize("S,B", [(64, 4), (32, 3), (128, 2)])
def test_discriminator_forward_backward(S, B):
    ref, new = make_pair("D", S)
    x = torch.rand(B, 3, S, S, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    xr = x.clone().requires_grad_(True)
    xn = x.clone().requires_grad_(True)
    pr, fr = ref(xr)
    pn, fn_ = new(xn)
    assert pn.shape == pr.shape == (B, 1, 1, 1) and pn.dtype == torch.float32
    assert len(fn_) == len(fr) == ref.n_down - 1
    assert torch.allclose(pn, pr, atol=3e-2)
    for a, b in zip(fn_, fr):
        assert a.shape == b.shape and a.dtype == torch.float32 and a.is_contiguous()
        assert rel_l2(a, b) < 5e-2
    # a loss touching the probability and every feature map
    def loss(p, feats):
        return p.log().mean() + sum((f.mean(0) ** 2).mean() for f in feats)
    loss(pr, fr).backward()
    loss(pn, fn_).backward()
    assert cos(xn.grad, xr.grad) > 0.9
    for (name, a), (_, b) in zip(new.named_parameters(), ref.named_parameters()):
        assert a.grad is not None and a.grad.shape == b.grad.shape, name
        assert cos(a.grad, b.grad) > 0.9, (name, cos(a.grad, b.grad))
        assert rel_l2(a.grad, b.grad) < 0.5, (name, rel_l2(a.grad, b.grad))
    for (name, a), (_, b) in zip(new.named_buffers(), ref.named_buffers()):
        if a.dtype == torch.int64:
            assert torch.equal(a, b), name
        else:
            assert torch.allclose(a, b, rtol=5e-2, atol=5e-3), name


@pytest.mark.parametrize("S,B", [(64, 4), (32, 3), (128, 2)])
def test_generator_forward_backward(S, B):
    ref, new = make_pair("G", S)
    x = torch.rand(B, 3, S, S, device="cuda", generator=torch.Generator(device="cuda").manual_seed(6))
    xr = x.clone().requires_grad_(True)
    xn = x.clone().requires_grad_(True)
    yr, yn = ref(xr), new(xn)
    assert yn.shape == yr.shape and yn.dtype == torch.float32
    assert rel_l2(yn, yr) < 5e-2
    t = torch.rand_like(yr)
    ((yr - t) ** 2).mean().backward()
    ((yn - t) ** 2).mean().backward()
    assert cos(xn.grad, xr.grad) > 0.85
    bad = []
    for (name, a), (_, b) in zip(new.named_parameters(), ref.named_parameters()):
        assert a.grad is not None, name
        c = cos(a.grad, b.grad)
        if c < 0.85:
            bad.append((name, c, rel_l2(a.grad, b.grad)))
    assert not bad, bad


def test_generator_eval_and_nograd():
    ref, new = make_pair("G", 64)
    x = torch.rand(5, 3, 64, 64, device="cuda")
    with torch.no_grad():      # train-mode no_grad forwards update running stats (save_sample_images semantics)
        ref(x); new(x)
    ref.eval(); new.eval()
    with torch.no_grad():
        yr, yn = ref(x[:1]), new(x[:1])       # batch 1 is legal in eval mode (inference.py:168-172)
    assert rel_l2(yn, yr) < 5e-2
    assert int(new.encoder[3].num_batches_tracked) == 1


def test_module_gradients_are_ordinary_autograd_gradients():
    """torch.autograd.grad w.r.t. parameters, parameter hooks and accumulation over two backward() calls work on the
    drop-in modules exactly as on the reference modules (the Functions return real parameter gradients)."""
    ref, new = make_pair("D", 64)
    x = torch.rand(4, 3, 64, 64, device="cuda")
    loss = lambda net: net(x)[0].log().mean()
    want = torch.autograd.grad(loss(ref), [ref.conv3.weight, ref.bn2.bias])
    got = torch.autograd.grad(loss(new), [new.conv3.weight, new.bn2.bias])
    for a, b in zip(got, want):
        assert a is not None and cos(a, b) > 0.95
    assert new.conv3.weight.grad is None                     # autograd.grad does not touch .grad
    seen = []
    h = new.conv2.weight.register_hook(lambda g: seen.append(g.clone()))
    loss(new).backward()
    g1 = new.conv2.weight.grad.clone()
    assert len(seen) == 1 and torch.equal(seen[0], g1)
    loss(new).backward()                                     # second call accumulates
    assert rel_l2(new.conv2.weight.grad, 2 * g1) < 2e-2
    h.remove()
    new.zero_grad(set_to_none=True)
    assert all(p.grad is None for p in new.parameters())
    # only the input gradient requested: no parameter buffers are produced
    xin = x.clone().requires_grad_(True)
    for p in new.parameters():
        p.requires_grad_(False)
    (gx,) = torch.autograd.grad(new(xin)[0].log().mean(), [xin])
    assert gx.shape == x.shape and all(p.grad is None for p in new.parameters())


def test_eval_forward_folded_batchnorm_matches_unfolded():
    """Eval-mode forwards fold BatchNorm + activation into the conv epilogues (default); same result as the separate
    BatchNorm pass up to one bf16 rounding per layer, both closer than 5e-2 to the oracle."""
    from discogan_modernized_b200 import ops
    for kind in ("G", "D"):
        ref, new = make_pair(kind, 64)
        x = torch.rand(6, 3, 64, 64, device="cuda")
        with torch.no_grad():
            ref(x); new(x)                      # one train-mode pass: running statistics away from their defaults
        ref.eval(); new.eval()
        outs = {}
        for fold in (True, False):
            ctx = ops.OpsContext()
            ctx.fold_eval_bn = fold
            with ops.use_context(ctx), torch.no_grad():
                outs[fold] = new(x)
        with torch.no_grad():
            want = ref(x)
        if kind == "G":
            assert rel_l2(outs[True], want) < 5e-2 and rel_l2(outs[False], want) < 5e-2
            assert rel_l2(outs[True], outs[False]) < 2e-2
        else:
            assert torch.allclose(outs[True][0], want[0], atol=3e-2) and torch.allclose(outs[True][0], outs[False][0], atol=2e-2)
            for a, b, c in zip(outs[True][1], outs[False][1], want[1]):
                assert rel_l2(a, c) < 5e-2 and rel_l2(a, b) < 2e-2


def test_error_behaviour():
    from discogan_modernized_b200 import model
    D = model.Discriminator(64).cuda()
    with pytest.raises(RuntimeError):
        D(torch.rand(2, 3, 32, 32, device="cuda"))           # wrong spatial size
    with pytest.raises(RuntimeError):
        D(torch.rand(2, 3, 64, 64))                            # CPU tensor: no CPU path
    with pytest.raises(ValueError):
        model.Generator(image_size=64).cuda()(torch.rand(1, 3, 64, 64, device="cuda"))   # B=1 in training
    with pytest.raises(ValueError):
        model.Generator(image_size=48)


@pytest.mark.parametrize("variant,arch", [("image_translation", "discogan"), ("angle_pairing", "discogan"),
                                          ("image_translation", "recongan"), ("image_translation", "gan")])
def test_train_step_matches_oracle(variant, arch):
    from discogan_modernized_b200 import DiscoGANTrainer, model
    from oracle.step import OracleStep, build_nets, synthetic_batch
    S, B, steps = 64, 8, 6
    ref_nets = build_nets(S, seed=1234, device="cuda")
    torch.manual_seed(1234)
    nets = [model.Generator(True, S), model.Generator(True, S), model.Discriminator(S), model.Discriminator(S)]
    for n, r in zip(nets, ref_nets):
        n.load_state_dict(r.state_dict())
    tr = DiscoGANTrainer(image_size=S, nets=nets, model_arch=arch, variant=variant)
    ref = OracleStep(ref_nets, model_arch=arch, variant=variant, device="cuda")
    for it in range(steps):
        A, Bt = synthetic_batch(B, S, step=it, device="cuda")
        was_dis = tr.step(A, Bt)
        want = ref.step(A, Bt)
        assert was_dis == want["is_dis_step"]
        got = tr.losses()
        # the two trajectories are independent trainings from step 0: rounding differences compound through the
        # updates, so the band widens after the first D,G,G cycle
        # (the order of the fused-statistics and split-K reductions is not reproducible, so neither is the last digit)
        rel, ab = (0.05, 0.02) if it < 2 else ((0.08, 0.03) if it == 2 else (0.12, 0.04))
        for k, v in got.items():
            assert abs(v - want[k]) <= rel * abs(want[k]) + ab, (it, k, v, want[k])
    # Parameters moved the same way.  Adam's early steps move every weight by ~lr whatever the gradient's size, so
    # elements whose gradient is below the bf16 noise floor may step the other way: compare the direction of the
    # accumulated update and bound the distance by the update length itself.
    init = build_nets(S, seed=1234, device="cuda")
    for new, old, idx, key in ((tr.D_A, ref_nets[2], 2, "conv4.weight"), (tr.G_B, ref_nets[1], 1, "decoder.0.weight")):
        if arch != "discogan" and new is tr.D_A:
            continue
        a, b = dict(new.named_parameters())[key], dict(old.named_parameters())[key]
        w0 = dict(init[idx].named_parameters())[key]
        assert cos(a - w0, b - w0) > 0.6, (key, cos(a - w0, b - w0))
        assert rel_l2(a, b) < 0.1, (key, rel_l2(a, b))
    if arch == "gan":   # G_A and D_A never receive a gradient: untouched by Adam (and by weight decay)
        torch.manual_seed(1234)
        fresh = model.Generator(True, S)
        assert torch.equal(tr.G_A.encoder[0].weight.cpu(), fresh.encoder[0].weight)


def test_golden_family64(golden_dir):
    """The committed CPU golden (oracle at 64^2, B=8, 3 iterations) replayed on the kernels."""
    from discogan_modernized_b200 import DiscoGANTrainer, model
    from oracle.step import synthetic_batch
    g = torch.load(golden_dir / "family64_step_image_translation_discogan.pt")
    torch.manual_seed(1234)
    nets = [model.Generator(True, 64), model.Generator(True, 64), model.Discriminator(64), model.Discriminator(64)]
    tr = DiscoGANTrainer(image_size=64, nets=nets)
    for it in range(3):
        A, B = synthetic_batch(8, 64, step=it, device="cuda")
        tr.step(A, B)
        got = tr.losses()
        for k, v in g["logs"][it].items():
            if k in got:
                assert abs(got[k] - v) <= 0.05 * abs(v) + 0.02, (it, k, got[k], v)


def test_full_size_512_step_matches_oracle():
    """BASELINE's full-size topology (512x512, the reference model.py itself), B=2: one D step and one G step of the
    fused trainer against the oracle on the same device -- exercises the 2048-channel layers, the direct wgrad
    epilogue and the deep split-free GEMMs that the 64x64 family never reaches."""
    from discogan_modernized_b200 import DiscoGANTrainer, model
    from oracle.step import OracleStep, build_nets, synthetic_batch
    S, B = 512, 2
    ref_nets = build_nets(S, seed=1234, device="cuda")
    nets = []
    with torch.device("meta"):
        shells = [model.Generator(True, S), model.Generator(True, S), model.Discriminator(S), model.Discriminator(S)]
    for shell, r in zip(shells, ref_nets):
        n = shell.to_empty(device="cuda")
        n.load_state_dict(r.state_dict())
        nets.append(n)
    tr = DiscoGANTrainer(image_size=S, nets=nets)
    ref = OracleStep(ref_nets, device="cuda")
    for it in range(2):
        A, Bt = synthetic_batch(B, S, step=it, device="cuda")
        tr.step(A, Bt)
        want = ref.step(A, Bt)
        got = tr.losses()
        # (B = 2: the 1x1 BatchNorm(100) bottleneck normalises over two samples, the noisiest configuration there is; the
        # benchmarked batch is covered by tests/test_parity_gpu.py)
        rel, ab = (0.06, 0.03) if it == 0 else (0.15, 0.05)
        for k, v in got.items():
            assert abs(v - want[k]) <= rel * abs(want[k]) + ab, (it, k, v, want[k])
    a = dict(tr.D_A.named_parameters())["conv7.weight"]
    b = dict(ref_nets[2].named_parameters())["conv7.weight"]
    assert rel_l2(a, b) < 0.05
    tr.close()


def test_full_size_512_generator_forward():
    """Generator forward at 512x512 against the golden minted from the REFERENCE model.py on CPU
    (tests/golden/ref512_forward.pt): same seed-1234 construction order, same synthetic batch."""
    from pathlib import Path
    from discogan_modernized_b200 import model
    from oracle.step import synthetic_batch
    g = torch.load(Path(__file__).resolve().parent / "golden" / "ref512_forward.pt")
    torch.manual_seed(1234)
    G = model.Generator(extra_layers=True)            # default image_size = 512
    D = model.Discriminator()
    G, D = G.cuda(), D.cuda()
    A, Bt = synthetic_batch(2, 512, step=0, device="cuda")
    with torch.no_grad():
        y = G(A)
        p, feats = D(Bt)
    got = y.flatten().cpu()[g["G_out_idx"]]
    assert float((got - g["G_out_samples"]).abs().max()) < 3e-2       # sigmoid outputs in (0,1), bf16 chain
    assert abs(float(y.mean()) - float(g["G_out_mean"])) < 5e-3
    assert torch.allclose(p.flatten().cpu(), g["D_prob"], atol=3e-2)
    assert [list(f.shape) for f in feats] == g["feat_shapes"]
    assert torch.allclose(torch.stack([f.mean() for f in feats]).cpu(), g["feat_means"], rtol=5e-2, atol=5e-3)
    assert torch.allclose(G.state_dict()["encoder.3.running_mean"].cpu(), g["G_running_mean_3"], rtol=5e-2, atol=1e-3)
