"""The oracle pinned against the vectors minted from the reference's own Python
(oracle/make_golden.py, run in the build container where /root/reference exists)."""
import json

import pytest
import torch

import oracle
from oracle.step import OracleStep, build_nets, synthetic_batch


def test_family_matches_reference_structure(golden_dir):
    s = json.loads((golden_dir / "ref512_structure.json").read_text())
    torch.manual_seed(0)
    with torch.device("meta"):
        G = oracle.Generator(extra_layers=True, image_size=512)
        D = oracle.Discriminator(image_size=512)
    assert {k: list(v.shape) for k, v in G.state_dict().items()} == s["generator"]
    assert {k: list(v.shape) for k, v in D.state_dict().items()} == s["discriminator"]
    assert [n for n, _ in G.named_parameters()] == s["generator_param_order"]
    assert [n for n, _ in D.named_parameters()] == s["discriminator_param_order"]
    assert sum(p.numel() for p in G.parameters()) == s["generator_params"] == 230192968
    assert sum(p.numel() for p in D.parameters()) == s["discriminator_params"] == 111852288
    assert G.main is None


def test_family_64_shapes():
    assert oracle.family_channels(64) == [64, 128, 256, 512]
    assert oracle.family_channels(512) == [64, 128, 256, 512, 1024, 2048, 2048]
    G = oracle.Generator(image_size=64)
    D = oracle.Discriminator(image_size=64)
    assert sum(p.numel() for p in G.parameters()) == 7153480
    assert sum(p.numel() for p in D.parameters()) == 2765568
    x = torch.rand(2, 3, 64, 64)
    p, feats = D(x)
    assert p.shape == (2, 1, 1, 1) and [tuple(f.shape[1:]) for f in feats] == [(128, 16, 16), (256, 8, 8), (512, 4, 4)]
    assert G(x).shape == x.shape
    for bad in (63, 8, 96):
        with pytest.raises(ValueError):
            oracle.family_channels(bad)


def test_extra_layers_is_noop():
    torch.manual_seed(3); a = oracle.Generator(extra_layers=True, image_size=64)
    torch.manual_seed(3); b = oracle.Generator(extra_layers=False, image_size=64)
    assert repr(a) == repr(b)
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)


def test_losses_match_reference(golden_dir):
    g = torch.load(golden_dir / "ref_losses.pt")
    bce, hinge = torch.nn.BCELoss(), torch.nn.HingeEmbeddingLoss()
    dl, gl = oracle.get_gan_loss(g["dr"], g["df"], bce, "cpu")
    assert torch.equal(dl, g["dis_loss"]) and torch.equal(gl, g["gen_loss"])
    assert float(gl) > 100.0 / 6 - 1e-3  # the p==0 sample hits the -100 clamp
    assert torch.allclose(oracle.get_fm_loss(g["rf"], g["ff"], hinge, "cpu"), g["fm"], rtol=1e-6, atol=0)
    assert torch.allclose(oracle.get_fm_loss_angle(g["rf"], g["ff"], hinge, "cpu"), g["fm_angle"], rtol=1e-6, atol=0)


@pytest.mark.slow
def test_family512_forward_matches_reference(golden_dir):
    g = torch.load(golden_dir / "ref512_forward.pt")
    torch.manual_seed(1234)
    G = oracle.Generator(extra_layers=True, image_size=512)
    D = oracle.Discriminator(image_size=512)
    A, B = synthetic_batch(2, 512, step=0)
    with torch.no_grad():
        y = G(A)
        p, feats = D(B)
    assert torch.allclose(y.flatten()[g["G_out_idx"]], g["G_out_samples"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(p.flatten(), g["D_prob"], rtol=1e-5, atol=1e-6)
    assert [list(f.shape) for f in feats] == g["feat_shapes"]
    assert torch.allclose(torch.stack([f.mean() for f in feats]), g["feat_means"], rtol=1e-4, atol=1e-6)
    assert torch.allclose(G.state_dict()["encoder.3.running_mean"], g["G_running_mean_3"], rtol=1e-5, atol=1e-7)
    G.eval()
    with torch.no_grad():
        ye = G(A)
    assert torch.allclose(ye.flatten()[g["G_out_idx"]], g["G_eval_samples"], rtol=1e-5, atol=1e-6)


@pytest.mark.slow
def test_step512_matches_reference_classes(golden_dir):
    """D step + G step at 512^2 B=2: oracle family + restated losses vs the run driven
    by the reference's model classes and loss helpers."""
    g = torch.load(golden_dir / "ref512_step.pt")
    nets = build_nets(512)
    st = OracleStep(nets)
    for it in range(2):
        A, B = synthetic_batch(2, 512, step=it)
        log = st.step(A, B)
        for k, v in g["logs"][it].items():
            assert log[k] == pytest.approx(v, rel=2e-4, abs=1e-6), (it, k)
    w = nets[2].state_dict()["conv2.weight"].flatten()[g["D_A_conv2_idx"]]
    assert torch.allclose(w, g["D_A_conv2_samples"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("variant,arch", [("image_translation", "discogan"), ("angle_pairing", "discogan"),
                                          ("image_translation", "recongan"), ("image_translation", "gan")])
def test_family64_step_golden(golden_dir, variant, arch):
    g = torch.load(golden_dir / f"family64_step_{variant}_{arch}.pt")
    nets = build_nets(64)
    st = OracleStep(nets, model_arch=arch, variant=variant)
    for it in range(3):
        A, B = synthetic_batch(8, 64, step=it)
        log = st.step(A, B)
        assert log["is_dis_step"] == (it == 0)
        for k, v in g["logs"][it].items():
            assert log[k] == pytest.approx(v, rel=2e-4, abs=1e-6), (it, k)
    assert torch.allclose(nets[0].state_dict()["encoder.0.weight"].flatten()[:64], g["G_A_enc0"], rtol=1e-4, atol=1e-7)
    assert torch.allclose(nets[3].state_dict()["conv1.weight"].flatten()[:64], g["D_B_conv1"], rtol=1e-4, atol=1e-7)


def test_preprocess_restatement_matches_cv2():
    """The numpy restatement of the reference's image arithmetic (dataset.py:50-66; OpenCV's fixed-point bilinear resize,
    3x3 dilate) against the reference's own dependency (cv2), bit for bit, on seeded images of the dataset geometries."""
    import numpy as np
    from oracle.preprocess import preprocess_cv2, preprocess_restated
    rng = np.random.default_rng(11)
    for H, W, dom, S in [(256, 512, "A", 64), (256, 512, "B", 64), (218, 178, None, 64), (256, 512, "A", 128),
                         (97, 131, None, 128), (300, 300, None, 512), (256, 600, "B", 64), (64, 64, None, 64), (33, 300, "A", 64)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        a, b = preprocess_restated(img, dom, S), preprocess_cv2(img, dom, S)
        assert a.dtype == b.dtype == np.float32 and a.shape == (3, S, S)
        assert np.array_equal(a, b), (H, W, dom, S, float(np.abs(a - b).max()))


def test_bf16_emulation_is_a_small_perturbation_of_the_oracle():
    """oracle/bf16_emul.py: same parameters, same losses to ~1e-3, running statistics advance identically; its noise-floor
    twin differs from it end to end by about as much as it differs from fp32 (the rounding cascade its docstring states)."""
    import torch
    from oracle.bf16_emul import emulate
    from oracle.step import OracleStep, build_nets, synthetic_batch
    torch.set_num_threads(4)
    S, B = 32, 8
    A, Bt = synthetic_batch(B, S)
    nets = {k: build_nets(S) for k in ("fp32", "e", "e2")}
    st = {"fp32": OracleStep(nets["fp32"]), "e": OracleStep(emulate(nets["e"])), "e2": OracleStep(emulate(nets["e2"], perturb=1e-6))}
    logs = {k: s.backward(A, Bt) for k, s in st.items()}
    for k in ("dis_loss_A", "gen_loss_B", "fm_loss_A", "recon_loss_B"):
        assert abs(logs["e"][k] - logs["fp32"][k]) <= 0.02 * abs(logs["fp32"][k]) + 2e-3, k
    assert int(nets["e"][0].encoder[3].num_batches_tracked) == int(nets["fp32"][0].encoder[3].num_batches_tracked) == 2
    rel = lambda a, b: float((a - b).norm() / b.norm())
    g = {k: nets[k][2].conv2.weight.grad for k in nets}
    gap, floor = rel(g["e"], g["fp32"]), rel(g["e2"], g["e"])
    assert 1e-3 < gap < 0.5 and 0.3 * gap < floor < 2.0 * gap, (gap, floor)
    # the optimiser sees the wrapped net's own parameters
    before = nets["e"][2].conv2.weight.detach().clone()
    st["e"].apply()
    assert not torch.equal(nets["e"][2].conv2.weight, before)
