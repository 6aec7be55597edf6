"""The callers and data formats either side of the step (N2 checkpoints + resume, N3 batched
inference, N4 logging / sample dumps, a15 save_sample_images semantics), through the re-hosted entry points."""
import re

import pytest
import torch

pytestmark = pytest.mark.gpu

# the regexes hyperparameter_search.py:269-271 applies to the training log
LOG_PATTERNS = (r"GEN: (\d+\.\d+)/(\d+\.\d+)", r"RECON: (\d+\.\d+)/(\d+\.\d+)", r"DIS: (\d+\.\d+)/(\d+\.\d+)")


def fresh_trainer(S=64, **kw):
    from discogan_modernized_b200 import DiscoGANTrainer
    return DiscoGANTrainer(image_size=S, seed=1234, data_parallel=False, **kw)


def flat_state(tr):
    parts = [tr.flat[n].flat_p for n in tr.nets()] + [tr.flat[n].exp_avg for n in tr.nets()] + \
            [tr.flat[n].exp_avg_sq for n in tr.nets()]
    bufs = [b.float().flatten() for n in tr.nets() for b in n.buffers()]
    return torch.cat([p.flatten() for p in parts] + bufs)


def test_resume_continues_bit_for_bit(tmp_path):
    """N2: train 4 iterations, save (reference-named .pth files + train_state), load into a fresh trainer, 2 more
    iterations == 6 uninterrupted iterations, bit for bit (deterministic mode, eager launches).  The saved state covers
    Adam's moments and step count and the iteration counter, so the D:G:G schedule and the curriculum continue."""
    from discogan_modernized_b200._cli import load_models, save_models
    from oracle.step import synthetic_batch
    S, B = 64, 8
    batches = [synthetic_batch(B, S, step=i, device="cuda") for i in range(6)]
    a = fresh_trainer(S, deterministic=True, use_graphs=False, gan_curriculum=5)    # the rate switches at iteration 5
    for i in range(6):
        a.step(*batches[i])
    ref_state, ref_losses = flat_state(a).clone(), a.losses()
    a2 = fresh_trainer(S, deterministic=True, use_graphs=False, gan_curriculum=5)   # determinism itself
    for i in range(6):
        a2.step(*batches[i])
    assert torch.equal(flat_state(a2), ref_state), "deterministic mode is not bit-reproducible"
    b = fresh_trainer(S, deterministic=True, use_graphs=False, gan_curriculum=5)
    for i in range(4):
        b.step(*batches[i])
    save_models(b, tmp_path, 4)
    names = sorted(p.name for p in tmp_path.iterdir())
    assert names == ["dis_A_4.pth", "dis_B_4.pth", "gen_A_4.pth", "gen_B_4.pth", "train_state_4.pth"]
    sd = torch.load(tmp_path / "gen_A_4.pth")
    assert list(sd) == list(b.G_A.state_dict()) and all(v.device.type == "cpu" for v in sd.values())
    torch.manual_seed(999)                                            # a different init: everything must come from disk
    from discogan_modernized_b200 import DiscoGANTrainer
    c = DiscoGANTrainer(image_size=S, data_parallel=False, deterministic=True, use_graphs=False, gan_curriculum=5)
    assert load_models(c, tmp_path, 4) is True
    assert c.iters == 4
    for i in range(4, 6):
        assert c.step(*batches[i]) == (i % 3 == 0)
    assert torch.equal(flat_state(c), ref_state)
    assert c.losses() == ref_losses
    # weights-only resume (what the reference does, distributed_image_translation.py:379-393): optimiser restarts
    (tmp_path / "train_state_4.pth").unlink()
    d = DiscoGANTrainer(image_size=S, data_parallel=False, deterministic=True, use_graphs=False)
    assert load_models(d, tmp_path, 4) is False and d.iters == 0
    assert torch.equal(d.flat[d.G_A].flat_p, b.flat[b.G_A].flat_p) and float(d.flat[d.G_A].exp_avg.abs().max()) == 0.0
    for t in (a, a2, b, c, d):
        t.close()


def test_resume_with_graphs_tracks(tmp_path):
    """The default (graph-replayed, non-deterministic reductions) trainer resumes onto the same trajectory up to
    reduction-order noise."""
    from discogan_modernized_b200._cli import load_models, save_models
    from oracle.step import synthetic_batch
    S, B = 64, 16
    batches = [synthetic_batch(B, S, step=i, device="cuda") for i in range(8)]
    a = fresh_trainer(S)
    for i in range(8):
        a.step(*batches[i])
    b = fresh_trainer(S)
    for i in range(5):
        b.step(*batches[i])
    save_models(b, tmp_path, "x")
    c = fresh_trainer(S)
    load_models(c, tmp_path, "x")
    for i in range(5, 8):
        c.step(*batches[i])
    la, lc = a.losses(), c.losses()
    for k in la:
        assert abs(la[k] - lc[k]) <= 0.03 * abs(la[k]) + 0.01, (k, la[k], lc[k])
    for t in (a, b, c):
        t.close()


def test_sample_dump_semantics(tmp_path):
    """a15: reference mode = train-mode no_grad passes that advance the BatchNorm running statistics (each generator is
    called twice: +2 on num_batches_tracked); eval mode leaves the training state untouched."""
    from discogan_modernized_b200._cli import save_sample_grid
    from PIL import Image
    S, n = 64, 5
    tr = fresh_trainer(S)
    A, B = torch.rand(n, 3, S, S, device="cuda"), torch.rand(n, 3, S, S, device="cuda")
    bn = tr.G_A.encoder[3]
    rm0, nbt0 = bn.running_mean.clone(), int(bn.num_batches_tracked)
    out = save_sample_grid(tr, A, B, tmp_path / "samples", 0, n_samples=n, mode="eval")
    assert torch.equal(bn.running_mean, rm0) and int(bn.num_batches_tracked) == nbt0
    assert tr.G_A.training and tr.G_B.training
    img = Image.open(out)
    assert img.size == (6 * S, n * S) and out.name == "samples_iter_0.png"
    save_sample_grid(tr, A, B, tmp_path / "samples", 1000, n_samples=n, mode="reference")
    assert int(bn.num_batches_tracked) == nbt0 + 2 and not torch.equal(bn.running_mean, rm0)
    # the dumped passes equal the oracle's no_grad train-mode passes (image_translation.py:170-176)
    from oracle.step import build_nets
    ref = build_nets(S, seed=1234, device="cuda")
    tr2 = fresh_trainer(S)
    with torch.no_grad():
        AB, BA = ref[1](A), ref[0](B)
        ABA = ref[0](AB)
    mine = tr2.sample_images(A, B, "reference")
    rel = lambda a, b: float((a - b).norm() / b.norm())
    # (five samples: the 1x1 BatchNorm(100) bottleneck normalises over 5 values, which amplifies bf16 rounding)
    assert rel(mine[0], AB) < 3e-2 and rel(mine[2], ABA) < 6e-2, (rel(mine[0], AB), rel(mine[2], ABA))
    rm_ref, rm_new = ref[0].encoder[3].running_mean, tr2.G_A.encoder[3].running_mean
    assert torch.allclose(rm_new, rm_ref, rtol=2e-2, atol=2e-3)
    tr.close(); tr2.close()


def test_entry_point_end_to_end(tmp_path, capsys):
    """image_translation.main on synthetic batches: log lines in the reference format (parsed with the
    hyperparameter_search.py regexes), sample grids at iteration 0 and every interval, checkpoints at iteration 0 and every
    interval + final, then --resume continues the iteration count."""
    from discogan_modernized_b200 import image_translation
    common = ["--synthetic", "--image_size", "64", "--batch_size", "8", "--epochs", "1", "--iters_per_epoch", "50",
              "--results_dir", str(tmp_path / "results"), "--models_dir", str(tmp_path / "models"), "--task_name", "unit",
              "--log_interval", "2", "--image_save_interval", "3", "--model_save_interval", "3", "--n_samples", "2",
              "--style_A", "Male", "--constraint", "Young"]
    tr = image_translation.main(common + ["--max_iters", "7"])
    err = capsys.readouterr().err
    assert "--constraint" in err and "ignores it" in err                   # accepted-but-unused flags say so
    log = (tr.result_path / "training_log.txt").read_text()
    lines = [l for l in log.splitlines() if l.startswith("Iter [")]
    assert [int(re.match(r"Iter \[(\d+)/50\]", l).group(1)) for l in lines] == [0, 2, 4, 6]
    for pat in LOG_PATTERNS:
        m = re.findall(pat, log)
        assert len(m) == 4 and all(float(x) >= 0 for pair in m for x in pair)
    assert sorted(p.name for p in (tr.result_path / "samples").iterdir()) == [f"samples_iter_{i}.png" for i in (0, 3, 6)]
    saved = sorted(p.name for p in tr.model_path.iterdir())
    for tag in ("0", "3", "6", "final"):
        for n in ("gen_A", "gen_B", "dis_A", "dis_B", "train_state"):
            assert f"{n}_{tag}.pth" in saved, (n, tag, saved)
    assert "Male" in str(tr.model_path)
    w_final = tr.flat[tr.G_B].flat_p.clone()
    first = tr.model_path
    tr.close()
    tr2 = image_translation.main(common + ["--max_iters", "2", "--resume", str(first)])
    assert tr2.iters == 9                                                    # 7 done before, 2 more
    assert not torch.equal(tr2.flat[tr2.G_B].flat_p, w_final)
    tr2.close()


def test_inference_from_checkpoint(tmp_path):
    """N3: gen_B_final.pth written by the trainer -> inference.load_generator -> batched eval-mode translate, against the
    oracle's eval forward on the same weights (inference.py:126-172)."""
    from discogan_modernized_b200 import inference
    from discogan_modernized_b200._cli import save_models
    from oracle import Generator as OracleGenerator
    from oracle.step import synthetic_batch
    S = 64
    tr = fresh_trainer(S)
    for i in range(4):                                                    # running statistics away from their defaults
        tr.step(*synthetic_batch(8, S, step=i, device="cuda"))
    save_models(tr, tmp_path, "final")
    g = inference.load_generator(tmp_path / "gen_B_final.pth", S)
    assert not g.training
    ref = OracleGenerator(True, S).cuda()
    ref.load_state_dict(torch.load(tmp_path / "gen_B_final.pth"))
    ref.eval()
    imgs = torch.rand(11, 3, S, S)                                        # ragged last batch, host tensors
    out = inference.translate(g, imgs, batch_size=4)
    with torch.no_grad():
        want = ref(imgs.cuda())
    assert out.shape == want.shape and float((out - want).abs().max()) < 3e-2
    one = inference.translate(g, imgs[:1], batch_size=1)                  # batch 1 is legal in eval mode
    assert float((one - want[:1]).abs().max()) < 3e-2
    tr.close()


def test_batch_size_tool(tmp_path):
    """N4: the re-hosted batch_size_optimization.py bisects over full train steps and writes the reference's report keys."""
    from discogan_modernized_b200 import batch_size_optimization as bso
    out = tmp_path / "bs.json"
    res = bso.main(["--min_batch", "8", "--max_batch", "32", "--step", "8", "--image_size", "64", "--output", str(out)])
    assert out.exists()
    for k in ("gpu_id", "total_memory_mb", "image_size", "model_arch", "extra_layers", "optimal_batch_size",
              "safe_batch_size", "safety_margin", "memory_usages"):
        assert k in res
    assert res["optimal_batch_size"] == 32 and res["safe_batch_size"] == 24          # 32 * 0.9 -> 28 -> multiple of 8
    assert all(v > 0 for v in res["memory_usages"].values())


def test_graphed_inference_matches_eager():
    from discogan_modernized_b200 import inference
    from discogan_modernized_b200.model import Generator
    torch.manual_seed(3)
    g = Generator(True, 64).cuda().eval()
    x = torch.rand(9, 3, 64, 64, device="cuda")
    with torch.no_grad():
        want = g(x)
    gg = inference.GraphedGenerator(g)
    assert torch.equal(gg(x).clone(), want)
    assert float((gg(x[:4]).clone() - want[:4]).abs().max()) < 3e-2      # another batch size may pick another split-K plan
    assert torch.equal(gg(x).clone(), want)                              # the first graph is still valid
    out = inference.translate(g, x.cpu(), batch_size=4)                  # 4 + 4 + 1: three graphs, outputs cloned
    assert float((out - want).abs().max()) < 3e-2
    with torch.no_grad():                                                # new weights: refresh() re-packs in place
        for p in g.parameters():
            p.mul_(0.5)
        gg.refresh()
        want2 = g(x)
    assert torch.equal(gg(x).clone(), want2) and not torch.equal(want2, want)
    with pytest.raises(ValueError):
        inference.GraphedGenerator(Generator(True, 64).cuda())     # train mode
