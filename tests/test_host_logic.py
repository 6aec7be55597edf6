"""Host-side logic that needs no GPU: model structure / state-dict boundary, loss weights, flat buffers and the
data-parallel reducer (gloo, world_size 2)."""
import json
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from discogan_modernized_b200 import model
from discogan_modernized_b200.train_step import FlatNet, GradReducer, loss_coefficients, plan_buckets


def test_state_dict_matches_reference(golden_dir):
    s = json.loads((golden_dir / "ref512_structure.json").read_text())
    with torch.device("meta"):
        G = model.Generator(extra_layers=True)          # default image_size = 512, the reference topology
        D = model.Discriminator()
    assert {k: list(v.shape) for k, v in G.state_dict().items()} == s["generator"]
    assert {k: list(v.shape) for k, v in D.state_dict().items()} == s["discriminator"]
    assert [n for n, _ in G.named_parameters()] == s["generator_param_order"]
    assert [n for n, _ in D.named_parameters()] == s["discriminator_param_order"]
    assert G.main is None and hasattr(G, "encoder") and hasattr(G, "decoder")


def test_default_init_equals_reference_init():
    """Same RNG consumption as the reference constructors: seeding then constructing gives identical weights."""
    for S in (64, 128):
        torch.manual_seed(7); a = model.Generator(True, S); b = model.Discriminator(S)
        torch.manual_seed(7); c = oracle.Generator(True, S); d = oracle.Discriminator(S)
        for x, y in ((a, c), (b, d)):
            for (k1, v1), (k2, v2) in zip(x.state_dict().items(), y.state_dict().items()):
                assert k1 == k2 and torch.equal(v1, v2)


def test_cpu_input_raises():
    D = model.Discriminator(64)
    with pytest.raises(RuntimeError, match="CUDA"):
        D(torch.rand(2, 3, 64, 64))
    with pytest.raises(ValueError):
        model.family_channels(100)


def test_loss_coefficients_match_autograd():
    """d(gen_loss)/d(component) from image_translation.py:370-382 for every model_arch."""
    for arch in ("discogan", "recongan", "gan"):
        for rate in (0.01, 0.5, 0.9):
            comp = {k: torch.tensor(float(i + 1), requires_grad=True) for i, k in
                    enumerate(("gen_A", "fm_A", "gen_B", "fm_B", "recon_A", "recon_B"))}
            gA = (comp["fm_B"] * 0.9 + comp["gen_B"] * 0.1) * (1 - rate) + comp["recon_A"] * rate
            gB = (comp["fm_A"] * 0.9 + comp["gen_A"] * 0.1) * (1 - rate) + comp["recon_B"] * rate
            loss = {"discogan": gA + gB, "recongan": gA, "gan": comp["gen_B"] * 0.1 + comp["fm_B"] * 0.9}[arch]
            loss.backward()
            co = loss_coefficients(arch, rate)
            for k, v in comp.items():
                want = 0.0 if v.grad is None else float(v.grad)
                assert co[k] == pytest.approx(want, abs=1e-7), (arch, rate, k)
    with pytest.raises(ValueError):
        loss_coefficients("wgan", 0.5)


def test_flatnet_views():
    net = model.Discriminator(16)
    before = {k: v.clone() for k, v in net.state_dict().items()}
    fn = FlatNet(net)
    for k, v in net.state_dict().items():
        assert torch.equal(v, before[k])
    for p, o in zip(fn.params, fn.offsets):
        assert p.data_ptr() == fn.flat_p.data_ptr() + 4 * o and o % 64 == 0
        assert p.grad.data_ptr() == fn.flat_g.data_ptr() + 4 * o
    fn.flat_g.fill_(1.0)
    assert all(float(p.grad.sum()) == p.numel() for p in fn.params)
    net.conv1.weight.grad = None
    fn.zero_grad()          # no memset: gradients are re-attached and marked fresh (first write overwrites)
    assert net.conv1.weight.grad is not None and all(p._dg_fresh for p in fn.params)
    buf, beta = model._grad_buf(net.conv1.weight)
    assert beta == 0.0 and buf.data_ptr() == fn.flat_g.data_ptr()
    assert model._grad_buf(net.conv1.weight)[1] == 1.0


def test_plan_buckets_backward_order():
    """Buckets are contiguous parameter ranges, first-finishing (= last-registered) first, each closed at the cap."""
    sizes = [64, 1024, 64, 64, 4096, 64, 64, 2048, 64]
    b = plan_buckets(sizes, 2000)
    assert b == [(7, 8), (4, 6), (0, 3)]
    assert sorted(i for lo, hi in b for i in range(lo, hi + 1)) == list(range(len(sizes)))      # a partition
    assert plan_buckets(sizes, 10 ** 9) == [(0, 8)]                                               # one bucket
    assert plan_buckets(sizes, 1) == [(i, i) for i in range(8, -1, -1)]                           # one per parameter
    # the reference topology: 25 MiB buckets (DDP's default) over the 230 M-parameter generator
    with torch.device("meta"):
        G = model.Generator(True, 512)
    sizes = [(p.numel() + 63) // 64 * 64 for p in G.parameters()]
    b = plan_buckets(sizes, (25 << 20) // 4)
    assert 6 <= len(b) <= 12 and b[0][1] == len(sizes) - 1 and b[-1][0] == 0      # a layer (up to 256 MB) is never split


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                       # different weights per rank before the broadcast
    net = model.Discriminator(16)
    fn = FlatNet(net)
    red = GradReducer(bucket_bytes=8 << 10)             # small buckets: several per network
    assert red.enabled and red.world == world and red.grad_scale == 1.0 / world
    red.broadcast_params([fn])
    red.prepare([fn])
    assert len(fn._buckets) >= 3
    fn.flat_g.fill_(float(rank + 1))
    # two backward passes report every parameter twice, last-registered layer first: a bucket goes out only when its
    # last parameter has been reported for the second time
    red.begin(fn, passes=2, slot=0)
    hook = red.hook(fn)
    order = list(reversed(fn.params))
    for p in order:
        hook([p])
    assert red.launched == []
    seen = []
    for p in order:
        hook([p])
        seen.append(len(red.launched))
    red.finish(fn)
    red.join()
    assert len(red.launched) == len(fn._buckets) and seen[-1] == len(fn._buckets) and seen[0] <= 1
    los = [lo for _, lo, _ in red.launched]
    assert los == sorted(los, reverse=True)             # deep (late-registered) buckets first
    covered = sum(hi - lo for _, lo, hi in red.launched)
    out[rank] = (fn.flat_p[:256].clone(), float(fn.flat_g[0]), float(fn.flat_g[-1]), covered == fn.numel,
                 bool((fn.flat_g == 3.0).all()))
    dist.destroy_process_group()


def test_grad_reducer_gloo_world2():
    world, port = 2, 29531
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_dp_worker, args=(world, port, out), nprocs=world, join=True)
        (p0, g0a, g0b, c0, all0), (p1, g1a, g1b, c1, all1) = out[0], out[1]
    assert torch.equal(p0, p1)                          # rank 0's weights everywhere
    assert g0a == g0b == g1a == g1b == 3.0              # summed; Adam applies grad_scale = 1/world
    assert c0 and c1 and all0 and all1                  # the buckets tile the whole flat gradient, each reduced once


def _dp_oracle_worker(rank, world, port, out):
    """R-rank data parallel == each rank steps on its own shard with per-rank BN/FM and averaged gradients."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.step import OracleStep, build_nets, synthetic_batch
    torch.set_num_threads(2)
    st = OracleStep(build_nets(16))

    def avg(nets):
        for n in nets:
            for p in n.parameters():
                if p.grad is not None:
                    dist.all_reduce(p.grad)
                    p.grad /= world
    for it in range(2):
        A, B = synthetic_batch(4, 16, step=it, rank=rank)
        st.step(A, B, grad_hook=avg)
    out[rank] = (st.D_A.conv1.weight.detach().clone(), st.G_A.encoder[3].running_mean.clone())
    dist.destroy_process_group()


def test_dp_oracle_semantics_gloo_world2():
    world, port = 2, 29532
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_dp_oracle_worker, args=(world, port, out), nprocs=world, join=True)
        (w0, rm0), (w1, rm1) = out[0], out[1]
    assert torch.allclose(w0, w1, atol=1e-7)            # weights stay synchronised
    assert not torch.equal(rm0, rm1)                    # BatchNorm statistics stay per rank


def test_log_line_matches_the_reference_parsers():
    """The log line is parsed by hyperparameter_search.py:269-271 with these regexes; the iteration prefix is
    image_translation.py:394."""
    import re
    from discogan_modernized_b200.train_step import LOSS_NAMES, format_log_line
    losses = {k: 0.1 * (i + 1) for i, k in enumerate(LOSS_NAMES)}
    losses["gen_loss_A"] = 12.34567
    line = format_log_line(150, 2000, losses)
    assert line.startswith("Iter [150/2000] GEN: 12.3457/")
    assert re.findall(r"GEN: (\d+\.\d+)/(\d+\.\d+)", line) == [("12.3457", f"{losses['gen_loss_B']:.4f}")]
    assert re.findall(r"RECON: (\d+\.\d+)/(\d+\.\d+)", line) == [(f"{losses['recon_loss_A']:.4f}", f"{losses['recon_loss_B']:.4f}")]
    assert re.findall(r"DIS: (\d+\.\d+)/(\d+\.\d+)", line) == [(f"{losses['dis_loss_A']:.4f}", f"{losses['dis_loss_B']:.4f}")]
    assert re.findall(r"FM: (\d+\.\d+)/(\d+\.\d+)", line) == [(f"{losses['fm_loss_A']:.4f}", f"{losses['fm_loss_B']:.4f}")]


def test_sampler_indices_equal_distributed_sampler():
    from torch.utils.data.distributed import DistributedSampler
    from discogan_modernized_b200.dataset import domain_crop, sampler_indices, task_domains
    for length, world in ((20, 2), (23, 4), (7, 8), (64, 1)):
        for epoch in (0, 5):
            for rank in range(world):
                ref = DistributedSampler(range(length), num_replicas=world, rank=rank, shuffle=True, seed=0)
                ref.set_epoch(epoch)
                assert sampler_indices(length, rank, world, epoch).tolist() == list(ref)
    assert task_domains("edges2shoes") == ("A", "B") and task_domains("handbags2shoes") == ("B", "B")
    assert task_domains("celebA") == (None, None)
    assert domain_crop("A", 512) == (0, 256, 1) and domain_crop("B", 512) == (256, 256, 0) and domain_crop(None, 178) == (0, 178, 0)
    with pytest.raises(ValueError):
        domain_crop("B", 200)


def test_hyperparameter_search_host_logic(tmp_path):
    """Sampling space, trial command and log parsing of the re-hosted search tool (hyperparameter_search.py:47-58,253-292)."""
    import random
    from discogan_modernized_b200 import hyperparameter_search as hs
    from discogan_modernized_b200.train_step import LOSS_NAMES, format_log_line
    hps = hs.sample_hyperparameters(30, random.Random(1))
    assert len(hps) == 30 and len({tuple(h.values()) for h in hps}) == 30
    assert all(h[k] in v for h in hps for k, v in hs.PARAM_RANGES.items())
    log = tmp_path / "train.log"
    lines = [format_log_line(i * 50, 500, {k: 0.5 / (i + 1) + 0.01 * j for j, k in enumerate(LOSS_NAMES)}) for i in range(4)]
    log.write_text("Training started\n" + "\n".join(lines) + "\n")
    m = hs.extract_metrics(log)
    last = {k: 0.5 / 4 + 0.01 * j for j, k in enumerate(LOSS_NAMES)}
    assert m["final_recon_loss_A"] == pytest.approx(last["recon_loss_A"], abs=1e-4)
    assert m["final_gen_loss_B"] == pytest.approx(last["gen_loss_B"], abs=1e-4)
    assert m["avg_recon_loss"] == pytest.approx((last["recon_loss_A"] + last["recon_loss_B"]) / 2, abs=1e-4)
    args = hs.parse_args(["--trials", "2", "--output_dir", str(tmp_path), "--base_epochs", "1"])
    cmd = hs.trial_command(args, hps[0], tmp_path / "t")
    assert "discogan_modernized_b200.image_translation" in cmd and "--synthetic" in cmd
    assert cmd[cmd.index("--learning_rate") + 1] == str(hps[0]["learning_rate"])
    info = {"log_file": str(log), "_best": float("inf"), "_stale": 0, "_seen": 0}
    args.early_stopping, args.patience = True, 2
    assert hs.check_early_stop(args, info) is False and info["_seen"] == 4          # loss improves on every line
    s = hs.analyze_results([{"trial_id": 0, "hyperparameters": hps[0], "metrics": m},
                            {"trial_id": 1, "hyperparameters": hps[1], "metrics": {"avg_recon_loss": 0.01}}], tmp_path)
    assert s["best"]["trial_id"] == 1 and (tmp_path / "summary.json").exists()


def test_entry_point_flags_and_defaults_match_the_reference(golden_dir):
    """Entry-point contract: every flag the reference parsers accept is accepted by the
    re-hosted parsers with the same default -- against tests/golden/ref_cli_defaults.json, minted by
    oracle/make_cli_golden.py from the reference's own parse_args()."""
    from discogan_modernized_b200 import angle_pairing, distributed_image_translation, image_translation
    ref = json.loads((golden_dir / "ref_cli_defaults.json").read_text())
    for name, mod in (("image_translation", image_translation), ("angle_pairing", angle_pairing),
                      ("distributed_image_translation", distributed_image_translation)):
        mine = vars(mod.parse_args([]))
        for flag, default in ref[name].items():
            assert flag in mine, (name, flag)
            assert mine[flag] == default, (name, flag, mine[flag], default)
    # the log line keeps the reference's field order (image_translation.py:394-398)
    src = ref["log_line_source"]
    order = [src.index(k) for k in ("Iter [", "GEN:", "FM:", "RECON:", "DIS:")]
    assert order == sorted(order)
    from discogan_modernized_b200.train_step import LOSS_NAMES, format_log_line
    line = format_log_line(1, 2, {k: 0.0 for k in LOSS_NAMES})
    assert [line.index(k) for k in ("Iter [", "GEN:", "FM:", "RECON:", "DIS:")] == sorted(line.index(k) for k in ("Iter [", "GEN:", "FM:", "RECON:", "DIS:"))
