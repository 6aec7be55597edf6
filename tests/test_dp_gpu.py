"""Data-parallel step on >= 2 GPUs (NCCL, bucketed all-reduce overlapped with backward) against the ORACLE's R-shard
emulation of the reference DDP step (distributed_image_translation.py:396-404,465-518: per-rank BatchNorm statistics and
feature-matching means, gradients averaged over ranks, identical Adam everywhere).  Skipped on single-GPU boxes, where
tests/test_parity_gpu.py::test_data_parallel_semantics_on_one_gpu checks the same semantics with two in-process ranks;
the host-side reducer logic is covered on CPU with gloo in tests/test_host_logic.py."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
LOSSES = ("dis_loss_A", "gen_loss_A", "dis_loss_B", "gen_loss_B", "fm_loss_A", "fm_loss_B", "recon_loss_A", "recon_loss_B")


def rel_l2(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _worker(rank, world, port, use_graphs, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from discogan_modernized_b200 import DiscoGANTrainer, model
    from oracle.bf16_emul import emulate
    from oracle.step import OracleDataParallel, OracleStep, build_nets, synthetic_batch
    S, B, steps = 64, 32, 6
    dev = f"cuda:{rank}"
    src = build_nets(S, seed=1234, device=dev)
    nets = [model.Generator(True, S), model.Generator(True, S), model.Discriminator(S), model.Discriminator(S)]
    if rank == 0:                                       # other ranks keep their own random init: the trainer must
        for n, o in zip(nets, src):                     # broadcast rank 0's weights (DDP constructor semantics)
            n.load_state_dict(o.state_dict())
    tr = DiscoGANTrainer(image_size=S, device=dev, nets=nets, use_graphs=use_graphs)
    assert tr.reducer.enabled and tr.reducer.world == world
    dp = None
    if rank == 0:
        dp = {"fp32": OracleDataParallel(lambda r: OracleStep(build_nets(S, seed=1234, device=dev), device=dev), world),
              "bf16e": OracleDataParallel(lambda r: OracleStep(emulate(build_nets(S, seed=1234, device=dev)), device=dev), world)}
    res = {"loss_rows": [], "grad_rows": []}
    for it in range(steps):
        A, Bt = synthetic_batch(B, S, step=it, rank=rank, device=dev)
        tr.step(A, Bt)
        if rank == 0:
            shards = [synthetic_batch(B, S, step=it, rank=r, device=dev) for r in range(world)]
            logs = {k: d.step(shards) for k, d in dp.items()}
            got = tr.losses()
            for k in LOSSES:
                res["loss_rows"].append((it, k, got[k], logs["bf16e"][0][k], logs["fp32"][0][k]))
            if it == 0:         # identical weights: all-reduced sum / world == the oracle's averaged gradient (D step)
                for idx in (2, 3):
                    for (pn, p), (_, q), (_, f) in zip(tr.nets()[idx].named_parameters(),
                                                       dp["bf16e"].replicas[0].nets()[idx].inner.named_parameters(),
                                                       dp["fp32"].replicas[0].nets()[idx].named_parameters()):
                        mean = p.grad / world
                        res["grad_rows"].append((pn, rel_l2(mean, f.grad), rel_l2(q.grad, f.grad),
                                                 float(mean.norm() / f.grad.norm())))
    torch.cuda.synchronize()
    flat = torch.cat([tr.flat[n].flat_p for n in tr.nets()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    rm = tr.G_A.encoder[3].running_mean.clone()
    rms = [torch.empty_like(rm) for _ in range(world)]
    dist.all_gather(rms, rm)
    if rank == 0:
        res["sync"] = all(torch.equal(gathered[0], g) for g in gathered[1:])
        res["bn_identical"] = all(torch.equal(rms[0], r) for r in rms[1:])
        # per-rank BatchNorm: each rank's running mean follows its own replica of the oracle
        res["bn_err"] = [float((rms[r].cpu() - dp["bf16e"].replicas[r].G_A.inner.encoder[3].running_mean.cpu()).abs().max())
                         for r in range(world)]
        res["buckets"] = len(tr.reducer.launched)
        res["bucket_plan"] = {n: len(tr.flat[net]._buckets) for n, net in zip(("G_A", "G_B", "D_A", "D_B"), tr.nets())}
        out.update(res)
    tr.close()                       # graphs first: NCCL teardown blocks while captured graphs reference the communicator
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("use_graphs", [True, False])
def test_data_parallel_step_matches_sharded_oracle(use_graphs):
    world, port = 2, 29541 + int(use_graphs)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, use_graphs, out), nprocs=world, join=True)
        out = dict(out)
    assert out["sync"], "weights diverged across ranks"
    assert not out["bn_identical"], "BatchNorm statistics must stay per rank"
    assert max(out["bn_err"]) < 1e-2, out["bn_err"]          # six independent noisy steps apart
    assert all(v >= 2 for v in out["bucket_plan"].values()), out["bucket_plan"]     # really bucketed
    assert out["buckets"] > 0
    for it, k, got, e, f in out["loss_rows"]:       # independent trainings drift apart after the first D,G,G cycle (B = 32)
        re_, ae, rf, af = (0.02, 0.01, 0.05, 0.02) if it < 3 else (0.05, 0.02, 0.08, 0.03)
        assert abs(got - e) <= re_ * abs(e) + ae, (it, k, got, e)
        assert abs(got - f) <= rf * abs(f) + af, (it, k, got, f)
    for pn, rel_f, floor_f, ratio in out["grad_rows"]:
        assert 0.95 < ratio < 1.05, (pn, ratio)                                     # a sum instead of a mean would read 2.0
        assert rel_f <= 1.25 * floor_f + 0.02, (pn, rel_f, floor_f)
