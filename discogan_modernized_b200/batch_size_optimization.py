"""Re-hosted ``batch_size_optimization.py`` (reference :16-200): find the largest per-GPU batch whose memory footprint stays
under a target fraction of the device, by bisection over a batch-size grid, and write the same JSON report.

Differences, all deliberate: the probe is one full D,G,G cycle of the B200 train step (forward, backward, Adam -- what a
training run allocates) instead of the reference's eight autograd-tracked forwards (:62-75); memory is read with
``torch.cuda.mem_get_info`` / the allocator's peak counter instead of shelling out to nvidia-smi twice per probe (:36-45);
each probe also records the step time, so the report shows throughput per batch size.
"""
import argparse
import json
import time
from pathlib import Path

import torch


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="largest batch size under a GPU-memory target (B200 DiscoGAN step)")
    p.add_argument("--gpu", type=int, default=0)
    p.add_argument("--model_arch", default="discogan", choices=["discogan", "recongan", "gan"])
    p.add_argument("--image_size", type=int, default=64)
    p.add_argument("--min_batch", type=int, default=16)
    p.add_argument("--max_batch", type=int, default=512)
    p.add_argument("--step", type=int, default=16)
    p.add_argument("--target_memory", type=float, default=0.85)
    p.add_argument("--extra_layers", action="store_true")
    p.add_argument("--safety_margin", type=float, default=0.9)
    p.add_argument("--output", default="batch_size_results.json")
    return p.parse_args(argv)


def get_gpu_memory(gpu_id):
    """(total MB, free MB) of a device (reference :34-45)."""
    free, total = torch.cuda.mem_get_info(gpu_id)
    return total >> 20, free >> 20


def test_batch_size(gpu_id, batch_size, image_size, model_arch="discogan"):
    """One D,G,G cycle at this batch size -> (peak MB used by the step, ok, ms per step)."""
    from .train_step import DiscoGANTrainer
    import gc
    dev = f"cuda:{gpu_id}"
    gc.collect()                                  # garbage from earlier probes would otherwise be freed DURING this probe
    torch.cuda.empty_cache()                      # and make (peak - base) read low
    torch.cuda.reset_peak_memory_stats(dev)
    base = torch.cuda.memory_allocated(dev)
    tr = None
    try:
        tr = DiscoGANTrainer(image_size=image_size, device=dev, model_arch=model_arch, seed=1234, data_parallel=False,
                             use_graphs=False)
        A = torch.rand(batch_size, 3, image_size, image_size, device=dev)      # reference :62-63 (random batches)
        B = torch.rand(batch_size, 3, image_size, image_size, device=dev)
        for _ in range(3):
            tr.step(A, B)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(3):
            tr.step(A, B)
        torch.cuda.synchronize(dev)
        ms = (time.perf_counter() - t0) / 3 * 1e3
        used = max(1, (torch.cuda.max_memory_allocated(dev) - base) >> 20)
        return int(used), True, ms
    except (torch.cuda.OutOfMemoryError, RuntimeError) as e:
        if not isinstance(e, torch.cuda.OutOfMemoryError) and "out of memory" not in str(e).lower():
            print(f"error at batch {batch_size}: {e}")
        return 0, False, 0.0
    finally:
        if tr is not None:
            tr.close()
        del tr
        torch.cuda.empty_cache()


def find_optimal_batch_size(args):
    """Bisection over range(min_batch, max_batch + step, step) (reference :104-165)."""
    total, free = get_gpu_memory(args.gpu)
    print(f"GPU {args.gpu}: {total} MB total, {free} MB free")
    sizes = list(range(args.min_batch, args.max_batch + args.step, args.step))
    lo, hi, best = 0, len(sizes) - 1, args.min_batch
    usages, step_ms = {}, {}
    while lo <= hi:
        mid = (lo + hi) // 2
        bs = sizes[mid]
        used, ok, ms = test_batch_size(args.gpu, bs, args.image_size, args.model_arch)
        if ok:
            usages[bs], step_ms[bs] = used, round(ms, 3)
            frac = used / total
            print(f"batch {bs}: {used} MB ({frac:.1%}), {ms:.2f} ms/step, {bs / ms * 1e3:.0f} pairs/s")
            if frac <= args.target_memory:
                best, lo = bs, mid + 1
            else:
                hi = mid - 1
        else:
            print(f"batch {bs}: out of memory")
            hi = mid - 1
    safe = max(args.min_batch, int(best * args.safety_margin))
    safe = max(args.step, (safe // args.step) * args.step)
    return {"gpu_id": args.gpu, "total_memory_mb": total, "image_size": args.image_size, "model_arch": args.model_arch,
            "extra_layers": args.extra_layers, "optimal_batch_size": best, "safe_batch_size": safe,
            "safety_margin": args.safety_margin, "memory_usages": usages, "ms_per_step": step_ms}


def main(argv=None):
    args = parse_args(argv)
    res = find_optimal_batch_size(args)
    Path(args.output).write_text(json.dumps(res, indent=2))
    print(f"optimal batch {res['optimal_batch_size']}, safe batch {res['safe_batch_size']} -> {args.output}")
    return res


if __name__ == "__main__":
    main()
