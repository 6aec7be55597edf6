"""Re-hosted ``hyperparameter_search.py`` (reference :20-292; the reference file is truncated at :360 and does not
import): random search over the reference's hyper-parameter grid (:47-58), one training subprocess per
trial pinned to a GPU, metrics read back from the training log with the reference's regular expressions (:269-271),
optional early stopping on the reconstruction loss (:203-241), a ranked summary at the end.

Trials run this package's ``image_translation`` entry point, whose log line keeps the reference format.
"""
import argparse
import json
import os
import random
import re
import subprocess
import sys
import time
from datetime import datetime
from pathlib import Path

PARAM_RANGES = {                       # reference :50-58
    "learning_rate": [0.0001, 0.0002, 0.0003, 0.0005],
    "beta1": [0.5, 0.7, 0.9],
    "beta2": [0.9, 0.99, 0.999],
    "starting_rate": [0.01, 0.05, 0.1, 0.2],
    "default_rate": [0.3, 0.5, 0.7, 0.9],
    "gan_curriculum": [5000, 10000, 15000, 20000],
    "update_interval": [1, 2, 3, 5],
}


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="DiscoGAN hyper-parameter search (B200 step)")
    p.add_argument("--task_name", default="edges2shoes")
    p.add_argument("--model_arch", default="discogan")
    p.add_argument("--gpus", default="0", help="comma-separated GPU ids; one trial per GPU at a time")
    p.add_argument("--trials", type=int, default=20)
    p.add_argument("--base_epochs", type=int, default=20)
    p.add_argument("--style_A", default=None)
    p.add_argument("--style_B", default=None)
    p.add_argument("--output_dir", default="./hp_search")
    p.add_argument("--batch_size", type=int, default=64)
    p.add_argument("--early_stopping", action="store_true")
    p.add_argument("--patience", type=int, default=5)
    # additions: where the data comes from (the reference's trials read ./datasets through dataset.py)
    p.add_argument("--image_size", type=int, default=64)
    p.add_argument("--data_A", default=None)
    p.add_argument("--data_B", default=None)
    p.add_argument("--iters_per_epoch", type=int, default=100, help="synthetic data only")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--poll_seconds", type=float, default=30.0)
    return p.parse_args(argv)


def sample_hyperparameters(num_samples=10, rng=None):
    """Independent uniform draws from every range, without repeating a combination (reference :77-100)."""
    rng = rng or random
    total = 1
    for v in PARAM_RANGES.values():
        total *= len(v)
    seen, out = set(), []
    while len(out) < min(num_samples, total):
        hp = {k: rng.choice(v) for k, v in PARAM_RANGES.items()}
        key = tuple(hp.values())
        if key not in seen:
            seen.add(key)
            out.append(hp)
    return out


def extract_metrics(log_file):
    """Last logged GEN / RECON / DIS pairs of a training log (reference :253-292)."""
    m = {k: None for k in ("final_gen_loss_A", "final_gen_loss_B", "final_recon_loss_A", "final_recon_loss_B",
                           "final_dis_loss_A", "final_dis_loss_B")}
    try:
        text = Path(log_file).read_text()
    except OSError as e:
        print(f"metric extraction failed: {e}")
        return m
    for name, pat in (("gen", r"GEN: (\d+\.\d+)/(\d+\.\d+)"), ("recon", r"RECON: (\d+\.\d+)/(\d+\.\d+)"),
                      ("dis", r"DIS: (\d+\.\d+)/(\d+\.\d+)")):
        found = re.findall(pat, text)
        if found:
            m[f"final_{name}_loss_A"], m[f"final_{name}_loss_B"] = float(found[-1][0]), float(found[-1][1])
    if m["final_recon_loss_A"] is not None and m["final_recon_loss_B"] is not None:
        m["avg_recon_loss"] = (m["final_recon_loss_A"] + m["final_recon_loss_B"]) / 2
    return m


def trial_command(args, hp, result_dir):
    cmd = [sys.executable, "-m", "discogan_modernized_b200.image_translation", "--task_name", args.task_name,
           "--model_arch", args.model_arch, "--epochs", str(args.base_epochs), "--batch_size", str(args.batch_size),
           "--image_size", str(args.image_size), "--results_dir", str(result_dir / "results"),
           "--models_dir", str(result_dir / "models"), "--image_save_interval", "0"]
    for k, v in hp.items():
        cmd += [f"--{k}", str(v)]
    if args.style_A:
        cmd += ["--style_A", args.style_A]
    if args.style_B:
        cmd += ["--style_B", args.style_B]
    if args.data_A and args.data_B:
        cmd += ["--data_A", args.data_A, "--data_B", args.data_B]
    else:
        cmd += ["--synthetic", "--iters_per_epoch", str(args.iters_per_epoch)]
    return cmd


def run_trial(args, trial_id, gpu_id, hp):
    """Start one training subprocess on one GPU (reference :150-198)."""
    result_dir = Path(args.output_dir) / f"trial_{trial_id:03d}"
    result_dir.mkdir(parents=True, exist_ok=True)
    log_file = result_dir / "train.log"
    cmd = trial_command(args, hp, result_dir)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(gpu_id))
    root = str(Path(__file__).resolve().parent.parent)
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    f = open(log_file, "w")
    proc = subprocess.Popen(cmd, stdout=f, stderr=subprocess.STDOUT, env=env)
    info = {"trial_id": trial_id, "gpu_id": gpu_id, "hyperparameters": hp, "command": " ".join(cmd),
            "log_file": str(log_file), "start_time": datetime.now().strftime("%Y%m%d_%H%M%S"), "pid": proc.pid,
            "status": "running", "_t0": time.time(), "_file": f, "_best": float("inf"), "_stale": 0, "_seen": 0}
    return proc, result_dir, info


def check_early_stop(args, info):
    """True when the average reconstruction loss has not improved for `patience` new log lines (reference :211-241)."""
    if not args.early_stopping:
        return False
    try:
        found = re.findall(r"RECON: (\d+\.\d+)/(\d+\.\d+)", Path(info["log_file"]).read_text())
    except OSError:
        return False
    for a, b in found[info["_seen"]:]:
        avg = (float(a) + float(b)) / 2
        if avg < info["_best"]:
            info["_best"], info["_stale"] = avg, 0
        else:
            info["_stale"] += 1
    info["_seen"] = len(found)
    return info["_stale"] >= args.patience


def finish(proc, result_dir, info, status):
    info["_file"].close()
    info.update(status=status, end_time=datetime.now().strftime("%Y%m%d_%H%M%S"), duration=time.time() - info["_t0"],
                returncode=proc.returncode, metrics=extract_metrics(info["log_file"]))
    public = {k: v for k, v in info.items() if not k.startswith("_")}
    (result_dir / "trial_info.json").write_text(json.dumps(public, indent=2))
    return public


def analyze_results(trials, output_dir):
    """Rank the finished trials by average reconstruction loss and write the summary."""
    ok = [t for t in trials if t["metrics"].get("avg_recon_loss") is not None]
    ok.sort(key=lambda t: t["metrics"]["avg_recon_loss"])
    summary = {"n_trials": len(trials), "n_with_metrics": len(ok), "best": ok[0] if ok else None,
               "ranking": [{"trial_id": t["trial_id"], "avg_recon_loss": t["metrics"]["avg_recon_loss"],
                            "hyperparameters": t["hyperparameters"]} for t in ok]}
    Path(output_dir).mkdir(parents=True, exist_ok=True)
    (Path(output_dir) / "summary.json").write_text(json.dumps(summary, indent=2))
    return summary


def main(argv=None):
    args = parse_args(argv)
    rng = random.Random(args.seed)
    queue_ = list(enumerate(sample_hyperparameters(args.trials, rng)))
    gpus = [g.strip() for g in args.gpus.split(",") if g.strip()]
    running, done = {}, []
    while queue_ or running:
        for g in gpus:
            if g not in running and queue_:
                tid, hp = queue_.pop(0)
                running[g] = run_trial(args, tid, g, hp)
                print(f"started trial {tid} on GPU {g}: {hp}")
        time.sleep(min(args.poll_seconds, 1.0) if not args.early_stopping else args.poll_seconds)
        for g, (proc, rdir, info) in list(running.items()):
            if proc.poll() is not None:
                done.append(finish(proc, rdir, info, "completed" if proc.returncode == 0 else "failed"))
                del running[g]
            elif check_early_stop(args, info):
                proc.terminate()
                proc.wait()
                done.append(finish(proc, rdir, info, "early_stopped"))
                del running[g]
    summary = analyze_results(done, args.output_dir)
    if summary["best"]:
        print(f"best: trial {summary['best']['trial_id']} avg recon {summary['best']['metrics']['avg_recon_loss']:.4f} "
              f"{summary['best']['hyperparameters']}")
    return summary


if __name__ == "__main__":
    main()
