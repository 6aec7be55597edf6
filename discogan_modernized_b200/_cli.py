"""Shared command-line plumbing for the re-hosted reference entry points.

Flags and defaults follow the reference parsers (image_translation.py:21-81, angle_pairing.py:22-72,
distributed_image_translation.py:48-126).  Data: image files are
decoded once and kept in HBM, preprocessed per batch by one CUDA launch (``dataset.py`` of this package), so batches
are either synthetic (``--synthetic``, the default when no folders are given; precedent
batch_size_optimization.py:62-63) or read from two image folders (``--data_A DIR --data_B DIR``) with the reference's
read_images semantics (crop halves, edge thickening, bilinear resize, /255, CHW).
"""
import argparse
import sys
from datetime import datetime
from pathlib import Path

import torch

from .train_step import DiscoGANTrainer, LossReadback, format_log_line


def build_parser(kind):
    angle = kind == "angle_pairing"
    p = argparse.ArgumentParser(description=f"DiscoGAN ({kind}) on B200 kernels")
    p.add_argument("--device", default="cuda")
    p.add_argument("--task_name", default="car2car" if angle else "facescrub")
    p.add_argument("--results_dir", default="./results/")
    p.add_argument("--models_dir", default="./models/")
    p.add_argument("--model_arch", default="discogan", choices=["discogan", "recongan", "gan"])
    p.add_argument("--epochs", type=int, default=10 if angle else 100)
    p.add_argument("--batch_size", type=int, default=64)
    p.add_argument("--learning_rate", type=float, default=2e-4)
    p.add_argument("--beta1", type=float, default=0.5)
    p.add_argument("--beta2", type=float, default=0.999)
    p.add_argument("--image_size", type=int, default=64)
    p.add_argument("--gan_curriculum", type=int, default=10000)
    p.add_argument("--starting_rate", type=float, default=0.9 if angle else 0.01)
    p.add_argument("--default_rate", type=float, default=0.9 if angle else 0.5)
    if not angle:
        p.add_argument("--style_A", default=None)
        p.add_argument("--style_B", default=None)
        p.add_argument("--constraint", default=None)
        p.add_argument("--constraint_type", default=None)
    p.add_argument("--n_test", type=int, default=200)
    p.add_argument("--update_interval", type=int, default=3)
    p.add_argument("--log_interval", type=int, default=50)
    p.add_argument("--image_save_interval", type=int, default=500 if angle else 1000)
    p.add_argument("--model_save_interval", type=int, default=10000)
    if kind == "distributed":
        p.add_argument("--distributed", action="store_true")
        p.add_argument("--local_rank", type=int, default=0)
        p.add_argument("--world_size", type=int, default=4)
        for n in ("gen_A", "gen_B", "dis_A", "dis_B"):
            p.add_argument(f"--load_{n}", default=None)
    # additions (not in the reference)
    p.add_argument("--resume", default=None, metavar="DIR[:TAG]",
                   help="continue a run: loads gen_A/gen_B/dis_A/dis_B_{TAG}.pth and train_state_{TAG}.pth (Adam moments, "
                        "iteration counter) from DIR; TAG defaults to 'final'")
    p.add_argument("--sample_mode", default="reference", choices=["reference", "eval"],
                   help="sample dumps: 'reference' keeps the generators in train mode under no_grad like the reference's "
                        "save_sample_images (advances BatchNorm running statistics); 'eval' leaves training untouched")
    p.add_argument("--n_samples", type=int, default=5, help="rows of the sample grid (reference: 5)")
    p.add_argument("--deterministic", action="store_true", help="bit-reproducible iterations (slightly slower)")
    p.add_argument("--synthetic", action="store_true", help="uniform-random A/B batches")
    p.add_argument("--data_A", default=None, help="folder of domain-A images (.jpg/.png)")
    p.add_argument("--data_B", default=None, help="folder of domain-B images")
    p.add_argument("--domain_A_type", default="auto", choices=["auto", "A", "B", "none"],
                   help="read_images domain handling for A: 'A' = left half of a side-by-side pair + edge thickening, "
                        "'B' = right half, 'none' = whole image; auto = by task name as in the reference")
    p.add_argument("--domain_B_type", default="auto", choices=["auto", "A", "B", "none"])
    p.add_argument("--iters_per_epoch", type=int, default=100, help="synthetic data only")
    p.add_argument("--max_iters", type=int, default=None)
    return p


class Batches:
    """Per-epoch A/B device batches.
    Folders (``--data_A DIR --data_B DIR``): ``dataset.DiscoGANDataset`` -- files decoded once into HBM, each batch one
    preprocessing launch with the reference's read_images arithmetic; domain types follow the task name
    (image_translation.py:243-251) unless --domain_A_type/--domain_B_type say otherwise.  The single-device entry points
    shuffle the two domains independently and drop the ragged tail like image_translation.py:296,309-319; the distributed
    one pairs index i of A with index i of B under the DistributedSampler index stream (distributed_...:182-226).
    Synthetic (``--synthetic`` or no folders): uniform-random batches (batch_size_optimization.py:62-63)."""

    def __init__(self, args, rank=0, world=1, kind="image_translation", device="cuda"):
        self.bs, self.S, self.rank, self.world, self.kind = args.batch_size, args.image_size, rank, world, kind
        self.device = device
        self.A = self.B = None
        self.ds = None
        if args.data_A and args.data_B and not args.synthetic:
            from . import dataset
            da, db = dataset.task_domains(args.task_name)
            da = args.domain_A_type if args.domain_A_type != "auto" else da
            db = args.domain_B_type if args.domain_B_type != "auto" else db
            fa, fb = dataset.list_images(args.data_A), dataset.list_images(args.data_B)
            n_test = min(args.n_test, len(fa) // 5, len(fb) // 5)          # held-out tail (dataset.py:113-116,165-168)
            self.test_files = (fa[len(fa) - n_test:], fb[len(fb) - n_test:], da, db) if n_test else None
            fa, fb = fa[:len(fa) - n_test], fb[:len(fb) - n_test]
            self.ds = dataset.DiscoGANDataset(fa, fb, {"none": None}.get(da, da), {"none": None}.get(db, db), self.S,
                                              device=device)
            if kind == "distributed":
                self.n_batches = -(-(-(-len(self.ds) // world)) // self.bs)
            else:
                self.n_batches = min(len(fa), len(fb)) // self.bs            # image_translation.py:296
            if self.n_batches == 0:
                raise ValueError(f"fewer images ({len(fa)}, {len(fb)}) than one batch of {self.bs}")
        else:
            self.test_files = None
            self.n_batches = args.iters_per_epoch

    def test_tensors(self, n):
        """The fixed test batch of the sample dumps (image_translation.py:236-252)."""
        if self.ds is not None and self.test_files:
            from . import dataset
            fa, fb, da, db = self.test_files
            return (dataset.read_images(fa[:n], {"none": None}.get(da, da), self.S, self.device),
                    dataset.read_images(fb[:n], {"none": None}.get(db, db), self.S, self.device))
        g = torch.Generator().manual_seed(4321)
        return (torch.rand(n, 3, self.S, self.S, generator=g).to(self.device),
                torch.rand(n, 3, self.S, self.S, generator=g).to(self.device))

    def epoch(self, epoch):
        if self.ds is not None:
            if self.kind == "distributed":
                yield from self.ds.batches(self.bs, epoch=epoch, rank=self.rank, world=self.world)
            else:
                yield from self.ds.batches(self.bs, epoch=epoch, independent=True, drop_last=True)
            return
        g = torch.Generator().manual_seed(1000 * self.rank + epoch)
        for _ in range(self.n_batches):
            yield (torch.rand(self.bs, 3, self.S, self.S, generator=g).pin_memory().to(self.device, non_blocking=True),
                   torch.rand(self.bs, 3, self.S, self.S, generator=g).pin_memory().to(self.device, non_blocking=True))


def save_models(tr, model_path, tag):
    """gen_A_{tag}.pth ... -- the reference's checkpoint names and state-dict layout (image_translation.py:420-432) --
    plus train_state_{tag}.pth: Adam moments / step state and the iteration counter, so that --resume continues the run
    (the reference restarts the optimiser and the GAN curriculum)."""
    model_path = Path(model_path)
    model_path.mkdir(parents=True, exist_ok=True)
    for name, net in (("gen_A", tr.G_A), ("gen_B", tr.G_B), ("dis_A", tr.D_A), ("dis_B", tr.D_B)):
        torch.save({k: v.detach().cpu().clone() for k, v in net.state_dict().items()}, model_path / f"{name}_{tag}.pth")
    torch.save(tr.training_state(), model_path / f"train_state_{tag}.pth")


def load_models(tr, model_path, tag="final", with_state=True):
    """Inverse of save_models.  Weights-only checkpoints written by the reference load too (with_state is then skipped
    with a note, and Adam / the curriculum start over exactly as in the reference)."""
    model_path = Path(model_path)
    dev = tr.device
    tr.load_weights({n: torch.load(model_path / f"{n}_{tag}.pth", map_location=dev)
                     for n in ("gen_A", "gen_B", "dis_A", "dis_B")})
    state = model_path / f"train_state_{tag}.pth"
    if with_state and state.exists():
        tr.load_training_state(torch.load(state, map_location="cpu"))
        return True
    if with_state:
        print(f"note: {state} not found -- weights restored, optimiser state and iteration counter start over",
              file=sys.stderr)
    return False


def warn_ignored(args):
    """Flags of the reference parsers that select dataset subsets this re-host has no use for."""
    for flag in ("style_B", "constraint", "constraint_type"):
        if getattr(args, flag, None):
            print(f"warning: --{flag} selects a CelebA attribute subset in the reference's dataset.py; this entry point reads "
                  "--data_A/--data_B folders (or --synthetic) and ignores it", file=sys.stderr)
    if getattr(args, "style_A", None):
        print("note: --style_A only names the results/models sub-directory here (reference: also a CelebA attribute)",
              file=sys.stderr)


def test_batches(args, data, device):
    """The fixed test tensors of the sample dumps (image_translation.py:236-252: held-out images per domain)."""
    return data.test_tensors(max(args.n_samples, 2))     # train-mode BatchNorm over a 1x1 map needs >= 2 images


def save_sample_grid(tr, test_A, test_B, save_dir, iteration, n_samples=5, mode="reference"):
    """``save_sample_images`` (image_translation.py:170-208): rows of A, B, A->B, B->A, A->B->A, B->A->B written to
    samples_iter_{iteration}.png (a plain image grid; the reference renders the same six columns with matplotlib)."""
    from PIL import Image
    AB, BA, ABA, BAB = tr.sample_images(test_A, test_B, mode)
    n = min(n_samples, test_A.shape[0])
    rows = [torch.cat([t[i].float().clamp(0, 1) for t in (test_A, test_B, AB, BA, ABA, BAB)], dim=2) for i in range(n)]
    grid = torch.cat(rows, dim=1).mul(255).round().byte().permute(1, 2, 0).cpu().numpy()
    save_dir = Path(save_dir)
    save_dir.mkdir(parents=True, exist_ok=True)
    out = save_dir / f"samples_iter_{iteration}.png"
    Image.fromarray(grid).save(out)
    return out


def run_training(args, kind, rank=0, world=1, process_group=None):
    variant = "angle_pairing" if kind == "angle_pairing" else "image_translation"
    device = f"cuda:{rank}" if kind == "distributed" else ("cuda" if args.device == "cuda" else args.device)
    ts = datetime.now().strftime("%Y%m%d_%H%M%S") + (f"_rank{rank}" if kind == "distributed" else "")
    sub = Path(args.task_name) / (getattr(args, "style_A", None) or "") / args.model_arch / ts
    result_path, model_path = Path(args.results_dir) / sub, Path(args.models_dir) / sub
    if rank == 0:
        warn_ignored(args)
    if kind == "distributed":
        torch.manual_seed(1234)                       # distributed_image_translation.py:372
    tr = DiscoGANTrainer(image_size=args.image_size, device=device, model_arch=args.model_arch,
                         learning_rate=args.learning_rate, beta1=args.beta1, beta2=args.beta2,
                         update_interval=args.update_interval, gan_curriculum=args.gan_curriculum,
                         starting_rate=args.starting_rate, default_rate=args.default_rate, variant=variant,
                         process_group=process_group, deterministic=args.deterministic)
    if kind == "distributed":                          # reference :379-393: weights only
        tr.load_weights({n: torch.load(getattr(args, f"load_{n}"), map_location=device) if getattr(args, f"load_{n}") else None
                         for n in ("gen_A", "gen_B", "dis_A", "dis_B")})
    if args.resume:
        path, _, tag = args.resume.partition(":")
        load_models(tr, path, tag or "final")
    data = Batches(args, rank, world, kind, device)
    total = args.epochs * data.n_batches
    log = None
    test_A = test_B = None
    if rank == 0:
        result_path.mkdir(parents=True, exist_ok=True)
        log = open(result_path / "training_log.txt", "a" if args.resume else "w")
        log.write(f"Training started at {ts}\nTask: {args.task_name}, Model: {args.model_arch}\n"
                  f"Batch size: {args.batch_size}, Learning rate: {args.learning_rate}\n\n")
        if args.image_save_interval > 0:
            test_A, test_B = test_batches(args, data, device)
    readback = LossReadback(tr.loss_buf)

    def emit(done):
        for it, losses in done:
            line = format_log_line(it, total, losses)
            print(line, flush=True)
            log.write(line + "\n")

    start_iter, done = tr.iters, False
    for epoch in range(args.epochs):
        for A, B in data.epoch(epoch):
            it = tr.iters
            tr.step(A, B)
            if rank == 0:
                if it % args.log_interval == 0:            # image_translation.py:393-398, without stalling the GPU
                    emit(readback.request(it))
                emit(readback.poll())
                if args.image_save_interval > 0 and it % args.image_save_interval == 0:     # :411-417 (incl. iteration 0)
                    save_sample_grid(tr, test_A, test_B, result_path / "samples", it, args.n_samples, args.sample_mode)
                if it % args.model_save_interval == 0 and (it > start_iter or it == 0):     # :420-424 (incl. iteration 0)
                    save_models(tr, model_path, it)
            if args.max_iters is not None and tr.iters - start_iter >= args.max_iters:
                done = True
                break
        if done:
            break
    if rank == 0:
        emit(readback.drain())
        save_models(tr, model_path, "final")
        log.close()
        print(f"Training completed. Final models saved to {model_path}")
        print(f"Results and logs saved to {result_path}")
    tr.result_path, tr.model_path = result_path, model_path
    return tr
