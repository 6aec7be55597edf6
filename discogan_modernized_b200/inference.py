"""Re-hosted ``inference.py`` (reference :114-195): eval-mode generator forward from ``gen_B_final.pth`` /
``gen_A_final.pth``.  Both generators are loaded once (the reference reloads the reverse one per image, :183-187)
and images can be batched."""
import argparse
from pathlib import Path

import torch

from .model import Generator


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--model_path", required=True)
    p.add_argument("--input_path", required=True)
    p.add_argument("--output_dir", default="./inference_results")
    p.add_argument("--image_size", type=int, default=64)
    p.add_argument("--direction", default="AtoB", choices=["AtoB", "BtoA"])
    p.add_argument("--use_extra_layers", action="store_true")
    p.add_argument("--dataset_type", default=None)
    p.add_argument("--domain", default=None)
    p.add_argument("--batch_size", type=int, default=16)
    return p.parse_args(argv)


def load_generator(path, image_size, extra_layers=False, device="cuda"):
    g = Generator(extra_layers=extra_layers, image_size=image_size).to(device)
    g.load_state_dict(torch.load(path, map_location=device))
    return g.eval()


class GraphedGenerator:
    """Eval-mode generator forward replayed from a CUDA graph (one graph per batch size).  A 64x64 forward is ~35 kernels of
    a few microseconds each: launched one by one from Python it is host-bound (~0.5 ms per batch whatever the batch size);
    replayed it costs what the kernels cost.  The input is copied into a static buffer, the output is a static buffer that
    the next call with the same batch size overwrites (``translate`` clones what it keeps).  The graphs read the bf16
    GEMM copies of the weights: after changing the generator's parameters (``load_state_dict``, further training) call
    ``refresh()``, which re-packs them in place; BatchNorm running statistics are read by the replayed kernels directly."""

    def __init__(self, generator):
        from . import ops
        if generator.training:
            raise ValueError("GraphedGenerator is for eval-mode generators (call .eval() first)")
        self.g = generator
        self.ctx = ops.OpsContext()
        self._graphs = {}
        self._pool = None

    def refresh(self):
        """Re-pack the bf16 GEMM weights from the current fp32 parameters (in place: captured graphs stay valid)."""
        self.g._packed.refresh()

    @torch.no_grad()
    def __call__(self, x):
        from . import ops
        n = x.shape[0]
        ent = self._graphs.get(n)
        with ops.use_context(self.ctx):
            if ent is None:
                static_in = torch.empty_like(x)
                static_in.copy_(x)
                self.g(static_in)                                 # eager once: sizes scratch buffers, packs the weights
                gen = self.ctx.generation
                graph = torch.cuda.CUDAGraph()
                if self._pool is None:
                    self._pool = torch.cuda.graph_pool_handle()
                torch.cuda.synchronize()
                with torch.cuda.graph(graph, pool=self._pool):
                    static_out = self.g(static_in)
                if self.ctx.generation != gen:                    # a scratch buffer moved while capturing: start over
                    self._graphs.clear()
                    return self(x)
                ent = self._graphs[n] = (graph, static_in, static_out)
            graph, static_in, static_out = ent
            static_in.copy_(x, non_blocking=True)
            graph.replay()
        return static_out


@torch.no_grad()
def translate(generator, images, batch_size=16, graphs=True):
    """images: fp32 [N,3,S,S] in [0,1] (any device) -> generated images on the GPU.  ``generator``: an eval-mode
    ``Generator`` or a ``GraphedGenerator``; with ``graphs`` a plain generator is wrapped (and the wrapper cached on it)."""
    if graphs and not isinstance(generator, GraphedGenerator) and not generator.training:
        wrapped = getattr(generator, "_graphed", None)
        if wrapped is None:
            wrapped = generator._graphed = GraphedGenerator(generator)
        generator = wrapped
    outs = []
    for i in range(0, len(images), batch_size):
        y = generator(images[i:i + batch_size].cuda(non_blocking=True).float().contiguous())
        outs.append(y.clone() if isinstance(generator, GraphedGenerator) else y)
    return torch.cat(outs)


def main(argv=None):
    from PIL import Image
    import numpy as np
    args = parse_args(argv)
    fwd, rev = ("gen_B_final.pth", "gen_A_final.pth") if args.direction == "AtoB" else ("gen_A_final.pth", "gen_B_final.pth")
    gen = load_generator(Path(args.model_path) / fwd, args.image_size, args.use_extra_layers)
    rev_path = Path(args.model_path) / rev
    rgen = load_generator(rev_path, args.image_size, args.use_extra_layers) if rev_path.exists() else None
    inp = Path(args.input_path)
    files = sorted(list(inp.glob("*.jpg")) + list(inp.glob("*.png"))) if inp.is_dir() else [inp]
    out_dir = Path(args.output_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    imgs = torch.stack([torch.from_numpy(np.asarray(Image.open(f).convert("RGB").resize((args.image_size,) * 2)).copy())
                        .permute(2, 0, 1).float() / 255.0 for f in files])
    gen_out = translate(gen, imgs, args.batch_size)
    rec = translate(rgen, gen_out, args.batch_size) if rgen is not None else None
    for i, f in enumerate(files):
        panels = [imgs[i], gen_out[i].cpu()] + ([rec[i].cpu()] if rec is not None else [])
        strip = torch.cat(panels, dim=2).clamp(0, 1).mul(255).byte().permute(1, 2, 0).numpy()
        Image.fromarray(strip).save(out_dir / f"{f.stem}_result.png")
        print(f"saved {out_dir / (f.stem + '_result.png')}")


if __name__ == "__main__":
    main()
