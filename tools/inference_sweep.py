"""Eval-mode generator (AtoB) throughput sweep over batch sizes (BASELINE config 5, inference.py:149-172).
    python tools/inference_sweep.py [image_size] > gpurun_out/inference_sweep.json"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from discogan_modernized_b200 import model  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(1234)
G = model.Generator(extra_layers=True, image_size=S).cuda()
x = torch.rand(8, 3, S, S, device="cuda")
with torch.no_grad():
    G(x)                                      # one train-mode pass so the running statistics are not the init values
G.eval()
out = {"image_size": S, "sweep": []}
for B in (1, 2, 4, 8, 16, 32, 64, 128, 256):
    if S >= 512 and B > 64:
        break
    x = torch.rand(B, 3, S, S, device="cuda")
    g = torch.cuda.CUDAGraph()
    with torch.no_grad():
        for _ in range(3):
            y = G(x)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            y = G(x)
    iters = 50 if S < 512 else 10
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    out["sweep"].append({"batch": B, "ms": round(ms, 4), "images_per_s": round(B / ms * 1e3, 1)})
print(json.dumps(out))
