"""One D,G,G cycle of the train step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from discogan_modernized_b200 import DiscoGANTrainer  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else (32 if S == 512 else 64)
tr = DiscoGANTrainer(image_size=S, seed=1234, use_graphs=False)
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.rand(B, 3, S, S, device="cuda", generator=g)
Bt = torch.rand(B, 3, S, S, device="cuda", generator=g)
for _ in range(3):
    tr.step(A, Bt)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(3):
    tr.step(A, Bt)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("cycle done", tr.losses())
