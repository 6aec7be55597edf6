#!/bin/bash
mkdir -p gpurun_out /tmp/dg && cd /tmp/dg && export PYTHONPATH=$GRAFT_REPO_ROOT
run() { name=$1; shift; echo "=== $name"; t0=$SECONDS; timeout 600 "$@" > $GRAFT_REPO_ROOT/gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s"; tail -n 4 $GRAFT_REPO_ROOT/gpurun_out/$name.log | cut -c1-250; }
run entry_it python -m discogan_modernized_b200.image_translation --synthetic --image_size 64 --batch_size 16 --epochs 1 --iters_per_epoch 160 --log_interval 50 --model_save_interval 100
run entry_ap python -m discogan_modernized_b200.angle_pairing --synthetic --image_size 64 --batch_size 8 --epochs 1 --iters_per_epoch 60 --log_interval 20
python - <<'PY'
import numpy as np
from PIL import Image
from pathlib import Path
Path("imgs").mkdir(exist_ok=True)
rng = np.random.default_rng(0)
for i in range(5):
    Image.fromarray(rng.integers(0, 255, (80, 80, 3), dtype=np.uint8)).save(f"imgs/im{i}.png")
PY
MP=$(ls -d models/facescrub/discogan/* | head -1)
run entry_inf python -m discogan_modernized_b200.inference --model_path $MP --input_path imgs --output_dir out --image_size 64 --batch_size 4
ls out | head -3; ls $MP
run entry_folder python -m discogan_modernized_b200.image_translation --data_A imgs --data_B imgs --image_size 64 --batch_size 2 --epochs 2 --log_interval 1
