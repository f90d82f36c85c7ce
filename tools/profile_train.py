"""Minimal driver for ncu launch lists of the train steps (BASELINE config 5 and its stage-1 analogue): bench.py's own train_steps().
   ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file X.csv python tools/profile_train.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda:0")
cfg, net, rend, conf, ps = bench.build_models(dev, "tc_two_level")
for rep in range(int(os.environ.get("PROFILE_TRAIN_REPS", "1"))):  # repeated: a one-off stall (lazy module load, allocator) shows up as an outlier
    print(json.dumps(bench.train_steps(dev, rend, ps, bench.scene(0), 1, 0)))
