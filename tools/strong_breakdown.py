"""Where the time of one ray shard of the relit-view chain goes (1 GPU): the chain on 1/W of the rays, the row packing, and the
per-kernel sums - to separate kernel scaling losses from host-side overhead in the strong-scaling numbers.
   python tools/strong_breakdown.py [W ...]"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from psnerf_b200 import _binding as B, pipeline, sharding  # noqa: E402
from psnerf_b200.stage1.common import arange_pixels  # noqa: E402

dev = torch.device("cuda:0")
lib = B.load()
cfg, net, rend, conf, ps = bench.build_models(dev, "tc_two_level")
K, pose = bench.scene(0)
lights = bench.scene_lights(pose).to(dev)
H = W = bench.H
res = {}
for world in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
    idx = sharding.shard_indices(H * W, 0, world)
    pix = arange_pixels((H, W))[0][:, idx].to(dev)

    def chain():
        return pipeline.extract_and_shade(rend, ps, H, W, K, pose, lights, pixels=pix)

    def rows():
        return pipeline.extract_and_shade_rows(rend, ps, H, W, K, pose, lights, pix)

    def sharded():
        return pipeline.extract_and_shade_sharded(rend, ps, H, W, K, pose, lights, 0, 1) if world == 1 else rows()
    out = {}
    for name, fn in (("chain", chain), ("chain+rows", rows)):
        fn()
        lib.psn_profile_enable(1)
        ms = bench._time_cuda(fn, reps=3)
        kern = bench.collect_kernels(lib, 4)
        lib.psn_profile_enable(0)
        out[name] = {"ms": ms, "kernel_ms": sum(v["ms_per_step"] for v in kern.values()),
                     "kernels": {k: round(v["ms_per_step"], 2) for k, v in kern.items()}}
    res["1/%d of the rays" % world] = out
print(json.dumps(res, indent=1))
