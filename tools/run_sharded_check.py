"""torchrun check of the multi-GPU paths under NCCL: the ray-sharded stage-1 render, the ray-sharded relit-view chain
(pipeline.extract_and_shade_sharded) and the surface-balanced stage-2 view (BASELINE config 4) each equal their single-GPU result.
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/run_sharded_check.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psnerf_b200 import pipeline, synth  # noqa: E402
from psnerf_b200.stage1 import NeuralNetwork, Renderer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
h = w = 96
cfg = synth.stage1_cfg(num_points_in=24, num_points_out=8, ray_marching_steps=128)
torch.manual_seed(0)
net = NeuralNetwork(cfg).eval()
net.precision = os.environ.get("PSN_PRECISION", "tc_two_level")
r = Renderer(net, cfg, device=dev)
K, pose = synth.intrinsics(h, w), synth.look_at_pose(20.0, 10.0)
full = pipeline.render_stage1_view_sharded(r, h, w, K, pose, rank, world)
one = pipeline.render_stage1_view_sharded(r, h, w, K, pose, 0, 1)
err = float((full - one).abs().max())
t = torch.tensor([err], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("stage-1 render, sharded-vs-single max abs diff over %d ranks: %.3e" % (world, float(t)))
    assert float(t) < 1e-6, "sharded render differs from the single-GPU render"

# the relit-view chain (headline workload) and the stage-2 view of config 4
from psnerf_b200.stage2 import PSNetwork  # noqa: E402
torch.manual_seed(0)
ps = PSNetwork(synth.stage2_conf()).to(dev).eval()
ps.precision = net.precision
lights = synth.lights(12, axis=tuple((-pose[0, :3, 2]).tolist())).to(dev)
shp, out = pipeline.extract_and_shade_sharded(r, ps, h, w, K, pose, lights, rank, world)
shp1, out1 = pipeline.extract_and_shade(r, ps, h, w, K, pose, lights)
errs = {"mask": float((shp["mask"] != shp1["mask"]).sum()), "points": float((shp["points"] - shp1["points"]).abs().max()),
        "shadow": float((shp["visibility"] - shp1["visibility"]).abs().max()),
        "rgb": float((out["sg_rgb_values"] - out1["sg_rgb_values"]).abs().max()),
        "normal_pred": float((out["normal_pred"] - out1["normal_pred"].reshape(out["normal_pred"].shape)).abs().max())}
inp = synth.stage2_input(h, w, 12, all_surface=False, seed=9, mask_frac=0.3)
inp = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inp.items()}
a = pipeline.render_stage2_view_sharded(ps, inp, lights, rank, world)
b = pipeline.render_stage2_view(ps, inp, lights)
errs["config4_rgb"] = float((a["sg_rgb_values"] - b["sg_rgb_values"]).abs().max())
errs["config4_vis"] = float((a["visibility"] - b["visibility"]).abs().max())
t = torch.tensor(list(errs.values()), device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("relit chain / config 4, sharded-vs-single max abs diffs:", dict(zip(errs, t.tolist())))
    assert float(t.max()) < 1e-6, "sharded result differs from the single-GPU result"
    print("OK")
dist.destroy_process_group()
