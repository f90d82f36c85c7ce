"""torchrun check of the multi-GPU path: ray-sharded stage-1 render + one NCCL all_gather == single-GPU render.
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
net.precision = os.environ.get("PSN_PRECISION", "tc")
r = Renderer(net, cfg, device=dev)
K, pose = synth.intrinsics(h, w), synth.look_at_pose(20.0, 10.0)
full = pipeline.render_stage1_view_sharded(r, h, w, K, pose, rank, world)
one = pipeline.render_stage1_view_sharded(r, h, w, K, pose, 0, 1)
err = float((full - one).abs().max())
t = torch.tensor([err], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("sharded-vs-single max abs diff over %d ranks: %.3e" % (world, float(t)))
    assert float(t) < 1e-6, "sharded render differs from the single-GPU render"
    print("OK")
dist.destroy_process_group()
