"""Minimal driver for ncu captures: N full unisurf steps (512x512x128spp) and one stage-2 shade (512x512, 96 lights)
on cuda:0.   python tools/profile_step.py [--steps 2] [--precision tc] [--stage2]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psnerf_b200 import synth  # noqa: E402
from psnerf_b200.stage1 import NeuralNetwork, Renderer  # noqa: E402
from psnerf_b200.stage2 import PSNetwork  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--precision", default="tc")
ap.add_argument("--stage2", action="store_true")
ap.add_argument("--res", type=int, default=512)
a = ap.parse_args()
dev = torch.device("cuda:0")
H = W = a.res
if not a.stage2:
    cfg = synth.stage1_cfg(num_points_in=96, num_points_out=32, ray_marching_steps=256)
    torch.manual_seed(0)
    net = NeuralNetwork(cfg).eval()
    net.precision = a.precision
    r = Renderer(net, cfg, device=dev)
    pix = synth.pixel_grid_xmajor(H, W).to(dev)
    K, pose = synth.intrinsics(H, W), synth.look_at_pose(20.0, 10.0)
    for _ in range(a.steps):
        out = r(pix, K, pose, None, "unisurf", add_noise=False, eval_=True, it=100000)
    torch.cuda.synchronize()
    print("hit rays:", int(out["mask_pred"].sum()), "rgb mean", float(out["rgb"].mean()))
else:
    conf = synth.stage2_conf()
    torch.manual_seed(0)
    m = PSNetwork(conf).to(dev).eval()
    m.precision = a.precision
    inp = synth.stage2_input(H, W, 96, all_surface=True)
    inp = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inp.items()}
    for _ in range(a.steps):
        out = m(inp)
    torch.cuda.synchronize()
    print("rgb mean", float(out["sg_rgb_values"].mean()))
