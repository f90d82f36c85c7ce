"""Minimal driver for ncu captures: N full unisurf steps (512x512x128spp) and one stage-2 shade (512x512, 96 lights)
on cuda:0, or (--relit) the bench's headline chain: one relit view 512x512 x 128 x 96 lights through pipeline.extract_and_shade.
   python tools/profile_step.py [--steps 2] [--precision tc_two_level] [--stage2 | --relit]"""
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
ap.add_argument("--precision", default="tc_two_level")
ap.add_argument("--stage2", action="store_true")
ap.add_argument("--relit", action="store_true")
ap.add_argument("--res", type=int, default=512)
a = ap.parse_args()
dev = torch.device("cuda:0")
H = W = a.res
if a.relit:
    from psnerf_b200 import pipeline
    cfg = synth.stage1_cfg(num_points_in=96, num_points_out=32, ray_marching_steps=256)
    torch.manual_seed(0)
    net = NeuralNetwork(cfg).eval()
    net.precision = a.precision
    r = Renderer(net, cfg, device=dev)
    torch.manual_seed(0)
    ps = PSNetwork(synth.stage2_conf()).to(dev).eval()
    ps.precision = a.precision
    K, pose = synth.intrinsics(H, W), synth.look_at_pose(20.0, 10.0)
    lights = synth.lights(96, axis=tuple((-pose[0, :3, 2]).tolist())).to(dev)
    for _ in range(a.steps):
        shp, out = pipeline.extract_and_shade(r, ps, H, W, K, pose, lights)
    torch.cuda.synchronize()
    print("surface points:", int(shp["mask"].sum()), "vis mean", float(shp["visibility"].mean()), "rgb mean", float(out["sg_rgb_values"].mean()))
elif not a.stage2:
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
