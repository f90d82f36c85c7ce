"""smoke(): one tiny pass of every part of the hot path on cuda:0, checked against the CPU oracle.
(The oracle is imported here only as the checker - see oracle/psnerf_oracle.py header; this file sits next to __graft_entry__.py,
outside the product package.)

  1. stage-1 unisurf render, 16 x 16 view (BASELINE configs[0] / [1] path) under the fp32 kernels and the default tensor-core program
  2. the headline chain, pipeline.extract_and_shade: surface search + normals + shadow-ray visibility (box-culled lists) + stage-2
     shading under 5 lights, against the oracle's shape_extract / light_visibility / PSNetwork.forward
Gates: 1e-4 max-abs on the pixels whose hit / miss decision agrees (the north star's tolerance)."""
import os
import sys

import torch


def run_smoke(verbose=True):
    root = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(root, "oracle"))
    import psnerf_oracle as O
    from psnerf_b200 import engine, pipeline, synth
    from psnerf_b200.stage1 import NeuralNetwork, Renderer
    from psnerf_b200.stage2 import PSNetwork

    dev = torch.device("cuda:0")
    res = {}
    assert engine.tc_available(), "sm_100a build without the tcgen05 kernels"
    precisions = ["fp32", "tc_two_level"]
    # ---- stage 1: 16x16 view, 64 march steps, 12+4 samples per ray
    cfg = synth.stage1_cfg(num_points_in=12, num_points_out=4, ray_marching_steps=64)
    torch.manual_seed(0)
    net = NeuralNetwork(cfg).eval()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    h = w = 16
    pix, K, pose = synth.pixel_grid_xmajor(h, w), synth.intrinsics(h, w), synth.look_at_pose(15.0, 10.0)
    ref = O.unisurf_render(sd, cfg, pix, K, pose, it=100000)
    rend = Renderer(net, cfg, device=dev)
    for prec in precisions:
        net.precision = prec
        out = rend(pix.to(dev), K, pose, None, "unisurf", add_noise=False, eval_=True, it=100000)
        agree = out["mask_pred"].cpu() == ref["mask_pred"]
        err = float((out["rgb"][0].cpu() - ref["rgb"][0])[agree].abs().max())
        nerr = float((out["normal_pred"][0].cpu() - ref["normal_pred"][0])[agree].abs().max())
        res["stage1_%s" % prec] = {"mask_agree": float(agree.float().mean()), "rgb": err, "normal": nerr}
        assert agree.float().mean() > 0.98 and err < 1e-4 and nerr < 1e-4, ("stage-1 smoke failed", prec, res)
    # ---- the relit-view chain: 16x16 pixels, 5 lights, shadow pass on
    conf = synth.stage2_conf()
    torch.manual_seed(0)
    ps = PSNetwork(conf)
    sd2 = {k: v.detach().clone() for k, v in ps.state_dict().items()}
    ps = ps.to(dev).eval()
    lights = synth.lights(5, seed=8, axis=tuple((-pose[0, :3, 2]).tolist()))
    shp_ref = O.shape_extract(sd, cfg, pix, K, pose)
    for prec in precisions:
        net.precision = prec
        ps.precision = prec
        shp, out2 = pipeline.extract_and_shade(rend, ps, h, w, K, pose, lights)
        agree = shp["mask"].cpu() == shp_ref["mask"]
        m = shp["mask"][0].cpu()
        with torch.no_grad():
            vis_ref = O.light_visibility(sd, cfg["model"], shp["points"][0].cpu()[m], lights).view(5, -1)
            K4 = torch.eye(4).unsqueeze(0)
            K4[0, 0, 0] = K4[0, 1, 1] = K[0, 0, 0]
            K4[0, 0, 2], K4[0, 1, 2] = K[0, 0, 2], K[0, 1, 2]
            inp = {"intrinsics": K4, "uv": pix.float(), "pose": pose, "object_mask": shp["mask"].cpu(), "surface_mask": shp["mask"].cpu(),
                   "points": shp["points"].cpu(), "normal": shp["normal"].cpu(), "light_direction": lights}
            ref2 = O.psnetwork_forward(sd2, conf, inp)
        e_vis = float((shp["visibility"].cpu()[:, m] - vis_ref).abs().max()) if int(m.sum()) else 0.0
        e_rgb = float((out2["sg_rgb_values"].cpu() - ref2["sg_rgb_values"]).abs().max())
        e_pts = float((shp["points"][0].cpu() - shp_ref["points"][0])[agree[0] & m].abs().max())
        res["relit_%s" % prec] = {"mask_agree": float(agree.float().mean()), "points": e_pts, "shadow_visibility": e_vis, "rgb": e_rgb,
                                  "surface_points": int(m.sum())}
        assert agree.float().mean() > 0.98 and e_vis < 1e-4 and e_rgb < 1e-4 and e_pts < 1e-4, ("relit smoke failed", prec, res)
    torch.cuda.synchronize()
    if verbose:
        print("psnerf_b200 smoke OK:", res)
    return res
