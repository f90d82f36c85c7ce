"""smoke(): one tiny stage-1 render and one stage-2 shade on cuda:0, checked against the CPU oracle.
(The oracle is imported here only as the checker - see oracle/psnerf_oracle.py header.)"""
import os
import sys

import torch


def run_smoke(verbose=True):
    root = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(root, "oracle"))
    import psnerf_oracle as O
    from psnerf_b200 import synth
    from psnerf_b200.stage1 import NeuralNetwork, Renderer
    from psnerf_b200.stage2 import PSNetwork

    dev = torch.device("cuda:0")
    res = {}
    from psnerf_b200 import engine
    precisions = ["fp32", "tc"] if engine.tc_available() else ["fp32"]
    # ---- stage 1: 16x16 view, 64 march steps, 12+4 samples per ray
    cfg = synth.stage1_cfg(num_points_in=12, num_points_out=4, ray_marching_steps=64)
    torch.manual_seed(0)
    net = NeuralNetwork(cfg).eval()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    h = w = 16
    pix, K, pose = synth.pixel_grid_xmajor(h, w), synth.intrinsics(h, w), synth.look_at_pose(15.0, 10.0)
    ref = O.unisurf_render(sd, cfg, pix, K, pose, it=100000)
    rend = Renderer(net, cfg, device=dev)
    for prec in precisions + (["tc_mixed"] if "tc" in precisions else []):  # tc_mixed: the radiance program bench.py runs by default
        net.precision = prec
        out = rend(pix.to(dev), K, pose, None, "unisurf", add_noise=False, eval_=True, it=100000)
        agree = out["mask_pred"].cpu() == ref["mask_pred"]
        err = float((out["rgb"][0].cpu() - ref["rgb"][0])[agree].abs().max())
        res["stage1_%s" % prec] = (float(agree.float().mean()), err)
        assert agree.float().mean() > 0.98 and err < 2e-3, ("stage-1 smoke failed", prec, res)
    # ---- stage 2: 12x12 pixels, 5 lights
    conf = synth.stage2_conf()
    torch.manual_seed(0)
    ps = PSNetwork(conf)
    sd2 = {k: v.detach().clone() for k, v in ps.state_dict().items()}
    inp = synth.stage2_input(12, 12, 5, all_surface=False)
    with torch.no_grad():
        ref2 = O.psnetwork_forward(sd2, conf, inp)
    ps = ps.to(dev).eval()
    for prec in precisions:
        ps.precision = prec
        out2 = ps({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inp.items()})
        err2 = float((out2["sg_rgb_values"].cpu() - ref2["sg_rgb_values"]).abs().max())
        res["stage2_%s" % prec] = err2
        assert err2 < 2e-3, ("stage-2 smoke failed", prec, res)
    torch.cuda.synchronize()
    if verbose:
        print("psnerf_b200 smoke OK:", res)
    return res
