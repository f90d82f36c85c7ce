"""Parity at the BASELINE configurations' own parameters, at the north star's gate (1e-4).

The fixture tests run 14 x 14 ... 24 x 24 views with few samples; here the kernels run the FULL 512 x 512 views of
  configs[1]  stage-1 unisurf render, 96 + 32 samples per ray, 256 march steps
  configs[2]  the relit view: shape_extract (512 march steps) + shadow-ray visibility (96 lights x 128 steps) + stage-2 shading (96 lights)
and are compared with the CPU oracle on a strided sub-grid of the same view (rays are independent, so the oracle only has to run on
the sub-grid; the GPU values are taken from the full-view result at those pixels).  Both weight sets: the reference constructors'
("init": geometric-init sphere) and the perturbed "trained" variant; every tensor-core program the package ships.

Gates: max-abs over O(1) quantities on the pixels whose discrete hit / miss decision agrees.  1e-4 is the north star's number;
where a quantity is held tighter the gate is 3 x the error measured on the B200 (profiles/r2_parity_errlog.jsonl)."""
import pytest
import torch

import psnerf_oracle as O
import util
from psnerf_b200 import pipeline, synth

pytestmark = pytest.mark.gpu

H = W = 512
GRID = 16          # 256 sample pixels, every 32nd pixel in x and y
L = 96
PRECS = ["tc", "tc_mixed", "tc_two_level"]
# measured maxima (run 4, all precisions and weight sets): rgb 1.0e-5, acc 1.6e-5, normal 1.4e-5, points 2.1e-6, shadow 1.4e-5;
# stage 2 on identical inputs: rgb 8e-7, albedo 6e-8, normal 1.5e-6, visibility 1.3e-7, specular 2.3e-5 (exp(lambda (h.n - 1)) with lambda = e^10)
GATE = {"rgb": 5e-5, "acc": 5e-5, "normal": 5e-5, "points": 1e-5, "shadow": 5e-5, "s2_rgb": 1e-5, "s2_albedo": 1e-5, "s2_normal": 1e-5,
        "s2_vis": 1e-5, "s2_spec": 1e-4}


def _sample():
    step = W // GRID
    xs = torch.arange(GRID) * step + step // 2 + 3
    gx, gy = torch.meshgrid(xs, xs, indexing="ij")
    pix = torch.stack([gx, gy], -1).long().view(1, -1, 2)
    return pix, pix[0, :, 0] * H + pix[0, :, 1]


def _cfg():
    return synth.stage1_cfg(num_points_in=96, num_points_out=32, ray_marching_steps=256)


def _view():
    pose = synth.look_at_pose(20.0, 10.0)
    return synth.intrinsics(H, W), pose, synth.lights(L, axis=tuple((-pose[0, :3, 2]).tolist()))


@pytest.fixture(scope="module")
def oracle_results():
    """Oracle outputs on the sample pixels for both weight sets (about 10 s of CPU work each)."""
    _, s1 = util.stage1_state_dicts()
    conf, s2 = util.stage2_state_dicts()
    cfg = _cfg()
    K, pose, lights = _view()
    pix, _ = _sample()
    res = {}
    for variant in ("init", "trained"):
        render = O.unisurf_render(s1[variant], cfg, pix, K, pose, it=100000)
        shp = O.shape_extract(s1[variant], cfg, pix, K, pose, visibility=True, light_dir=lights)
        Ks = torch.eye(4).unsqueeze(0)
        Ks[0, 0, 0] = Ks[0, 1, 1] = K[0, 0, 0]
        Ks[0, 0, 2], Ks[0, 1, 2] = K[0, 0, 2], K[0, 1, 2]
        inp = {"intrinsics": Ks, "uv": pix.float(), "pose": pose, "object_mask": shp["mask"], "surface_mask": shp["mask"],
               "points": shp["points"], "normal": shp["normal"], "light_direction": lights}
        with torch.no_grad():
            s2out = O.psnetwork_forward(s2[variant], conf, inp)
        res[variant] = (render, shp, s2out)
    return res


def _models(variant, prec):
    from psnerf_b200.stage1 import NeuralNetwork, Renderer
    from psnerf_b200.stage2 import PSNetwork
    _, s1 = util.stage1_state_dicts()
    conf, s2 = util.stage2_state_dicts()
    cfg = _cfg()
    net = NeuralNetwork(cfg)
    net.load_state_dict(s1[variant])
    net = net.cuda().eval()
    net.precision = prec
    ps = PSNetwork(conf)
    ps.load_state_dict(s2[variant])
    ps = ps.cuda().eval()
    ps.precision = prec
    return Renderer(net, cfg, device=torch.device("cuda")), ps


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("variant", ["init", "trained"])
def test_stage1_render_512x512x128spp_vs_oracle(oracle_results, variant, prec):
    """BASELINE configs[1] at full size against the oracle on the sample pixels."""
    ref = oracle_results[variant][0]
    rend, _ = _models(variant, prec)
    K, pose, _ = _view()
    _, idx = _sample()
    out = rend(synth.pixel_grid_xmajor(H, W).cuda(), K, pose, None, "unisurf", add_noise=False, eval_=True, it=100000)
    agree = out["mask_pred"][idx].cpu() == ref["mask_pred"]
    assert agree.float().mean() >= 0.99
    tag = "at_size/stage1_render/%s/%s/" % (prec, variant)
    util.bound(tag + "rgb", util.max_abs(out["rgb"][0][idx].cpu()[agree], ref["rgb"][0][agree]), GATE["rgb"])
    util.bound(tag + "acc", util.max_abs(out["acc_map"][0][idx].cpu()[agree], ref["acc_map"][0][agree]), GATE["acc"])
    util.bound(tag + "normal", util.max_abs(out["normal_pred"][0][idx].cpu()[agree], ref["normal_pred"][0][agree]), GATE["normal"])
    assert O.psnr(out["rgb"][0][idx].cpu()[agree], ref["rgb"][0][agree]) > 70.0


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("variant", ["init", "trained"])
def test_relit_view_512x512x128x96L_vs_oracle(oracle_results, variant, prec):
    """The headline chain (pipeline.extract_and_shade with the shadow pass) at full size against the oracle's two-stage chain."""
    _, shp, s2out = oracle_results[variant]
    rend, ps = _models(variant, prec)
    K, pose, lights = _view()
    _, idx = _sample()
    gshp, gout = pipeline.extract_and_shade(rend, ps, H, W, K, pose, lights)
    agree = gshp["mask"][0][idx].cpu() == shp["mask"][0]
    both = agree & shp["mask"][0]
    assert agree.float().mean() >= 0.99 and int(both.sum()) > 20
    tag = "at_size/relit/%s/%s/" % (prec, variant)
    util.bound(tag + "points", util.max_abs(gshp["points"][0][idx].cpu()[both], shp["points"][0][both]), GATE["points"])
    util.bound(tag + "normal", util.max_abs(gshp["normal"][0][idx].cpu()[both], shp["normal"][0][both]), GATE["normal"])
    util.bound(tag + "shadow", util.max_abs(gshp["visibility"][:, idx].cpu()[:, both], shp["visibility"][:, both]), GATE["shadow"])
    # Stage 2 on IDENTICAL inputs (north star: "within 1e-4 relative on identical inputs"): the oracle's PSNetwork.forward is fed with
    # the kernels' own surface at the sample pixels.  Chained to the ORACLE's surface instead, the 2^9-frequency point encoding of
    # stage 2 turns the ~1e-6 difference between the two secant depths into ~1e-4 of normal_pred / specular - input conditioning, not
    # kernel error (logged below as s2chain_*, gated only loosely).
    conf, s2 = util.stage2_state_dicts()
    K4 = torch.eye(4).unsqueeze(0)
    K4[0, 0, 0] = K4[0, 1, 1] = K[0, 0, 0]
    K4[0, 0, 2], K4[0, 1, 2] = K[0, 0, 2], K[0, 1, 2]
    pix, _ = _sample()
    gm = gshp["mask"][:, idx].cpu()
    inp = {"intrinsics": K4, "uv": pix.float(), "pose": pose, "object_mask": gm, "surface_mask": gm,
           "points": gshp["points"][:, idx].cpu(), "normal": gshp["normal"][:, idx].cpu(), "light_direction": lights}
    with torch.no_grad():
        same = O.psnetwork_forward(s2[variant], conf, inp)
    for key, name in (("sg_rgb_values", "s2_rgb"), ("sg_diffuse_albedo_values", "s2_albedo"), ("normal_pred", "s2_normal"),
                      ("visibility", "s2_vis"), ("sg_specular_rgb_values", "s2_spec")):
        got = gout[key][:, idx].cpu()
        util.bound(tag + name, util.max_abs(got, same[key].reshape(got.shape)), GATE[name])
        util.bound(tag + name.replace("s2_", "s2chain_"), util.max_abs(got[:, agree], s2out[key].reshape(got.shape)[:, agree]), 2e-3)
    assert O.psnr(gout["sg_rgb_values"][:, idx].cpu(), same["sg_rgb_values"]) > 70.0
