"""Entry-point loops and the fused extract->shade pipeline (host code over the CUDA path)."""
import pytest
import torch

import psnerf_oracle as O
import util
from psnerf_b200 import pipeline, synth

pytestmark = pytest.mark.gpu


def _models(prec="fp32"):
    from psnerf_b200.stage1 import NeuralNetwork, Renderer
    from psnerf_b200.stage2 import PSNetwork
    cfg, s1 = util.stage1_state_dicts()
    conf, s2 = util.stage2_state_dicts()
    net = NeuralNetwork(cfg)
    net.load_state_dict(s1["init"])
    net.eval()
    net.precision = prec
    r = Renderer(net, cfg, device=torch.device("cuda"))
    ps = PSNetwork(conf)
    ps.load_state_dict(s2["trained"])
    ps = ps.cuda().eval()
    ps.precision = prec
    return cfg, s1["init"], r, conf, s2["trained"], ps


def test_stage1_view_layout_matches_reference_eval():
    cfg, sd, r, *_ = _models()
    h, w = 12, 20  # non-square: exercises the x-major -> [h,w] transpose of stage1/eval.py:22
    K, pose = synth.intrinsics(h, w), synth.look_at_pose(15.0, 10.0)
    img = pipeline.render_stage1_view(r, h, w, K, pose)
    ref = O.unisurf_render(sd, cfg, O.arange_pixels((h, w))[0], K, pose, it=100000)
    ref_rgb = ref["rgb"][0].reshape(w, h, 3).permute(1, 0, 2)
    ref_mask = ref["mask_pred"].reshape(w, h).permute(1, 0)
    agree = img["mask"].cpu() == ref_mask
    assert img["rgb"].shape == (h, w, 3) and agree.float().mean() > 0.98
    assert util.max_abs(img["rgb"].cpu()[agree], ref_rgb[agree]) < 2e-4


@pytest.mark.parametrize("prec", ["fp32", "tc"])
def test_extract_and_shade_equals_two_stage_oracle(prec):
    cfg, sd1, r, conf, sd2, ps = _models(prec)
    h = w = 20
    K, pose = synth.intrinsics(h, w), synth.look_at_pose(-20.0, 25.0)
    lights = synth.lights(6, seed=8, axis=tuple((-pose[0, :3, 2]).tolist()))
    shp, out = pipeline.extract_and_shade(r, ps, h, w, K, pose, lights, light_batch=4)
    pix = O.arange_pixels((h, w))[0]
    ref_shape = O.shape_extract(sd1, cfg, pix, K, pose)
    agree = shp["mask"].cpu() == ref_shape["mask"]
    assert agree.float().mean() > 0.98
    # feed the oracle's stage 2 with the kernel's own surface so that only the shading is compared
    inp = {"intrinsics": torch.eye(4).unsqueeze(0), "uv": pix.float(), "pose": pose, "object_mask": shp["mask"].cpu(),
           "surface_mask": shp["mask"].cpu(), "points": shp["points"].cpu(), "normal": shp["normal"].cpu(), "light_direction": lights}
    inp["intrinsics"][0, 0, 0] = inp["intrinsics"][0, 1, 1] = K[0, 0, 0]
    inp["intrinsics"][0, 0, 2], inp["intrinsics"][0, 1, 2] = K[0, 0, 2], K[0, 1, 2]
    with torch.no_grad():
        ref = O.psnetwork_forward(sd2, conf, inp)
    assert out["sg_rgb_values"].shape == (6, h * w, 3)
    util.bound("extract_and_shade/%s/rgb" % prec, util.max_abs(out["sg_rgb_values"].cpu(), ref["sg_rgb_values"]), 1e-5)
    util.bound("extract_and_shade/%s/normal_pred" % prec, util.max_abs(out["normal_pred"].cpu(), ref["normal_pred"]), 1e-5)
    # the shadow pass is part of the fused call (SURVEY.md 8f-1 lists rendering.py:378-408): [L, N] transmittances, 1 off the surface
    assert shp["visibility"].shape == (6, h * w) and float((shp["visibility"][:, ~shp["mask"][0]] - 1).abs().max()) == 0.0
    with torch.no_grad():
        vref = O.light_visibility(sd1, cfg["model"], shp["points"][0][shp["mask"][0]].cpu(), lights).view(6, -1)
    util.bound("extract_and_shade/%s/shadow" % prec, util.max_abs(shp["visibility"][:, shp["mask"][0]].cpu(), vref), 5e-5)


def test_sharded_render_single_rank_is_identity():
    cfg, sd, r, *_ = _models()
    h = w = 16
    K, pose = synth.intrinsics(h, w), synth.look_at_pose(15.0, 10.0)
    full = pipeline.render_stage1_view_sharded(r, h, w, K, pose, rank=0, world=1)
    one = pipeline.render_stage1_view(r, h, w, K, pose, pixels=O.arange_pixels((h, w))[0].cuda())
    assert torch.equal(full[:, :3], one["rgb"][0])


def test_stage2_sharded_view_matches_unsharded():
    """render_stage2_view_sharded with the real PSNetwork: a single rank reproduces render_stage2_view bit for bit, and the two
    surface-balanced shards of a 2-rank deal, rendered one after the other on this GPU, reassemble to the same view."""
    from psnerf_b200 import sharding
    *_, conf, sd2, ps = _models("fp32")
    inp = synth.stage2_input(14, 12, 1, all_surface=False, seed=9, mask_frac=0.5)
    inp = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in inp.items()}
    lights = synth.lights(5).cuda()
    want = pipeline.render_stage2_view(ps, inp, lights, light_batch=2)
    one = pipeline.render_stage2_view_sharded(ps, inp, lights, 0, 1, light_batch=2)
    n = inp["uv"].shape[1]
    for k in ("sg_rgb_values", "visibility", "normal_pred", "sg_diffuse_albedo_values"):
        assert torch.equal(one[k], want[k].reshape(one[k].shape)), k
    full = torch.zeros(5, n, 3, device="cuda")
    for r in range(2):
        idx = sharding.shard_indices_by_mask(inp["surface_mask"][0], r, 2, tile=16).cuda()
        sub = dict(inp)
        for k in pipeline.PER_PIXEL_INPUTS:
            if k in inp and torch.is_tensor(inp[k]) and inp[k].dim() >= 2 and inp[k].shape[1] == n:
                sub[k] = torch.index_select(inp[k], 1, idx)
        full[:, idx] = pipeline.render_stage2_view(ps, sub, lights, light_batch=2)["sg_rgb_values"]
    assert util.max_abs(full.cpu(), want["sg_rgb_values"].cpu()) < 1e-6


@pytest.mark.parametrize("prec", ["fp32", "tc"])
def test_envmap_relighting_matches_oracle_loop(prec):
    """stage2/eval.py:173-231: ragged light batches over a small lat-long grid, RGB intensities, sum + clip / mean."""
    *_, conf, sd2, ps = _models(prec)
    inp = synth.stage2_input(10, 12, 1, all_surface=False, seed=5, mask_frac=0.6)
    xyz, _ = synth.latlong_light_grid(4)  # 4 x 8 = 32 lights
    env = torch.rand(32, 3, generator=torch.Generator().manual_seed(6)) * 0.2
    inp_cuda = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in inp.items()}
    got = pipeline.render_envmap_view(ps, inp_cuda, env, xyz, light_batch=12)
    with torch.no_grad():
        ref = O.envmap_relight(sd2, conf, inp, env, xyz, light_batch=12)
    tol = 2e-5 if prec == "fp32" else 2e-4
    assert util.max_abs(got["rgb"].cpu(), ref["rgb"]) < tol
    assert util.max_abs(got["visibility"].cpu(), ref["visibility"]) < 5 * tol


@pytest.mark.parametrize("prec", ["fp32", "tc"])
def test_occupancy_grid_logits_vs_oracle(prec):
    """Dense -logit lattice of the mesh-export path (extracting.py:84-96): 20^3 points (not a multiple of the 128-row tile)."""
    cfg, sd, r, *_ = _models(prec)
    nx, pad = 20, 0.1
    grid = pipeline.occupancy_grid_logits(r.model, nx, padding=pad)
    ax = torch.linspace(-0.5, 0.5, nx)
    p = (2.0 + pad) * torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
    with torch.no_grad():
        ref = O.network_forward(sd, cfg["model"], p, return_logits=True).reshape(nx, nx, nx)
    assert grid.shape == (nx, nx, nx)
    assert util.max_abs(grid.cpu(), ref) < (2e-5 if prec == "fp32" else 1e-4)
    # the level set the mesh extractor thresholds (occupancy 0.5 <=> logit 0) has the same sign pattern
    assert ((grid.cpu() > 0) == (ref > 0)).float().mean() > 0.999
