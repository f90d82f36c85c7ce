"""GPU parity of the stage-1 path (through the C ABI via the drop-in modules) against the golden fixtures of
the real reference and against the CPU oracle on seeded inputs.

Tolerances: fp32 kernels re-associate sums, so continuous outputs are held to 2e-5 (abs, on O(1) quantities) /
1e-5 relative-L2 - tighter than the north star's 1e-4.  Discrete decisions (surface hit / miss, which march
interval) may flip on rays whose occupancy is within rounding of 0.5; those are counted separately and bounded."""
import numpy as np
import pytest
import torch

import psnerf_oracle as O
import util
from psnerf_b200 import synth

pytestmark = pytest.mark.gpu
def _precisions():
    from psnerf_b200 import engine
    try:
        return ["fp32", "tc", "tc_two_level"] if engine.tc_available() else ["fp32"]
    except Exception:
        return ["fp32"]


PRECISIONS = _precisions()
TOL = {"fp32": dict(rel=1e-5, abs=2e-5), "tc": dict(rel=5e-5, abs=1e-4), "tc_two_level": dict(rel=5e-5, abs=1e-4)}


# Gates on rendered quantities against the real reference's fixtures (all O(1) quantities, max-abs over the pixels whose discrete
# hit / miss decision agrees): the north star's 1e-4 or tighter - about 3 x the largest error measured on the B200 over every case,
# weight set and precision (profiles/r2_parity_errlog_final.jsonl: rgb 8.8e-6, acc 1.4e-5, normal 2.3e-5, points 1.2e-6, visibility
# 1.3e-5).  Round 1 needed 5e-4 / 2e-3 here; the secant refinement and the normal output now run on the fp32 kernels under every
# precision (csrc/api_stage1.cu accuracy policy), which took the surface points from 3.5e-5 to 1.2e-6 and everything downstream with them.
GATE = {"rgb": 3e-5, "acc": 5e-5, "normal": 1e-4, "points": 1e-5, "visibility": 5e-5, "light_visibility": 5e-5}


@pytest.fixture(scope="module")
def s1():
    return util.stage1_state_dicts()


def make_model(cfg, sd, prec):
    from psnerf_b200.stage1 import NeuralNetwork
    m = NeuralNetwork(cfg)
    m.load_state_dict(sd)
    m = m.cuda().eval()  # inference tests: train() mode with gradients enabled selects the differentiable path
    m.precision = prec
    return m


@pytest.mark.parametrize("prec", PRECISIONS)
@pytest.mark.parametrize("variant", ["init", "trained"])
def test_network_vs_golden(s1, variant, prec):
    cfg, sds = s1
    m = make_model(cfg, sds[variant], prec)
    g = util.golden("stage1_net")
    t = TOL[prec]
    pts, views = torch.from_numpy(g["pts"]).cuda(), torch.from_numpy(g["views"]).cuda()
    assert util.rel_l2(m.infer_occ(pts).cpu(), g[variant + "_infer_occ"]) < t["rel"]
    assert util.max_abs(m(pts, only_occupancy=True).cpu(), g[variant + "_alpha"]) < t["abs"]
    assert util.max_abs(m(pts, return_logits=True).cpu(), g[variant + "_neg_logit"]) < t["abs"] * 5
    assert util.rel_l2(m.gradient(pts).cpu(), g[variant + "_grad"]) < t["rel"] * 3
    rgb, a = m(pts, views, return_addocc=True)
    assert util.max_abs(rgb.cpu(), g[variant + "_rgb"]) < t["abs"]
    assert util.max_abs(a.cpu(), g[variant + "_rgb_alpha"]) < t["abs"]


@pytest.mark.parametrize("prec", PRECISIONS)
@pytest.mark.parametrize("M", [1, 63, 64, 65, 129, 1000, 4133])
def test_network_vs_oracle_ragged_sizes(s1, M, prec):
    cfg, sds = s1
    sd = sds["trained"]
    m = make_model(cfg, sd, prec)
    t = TOL[prec]
    g = torch.Generator().manual_seed(100 + M)
    pts = torch.rand(M, 3, generator=g) * 3.0 - 1.5
    views = torch.randn(M, 3, generator=g)
    with torch.no_grad():
        a_ref = O.network_forward(sd, cfg["model"], pts, only_occupancy=True)
    g_ref = O.geo_gradient(sd, pts, cfg["model"])
    rgb_ref, _ = O.network_forward(sd, cfg["model"], pts, views, return_addocc=True)
    assert util.max_abs(m(pts.cuda(), only_occupancy=True).cpu(), a_ref) < t["abs"]
    assert util.rel_l2(m.gradient(pts.cuda()).cpu(), g_ref) < t["rel"] * 3
    rgb, _ = m(pts.cuda(), views.cuda(), return_addocc=True)  # un-normalised view dirs are normalised inside
    assert util.max_abs(rgb.cpu(), rgb_ref.detach()) < t["abs"]


def test_empty_inputs(s1):
    cfg, sds = s1
    m = make_model(cfg, sds["init"], "fp32")
    z = torch.zeros(0, 3, device="cuda")
    assert m(z, only_occupancy=True).shape == (0, 1)
    assert m.gradient(z).shape == (0, 1, 3)
    rgb, a = m(z, z, return_addocc=True)
    assert rgb.shape == (0, 3) and a.shape == (0, 1)


@pytest.mark.parametrize("prec", PRECISIONS)
@pytest.mark.parametrize("variant", ["init", "trained"])
@pytest.mark.parametrize("case", list(util.STAGE1_CASES))
def test_unisurf_vs_golden(s1, variant, case, prec):
    from psnerf_b200.stage1 import Renderer
    _, sds = s1
    h, w, s_in, s_out, msteps, it = util.STAGE1_CASES[case]
    cfg = synth.stage1_cfg(num_points_in=s_in, num_points_out=s_out, ray_marching_steps=msteps)
    r = Renderer(make_model(cfg, sds[variant], prec), cfg, device=torch.device("cuda"))
    g = util.golden("stage1_render")
    t = TOL[prec]
    pose = torch.from_numpy(g["pose"])
    pix, K = synth.pixel_grid_xmajor(h, w), synth.intrinsics(h, w)
    key = "%s_%s_" % (variant, case)
    # ray directions and the surface search
    ray0 = pose[:, :3, 3].unsqueeze(1).repeat(1, h * w, 1).cuda()
    _, dirs = r._rays(pix, K, pose)
    assert util.max_abs(dirs.cpu(), g[key + "dirs"][0]) < 1e-6
    d = r.ray_marching(ray0, dirs.unsqueeze(0), n_steps=[msteps, msteps + 1], n_secant_steps=8,
                       depth_range=r.depth_range, rad=cfg["rendering"]["radius"]).cpu().numpy()
    d_ref = g[key + "d_i"]
    same = np.isfinite(d) == np.isfinite(d_ref)
    assert same.mean() >= 0.99
    both = np.isfinite(d) & np.isfinite(d_ref)
    close = np.abs(d[both] - d_ref[both]) < 1e-4
    assert close.mean() >= 0.98, "secant depths differ on %d rays" % int((~close).sum())
    # the full render
    out = r(pix.cuda(), K.cuda(), pose.cuda(), None, "unisurf", add_noise=False, eval_=True, it=it)
    mask_ref = g[key + "mask"]
    agree = out["mask_pred"].cpu().numpy() == mask_ref
    assert agree.mean() >= 0.99
    ok = torch.from_numpy(agree & (same[0]))
    tag = "unisurf_golden/%s/%s/%s/" % (prec, variant, case)
    util.bound(tag + "rgb", util.max_abs(out["rgb"][0].cpu()[ok], g[key + "rgb"][0][ok.numpy()]), GATE["rgb"])
    util.bound(tag + "acc", util.max_abs(out["acc_map"][0].cpu()[ok], g[key + "acc"][0][ok.numpy()]), GATE["acc"])
    util.bound(tag + "normal", util.max_abs(out["normal_pred"][0].cpu()[ok], g[key + "normal"][0][ok.numpy()]), GATE["normal"])
    assert O.psnr(out["rgb"].cpu(), torch.from_numpy(g[key + "rgb"])) > 50.0
    assert out["diff_norm"] is None


@pytest.mark.parametrize("prec", PRECISIONS)
def test_sample_plan_and_jitter_vs_oracle(s1, prec):
    """Interval sampling incl. the sorted outside samples and caller-supplied jitter (rendering.py:110-168)."""
    import ctypes as C
    from psnerf_b200 import _binding as B, engine
    from psnerf_b200.stage1 import Renderer
    _, sds = s1
    h = w = 18
    cfg = synth.stage1_cfg(num_points_in=16, num_points_out=8, ray_marching_steps=96)
    sd = sds["init"]
    r = Renderer(make_model(cfg, sd, prec), cfg, device=torch.device("cuda"))
    pose = synth.look_at_pose(-30.0, 20.0)
    pix, K = synth.pixel_grid_xmajor(h, w), synth.intrinsics(h, w)
    g, a = r._geo_app()
    origin, dirs = r._rays(pix, K, pose)
    delta = 0.446
    prm = B.UnisurfParams(2.0, 2.0, delta, 0.5, 96, 8, 16, 8, 1)
    gen = torch.Generator().manual_seed(7)
    noise = torch.rand(h * w, 24, generator=gen)
    for nz in (None, noise):
        out = engine.render_unisurf(g, a, origin, dirs, prm, noise=None if nz is None else nz.cuda(), want_sample_depth=True,
                                    precision=r.model._prec())
        mask = out["mask"].cpu()
        # oracle with the same discrete decisions is awkward to force; compare rays where the masks agree
        onz = None if nz is None else {"miss": None, "hit": None}
        ref = O.unisurf_render(sd, _cfg_with_delta(cfg, delta), pix, K, pose, it=100000, return_aux=True,
                               noise=None if nz is None else _split_noise(nz, sd, cfg, pix, K, pose))
        agree = mask == ref["mask_pred"]
        assert agree.float().mean() >= 0.99
        dd = (out["sample_depth"].cpu() - ref["aux"]["depth"]).abs()[agree]
        assert float(dd.max()) < 2e-4
        assert util.max_abs(out["rgb"].cpu()[agree], ref["rgb"][0][agree]) < 5 * TOL[prec]["abs"]


def _cfg_with_delta(cfg, delta):
    import copy, math
    c = copy.deepcopy(cfg)
    # choose interval_end so that max(start*exp(-decay*it), end) == delta at it=100000
    c["rendering"]["interval_start"] = delta / math.exp(-1.5)
    c["rendering"]["interval_end"] = 0.0
    return c


def _split_noise(nz, sd, cfg, pix, K, pose):
    ray0, rayd = O.pixels_to_rays(pix, K, pose)
    d_i = O.ray_marching(sd, cfg["model"], ray0, rayd, int(cfg["rendering"]["ray_marching_steps"]), 2.0, 2.0)
    obj = ((d_i.abs() != np.inf) & (d_i != 0))[0]
    return {"miss": nz[~obj], "hit": nz[obj]}


@pytest.mark.parametrize("prec", PRECISIONS)
@pytest.mark.parametrize("variant", ["init", "trained"])
def test_shape_extract_and_shadow_vs_golden(s1, variant, prec):
    from psnerf_b200.stage1 import Renderer
    _, sds = s1
    cfg = synth.stage1_cfg()
    r = Renderer(make_model(cfg, sds[variant], prec), cfg, device=torch.device("cuda"))
    g = util.golden("stage1_render")
    pose = torch.from_numpy(g["pose"])
    key = "%s_extract_" % variant
    h = w = 14
    lights = torch.from_numpy(g[key + "lights"])
    out = r(synth.pixel_grid_xmajor(h, w).cuda(), synth.intrinsics(h, w), pose, None, "shape_extract", visibility=True,
            light_dir=lights)
    agree = out["mask"].cpu().numpy() == g[key + "mask"]
    assert agree.mean() >= 0.99
    ok = torch.from_numpy(agree[0])
    t = TOL[prec]
    tag = "extract_golden/%s/%s/" % (prec, variant)
    util.bound(tag + "points", util.max_abs(out["points"][0].cpu()[ok], g[key + "points"][0][ok.numpy()]), GATE["points"])
    util.bound(tag + "normal", util.max_abs(out["normal"][0].cpu()[ok], g[key + "normal"][0][ok.numpy()]), GATE["normal"])
    util.bound(tag + "visibility", util.max_abs(out["visibility"].cpu()[:, ok], g[key + "visibility"][:, ok.numpy()]), GATE["visibility"])
    # direct light_visibility on the golden surface points (no discrete decisions involved)
    surf = torch.from_numpy(g[key + "points"][0][g[key + "mask"][0]])
    vis = r.light_visibility(surf=surf.cuda(), light_dir=lights.cuda()).cpu()
    with torch.no_grad():
        ref = O.light_visibility(sds[variant], cfg["model"], surf, lights)
    util.bound(tag + "light_visibility", util.max_abs(vis, ref), GATE["light_visibility"])


@pytest.mark.parametrize("prec", PRECISIONS)
@pytest.mark.parametrize("Ns,L", [(1, 1), (37, 7), (1000, 5)])
def test_box_culled_shadow_pass(s1, prec, Ns, L, monkeypatch):
    """The shadow pass evaluates the occupancy MLP only at the in-box steps of every shadow ray (k_shadow_plan / GEN_SHADOW_LIST), and
    past the first 16 of them only while the ray's transmittance is still >= 1e-6 (k_shadow_plan_b): equal to the evaluation of every
    step (what the reference executes, PSNERF_B200_SHADOW_UNCULLED=1) up to the order of the transmittance product and the < 1e-6 of
    the dropped tail, and to the oracle; points outside the +-1.1 cube (no in-box step at all) give visibility exactly 1."""
    from psnerf_b200 import engine
    from psnerf_b200.stage1 import Renderer
    _, sds = s1
    cfg = synth.stage1_cfg()
    r = Renderer(make_model(cfg, sds["trained"], prec), cfg, device=torch.device("cuda"))
    g, _ = r._geo_app()
    gen = torch.Generator().manual_seed(500 + Ns)
    surf = torch.nn.functional.normalize(torch.randn(Ns, 3, generator=gen), dim=-1) * (0.3 + 0.9 * torch.rand(Ns, 1, generator=gen))
    if Ns > 2:
        surf[0] = torch.tensor([5.0, 5.0, 5.0])     # no step of any unit direction (t <= 3.5) reaches the box
        surf[1] = torch.tensor([1.09, 1.09, 1.09])  # leaves the box after a few steps whatever the light
    lights = synth.lights(L, seed=21)
    vis, st = engine.shadow_visibility(g, surf.cuda(), lights.cuda(), precision=r.model._prec(), return_stats=True)
    assert st["culled"] and st["nominal"] == Ns * L * 128 and 0 <= st["evaluated"] < st["nominal"]
    monkeypatch.setenv("PSNERF_B200_SHADOW_UNCULLED", "1")
    vis_all, st_all = engine.shadow_visibility(g, surf.cuda(), lights.cuda(), precision=r.model._prec(), return_stats=True)
    monkeypatch.delenv("PSNERF_B200_SHADOW_UNCULLED")
    assert not st_all["culled"] and st_all["evaluated"] == st_all["nominal"]
    util.bound("shadow_culled_vs_every_step/%s/%d" % (prec, Ns), util.max_abs(vis.cpu(), vis_all.cpu()), 3e-6)
    with torch.no_grad():
        ref = O.light_visibility(sds["trained"], cfg["model"], surf, lights).view(L, Ns)
    util.bound("shadow_culled_vs_oracle/%s/%d" % (prec, Ns), util.max_abs(vis.cpu(), ref), GATE["light_visibility"])
    if Ns > 2:
        assert float((vis[:, 0] - 1).abs().max()) == 0.0
    # never more than the oracle's own box mask admits, never less than the first 16 in-box steps of every ray
    t = torch.linspace(0, 1, 128)
    p = surf[None, :, None, :] + lights[:, None, None, :] * (0.1 * (1 - t) + 3.5 * t)[None, None, :, None]
    inside = ((p <= 1.1) & (p >= -1.1)).all(-1)
    slack = max(2, int(0.001 * int(inside.sum())))  # a step exactly on the box face may round either way
    assert int(inside.sum(-1).clamp(max=16).sum()) - slack <= st["evaluated"] <= int(inside.sum()) + slack


def test_composite_properties():
    """Size-independent properties of the compositing integral at a BASELINE-sized ray count."""
    from psnerf_b200 import engine
    N, S = 100000, 128
    g = torch.Generator().manual_seed(3)
    alpha = torch.rand(N, S, generator=g).cuda() * 0.2
    rgb_s = torch.rand(N, S, 3, generator=g).cuda()
    rgb, acc = engine.composite(rgb_s, alpha, white_background=False)
    w = O.composite(alpha[:2000].cpu())
    assert util.max_abs(acc[:2000].cpu(), w.sum(-1)) < 1e-5
    assert util.max_abs(rgb[:2000].cpu(), (w.unsqueeze(-1) * rgb_s[:2000].cpu()).sum(-2)) < 1e-5
    assert float(acc.min()) >= 0 and float(acc.max()) <= 1 + 1e-3
    # linearity in the radiance, and alpha = 0 -> empty ray -> white background
    rgb2, _ = engine.composite(rgb_s * 0.5, alpha, white_background=False)
    assert util.max_abs(rgb2.cpu(), rgb.cpu() * 0.5) < 1e-6
    rgb0, acc0 = engine.composite(rgb_s, torch.zeros_like(alpha), white_background=True)
    assert float(acc0.abs().max()) == 0 and float((rgb0 - 1).abs().max()) == 0
    # an opaque first sample hides everything behind it
    a1 = alpha.clone()
    a1[:, 0] = 1.0
    rgb1, acc1 = engine.composite(rgb_s, a1, white_background=False)
    assert util.max_abs(rgb1.cpu(), rgb_s[:, 0].cpu()) < 3e-4


def test_phong_renderer_runs(s1):
    from psnerf_b200.stage1 import Renderer
    _, sds = s1
    cfg = synth.stage1_cfg()
    r = Renderer(make_model(cfg, sds["init"], "fp32"), cfg, device=torch.device("cuda"))
    out = r(synth.pixel_grid_xmajor(12, 12).cuda(), synth.intrinsics(12, 12), synth.look_at_pose(10.0, 5.0), None,
            "phong_renderer")
    ref = O.phong_render(sds["init"], cfg, synth.pixel_grid_xmajor(12, 12), synth.intrinsics(12, 12), synth.look_at_pose(10.0, 5.0))
    same = (out["rgb"].cpu() == 1).all(-1) == (ref["rgb"] == 1).all(-1)
    assert same.float().mean() > 0.98
    assert util.max_abs(out["rgb"].cpu()[same], ref["rgb"][same]) < 1e-3


@pytest.mark.parametrize("prec", PRECISIONS)
def test_full_size_render_properties(s1, prec):
    """BASELINE configs[1] at full size (512 x 512 rays, 128 samples, 256 march steps): properties that need no CPU reference.
    Rays are independent, so a render of a ray subset must reproduce the full render at those rays bit for bit (different tile
    composition, different CTA pairing), two renders of the same view are identical (fixed-order reductions), the opacity is a
    convex weight sum, hit rays carry unit normals and missed rays none."""
    from psnerf_b200.stage1 import Renderer
    _, sds = s1
    cfg = synth.stage1_cfg(num_points_in=96, num_points_out=32, ray_marching_steps=256)
    r = Renderer(make_model(cfg, sds["init"], prec), cfg, device=torch.device("cuda"))
    H = W = 512
    pix = synth.pixel_grid_xmajor(H, W).cuda()
    K, pose = synth.intrinsics(H, W), synth.look_at_pose(20.0, 10.0)
    full = r(pix, K, pose, None, "unisurf", add_noise=False, eval_=True, it=100000)
    again = r(pix, K, pose, None, "unisurf", add_noise=False, eval_=True, it=100000)
    for k in ("rgb", "normal_pred", "acc_map"):
        assert torch.equal(full[k], again[k]), k
    idx = torch.arange(37, H * W, 61, device="cuda")[:4099]  # ragged count, strided over the image
    part = r(pix[:, idx], K, pose, None, "unisurf", add_noise=False, eval_=True, it=100000)
    assert torch.equal(part["mask_pred"], full["mask_pred"][idx])
    for k in ("rgb", "normal_pred", "acc_map"):
        assert torch.equal(part[k][0], full[k][0][idx]), k
    acc, rgb, nrm, mask = full["acc_map"][0], full["rgb"][0], full["normal_pred"][0], full["mask_pred"]
    assert torch.isfinite(rgb).all() and float(acc.min()) >= 0 and float(acc.max()) <= 1 + 1e-3
    assert float(rgb.min()) >= -1e-5 and float(rgb.max()) <= 1 + 1e-3
    n_hit = int(mask.sum())
    assert 0.05 * H * W < n_hit < 0.5 * H * W            # the sphere-like init field covers part of the view
    ln = nrm.norm(dim=-1)
    assert float((ln[mask] - 1).abs().max()) < 1e-3 and float(ln[~mask].max()) == 0.0
    assert float(acc[mask].mean()) > 0.9  # hit rays are opaque (missed rays of the soft init field still gather some opacity)
