"""PSN_PREC_TC_MIXED ('tc_mixed'): the radiance program with the appearance side in single fp16 passes (tc_rad.cu header,
tests/precision_study.py).  Opt-in (the default tensor-core precision stays 'tc').

What must hold: alpha, depth, masks and surface normals are BIT-IDENTICAL to 'tc' (the same three-pass programs produce them);
rgb stays inside the 'tc' gate against the reference fixtures (rel-L2 5e-5 / max-abs 1e-4; the CPU emulation predicts 6e-6 / 2e-5
per sample and 1e-6 / 4e-6 per rendered pixel)."""
import pytest
import torch

import psnerf_oracle as O
import util
from psnerf_b200 import synth

pytestmark = pytest.mark.gpu
TOL = dict(rel=5e-5, abs=1e-4)


@pytest.fixture(scope="module")
def s1():
    return util.stage1_state_dicts()


def make_model(cfg, sd, prec):
    from psnerf_b200.stage1 import NeuralNetwork
    m = NeuralNetwork(cfg)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    m.precision = prec
    return m


@pytest.mark.parametrize("variant", ["init", "trained"])
def test_network_vs_golden_and_tc(s1, variant):
    cfg, sds = s1
    m, m_tc = make_model(cfg, sds[variant], "tc_mixed"), make_model(cfg, sds[variant], "tc")
    g = util.golden("stage1_net")
    pts, views = torch.from_numpy(g["pts"]).cuda(), torch.from_numpy(g["views"]).cuda()
    rgb, a = m(pts, views, return_addocc=True)
    rgb_tc, a_tc = m_tc(pts, views, return_addocc=True)
    assert torch.equal(a, a_tc)
    assert util.max_abs(rgb.cpu(), g[variant + "_rgb"]) < TOL["abs"]
    assert util.rel_l2(rgb.cpu(), g[variant + "_rgb"]) < TOL["rel"]
    assert util.max_abs(rgb.cpu(), rgb_tc.cpu()) < 5e-5
    assert not torch.equal(rgb, rgb_tc)  # the single-pass program really ran
    assert torch.equal(m.gradient(pts), m_tc.gradient(pts))  # the normal output never uses the mixed program


@pytest.mark.parametrize("M", [1, 127, 128, 129, 300, 4133, 148 * 128 * 2 + 5])
def test_ragged_sizes_vs_oracle(s1, M):
    """Masked rows, a lone tile (the odd CTA of the pair runs a dummy), several tiles per CTA."""
    cfg, sds = s1
    sd = sds["trained"]
    m = make_model(cfg, sd, "tc_mixed")
    g = torch.Generator().manual_seed(200 + M)
    pts = torch.rand(M, 3, generator=g) * 3.0 - 1.5
    views = torch.randn(M, 3, generator=g)
    n = min(M, 2048)  # the oracle is evaluated on a prefix; the tail is compared with 'tc'
    rgb_ref, a_ref = O.network_forward(sd, cfg["model"], pts[:n], views[:n], return_addocc=True)
    rgb, a = m(pts.cuda(), views.cuda(), return_addocc=True)
    assert util.max_abs(rgb[:n].cpu(), rgb_ref.detach()) < TOL["abs"]
    assert util.max_abs(a[:n].cpu(), a_ref.detach()) < TOL["abs"]
    rgb_tc, a_tc = make_model(cfg, sd, "tc")(pts.cuda(), views.cuda(), return_addocc=True)
    assert torch.equal(a, a_tc)
    assert util.max_abs(rgb.cpu(), rgb_tc.cpu()) < 5e-5


@pytest.mark.parametrize("variant", ["init", "trained"])
@pytest.mark.parametrize("case", list(util.STAGE1_CASES))
def test_unisurf_vs_golden_and_tc(s1, variant, case):
    from psnerf_b200.stage1 import Renderer
    _, sds = s1
    h, w, s_in, s_out, msteps, it = util.STAGE1_CASES[case]
    cfg = synth.stage1_cfg(num_points_in=s_in, num_points_out=s_out, ray_marching_steps=msteps)
    g = util.golden("stage1_render")
    pose = torch.from_numpy(g["pose"])
    pix, K = synth.pixel_grid_xmajor(h, w), synth.intrinsics(h, w)
    outs = {}
    for prec in ("tc", "tc_mixed"):
        r = Renderer(make_model(cfg, sds[variant], prec), cfg, device=torch.device("cuda"))
        outs[prec] = r(pix.cuda(), K.cuda(), pose.cuda(), None, "unisurf", add_noise=False, eval_=True, it=it)
    a, b = outs["tc_mixed"], outs["tc"]
    assert torch.equal(a["mask_pred"], b["mask_pred"])
    assert torch.equal(a["acc_map"], b["acc_map"])
    assert torch.equal(a["normal_pred"], b["normal_pred"])
    assert util.max_abs(a["rgb"].cpu(), b["rgb"].cpu()) < 2e-5
    key = "%s_%s_" % (variant, case)
    same = a["mask_pred"].cpu().numpy().reshape(-1) == g[key + "mask"].reshape(-1)
    assert same.mean() >= 0.99
    d = (a["rgb"][0].cpu().numpy() - g[key + "rgb"][0])[same]
    util.bound("tc_mixed_unisurf_golden/%s/%s/rgb" % (variant, case), float(abs(d).max()), 3e-5)  # measured 8.8e-6
    assert O.psnr(a["rgb"].cpu(), torch.from_numpy(g[key + "rgb"])) > 50.0
