"""PSN_PREC_TC_TWOLEVEL ('tc_two_level', the engine default): tc_mixed plus the two-level surface march (api_stage1.cu
raymarch_impl).  Verified on hardware in round 2 (profiles/r2_bringup.md): march 152 -> 88 ms at 512 x 512 x 256.

What must hold: every output is BIT-IDENTICAL to 'tc_mixed' (and every march depth to 'tc'): the full program re-evaluates all
proposal points the scan can tell apart, the single-pass values elsewhere only contribute their sign."""
import pytest
import torch

import util
from psnerf_b200 import engine, synth

pytestmark = pytest.mark.gpu


def make_model(cfg, sd, prec):
    from psnerf_b200.stage1 import NeuralNetwork
    m = NeuralNetwork(cfg)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    m.precision = prec
    return m


@pytest.mark.parametrize("variant", ["init", "trained"])
@pytest.mark.parametrize("steps", [64, 256, 512])
def test_march_depths_equal_full_program(variant, steps):
    from psnerf_b200.stage1 import Renderer
    cfg0, sds = util.stage1_state_dicts()
    cfg = synth.stage1_cfg(ray_marching_steps=steps)
    h = w = 40
    pix, K, pose = synth.pixel_grid_xmajor(h, w), synth.intrinsics(h, w), synth.look_at_pose(-35.0, 25.0)
    d = {}
    for prec in ("tc", "tc_two_level"):
        r = Renderer(make_model(cfg, sds[variant], prec), cfg, device=torch.device("cuda"))
        g, _ = r._geo_app()
        origin, dirs = r._rays(pix, K, pose)
        d[prec] = engine.raymarch(g, origin, dirs, 2.0, 2.0, steps, 8, 0.5, r.model._prec())
    assert torch.equal(d["tc"], d["tc_two_level"])
    assert int(torch.isfinite(d["tc"]).sum()) > 50  # the view does hit the surface


@pytest.mark.parametrize("variant", ["init", "trained"])
@pytest.mark.parametrize("case", list(util.STAGE1_CASES))
def test_unisurf_equals_tc_mixed(variant, case):
    from psnerf_b200.stage1 import Renderer
    _, sds = util.stage1_state_dicts()
    h, w, s_in, s_out, msteps, it = util.STAGE1_CASES[case]
    cfg = synth.stage1_cfg(num_points_in=s_in, num_points_out=s_out, ray_marching_steps=msteps)
    g = util.golden("stage1_render")
    pose = torch.from_numpy(g["pose"])
    pix, K = synth.pixel_grid_xmajor(h, w), synth.intrinsics(h, w)
    outs = {}
    for prec in ("tc_mixed", "tc_two_level"):
        r = Renderer(make_model(cfg, sds[variant], prec), cfg, device=torch.device("cuda"))
        outs[prec] = r(pix.cuda(), K.cuda(), pose.cuda(), None, "unisurf", add_noise=False, eval_=True, it=it)
    for k in ("rgb", "mask_pred", "acc_map", "normal_pred"):
        assert torch.equal(outs["tc_mixed"][k], outs["tc_two_level"][k]), k


@pytest.mark.parametrize("variant", ["init", "trained"])
@pytest.mark.parametrize("prec", ["fp32", "tc", "tc_two_level"])
def test_fused_secant_equals_one_launch_per_iteration(variant, prec, monkeypatch):
    """Renderer.secant (rendering.py:525-555) runs as ONE launch of the fp32 occupancy kernel with the bracket kept on chip (under
    every precision: csrc/api_stage1.cu accuracy policy); PSNERF_B200_SECANT_UNFUSED=1 runs one evaluation launch + one update
    kernel per iteration: same bits."""
    from psnerf_b200.stage1 import Renderer
    _, sds = util.stage1_state_dicts()
    cfg = synth.stage1_cfg(ray_marching_steps=128)
    h, w = 37, 29  # 1073 rays: several tiles, a ragged last one
    pix, K, pose = synth.pixel_grid_xmajor(h, w), synth.intrinsics(h, w), synth.look_at_pose(40.0, -15.0)
    r = Renderer(make_model(cfg, sds[variant], prec), cfg, device=torch.device("cuda"))
    g, _ = r._geo_app()
    origin, dirs = r._rays(pix, K, pose)
    fused = engine.raymarch(g, origin, dirs, 2.0, 2.0, 128, 8, 0.5, r.model._prec())
    monkeypatch.setenv("PSNERF_B200_SECANT_UNFUSED", "1")
    unfused = engine.raymarch(g, origin, dirs, 2.0, 2.0, 128, 8, 0.5, r.model._prec())
    assert int(torch.isfinite(fused).sum()) > 50
    assert torch.equal(fused, unfused)
    for n_sec in (0, 1, 3):
        monkeypatch.delenv("PSNERF_B200_SECANT_UNFUSED", raising=False)
        a = engine.raymarch(g, origin, dirs, 2.0, 2.0, 128, n_sec, 0.5, r.model._prec())
        monkeypatch.setenv("PSNERF_B200_SECANT_UNFUSED", "1")
        assert torch.equal(a, engine.raymarch(g, origin, dirs, 2.0, 2.0, 128, n_sec, 0.5, r.model._prec())), n_sec
