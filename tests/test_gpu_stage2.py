"""GPU parity of the stage-2 shading path against the real-reference golden fixtures and the CPU oracle."""
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
        return ["fp32", "tc"] if engine.tc_available() else ["fp32"]
    except Exception:
        return ["fp32"]


PRECISIONS = _precisions()
TOL = {"fp32": 2e-5, "tc": 1e-4}
SHADE_GATE = 2e-5  # test_shading_vs_golden: measured on the B200 (profiles/r2_parity_errlog_final.jsonl) <= 6.1e-6 for every output, both precisions
KEYS = ("sg_rgb_values", "sg_specular_rgb_values", "visibility", "normal_pred", "sg_diffuse_albedo_values", "sg_weight",
        "vis_train")


def make_model(conf, sd, prec):
    from psnerf_b200.stage2 import PSNetwork
    m = PSNetwork(conf)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    m.precision = prec
    return m


def to_cuda(inp):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in inp.items()}


@pytest.mark.parametrize("prec", PRECISIONS)
@pytest.mark.parametrize("variant", ["init", "trained"])
@pytest.mark.parametrize("case", list(util.STAGE2_CASES))
def test_shading_vs_golden(variant, case, prec):
    conf, sds = util.stage2_state_dicts()
    m = make_model(conf, sds[variant], prec)
    g = util.golden("stage2_shade")
    out = m(to_cuda(util.stage2_case_input(case)))
    key = "%s_%s_" % (variant, case)
    for k in KEYS:
        if key + k in g.files:
            assert tuple(out[k].shape) == g[key + k].shape, k
            util.bound("shading_golden/%s/%s/%s/%s" % (prec, variant, case, k), util.max_abs(out[k].cpu(), g[key + k]),
                       SHADE_GATE)
    assert set(["points", "object_mask", "network_object_mask", "normal_values", "albedo_jitter", "rough_jitter"]) <= set(out)


@pytest.mark.parametrize("prec", PRECISIONS)
@pytest.mark.parametrize("case", ["multi", "single"])
def test_microfacet_vs_golden(case, prec):
    """render_model = microfacet (psn_shade_params.render_model = 1) vs the real reference's outputs."""
    conf, sd = util.stage2_micro_state_dict()
    m = make_model(conf, sd, prec)
    g = util.golden("stage2_edit")
    out = m(to_cuda(util.micro_case_input(case)))
    assert "sg_weight" not in out
    for k in util.MICRO_KEYS:
        assert tuple(out[k].shape) == g["micro_%s_%s" % (case, k)].shape, k
        # the GGX lobe divides by cos^2 terms: a few 1e-5 of headroom over the SG path on the rgb image
        assert util.max_abs(out[k].cpu(), g["micro_%s_%s" % (case, k)]) < TOL[prec] * (5 if k in ("visibility", "sg_rgb_values") else 1), k


@pytest.mark.parametrize("prec", PRECISIONS)
@pytest.mark.parametrize("case", list(util.EDIT_CASES))
def test_material_editing_vs_golden(case, prec):
    """albedo_new / basis_new (psn_shade_stage2_edit) with [L,3] intensities vs the real reference's outputs."""
    conf, sds = util.stage2_state_dicts()
    m = make_model(conf, sds["trained"], prec)
    g = util.golden("stage2_edit")
    albedo_new, basis_new = util.EDIT_CASES[case]
    out = m(to_cuda(util.edit_case_input()), albedo_new=None if albedo_new is None else np.asarray(albedo_new, np.float32),
            basis_new=basis_new)
    for k in util.EDIT_KEYS:
        assert tuple(out[k].shape) == g["%s_%s" % (case, k)].shape, k
        assert util.max_abs(out[k].cpu(), g["%s_%s" % (case, k)]) < TOL[prec] * (5 if k == "visibility" else 1), k


@pytest.mark.parametrize("prec", PRECISIONS)
@pytest.mark.parametrize("hw,L,frac", [((40, 50), 7, 0.3), ((33, 31), 96, 0.9), ((10, 10), 2, 0.0)])
def test_shading_vs_oracle(hw, L, frac, prec):
    conf, sds = util.stage2_state_dicts()
    sd = sds["trained"]
    m = make_model(conf, sd, prec)
    inp = synth.stage2_input(hw[0], hw[1], L, all_surface=False, seed=17, mask_frac=frac)
    if frac == 0.0:
        inp["surface_mask"][:] = False
    with torch.no_grad():
        ref = O.psnetwork_forward(sd, conf, inp)
        ref64 = O.psnetwork_forward({k: v.double() for k, v in sd.items()}, conf,
                                    {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in inp.items()})
    out = m(to_cuda(inp))
    for k in KEYS[:-1]:
        assert tuple(out[k].shape) == tuple(ref[k].shape), k
        tol = TOL[prec] * (5 if k == "visibility" else 1)
        if k in ("sg_specular_rgb_values", "sg_rgb_values"):
            # The SG lobes (lambda up to e^10, sgbasis.py:12) amplify fp32 rounding of h.n by 2e4: the reference's own fp32
            # result is only this close to the exact value, so hold the kernel to the same accuracy class vs an fp64 oracle.
            tol = max(tol, 4.0 * util.max_abs(ref[k], ref64[k]))
            assert util.max_abs(out[k].cpu(), ref64[k]) < tol, k
        else:
            assert util.max_abs(out[k].cpu(), ref[k]) < tol, k
    assert O.psnr(out["sg_rgb_values"].cpu(), ref["sg_rgb_values"]) > 70.0


def test_jitter_branch_with_supplied_noise():
    conf, sds = util.stage2_state_dicts()
    sd = sds["trained"]
    m = make_model(conf, sd, "fp32")
    inp = synth.stage2_input(16, 16, 3, all_surface=False, seed=5)
    ns = int(inp["surface_mask"].sum())
    z = torch.randn(ns, 3, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = O.psnetwork_forward(sd, conf, inp, noise={"xyz": z})
    out = m(to_cuda(inp), noise={"xyz": z})
    for k in ("albedo_jitter", "rough_jitter", "albedo_values", "rough_values"):
        assert util.max_abs(out[k].cpu(), ref[k]) < 2e-5, k


def test_properties_at_baseline_light_count():
    """rgb in [0,1]; doubling a scalar intensity doubles unclamped pixels; non-surface pixels stay 1.0."""
    conf, sds = util.stage2_state_dicts()
    m = make_model(conf, sds["trained"], "fp32")
    inp = to_cuda(synth.stage2_input(64, 64, 96, all_surface=False, seed=2, mask_frac=0.7))
    inp["light_intensity"] = torch.tensor([0.25], device="cuda")
    a = m(inp)
    inp["light_intensity"] = torch.tensor([0.5], device="cuda")
    b = m(inp)
    ra, rb = a["sg_rgb_values"], b["sg_rgb_values"]
    assert float(ra.min()) >= 0 and float(ra.max()) <= 1
    sm = inp["surface_mask"][0]
    assert float((ra[:, ~sm] - 1).abs().max()) == 0 and float((a["visibility"][:, ~sm] - 1).abs().max()) == 0
    un = (rb < 1) & (rb > 0)
    assert float((rb[un] - 2 * ra[un]).abs().max()) < 1e-5
    assert torch.equal(a["visibility"], b["visibility"])


@pytest.mark.parametrize("prec", PRECISIONS)
def test_editing_and_microfacet_edge_cases(prec):
    """Material editing with an empty surface mask (every output keeps its pre-fill) and the microfacet jitter branch."""
    conf, sds = util.stage2_state_dicts()
    m = make_model(conf, sds["trained"], prec)
    inp = synth.stage2_input(9, 7, 3, all_surface=False, seed=2, mask_frac=0.0)
    out = m(to_cuda(inp), albedo_new=np.asarray([0.1, 0.2, 0.3], np.float32), basis_new=2)
    assert float((out["sg_rgb_values"] - 1).abs().max()) == 0.0 and float(out["sg_weight"].abs().max()) == 0.0
    confm, sdm = util.stage2_micro_state_dict()
    mm = make_model(confm, sdm, prec)
    inp = synth.stage2_input(8, 8, 2, all_surface=False, seed=3, mask_frac=0.5)
    ns = int(inp["surface_mask"].sum())
    z = torch.randn(ns, 3, generator=torch.Generator().manual_seed(1))
    out = mm(to_cuda(inp), noise={"xyz": z})
    ref = O.psnetwork_forward(sdm, confm, inp, noise={"xyz": z})
    for k in ("albedo_jitter", "rough_jitter", "rough_values", "sg_rgb_values"):
        assert tuple(out[k].shape) == tuple(ref[k].shape), k
        assert util.max_abs(out[k].cpu(), ref[k]) < TOL[prec] * 5, k
