"""Why the tensor-core programs split their operands the way they do, as an executable check (CPU, no GPU): the emulation in
tests/precision_study.py of the operand-split schemes on the oracle's networks.  The north star's gate is 1e-4 relative.

* the softplus stack (alpha) and the stage-2 visibility MLP need the three-pass split: every cheaper scheme misses the gate;
* the appearance side of the radiance program does not: PSN_PREC_TC_MIXED ('mixed') leaves alpha untouched and moves rgb by ~6e-6."""
import torch

import precision_study as P
import util


def _points(n=1024):
    g = torch.Generator().manual_seed(0)
    return torch.rand(n, 3, generator=g) * 2.4 - 1.2, torch.randn(n, 3, generator=g)


def test_softplus_stack_needs_the_three_pass_split():
    cfg, sds = util.stage1_state_dicts()
    p, v = _points()
    ref = P.field(sds["trained"], cfg["model"], p, v, "fp32")
    a3, g3, c3 = (P.err(o, r) for o, r in zip(P.field(sds["trained"], cfg["model"], p, v, "3pass"), ref))
    assert a3[0] < 5e-6 and g3[0] < 1e-5 and c3[0] < 1e-6
    for scheme in ("a_hi", "w_hi", "1pass"):
        a, g, c = (P.err(o, r) for o, r in zip(P.field(sds["trained"], cfg["model"], p, v, scheme), ref))
        assert a[0] > 1.5e-4 and g[0] > 1.5e-4, (scheme, a, g)  # alpha and the normal miss the 1e-4 gate


def test_mixed_program_keeps_alpha_and_holds_rgb():
    cfg, sds = util.stage1_state_dicts()
    p, v = _points()
    for variant in ("init", "trained"):
        ref = P.field(sds[variant], cfg["model"], p, v, "fp32")
        full = P.field(sds[variant], cfg["model"], p, v, "3pass")
        mixed = P.field(sds[variant], cfg["model"], p, v, "mixed")
        assert P.err(mixed[0], ref[0])[0] < 2e-6 and P.err(full[0], ref[0])[0] < 2e-6  # alpha: the same three-pass layers feed it
        rel, mx = P.err(mixed[2], ref[2])
        assert rel < 2e-5 and mx < 5e-5, (variant, rel, mx)  # rgb: 5x inside the gate per sample (100x per rendered pixel)
        assert P.err(mixed[1], ref[1])[0] > 1e-4  # ... which is why the normal OUTPUT never comes from the mixed program


def test_stage2_visibility_mlp_needs_the_split_too():
    conf, sds = util.stage2_state_dicts()
    out = P.visibility_errors(sds["trained"], conf, points=512, lights=8)
    assert out["3pass"][0] < 5e-6
    for scheme in ("a_hi", "w_hi", "1pass"):
        assert out[scheme][0] > 2e-4, (scheme, out[scheme])


def test_two_level_march_reproduces_the_full_scan():
    """PSN_PREC_TC_TWOLEVEL: single-pass values everywhere, the full program only where the scan can tell the difference (< 2 % of
    the points here): crossing index and bracket values of every ray equal those of the full evaluation."""
    assert P.march_refine_study(R=16, n_steps=128, margin=0.02)
