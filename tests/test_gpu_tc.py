"""Tensor-core (tcgen05, fp16 hi/lo split operands) path: layer-level and kernel-level parity with the CPU oracle."""
import ctypes as C

import pytest
import torch

import psnerf_oracle as O
import util
from psnerf_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def s1():
    return util.stage1_state_dicts()


def _model(cfg, sd):
    from psnerf_b200.stage1 import NeuralNetwork
    m = NeuralNetwork(cfg)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    m.precision = "tc"
    return m


@pytest.mark.parametrize("variant", ["init", "trained"])
def test_every_geo_layer_matches_the_oracle(s1, variant):
    from psnerf_b200 import _binding as B, engine
    cfg, sds = s1
    sd = sds[variant]
    g, _ = _model(cfg, sd)._packed()
    M = 517  # ragged: 4 full tiles + 5 rows
    pts = torch.rand(M, 3, generator=torch.Generator().manual_seed(5)) * 2.4 - 1.2
    with torch.no_grad():
        out, pre = O.geo_forward(sd, pts, cfg["model"], return_pre=True)
    pe = O.positional_encoding(pts, 6)
    pc = pts.cuda().contiguous()
    lib = B.load()
    for l in range(8):
        ncmp = 256
        act = torch.nn.functional.softplus(pre[l], beta=100)
        if l == 3:  # columns >= 217 of the skip layer's input are the encoding, written separately (not part of the dump)
            act = torch.cat([act, pe], -1) / (2 ** 0.5)
            ncmp = 217
        dump = torch.full((M, 256), float("nan"), device="cuda")
        logits = torch.empty(M, device="cuda")
        B.check(lib.psn_tc_debug_layer(g.handle, C.c_void_p(pc.data_ptr()), M, l, C.c_void_p(dump.data_ptr()),
                                       C.c_void_p(logits.data_ptr()), engine._stream()), "psn_tc_debug_layer")
        assert util.rel_l2(dump.cpu()[:, :ncmp], act[:, :ncmp]) < 2e-5, "layer %d" % l
        assert util.max_abs(logits.cpu(), out[:, 0]) < 1e-4


@pytest.mark.parametrize("M", [1, 127, 128, 129, 1000, 40000])
def test_occupancy_kernel_sizes(s1, M):
    cfg, sds = s1
    sd = sds["trained"]
    m = _model(cfg, sd)
    pts = torch.rand(M, 3, generator=torch.Generator().manual_seed(M)) * 3.0 - 1.5
    with torch.no_grad():
        ref = O.network_forward(sd, cfg["model"], pts, only_occupancy=True)
        refl = O.network_forward(sd, cfg["model"], pts, return_logits=True)
    assert util.max_abs(m(pts.cuda(), only_occupancy=True).cpu(), ref) < 5e-5
    assert util.max_abs(m(pts.cuda(), return_logits=True).cpu(), refl) < 2e-4


def test_fused_shadow_pass_matches_unfused_and_oracle(s1):
    from psnerf_b200 import _binding as B, engine
    cfg, sds = s1
    sd = sds["trained"]
    g, _ = _model(cfg, sd)._packed()
    gen = torch.Generator().manual_seed(9)
    d = torch.randn(70, 3, generator=gen)
    surf = 0.6 * d / d.norm(dim=-1, keepdim=True)
    lights = synth.lights(5, seed=4)
    with torch.no_grad():
        ref = O.light_visibility(sd, cfg["model"], surf, lights).view(5, 70)
    tc = engine.shadow_visibility(g, surf.cuda(), lights.cuda(), precision=B.PREC_TC).cpu()
    f32 = engine.shadow_visibility(g, surf.cuda(), lights.cuda(), precision=B.PREC_FP32).cpu()
    assert util.max_abs(f32, ref) < 1e-4
    assert util.max_abs(tc, ref) < 5e-4
    tc64 = engine.shadow_visibility(g, surf.cuda(), lights.cuda(), n_steps=64, precision=B.PREC_TC).cpu()  # unfused fallback
    with torch.no_grad():
        ref64 = O.light_visibility(sd, cfg["model"], surf, lights, n_steps=64).view(5, 70)
    assert util.max_abs(tc64, ref64) < 5e-4
