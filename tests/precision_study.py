"""CPU study of the operand-split schemes the tensor-core kernels could run (no GPU needed).

The shipped kernels split BOTH operands into fp16 hi + lo and issue 3 MMA passes per product (A_hi W_hi + A_lo W_hi + A_hi W_lo,
fp32 accumulate).  The epilogue, not the MMA pipe, bounds them, and half of the epilogue's conversion / TMEM
store work exists only to produce A_lo.  This script emulates the candidates on the oracle's networks in fp32 torch and reports the
error of each against the fp32 oracle, on the quantities the parity tests gate (alpha, d logit / d p, rgb; TOL["tc"] in
tests/test_gpu_stage1.py: rel-L2 5e-5, max-abs 1e-4):

    3pass   A_hi W_hi + A_lo W_hi + A_hi W_lo      (shipped)
    a_hi    A_hi (W_hi + W_lo)                      2 passes, no A_lo: half the epilogue conversion / tcgen05.st work
    w_hi    (A_hi + A_lo) W_hi                      2 passes, epilogue unchanged, half the weight traffic
    1pass   A_hi W_hi                               plain fp16 (= the precision of one tf32 pass)
    X:Y     geo net under X, appearance net under Y;  X/Z:Y  geo forward X, reverse sweep Z, appearance Y
    mixed   geo layers 0-7 3pass (alpha and the sigma' stash need it), fp32 logit head, then feature head, reverse sweep and
            appearance net all 1pass - what a radiance-mode k_tc_rad could run when only rgb / alpha leave the kernel

    python tests/precision_study.py [--points 4096] [--render 24]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import psnerf_oracle as O  # noqa: E402
import util  # noqa: E402
from psnerf_b200 import synth  # noqa: E402

SCHEMES = ["3pass", "a_hi", "w_hi", "1pass", "3pass:a_hi", "3pass:w_hi", "3pass:1pass", "3pass/a_hi:1pass", "3pass/w_hi:1pass", "3pass/1pass:1pass", "mixed"]  # "geo:app" = different schemes per net


def hi(x):
    return x.half().float()


def lo(x):
    return (x - hi(x)).half().float()


def mm(x, W, scheme):
    """x [M,K] @ W[out,K]^T with the operands split as the scheme says; products accumulated in fp64 then rounded (an ideal fp32
    accumulator: isolates the operand error)."""
    xd, Wd = x.double(), W.double()
    if scheme == "fp32":
        return (xd @ Wd.t()).float()
    xh, xl, Wh, Wl = hi(x).double(), lo(x).double(), hi(W).double(), lo(W).double()
    acc = xh @ Wh.t()
    if scheme in ("3pass", "w_hi"):
        acc = acc + xl @ Wh.t()
    if scheme in ("3pass", "a_hi"):
        acc = acc + xh @ Wl.t()
    return acc.float()


def field(sd, mcfg, p, views, scheme):
    """(alpha [M], grad [M,3], rgb [M,3]) the way k_tc_rad computes them: forward stack, analytic reverse sweep, appearance MLP."""
    mixed = scheme == "mixed"  # the proposed radiance-mode program: s0-s7 3pass, fp32 logit head, EVERYTHING after it 1pass
    if mixed:
        scheme = "3pass/1pass:1pass"
    scheme, app_scheme = scheme.split(":") if ":" in scheme else (scheme, scheme)
    scheme, bwd_scheme = scheme.split("/") if "/" in scheme else (scheme, scheme)  # "fwd/bwd:app"
    layers = O.stage1_weights(sd, "lin", O.count_layers(sd, "lin"))
    nl = len(layers)
    L = mcfg["octaves_pe"]
    q = p / mcfg["rescale"]
    pe = O.positional_encoding(q, L)
    npe = pe.shape[-1]
    inv = float(1.0 / np.sqrt(2))
    x, pre = pe, []
    for l, (W, b) in enumerate(layers):
        if l in mcfg["skips"]:
            x = torch.cat([x, pe], -1) * inv
        if mixed and l == nl - 1:
            x = torch.cat([mm(x, W[:1], "fp32"), mm(x, W[1:], "1pass")], -1) + b
        else:
            x = mm(x, W, scheme) + b
        pre.append(x)
        if l < nl - 1:
            x = O.softplus100(x)
    out = x
    gx = layers[nl - 1][0][0:1, :].expand(p.shape[0], -1)
    gpe = torch.zeros_like(pe)
    for l in range(nl - 2, -1, -1):
        gz = gx * torch.sigmoid(100.0 * pre[l])
        gx = mm(gz, layers[l][0].t().contiguous(), bwd_scheme)
        if l in mcfg["skips"]:
            gx = gx * inv
            gpe = gpe + gx[:, -npe:]
            gx = gx[:, :-npe]
    gpe = gpe + gx
    g = gpe[:, 0:3].clone()
    for i in range(L):
        f = float(2 ** i)
        g = g + f * torch.cos(f * q) * gpe[:, 3 + 6 * i:6 + 6 * i] - f * torch.sin(f * q) * gpe[:, 6 + 6 * i:9 + 6 * i]
    g = g / mcfg["rescale"]
    alpha = torch.sigmoid(out[:, 0] * -10.0)
    v = views / views.norm(dim=-1, keepdim=True)
    x = torch.cat([p, O.positional_encoding(v, mcfg["octaves_pe_views"]), g, out[:, 1:]], -1)
    app = O.stage1_weights(sd, "lina", O.count_layers(sd, "lina"))
    for l, (W, b) in enumerate(app):
        x = mm(x, W, app_scheme) + b
        if l < len(app) - 1:
            x = torch.relu(x)
    return alpha, g, torch.tanh(x) * 0.5 + 0.5


def err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30)), float((a - b).abs().max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=4096)
    ap.add_argument("--render", type=int, default=0, help="also render an RxR unisurf view per scheme (slow on CPU)")
    a = ap.parse_args()
    torch.set_num_threads(8)
    cfg, sds = util.stage1_state_dicts()
    mcfg = cfg["model"]
    g = torch.Generator().manual_seed(0)
    p = torch.rand(a.points, 3, generator=g) * 2.4 - 1.2
    near = torch.nn.functional.normalize(torch.randn(a.points, 3, generator=g), dim=-1) * (0.5 + 0.02 * torch.randn(a.points, 1, generator=g))
    views = torch.randn(a.points, 3, generator=g)
    print("%-8s %-8s %-17s | %-22s | %-22s | %-22s" % ("weights", "points", "scheme", "alpha relL2 / maxabs", "grad relL2 / maxabs", "rgb relL2 / maxabs"))
    for wname, sd in sds.items():
        for pname, pts in (("volume", p), ("surface", near)):
            ref = field(sd, mcfg, pts, views, "fp32")
            for s in SCHEMES:
                out = field(sd, mcfg, pts, views, s)
                e = [err(o, r) for o, r in zip(out, ref)]
                print("%-8s %-8s %-17s | %.2e / %.2e    | %.2e / %.2e    | %.2e / %.2e" % (wname, pname, s, *e[0], *e[1], *e[2]))
    if a.render:
        R = a.render
        cfg_r = synth.stage1_cfg(num_points_in=32, num_points_out=8, ray_marching_steps=128)
        pix = synth.pixel_grid_xmajor(R, R)
        K, pose = synth.intrinsics(R, R), synth.look_at_pose(20.0, 10.0)
        sd = sds["trained"]
        ref = O.unisurf_render(sd, cfg_r, pix, K, pose, it=100000, return_aux=True)
        real_geo, real_grad, real_app = O.geo_forward, O.geo_gradient, O.network_forward
        print("\nunisurf %dx%d, trained weights: mask flips, rgb relL2 / maxabs, normal maxabs (rays hit in both)" % (R, R))
        for s in ["3pass", "1pass", "mixed", "a_hi|mixed", "w_hi|mixed"]:  # "X|Y": occupancy queries (march, secant) under X, radiance under Y
            def net(sd_, mcfg_, pp, ray_d=None, only_occupancy=False, return_logits=False, return_addocc=False, _s=s):
                _occ_s, _rad_s = _s.split("|") if "|" in _s else (_s, _s)
                al, gg, rgb = field(sd_, mcfg_, pp.reshape(-1, 3), ray_d if ray_d is not None else torch.ones_like(pp).reshape(-1, 3),
                                    _occ_s if only_occupancy else _rad_s)
                if only_occupancy:
                    return al.reshape(*pp.shape[:-1], 1)
                if ray_d is not None:
                    return (rgb, al.reshape(-1, 1)) if return_addocc else rgb
                raise NotImplementedError
            O.network_forward = net
            O.geo_gradient = lambda sd_, pp, mcfg_, _s=("3pass" if "mixed" in s else s): field(sd_, mcfg_, pp, torch.ones_like(pp), _s)[1].unsqueeze(1)  # normals: gradient-mode launch
            try:
                out = O.unisurf_render(sd, cfg_r, pix, K, pose, it=100000, return_aux=True)
            finally:
                O.geo_forward, O.geo_gradient, O.network_forward = real_geo, real_grad, real_app
            flips = int((out["mask_pred"] != ref["mask_pred"]).sum())
            both = (out["mask_pred"] & ref["mask_pred"]).reshape(-1)
            e = err(out["rgb"], ref["rgb"])
            dd = (out["aux"]["d_i"] - ref["aux"]["d_i"]).abs() if "aux" in out else None
            en = float((out["normal_pred"][0][both] - ref["normal_pred"][0][both]).abs().max()) if both.any() else 0.0
            fin = torch.isfinite(dd)
            print("%-17s flips %d / %d   rgb %.2e / %.2e   normal %.2e   depth maxabs %.2e" % (s, flips, R * R, *e, en, float(dd[fin].max()) if fin.any() else 0.0))


def visibility_errors(sd, conf, points=1024, lights=16):
    """{scheme: (relL2, maxabs)} of visibility_net (126 -> 256 x 8 -> 1, ReLU, cat[y, x] after the skip layer; renderer.py:17-49,193)
    against its fp32 evaluation."""
    nf = int(conf["brdf.net.n_freqs_xyz"])
    skip = int(conf["visibility.net.mlp_skip_at"])
    g = torch.Generator().manual_seed(3)
    surf = torch.nn.functional.normalize(torch.randn(points, 3, generator=g), dim=-1) * (0.5 + 0.5 * torch.rand(points, 1, generator=g))
    lt = synth.lights(lights)
    x = torch.cat([O.embed(surf, nf).tile(lights, 1), O.embed(lt[:, None].expand(-1, points, -1).reshape(-1, 3), nf)], -1)
    n = 0
    while ("visibility_net.linears.%d.bias" % n) in sd:
        n += 1
    outs = {}
    for s in ["fp32", "3pass", "a_hi", "w_hi", "1pass"]:
        y = x
        for li in range(n):
            y = mm(y, sd["visibility_net.linears.%d.weight" % li], s) + sd["visibility_net.linears.%d.bias" % li]
            if li != n - 1:
                y = torch.relu(y)
            if li == skip:
                y = torch.cat([y, x], -1)
        outs[s] = y
    res = {s: err(outs[s], outs["fp32"]) for s in ["3pass", "a_hi", "w_hi", "1pass"]}
    res["max"] = float(outs["fp32"].abs().max())
    return res


def study_stage2(points=1024, lights=16):
    """Gates of tests/test_gpu_stage2.py (tc): visibility max-abs 5e-4, rgb max-abs 1e-4 (rgb = brdf * intensity * cos * vis)."""
    conf, sds = util.stage2_state_dicts()
    print("\nstage-2 visibility_net, %d points x %d lights: relL2 / maxabs of the visibility" % (points, lights))
    for wname, sd in sds.items():
        r = visibility_errors(sd, conf, points, lights)
        for s in ["3pass", "a_hi", "w_hi", "1pass"]:
            print("%-8s %-17s %.2e / %.2e   (|vis| max %.2f)" % (wname, s, *r[s], r["max"]))


if __name__ == "__main__":
    main()
    study_stage2()


POLY3 = (1.42459315, -0.589203729, 0.165381165)  # lg2(1 + u) ~ u (C1 + C2 u + C3 u^2) on (0, 1]: tc_mlp.cuh softplus_scaled_cheap


def softplus_poly3(z):
    """softplus(beta = 100) the way the CHEAP occupancy program evaluates it: one ex2 and a cubic for lg2(1 + 2^-|t|)."""
    t = (z * (100.0 / np.log(2.0))).float()
    u = torch.exp2(-t.abs())
    q = (POLY3[2] * u + POLY3[1]) * u + POLY3[0]
    return (float(np.log(2.0) / 100.0) * (q * u + torch.clamp(t, min=0.0))).float()


def softplus_poly3_h2(z):
    """The same activation in packed fp16 arithmetic (tc_mlp.cuh softplus_scaled_cheap_h2): the accumulator is rounded to fp16 and every
    operation after it rounds to fp16 again (the fused HFMA2s round once; rounding twice here errs on the pessimistic side)."""
    def h(t):
        return torch.as_tensor(t, dtype=torch.float32).half().float()

    zs = h(z * (100.0 / np.log(2.0)))
    u = h(torch.exp2(-zs.abs()))
    q = h(h(h(POLY3[2]) * u) + h(POLY3[1]))
    q = h(h(q * u) + h(POLY3[0]))
    r = h(h(q * u) + torch.clamp(zs, min=0.0))
    return h(r * h(float(np.log(2.0) / 100.0)))


def alpha_only(sd, mcfg, p, scheme):
    """Occupancy probability of the geo net under `scheme` (what k_tc_occ computes: forward stack + fp32 logit head);
    "1pass_poly3" = the CHEAP program (single fp16 pass + softplus_poly3)."""
    act = act_last = O.softplus100
    if scheme == "1pass_poly3":
        scheme, act, act_last = "1pass", softplus_poly3, softplus_poly3
    elif scheme == "1pass_h2":  # the CHEAP program as built: layers 0-6 in packed fp16, layer 7 (feeds the fp32 logit head) in fp32
        scheme, act, act_last = "1pass", softplus_poly3_h2, softplus_poly3
    layers = O.stage1_weights(sd, "lin", O.count_layers(sd, "lin"))
    nl = len(layers)
    pe = O.positional_encoding(p / mcfg["rescale"], mcfg["octaves_pe"])
    x = pe
    inv = float(1.0 / np.sqrt(2))
    for l, (W, b) in enumerate(layers[:nl - 1]):
        if l in mcfg["skips"]:
            x = torch.cat([x, pe], -1) * inv
        x = (act if l < nl - 2 else act_last)(mm(x, W, scheme) + b)
    W, b = layers[nl - 1]
    logit = mm(x, W[:1], "fp32") + b[:1]
    return torch.sigmoid(logit[:, 0] * -10.0)


def march_refine_study(R=32, n_steps=256, margin=0.02, cheap="1pass_h2"):
    """Two-level surface march: every proposal point with the CHEAP program, then the full three-pass program only where the scan
    (rendering.py:443-470) can see the difference - points within `margin` of the threshold, points next to a sign change of the
    cheap values, and their neighbours.  The scan reads nothing else (signs everywhere, values only at the crossing), so the
    merged values must give the SAME crossing index and bracket values as the full evaluation, on every ray."""
    cfg, sds = util.stage1_state_dicts()
    mcfg = cfg["model"]
    pix = synth.pixel_grid_xmajor(R, R)
    K, pose = synth.intrinsics(R, R), synth.look_at_pose(20.0, 10.0)
    print("\ntwo-level march, %dx%d rays x %d steps, cheap = %s, margin = %g" % (R, R, n_steps, cheap, margin))
    all_same = True
    for wname, sd in sds.items():
        ray0, rayd = O.pixels_to_rays(pix, K, pose)
        N = ray0.shape[1]
        far = O.sphere_intersection(ray0[:, 0], rayd, r=2.0)[0][..., 1]
        t = torch.linspace(0, 1, steps=n_steps).view(1, 1, n_steps, 1)
        d_prop = 2.0 * (1.0 - t) + far.view(1, -1, 1, 1) * t
        p_prop = (ray0.unsqueeze(2) + rayd.unsqueeze(2) * d_prop).reshape(-1, 3)
        full = torch.cat([alpha_only(sd, mcfg, c, "3pass") for c in torch.split(p_prop, 65536)]).view(N, n_steps) - 0.5
        chp = torch.cat([alpha_only(sd, mcfg, c, cheap) for c in torch.split(p_prop, 65536)]).view(N, n_steps) - 0.5
        unsure = chp.abs() < margin
        flip = torch.zeros_like(unsure)
        flip[:, 1:] |= torch.sign(chp[:, 1:]) != torch.sign(chp[:, :-1])
        flip[:, :-1] |= torch.sign(chp[:, :-1]) != torch.sign(chp[:, 1:])
        sel = unsure | flip
        grow = sel.clone()
        grow[:, 1:] |= sel[:, :-1]
        grow[:, :-1] |= sel[:, 1:]
        merged = torch.where(grow, full, chp)

        def scan(val):
            sign = torch.cat([torch.sign(val[:, :-1] * val[:, 1:]), torch.ones(N, 1)], dim=-1)
            cost = sign * torch.arange(n_steps, 0, -1).float()
            values, idx = torch.min(cost, -1)
            ar = torch.arange(N)
            mask = (values < 0) & (val[ar, idx] < 0) & (val[:, 0] < 0)
            idx2 = torch.clamp(idx + 1, max=n_steps - 1)
            return mask, idx, val[ar, idx], val[ar, idx2], val[:, 0] < 0

        m_f, i_f, lo_f, hi_f, ff_f = scan(full)
        m_m, i_m, lo_m, hi_m, ff_m = scan(merged)
        m_c, i_c, _, _, _ = scan(chp)
        same = bool(torch.equal(m_f, m_m) and torch.equal(ff_f, ff_m) and torch.equal(i_f[m_f], i_m[m_f])
                    and torch.equal(lo_f[m_f], lo_m[m_f]) and torch.equal(hi_f[m_f], hi_m[m_f]))
        print("%-8s refined %.3f %% of the points (%d of %d); hit rays %d; merged scan identical to the full one: %s; "
              "cheap alone: %d masks / %d crossing indices differ; max |cheap - full| = %.2e"
              % (wname, 100.0 * float(grow.float().mean()), int(grow.sum()), grow.numel(), int(m_f.sum()), same,
                 int((m_c != m_f).sum()), int((i_c[m_f & m_c] != i_f[m_f & m_c]).sum()), float((chp - full).abs().max())))
        all_same = all_same and same and int(m_f.sum()) > 0
    return all_same
