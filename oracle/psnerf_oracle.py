"""CPU oracle for the PS-NeRF render/shading hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (not a copy) of the reference algorithm, written as stateless
functions over plain state-dicts so that it can travel to the GPU box (where /root/reference does
not exist).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import it; the product package ``psnerf_b200`` never does.

Pinning: the reference has no tests/golden vectors for this path, so the oracle
is pinned against outputs of the REAL reference run in the authoring container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``).

All citations are relative to /root/reference.  dtype/device follow the inputs (fp32 for parity,
fp64 to separate implementation error from fp32 noise).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

EPS_T = 1e-6  # stage1/model/rendering.py:8


# ----------------------------------------------------------------------------------------------
# Stage 1: field networks (stage1/model/network.py)
# ----------------------------------------------------------------------------------------------
def fold_weight_norm(g, v):
    """W = g * v / ||v||_row  (torch.nn.utils.weight_norm, dim=0; network.py:64,77)."""
    return torch._weight_norm(v, g, 0)


def stage1_weights(sd, prefix, n):
    """Effective (W, b) per layer from a reference state-dict (keys lin{l}.weight_g/_v/bias)."""
    out = []
    for l in range(n):
        g = sd["%s%d.weight_g" % (prefix, l)]
        v = sd["%s%d.weight_v" % (prefix, l)]
        out.append((fold_weight_norm(g, v), sd["%s%d.bias" % (prefix, l)]))
    return out


def count_layers(sd, prefix):
    n = 0
    while ("%s%d.bias" % (prefix, n)) in sd:
        n += 1
    return n


def positional_encoding(p, octaves):
    """[p, sin(2^0 p), cos(2^0 p), ..., sin(2^{L-1} p), cos(2^{L-1} p)], pi-factor 1.0 (network.py:141-150)."""
    parts = [p]
    for i in range(octaves):
        parts.append(torch.sin((2 ** i) * 1.0 * p))
        parts.append(torch.cos((2 ** i) * 1.0 * p))
    return torch.cat(parts, dim=-1)


def softplus100(x):
    return F.softplus(x, beta=100)  # nn.Softplus(beta=100), threshold 20 (network.py:68)


def geo_forward(sd, p, mcfg, return_pre=False):
    """Occupancy/feature MLP (network.py:85-95): out[...,0]=logit, out[...,1:]=feature."""
    layers = stage1_weights(sd, "lin", count_layers(sd, "lin"))
    pe = positional_encoding(p / mcfg["rescale"], mcfg["octaves_pe"])
    x = pe
    pre = []
    nl = len(layers)
    for l, (W, b) in enumerate(layers):
        if l in mcfg["skips"]:
            x = torch.cat([x, pe], -1) / np.sqrt(2)
        x = F.linear(x, W, b)
        if return_pre:
            pre.append(x)
        if l < nl - 1:
            x = softplus100(x)
    return (x, pre) if return_pre else x


def geo_gradient(sd, p, mcfg):
    """d logit / d p via autograd, [M,1,3] (network.py:108-120)."""
    with torch.enable_grad():
        q = p.detach().clone().requires_grad_(True)
        y = geo_forward(sd, q, mcfg)[..., :1]
        (g,) = torch.autograd.grad(y, q, torch.ones_like(y))
    return g.detach().unsqueeze(1)


def geo_gradient_analytic(sd, p, mcfg):
    """Hand-derived reverse pass equal to geo_gradient; this is the algorithm the CUDA kernels run.

    softplus'(z) = sigmoid(100 z); skip layer input is cat[x, pe]/sqrt(2); PE Jacobian is
    d/dp [p, sin(f p), cos(f p)] = [1, f cos(f p), -f sin(f p)].
    """
    layers = stage1_weights(sd, "lin", count_layers(sd, "lin"))
    nl = len(layers)
    L = mcfg["octaves_pe"]
    q = p / mcfg["rescale"]
    pe = positional_encoding(q, L)
    npe = pe.shape[-1]
    _, pre = geo_forward(sd, p, mcfg, return_pre=True)
    inv = 1.0 / np.sqrt(2)
    gx = layers[nl - 1][0][0:1, :].expand(p.shape[0], -1)  # d logit / d x_{nl-1}
    gpe = torch.zeros_like(pe)
    for l in range(nl - 2, -1, -1):
        gz = gx * torch.sigmoid(100.0 * pre[l])
        gx = gz @ layers[l][0]
        if l in mcfg["skips"]:
            gx = gx * inv
            gpe = gpe + gx[:, -npe:]
            gx = gx[:, :-npe]
    gpe = gpe + gx
    g = gpe[:, 0:3].clone()
    for i in range(L):
        f = float(2 ** i)
        g = g + f * torch.cos(f * q) * gpe[:, 3 + 6 * i:6 + 6 * i] - f * torch.sin(f * q) * gpe[:, 6 + 6 * i:9 + 6 * i]
    return (g / mcfg["rescale"]).unsqueeze(1)


def app_forward(sd, points, normals, view_pe, feat):
    """Appearance MLP on cat[p, PE(view), normal, feature]; ReLU hidden; tanh*0.5+0.5 (network.py:97-106)."""
    layers = stage1_weights(sd, "lina", count_layers(sd, "lina"))
    x = torch.cat([points, view_pe, normals.squeeze(-2), feat], dim=-1)
    for l, (W, b) in enumerate(layers):
        x = F.linear(x, W, b)
        if l < len(layers) - 1:
            x = torch.relu(x)
    return torch.tanh(x) * 0.5 + 0.5


def network_forward(sd, mcfg, p, ray_d=None, only_occupancy=False, return_logits=False, return_addocc=False):
    """Dispatcher equal to NeuralNetwork.forward (network.py:122-138)."""
    x = geo_forward(sd, p, mcfg)
    if only_occupancy:
        return torch.sigmoid(x[..., :1] * -10.0)
    if ray_d is not None:
        v = ray_d / torch.norm(ray_d, dim=-1, keepdim=True)
        v = positional_encoding(v, mcfg["octaves_pe_views"])
        n = geo_gradient(sd, p, mcfg)  # un-normalised gradient feeds the app MLP (network.py:130-132)
        rgb = app_forward(sd, p, n, v, x[..., 1:])
        if return_addocc:
            return rgb, torch.sigmoid(x[..., :1] * -10.0)
        return rgb
    if return_logits:
        return -1 * x[..., :1]
    return None


# ----------------------------------------------------------------------------------------------
# Stage 1: cameras and rays (stage1/model/common.py)
# ----------------------------------------------------------------------------------------------
def arange_pixels(resolution, image_range=(-1.0, 1.0)):
    """x-major integer pixel grid + scaled copy (common.py:55-93)."""
    h, w = resolution
    gx, gy = torch.meshgrid(torch.arange(0, w), torch.arange(0, h), indexing="ij")
    loc = torch.stack([gx, gy], dim=-1).long().view(1, -1, 2)
    sc = loc.clone().float()
    scale = image_range[1] - image_range[0]
    half = scale / 2
    sc[:, :, 0] = scale * sc[:, :, 0] / (w - 1) - half
    sc[:, :, 1] = scale * sc[:, :, 1] / (h - 1) - half
    return loc, sc


def pixels_to_rays(pixels, camera_mat, world_mat):
    """(origin [1,N,3], unit direction [1,N,3]); fx used for both axes (common.py:205-226, rendering.py:67-71)."""
    n = pixels.shape[1]
    origin = world_mat[:, :3, -1].unsqueeze(1).repeat(1, n, 1)
    pt = (pixels - camera_mat[0, :2, 2]) / camera_mat[0, 0, 0]
    pt = torch.cat([pt, torch.ones_like(pt[..., :1])], dim=2)
    d = torch.einsum("bij,bnj->bni", world_mat[:, :3, :3], pt)
    d = d / d.norm(2, 2).unsqueeze(-1)
    return origin, d


def sphere_intersection(cam_loc, ray_dirs, r=1.0):
    """near/far of ray-sphere, clamped >=0, zeros for misses (rendering.py:576-596)."""
    n_img, n_pix, _ = ray_dirs.shape
    rc = torch.bmm(ray_dirs, cam_loc.unsqueeze(-1)).squeeze()
    under = (rc ** 2 - (cam_loc.norm(2, 1) ** 2 - r ** 2)).reshape(-1)
    hit = under > 0
    out = torch.zeros(n_img * n_pix, 2, dtype=ray_dirs.dtype, device=ray_dirs.device)
    sq = torch.sqrt(under[hit]).unsqueeze(-1) * torch.tensor([-1.0, 1.0], dtype=ray_dirs.dtype)
    out[hit] = sq
    out[hit] -= rc.reshape(-1)[hit].unsqueeze(-1)
    out = out.reshape(n_img, n_pix, 2).clamp_min(0.0)
    return out, hit.reshape(n_img, n_pix)


# ----------------------------------------------------------------------------------------------
# Stage 1: surface search (stage1/model/rendering.py:410-555)
# ----------------------------------------------------------------------------------------------
def _occ(sd, mcfg, p, max_points=3500000):
    flat = p.reshape(-1, 3)
    outs = [network_forward(sd, mcfg, c, only_occupancy=True) for c in torch.split(flat, int(max_points), dim=0)]
    return torch.cat(outs, dim=0)


def secant(sd, mcfg, f_low, f_high, d_low, d_high, n_secant, o, d, tau):
    """Secant refinement on [d_low, d_high] (rendering.py:525-555)."""
    d_pred = -f_low * (d_high - d_low) / (f_high - f_low) + d_low
    for _ in range(n_secant):
        p_mid = o + d_pred.unsqueeze(-1) * d
        f_mid = _occ(sd, mcfg, p_mid)[..., 0] - tau
        lo = f_mid < 0
        d_low = torch.where(lo, d_pred, d_low)
        f_low = torch.where(lo, f_mid, f_low)
        d_high = torch.where(lo, d_high, d_pred)
        f_high = torch.where(lo, f_high, f_mid)
        d_pred = -f_low * (d_high - d_low) / (f_high - f_low) + d_low
    return d_pred


def ray_marching(sd, mcfg, ray0, ray_dir, n_steps, near, rad, n_secant=8, tau=0.5, return_aux=False):
    """First occupancy sign change along each ray + secant; inf = miss, 0 = first point occupied.

    rendering.py:410-523.  ray0/ray_dir: [1,N,3].  n_steps is the lower bound of the reference's
    randint(a, a+1) (rendering.py:441).
    """
    B, N, _ = ray0.shape
    dt = ray0.dtype
    far = sphere_intersection(ray0[:, 0], ray_dir, r=rad)[0][..., 1]
    t = torch.linspace(0, 1, steps=n_steps).view(1, 1, n_steps, 1).to(dt)
    d_prop = near * (1.0 - t) + far.view(1, -1, 1, 1) * t
    p_prop = ray0.unsqueeze(2) + ray_dir.unsqueeze(2) * d_prop
    val = (_occ(sd, mcfg, p_prop) - tau).view(B, N, n_steps)
    first_free = val[:, :, 0] < 0
    sign = torch.cat([torch.sign(val[:, :, :-1] * val[:, :, 1:]), torch.ones(B, N, 1, dtype=dt)], dim=-1)
    cost = sign * torch.arange(n_steps, 0, -1).to(dt)
    values, idx = torch.min(cost, -1)
    changed = values < 0
    n = B * N
    ar = torch.arange(n)
    v_flat = val.reshape(n, n_steps)
    d_flat = d_prop.reshape(n, n_steps)
    neg_to_pos = v_flat[ar, idx.view(n)].view(B, N) < 0
    mask = changed & neg_to_pos & first_free
    idx2 = torch.clamp(idx + 1, max=n_steps - 1)
    d_low = d_flat[ar, idx.view(n)].view(B, N)[mask]
    f_low = v_flat[ar, idx.view(n)].view(B, N)[mask]
    d_high = d_flat[ar, idx2.view(n)].view(B, N)[mask]
    f_high = v_flat[ar, idx2.view(n)].view(B, N)[mask]
    d_pred = secant(sd, mcfg, f_low, f_high, d_low, d_high, n_secant, ray0[mask], ray_dir[mask], tau)
    out = torch.ones(B, N, dtype=dt)
    out[mask] = d_pred
    out[mask == 0] = np.inf
    out[first_free == 0] = 0
    if return_aux:
        return out, {"val": val, "idx": idx, "mask": mask, "far": far}
    return out


def _surface_from_depth(d_i, ray0, ray_dir):
    """dists / object mask / surface points shared by unisurf, phong and shape_extract (rendering.py:88-108)."""
    zero_occ = d_i == 0
    finite = (d_i.abs() != np.inf) & ~torch.isnan(d_i)
    dists = torch.ones_like(d_i)
    dists[finite] = d_i[finite]
    dists[zero_occ] = 0.0
    obj = (finite & ~zero_occ)[0]
    dists = dists[0]
    o = ray0.reshape(-1, 3)
    d = ray_dir.reshape(-1, 3)
    pts = o + d * dists.unsqueeze(-1)
    return obj, dists, o, d, pts


def composite(alpha):
    """w_i = a_i * prod_{j<i}(1 - a_j + 1e-6) (rendering.py:196, :405)."""
    ones = torch.ones((alpha.shape[0], 1), dtype=alpha.dtype)
    return alpha * torch.cumprod(torch.cat([ones, 1.0 - alpha + EPS_T], -1), -1)[:, :-1]


def unisurf_render(sd, cfg_all, pixels, camera_mat, world_mat, it=100000, noise=None, return_aux=False):
    """Renderer.unisurf with add_noise=False, eval_=True (rendering.py:50-226).

    ``noise`` (optional dict with 'miss' [Nmiss,S] and 'hit' [Nhit,S] uniform samples) reproduces
    the add_noise=True stratified jitter with externally supplied randoms (rendering.py:135-140,159-164).
    """
    mcfg, rcfg = cfg_all["model"], cfg_all["rendering"]
    dt = world_mat.dtype
    B, N, _ = pixels.shape
    rad = rcfg["radius"]
    near = float(rcfg["near"])
    steps, steps_out = rcfg["num_points_in"], rcfg["num_points_out"]
    ray0, rayd = pixels_to_rays(pixels, camera_mat, world_mat)
    d_sph, _ = sphere_intersection(ray0[:, 0], rayd, r=rad)
    d_i = ray_marching(sd, mcfg, ray0, rayd, int(rcfg["ray_marching_steps"]), near, rad, n_secant=8)
    obj, dists, o, d, pts = _surface_from_depth(d_i, ray0, rayd)
    d_sph[:, :, 0] = 0.0
    d_sph = d_sph.reshape(-1, 2)
    d_hit = dists[obj]
    d_far_hit = d_sph[obj][:, 1]
    delta = torch.max(rcfg["interval_start"] * torch.exp(-1 * rcfg["interval_decay"] * it * torch.ones(1)),
                      rcfg["interval_end"] * torch.ones(1)).to(dt)
    near_t = torch.tensor(near, dtype=dt)
    dnp = d_hit - delta
    dfp = d_hit + delta
    dnp = torch.where(dnp < near_t, near_t, dnp)
    dfp = torch.where(dfp > d_far_hit, d_far_hit, dfp)
    full = steps + steps_out if (bool((dnp != 0.0).all()) and it > 5000) else steps
    far_miss = d_sph[~obj][:, 1]
    t_full = torch.linspace(0.0, 1.0, steps=full).to(dt).view(1, -1)
    d2 = near_t * (1.0 - t_full) + far_miss.view(-1, 1) * t_full
    t_in = torch.linspace(0.0, 1.0, steps=steps).to(dt).view(1, -1)
    d_int = dnp.view(-1, 1) * (1.0 - t_in) + dfp.view(-1, 1) * t_in
    if full != steps:
        t_out = torch.linspace(0.0, 1.0, steps=steps_out).to(dt).view(1, -1)
        d_b = near_t * (1.0 - t_out) + dnp.view(-1, 1) * t_out
        d1, _ = torch.sort(torch.cat([d_b, d_int], dim=-1), dim=-1)
    else:
        d1 = d_int

    def jitter(dd, u):
        mid = 0.5 * (dd[:, 1:] + dd[:, :-1])
        hi = torch.cat([mid, dd[:, -1:]], dim=-1)
        lo = torch.cat([dd[:, :1], mid], dim=-1)
        return lo + (hi - lo) * u

    if noise is not None:
        d2 = jitter(d2, noise["miss"])
        d1 = jitter(d1, noise["hit"])
    depth = torch.zeros(N, full, dtype=dt)
    depth[~obj] = d2
    depth[obj] = d1
    p_fg = (o.unsqueeze(1) + d.unsqueeze(1) * depth.unsqueeze(-1)).reshape(-1, 3)
    v_fg = (-1 * d.unsqueeze(1).repeat(1, full, 1)).reshape(-1, 3)
    rgbs, alphas = [], []
    for i in range(0, p_fg.shape[0], rcfg["n_max_network_queries"]):
        r_i, a_i = network_forward(sd, mcfg, p_fg[i:i + rcfg["n_max_network_queries"]],
                                   v_fg[i:i + rcfg["n_max_network_queries"]], return_addocc=True)
        rgbs.append(r_i)
        alphas.append(a_i)
    rgb_s = torch.cat(rgbs, 0).reshape(N, full, 3)
    alpha = torch.cat(alphas, 0).view(N, full)
    w = composite(alpha)
    rgb = torch.sum(w.unsqueeze(-1) * rgb_s, dim=-2)
    g = geo_gradient(sd, pts[obj], mcfg)
    nrm = g[:, 0, :] / (g[:, 0, :].norm(2, dim=1).unsqueeze(-1) + 10 ** (-5))
    normal = torch.zeros_like(rgb)
    normal[obj] = nrm
    acc = torch.sum(w, -1)
    if rcfg["white_background"]:
        rgb = rgb + (1.0 - acc.unsqueeze(-1))
    out = {"rgb": rgb.reshape(B, -1, 3), "mask_pred": obj, "diff_norm": None,
           "normal_pred": normal.reshape(B, -1, 3), "acc_map": acc.reshape(B, -1)}
    if return_aux:
        out["aux"] = {"d_i": d_i, "depth": depth, "alpha": alpha, "rgb_s": rgb_s, "points": pts, "full_steps": full}
    return out


def unisurf_train(sd, cfg_all, pixels, camera_mat, world_mat, it=100000, neigh_u=None, noise=None):
    """Renderer.unisurf with eval_=False (the training forward, rendering.py:50-226), differentiable w.r.t. the tensors of
    ``sd``: radiance samples with create_graph normals (network.py:117,130-132), compositing, surface normals and the
    normal-consistency term diff_norm (rendering.py:199-212).  The surface search and the sample depths carry no gradient
    (rendering.py:79-87 runs under no_grad).  ``neigh_u`` [Ns,3] replaces torch.rand_like (rendering.py:204)."""
    mcfg, rcfg = cfg_all["model"], cfg_all["rendering"]
    with torch.no_grad():
        sd0 = {k: v.detach() for k, v in sd.items()}
        base = unisurf_render(sd0, cfg_all, pixels, camera_mat, world_mat, it=it, noise=noise, return_aux=True)
        depth, pts, obj = base["aux"]["depth"], base["aux"]["points"], base["mask_pred"]
        ray0, rayd = pixels_to_rays(pixels, camera_mat, world_mat)
    o, d = ray0.reshape(-1, 3), rayd.reshape(-1, 3)
    N, S = depth.shape
    p_fg = (o.unsqueeze(1) + d.unsqueeze(1) * depth.unsqueeze(-1)).reshape(-1, 3)
    v_fg = (-1 * d.unsqueeze(1).repeat(1, S, 1)).reshape(-1, 3)
    x = geo_forward(sd, p_fg, mcfg)
    v = positional_encoding(v_fg / torch.norm(v_fg, dim=-1, keepdim=True), mcfg["octaves_pe_views"])
    n = geo_gradient_analytic(sd, p_fg, mcfg)  # differentiable restatement of gradient(p, create_graph=True)
    rgb_s = app_forward(sd, p_fg, n, v, x[..., 1:]).reshape(N, S, 3)
    alpha = torch.sigmoid(x[..., :1] * -10.0).view(N, S)
    w = composite(alpha)
    rgb = torch.sum(w.unsqueeze(-1) * rgb_s, dim=-2)
    sp = pts[obj]
    Ns = sp.shape[0]
    if neigh_u is None:
        neigh_u = torch.rand_like(sp)
    pp = torch.cat([sp, sp + (neigh_u - 0.5) * 0.01], 0)
    g = geo_gradient_analytic(sd, pp, mcfg)[:, 0, :]
    nrm = g / (g.norm(2, dim=1).unsqueeze(-1) + 10 ** (-5))
    normal = torch.zeros_like(rgb)
    normal[obj] = nrm[:Ns]
    diff_norm = torch.norm(nrm[:Ns] - nrm[Ns:], dim=-1)
    acc = torch.sum(w, -1)
    if rcfg["white_background"]:
        rgb = rgb + (1.0 - acc.unsqueeze(-1))
    return {"rgb": rgb.reshape(1, -1, 3), "mask_pred": obj, "diff_norm": diff_norm, "normal_pred": normal.reshape(1, -1, 3),
            "acc_map": acc.reshape(1, -1)}


def light_visibility(sd, mcfg, surf, light_dir, lnear=0.1, lfar=3.5, n_steps=128, box=1.1):
    """Shadow-ray transmittance per (light, surface point), light-major [L*Ns] (rendering.py:378-408)."""
    dt = surf.dtype
    L, Ns = light_dir.shape[0], surf.shape[0]
    t = torch.linspace(0, 1, steps=n_steps).view(1, 1, n_steps, 1).to(dt)
    dd = lnear * (1.0 - t) + lfar * t
    p = surf[None, :, None, :] + light_dir[:, None, None, :] * dd
    p = p.expand(L, Ns, n_steps, 3)
    alpha = _occ(sd, mcfg, p).view(-1, n_steps).clone()
    inside = torch.logical_and((p <= box).all(dim=-1), (p >= -box).all(dim=-1)).reshape(-1, n_steps)
    alpha[~inside] = 0
    return 1 - torch.sum(composite(alpha), -1)


def shape_extract(sd, cfg_all, pixels, camera_mat, world_mat, visibility=False, light_dir=None, ray_steps=512):
    """Surface points / F.normalize'd normals / mask (+ per-light visibility) (rendering.py:297-376)."""
    mcfg, rcfg = cfg_all["model"], cfg_all["rendering"]
    B, N, _ = pixels.shape
    ray0, rayd = pixels_to_rays(pixels, camera_mat, world_mat)
    d_i = ray_marching(sd, mcfg, ray0, rayd, ray_steps, float(rcfg["near"]), rcfg["radius"], n_secant=8)
    obj, dists, o, d, pts = _surface_from_depth(d_i, ray0, rayd)
    surf = pts[obj]
    normal = torch.zeros(B * N, 3, dtype=world_mat.dtype)
    if len(surf) > 0:
        normal[obj] = F.normalize(geo_gradient(sd, surf, mcfg)[:, 0, :], dim=-1)
    out = {"mask": obj.reshape(B, -1), "normal": normal.reshape(B, -1, 3), "points": pts.reshape(B, -1, 3)}
    if visibility and light_dir is not None:
        vis = torch.ones(light_dir.shape[0], N, dtype=world_mat.dtype)
        if len(surf) > 0:
            parts = [light_visibility(sd, mcfg, surf, light_dir[s:s + 96]) for s in range(0, len(light_dir), 96)]
            vis[obj[None, ].expand_as(vis)] = torch.cat(parts, dim=0)
        out["visibility"] = vis
    return out


def phong_render(sd, cfg_all, pixels, camera_mat, world_mat):
    """Debug Phong shading of the marched surface, 512 steps (rendering.py:228-293)."""
    mcfg, rcfg = cfg_all["model"], cfg_all["rendering"]
    B, N, _ = pixels.shape
    dt = world_mat.dtype
    ray0, rayd = pixels_to_rays(pixels, camera_mat, world_mat)
    src = ray0[0, 0]
    light = (src / src.norm(2)).unsqueeze(1)
    d_i = ray_marching(sd, mcfg, ray0, rayd, 512, float(rcfg["near"]), rcfg["radius"], n_secant=8)
    obj, dists, o, d, pts = _surface_from_depth(d_i, ray0, rayd)
    rgb = torch.ones_like(pts)
    g = geo_gradient(sd, pts[obj], mcfg)[:, 0, :]
    nrm = g / g.norm(2, 1, keepdim=True)
    diffuse = torch.mm(nrm, light).clamp_min(0).repeat(1, 3) * torch.tensor([0.7, 0.7, 0.7], dtype=dt).unsqueeze(0)
    rgb[obj] = (torch.tensor([0.3, 0.3, 0.3], dtype=dt).unsqueeze(0) + diffuse).clamp_max(1.0)
    return {"rgb": rgb.reshape(B, -1, 3)}


# ----------------------------------------------------------------------------------------------
# Stage 2: photometric-stereo shading (stage2/model/{renderer,sgbasis,embedder}.py)
# ----------------------------------------------------------------------------------------------
def embed(x, n_freqs):
    """NeRF embedding, include input, log-sampled 2^0..2^{n-1} (embedder.py:6-54)."""
    if n_freqs <= 0:
        return x
    bands = 2.0 ** torch.linspace(0.0, n_freqs - 1, steps=n_freqs)
    parts = [x]
    for f in bands:
        parts.append(torch.sin(x * f))
        parts.append(torch.cos(x * f))
    return torch.cat(parts, -1)


def s2_mlp(sd, prefix, x, skip_at, final):
    """Plain Linear stack, ReLU hidden, cat[y, x] AFTER layer skip_at (renderer.py:17-49)."""
    n = 0
    while ("%s.linears.%d.bias" % (prefix, n)) in sd:
        n += 1
    y = x
    for li in range(n):
        y = F.linear(y, sd["%s.linears.%d.weight" % (prefix, li)], sd["%s.linears.%d.bias" % (prefix, li)])
        if li != n - 1:
            y = torch.relu(y)
        elif final == "sigmoid":
            y = torch.sigmoid(y)
        if li in skip_at:
            y = torch.cat([y, x], -1)
    return y


def sg_basis(lobe, v, n, l, albedo, weights, specular_rgb, nbasis):
    """brdf = albedo + sum_k w_k exp(lambda_k (h.n - 1)) (sgbasis.py:16-32); nbasis = lobes per channel."""
    h = F.normalize(l + v, dim=-1)
    D = torch.exp(lobe[None, ].clamp(min=0) * ((h * n).sum(-1, keepdim=True) - 1))
    if specular_rgb:
        spec = (weights.view(-1, 3, nbasis) * D[:, None]).sum(-1).clamp(min=0.0)
    else:
        spec = (weights * D).sum(-1, keepdim=True).clamp(min=0.0)
    return albedo + spec.expand_as(albedo), spec


def _div_no_nan(x, y):
    a = x / (y + 1e-6)  # microfacet.py:20-24
    return torch.where(torch.isinf(a) | torch.isnan(a), torch.zeros_like(a), a)


def microfacet_brdf(l, v, n, albedo, rough, f0=0.05):
    """GGX microfacet BRDF (stage2/model/microfacet.py:35-114).  l [Np,L,3], v/n [Np,3], albedo [Np,3], rough [Np,1] -> [Np,L,3]."""
    l = F.normalize(l, dim=2, eps=1e-6)
    v = F.normalize(v, dim=1, eps=1e-6)
    n = F.normalize(n, dim=1, eps=1e-6)
    h = F.normalize(l + v[:, None, :], dim=2, eps=1e-6)
    f = f0 + (1 - f0) * (1 - torch.einsum("ijk,ijk->ij", l, h)) ** 5          # Schlick (:108-113)
    alpha = rough ** 2
    cm = torch.einsum("ijk,ik->ij", h, n)                                      # GGX distribution (:94-106)
    cm2 = cm ** 2
    tm2 = _div_no_nan(1 - cm2, cm2)
    d = _div_no_nan(alpha ** 2 * (cm > 0).to(l.dtype), np.pi * cm2 ** 2 * (alpha ** 2 + tm2) ** 2)
    cv = torch.einsum("ij,ij->i", n, v)                                        # GGX geometry term (:76-92)
    chi = (_div_no_nan(torch.einsum("ijk,ik->ij", h, v), cv[:, None]) > 0).to(l.dtype)
    cv2 = torch.clamp(cv ** 2, 0.0, 1.0)
    tv2 = torch.clamp(_div_no_nan(1 - cv2, cv2), 0.0, float("inf"))
    g = _div_no_nan(chi * 2, 1 + torch.sqrt(1 + alpha ** 2 * tv2[:, None]))
    ldn = torch.einsum("ijk,ik->ij", l, n)
    micro = _div_no_nan(f * g * d, 4 * ldn.abs() * cv.abs()[:, None])
    return micro[:, :, None].repeat(1, 1, 3) + (albedo / np.pi)[:, None, :]


def camera_params(uv, pose, intrinsics):
    """Unit ray dirs + camera centre, pose-matrix branch (stage2/utils/rend_util.py:90-147)."""
    cam_loc = pose[:, :3, 3]
    fx = intrinsics[:, 0, 0].unsqueeze(-1)
    fy = intrinsics[:, 1, 1].unsqueeze(-1)
    cx = intrinsics[:, 0, 2].unsqueeze(-1)
    cy = intrinsics[:, 1, 2].unsqueeze(-1)
    z = torch.ones_like(uv[:, :, 0])
    pc = torch.stack(((uv[:, :, 0] - cx) / fx * z, (uv[:, :, 1] - cy) / fy * z, z), dim=-1)
    dirs = torch.einsum("bij,bnj->bni", pose[:, :3, :3], pc)
    return F.normalize(dirs, dim=2), cam_loc


def psnetwork_forward(sd, conf, inp, noise=None, albedo_new=None, basis_new=None):
    """PSNetwork.forward for the shipped configuration family (renderer.py:110-266):
    render_model=sgbasis, shape_pregen, normal_mlp, visibility.  ``conf`` is a flat dict
    (see psnerf_b200.synth.stage2_conf).  ``noise`` optionally supplies 'xyz' [Ns,3] N(0,1) samples
    (scaled by xyz_jitter_std) in place of torch.normal (renderer.py:212).  ``albedo_new`` ([3]) / ``basis_new`` (lobe index)
    are the material-editing overrides of stage2/eval.py:116-132 (renderer.py:167-168,175-181).
    """
    micro = conf.get("train.render_model", "sgbasis") == "microfacet"
    nb = int(conf.get("train.nbasis", 9))
    spec_rgb = bool(conf.get("train.specular_rgb", False)) and not micro
    nf = int(conf["brdf.net.n_freqs_xyz"])
    nfn = int(conf["normal.net.n_freqs_xyz"])
    uv, pose, K = inp["uv"], inp["pose"], inp["intrinsics"]
    ray_dirs, _ = camera_params(uv, pose, K)
    smask = inp["surface_mask"]
    points = inp["points"]
    normals_in = inp["normal"]
    surf = points[smask]
    dt = points.dtype
    normal_pred = torch.ones_like(points)
    L = inp["light_direction"].shape[0]
    nbt = 1 if micro else (nb * 3 if spec_rgb else nb)
    rgb_v = torch.ones_like(points)
    alb_v = torch.ones_like(points)
    rough_v = torch.ones_like(points)
    w_v = torch.zeros(*points.shape[:-1], nbt, dtype=dt)
    vis_v = torch.ones_like(points)
    if L > 1:
        rgb_v = rgb_v.repeat(L, 1, 1)
        if not micro:  # renderer.py:156-157
            rough_v = rough_v.repeat(L, 1, 1)
        vis_v = vis_v.repeat(L, 1, 1)
    out_extra = {}
    Ns = surf.shape[0]
    if Ns > 0:
        n_out = F.normalize(s2_mlp(sd, "normal_net", embed(surf, nfn), [int(conf["normal.net.mlp_skip_at"])], None), dim=-1)
        normal_pred[smask] = n_out
        normal = normal_pred[smask]
        v = -ray_dirs[smask]
        me = smask.expand(L, -1)
        l = inp["light_direction"][:, None].expand(rgb_v.shape)[me]
        pemb = embed(surf, nf)
        albedo = s2_mlp(sd, "albedo_net", pemb, [int(conf["brdf.net.mlp_skip_at"])], "sigmoid")
        if micro:  # renderer.py:73-74: Network(dim_emb, 1, W, depth) with the albedo net's trunk shape, sigmoid output
            rough = s2_mlp(sd, "rough_net", pemb, [int(conf["brdf.net.mlp_skip_at"])], "sigmoid")
        else:
            rough = s2_mlp(sd, "rough_net", pemb, [int(conf.get("brdf.sgnet.mlp_skip_at", 2))], None)
        if albedo_new is not None:  # renderer.py:167-168
            albedo = torch.as_tensor(albedo_new, dtype=dt)[None].expand_as(albedo)
        if micro:  # renderer.py:171-172
            brdf = microfacet_brdf(l.view(L, -1, 3).permute(1, 0, 2), v, normal, albedo, rough,
                                   float(conf.get("brdf.fresnel_f0", 0.05))).permute(1, 0, 2).reshape(-1, 3)
            weights = rough
        else:
            weights = torch.relu(rough)
            if basis_new is not None:  # renderer.py:175-181: a single lobe with weight 2^k/100 in every colour channel
                wn = torch.zeros_like(weights)
                wn.view(-1, 3 if spec_rgb else 1, nb)[:, :, basis_new] = 2 ** basis_new / 100
                weights = wn.reshape(-1, nbt)
            if L > 1:
                brdf, spec = sg_basis(sd["sgbasis.lobe"], v.tile(L, 1), normal.tile(L, 1), l, albedo.tile(L, 1),
                                      weights.tile(L, 1), spec_rgb, nb)
            else:
                brdf, spec = sg_basis(sd["sgbasis.lobe"], v, normal, l, albedo, weights, spec_rgb, nb)
        w_v[smask] = weights
        cos = torch.einsum("lni,ni->ln", l.view(L, -1, 3), normal).reshape(-1, 1)
        inten = inp.get("light_intensity", float(conf.get("brdf.light_intensity", 4.0)))
        if torch.is_tensor(inten) and inten.shape[0] > 1:
            inten = inten.repeat_interleave(Ns, dim=0)
        lv = l.detach() if conf.get("train.light_vis_detach", False) else l  # renderer.py:192-195
        vis = s2_mlp(sd, "visibility_net", torch.cat([pemb.tile(L, 1), embed(lv, nf)], -1),
                     [int(conf["visibility.net.mlp_skip_at"])], None)
        vr = vis.detach() if conf.get("train.vis_rgb_detach", False) else vis  # renderer.py:196-199
        rgb = (brdf * inten * cos * vr.clamp(0, 1)).clamp(0, 1)
        vis_v[me] = vis.expand(rgb.shape)
        rgb_v[me] = rgb
        alb_v[smask] = albedo
        if micro:
            rough_v[smask] = rough.expand(-1, 3)  # renderer.py:206-207
        else:
            rough_v[me] = spec.expand(-1, 3)
        jstd = float(conf.get("brdf.net.xyz_jitter_std", 0))
        if jstd > 0 and noise is not None:
            pj = embed(surf + noise["xyz"] * jstd, nf)
            aj = torch.ones_like(points)
            aj[smask] = s2_mlp(sd, "albedo_net", pj, [int(conf["brdf.net.mlp_skip_at"])], "sigmoid")
            if micro:  # renderer.py:224-226
                rj = torch.ones_like(points)
                rj[smask] = s2_mlp(sd, "rough_net", pj, [int(conf["brdf.net.mlp_skip_at"])], "sigmoid").expand(-1, 3)
                out_extra.update({"albedo_values": alb_v, "albedo_jitter": aj, "rough_values": rough_v, "rough_jitter": rj})
            else:
                rj = torch.ones_like(w_v)
                rj[smask] = torch.relu(s2_mlp(sd, "rough_net", pj, [int(conf.get("brdf.sgnet.mlp_skip_at", 2))], None))
                out_extra.update({"albedo_values": alb_v, "albedo_jitter": aj, "rough_values": w_v, "rough_jitter": rj})
        if "light_vis_train" in inp:
            Lt = inp["light_vis_train"].shape[0]
            lt = inp["light_vis_train"][:, None].expand(-1, points.shape[1], -1)[smask.expand(Lt, -1)]
            if conf.get("train.light_vis_detach", False):
                lt = lt.detach()
            vt = torch.ones_like(points)
            if Lt > 1:
                vt = vt.repeat(Lt, 1, 1)
            vv = s2_mlp(sd, "visibility_net", torch.cat([pemb.tile(Lt, 1), embed(lt, nf)], -1),
                        [int(conf["visibility.net.mlp_skip_at"])], None)
            vt[smask.expand(Lt, -1)] = vv.expand(-1, 3)
            out_extra["vis_train"] = vt
    out = {"points": points, "object_mask": inp["object_mask"], "network_object_mask": smask,
           "sg_rgb_values": rgb_v, "normal_values": normals_in, "sg_diffuse_albedo_values": alb_v,
           "sg_specular_rgb_values": rough_v, "normal_pred": normal_pred, "visibility": vis_v, "sg_weight": w_v}
    if micro:
        del out["sg_weight"]  # renderer.py:263-264
    out.update(out_extra)
    return out


def main_loss(out, rgb_gt, inp, sg_rgb_weight=1.0, albedo_smooth_weight=0.05, rough_smooth_weight=0.01, vis_weight=1.0):
    """MainLoss with loss_type L1 (stage2/model/loss.py:6-98), device agnostic."""
    mask = out["network_object_mask"] & out["object_mask"]
    if mask.sum() == 0:
        return torch.zeros((), dtype=rgb_gt.dtype)
    me = mask.expand(rgb_gt.shape[0], -1)
    loss = sg_rgb_weight * F.l1_loss(out["sg_rgb_values"][me].reshape(-1, 3), rgb_gt[me].reshape(-1, 3))
    if "albedo_jitter" in out and albedo_smooth_weight > 0:
        m1 = mask.expand(out["albedo_values"].shape[0], -1)
        loss = loss + albedo_smooth_weight * F.l1_loss(out["albedo_values"][m1], out["albedo_jitter"][m1])
    if "rough_jitter" in out and rough_smooth_weight > 0:
        m1 = mask.expand(out["rough_values"].shape[0], -1)
        loss = loss + rough_smooth_weight * F.l1_loss(out["rough_values"][m1], out["rough_jitter"][m1])
    if "vis_train" in out and "vis_train_gt" in inp:
        mv = mask.expand(inp["vis_train_gt"].shape[0], -1)
        loss = loss + vis_weight * F.l1_loss(out["vis_train"][..., 0][mv].reshape(-1), inp["vis_train_gt"][mv].reshape(-1))
    return loss


def normal_loss(out, normal_weight=1.0):
    """NormalLoss without the (unused, jitter std 0) smoothness term (stage2/model/loss.py:102-141)."""
    mask = (out["network_object_mask"] & out["object_mask"])
    if mask.sum() == 0:
        return torch.zeros((), dtype=out["normal_pred"].dtype)
    gt = F.normalize(out["normal_values"], dim=-1)
    return normal_weight * F.mse_loss(out["normal_pred"][mask].reshape(-1, 3), gt[mask].reshape(-1, 3))


def latlong_light_grid(envmap_h, envmap_w, envmap_radius=1.0):
    """gen_light_xyz + sph2cart (stage2/utils/eval_utils.py:64-99,255-296): lat-long texel centres, poles excluded."""
    lat_step = np.pi / (envmap_h + 2)
    lng_step = 2 * np.pi / (envmap_w + 2)
    lats = np.linspace(np.pi / 2 - lat_step, -np.pi / 2 + lat_step, envmap_h)
    lngs = np.linspace(np.pi - lng_step, -np.pi + lng_step, envmap_w)
    lngs, lats = np.meshgrid(lngs, lats)
    z = envmap_radius * np.sin(lats)
    x = envmap_radius * np.cos(lats) * np.cos(lngs)
    y = envmap_radius * np.cos(lats) * np.sin(lngs)
    sin_colat = np.sin(np.pi / 2 - lats)
    return np.stack((x, y, z), -1).reshape(-1, 3), (4 * np.pi * sin_colat / np.sum(sin_colat)).reshape(-1)


def envmap_relight(sd, conf, inp, env_light, light_xyz, light_batch=64):
    """Envmap relighting loop of stage2/eval.py:196-219: per-light renders summed and clipped, visibility averaged."""
    rgbs, viss = [], []
    env_light = torch.as_tensor(env_light).float().reshape(-1, 3)
    xyz = torch.as_tensor(light_xyz).float().reshape(-1, 3)
    for s in range(0, xyz.shape[0], light_batch):
        i2 = dict(inp)
        i2["light_direction"] = F.normalize(xyz[s:s + light_batch], p=2, dim=-1)
        i2["light_intensity"] = env_light[s:s + light_batch]
        out = psnetwork_forward(sd, conf, i2)
        rgbs.append(out["sg_rgb_values"].numpy())
        viss.append(out["visibility"].numpy())
    return {"rgb": torch.from_numpy(np.concatenate(rgbs, 0).sum(0).clip(0, 1)), "visibility": torch.from_numpy(np.concatenate(viss, 0).mean(0))}


def split_input(model_input, total_pixels, n_pixels=1024):
    """1024-pixel chunking of per-pixel entries (stage2/utils/general.py:23-37)."""
    chunks = []
    for idx in torch.split(torch.arange(total_pixels), n_pixels, dim=0):
        d = dict(model_input)
        for k in ["uv", "object_mask", "gt_normal", "normal", "depth", "points", "surface_mask", "visibility"]:
            if k in model_input:
                d[k] = torch.index_select(model_input[k], 1, idx)
        chunks.append(d)
    return chunks


def merge_output(res, total_pixels, batch_size):
    """Concatenate chunk outputs along the pixel axis (stage2/utils/general.py:39-53)."""
    out = {}
    for k in res[0]:
        if res[0][k] is None:
            continue
        if len(res[0][k].shape) < 3:
            out[k] = torch.cat([r[k].reshape(batch_size, -1, 1) for r in res], 1).reshape(batch_size * total_pixels)
        else:
            out[k] = torch.cat([r[k].reshape(*r[k].shape[:-2], -1, r[k].shape[-1]) for r in res], -2
                               ).reshape(-1, res[0][k].shape[-1])
    return out


def psnr(a, b, mask=None):
    """10 log10(1/mse) over the (masked) pixels, 100 dB when identical (stage2/utils/metrics.py:38-51)."""
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    if mask is not None:
        m = torch.as_tensor(mask).bool()
        a, b = a[m], b[m]
    mse = float(torch.mean((a - b) ** 2))
    return 100.0 if mse == 0 else 10.0 * math.log10(1.0 / mse)


def mae(n1, n2, mask=None, normalize=True):
    """Mean angular error in degrees between two normal maps, and the per-pixel errors (stage2/utils/metrics.py:16-36): fp32
    normalisation by (|n| + 1e-5) with zero vectors kept zero, fp64 dot product clipped to [-1, 1]."""
    v1, v2 = np.array(n1, dtype=np.float32, copy=True), np.array(n2, dtype=np.float32, copy=True)
    if normalize:
        l1 = np.linalg.norm(v1.astype(np.float64), axis=-1)
        l2 = np.linalg.norm(v2.astype(np.float64), axis=-1)
        v1 /= l1[..., None] + 1e-5
        v2 /= l2[..., None] + 1e-5
        v1[l1 == 0] = 0
        v2[l2 == 0] = 0
    dot = (v1.astype(np.float64) * v2.astype(np.float64)).sum(-1).clip(-1, 1)
    if mask is not None:
        dot = dot[np.asarray(mask).astype(bool)]
    ang = np.arccos(dot) * 180.0 / math.pi
    return ang.mean(), ang


# ----------------------------------------------------------------------------------------------
# Optimizer steps of the train loops (the arithmetic lives in torch.optim, pinned pytorch=1.8.0:
# torch/optim/_functional.py adam() / sparse_adam(); call sites stage1/train.py:62,
# stage2/trainer.py:116,165).  Pinned against torch.optim.Adam / SparseAdam of this container
# in tests/test_oracle_golden.py.
# ----------------------------------------------------------------------------------------------
def adam_step(p, g, m, v, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """One torch.optim.Adam update (amsgrad=False) of numpy fp32 arrays; step = count after the increment.  Returns (p, m, v)."""
    f = np.float32
    b1, b2 = betas
    g = g.astype(f)
    if weight_decay != 0:
        g = g + f(weight_decay) * p
    m = m * f(b1) + g * f(1 - b1)
    v = v * f(b2) + g * g * f(1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = np.sqrt(v) / f(math.sqrt(bc2)) + f(eps)
    p = p - f(lr / bc1) * (m / denom)
    return p.astype(f), m.astype(f), v.astype(f)


def sparse_adam_step(p, rows, gv, m, v, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
    """One torch.optim.SparseAdam update of the [R,D] table p for the sparse gradient (rows [K], gv [K,D]); duplicate rows are
    summed first (coalesce), untouched rows keep p, m and v.  Returns (p, m, v)."""
    f = np.float32
    b1, b2 = betas
    p, m, v = p.copy(), m.copy(), v.copy()
    uniq = sorted(set(int(r) for r in rows))
    g = np.zeros((len(uniq), p.shape[1]), f)
    for k, r in enumerate(rows):
        g[uniq.index(int(r))] += gv[k]
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    step_size = f(lr * math.sqrt(bc2) / bc1)
    for i, r in enumerate(uniq):
        m[r] = m[r] + (g[i] - m[r]) * f(1 - b1)
        v[r] = v[r] + (g[i] * g[i] - v[r]) * f(1 - b2)
        p[r] = p[r] - step_size * (m[r] / (np.sqrt(v[r]) + f(eps)))
    return p, m, v
