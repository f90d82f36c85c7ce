"""Synthetic scenes for parity tests and benchmarks.

Datasets and checkpoints of the reference are download-only, so every measured configuration uses
reference-constructor weights under a fixed seed and the synthetic cameras / lights / surface
points defined here.  Pure host-side helpers: torch CPU tensors in, torch CPU tensors out.
"""
import copy
import math

import torch

# model / rendering sections of stage1/configs/bear.yaml:1-23 (bunny.yaml differs only in near/far)
STAGE1_MODEL = {
    "num_layers": 8, "hidden_dim": 256, "octaves_pe": 6, "octaves_pe_views": 4, "skips": [4],
    "geometric_init": True, "feat_size": 256, "rescale": 1.0,
}
STAGE1_RENDERING = {
    "type": "unisurf", "n_max_network_queries": 64000, "white_background": True,
    "near": 2, "far": 6, "radius": 2.0, "interval_start": 2.0, "interval_end": 0.1,
    "interval_decay": 0.000015, "num_points_in": 64, "num_points_out": 32,
    "ray_marching_steps": 256, "occ_prob_points": 64,
}

# stage2/confs/bear.conf:1-97 flattened (only keys read by the hot path)
STAGE2_CONF = {
    "train.render_model": "sgbasis", "train.nbasis": 9, "train.specular_rgb": True,
    "train.visibility": True, "train.light_vis_detach": True, "train.vis_rgb_detach": True,
    "train.normal_mlp": True, "train.normal_joint": True, "train.shape_pregen": True,
    "brdf.net.n_freqs_xyz": 10, "brdf.net.mlp_width": 128, "brdf.net.mlp_depth": 4,
    "brdf.net.mlp_skip_at": 2, "brdf.net.xyz_jitter_std": 0.01,
    "brdf.sgnet.mlp_width": 64, "brdf.sgnet.mlp_depth": 2, "brdf.sgnet.mlp_skip_at": -1,
    "brdf.fresnel_f0": 0.05, "brdf.light_intensity": 2.0,
    "normal.net.n_freqs_xyz": 10, "normal.net.mlp_width": 128, "normal.net.mlp_depth": 4,
    "normal.net.mlp_skip_at": 2, "normal.net.xyz_jitter_std": 0.0,
    "visibility.net.n_freqs_xyz": 10, "visibility.net.mlp_width": 256, "visibility.net.mlp_depth": 8,
    "visibility.net.mlp_skip_at": 4,
}


def stage1_cfg(num_points_in=64, num_points_out=32, ray_marching_steps=256, near=2, far=6, **over):
    cfg = {"model": copy.deepcopy(STAGE1_MODEL), "rendering": copy.deepcopy(STAGE1_RENDERING)}
    cfg["rendering"].update({"num_points_in": num_points_in, "num_points_out": num_points_out,
                             "ray_marching_steps": ray_marching_steps, "near": near, "far": far})
    cfg["rendering"].update(over)
    return cfg


def stage2_conf(**over):
    c = dict(STAGE2_CONF)
    c.update(over)
    return c


def look_at_pose(azim_deg, elev_deg, radius=4.0):
    """OpenCV camera-to-world pose (x right, y down, z forward) on an orbit, looking at the origin."""
    a, e = math.radians(azim_deg), math.radians(elev_deg)
    c = torch.tensor([radius * math.cos(e) * math.sin(a), -radius * math.sin(e), -radius * math.cos(e) * math.cos(a)])
    z = -c / c.norm()
    up = torch.tensor([0.0, -1.0, 0.0])
    x = torch.linalg.cross(-up, z)
    x = x / x.norm()
    y = torch.linalg.cross(z, x)
    pose = torch.eye(4)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = x, y, z, c
    return pose.unsqueeze(0).float()


def intrinsics(h, w):
    """fx = fy = 800 * (W/512), principal point at the image centre; 4x4 like the reference datasets."""
    K = torch.eye(4)
    K[0, 0] = K[1, 1] = 800.0 * (w / 512.0)
    K[0, 2], K[1, 2] = w / 2.0, h / 2.0
    return K.unsqueeze(0).float()


def pixel_grid_xmajor(h, w):
    """Integer pixel grid in the x-major order of stage1 arange_pixels (stage1/model/common.py:73)."""
    gx, gy = torch.meshgrid(torch.arange(0, w), torch.arange(0, h), indexing="ij")
    return torch.stack([gx, gy], dim=-1).long().view(1, -1, 2)


def uv_grid_rowmajor(h, w):
    """Float uv (x, y), row-major, as stage2/eval.py:320-322 builds it from np.mgrid."""
    gy, gx = torch.meshgrid(torch.arange(0, h), torch.arange(0, w), indexing="ij")
    return torch.stack([gx, gy], dim=-1).float().view(1, -1, 2)


def lights(n=96, seed=2, axis=(0.0, 0.0, -1.0)):
    """n unit vectors in the hemisphere about ``axis`` (default: towards a camera on the -z side)."""
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(n, 3, generator=g)
    v = v / v.norm(dim=-1, keepdim=True)
    ax = torch.tensor(axis)
    ax = ax / ax.norm()
    s = (v @ ax).unsqueeze(-1)
    v = torch.where(s < 0, v - 2 * s * ax, v)
    return v.float()


def shell_points(n, seed=3):
    """n points uniform on the radius 0.5..1 shell, normals = normalised position (stage-2 all-surface case)."""
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    r = 0.5 + 0.5 * torch.rand(n, 1, generator=g)
    return (d * r).float().unsqueeze(0), d.float().unsqueeze(0)


def perturb_state_dict(sd, rel=0.02, seed=1):
    """'Trained-like' weights: every float tensor += N(0, rel*std) (fixed generator), SURVEY.md §8d."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(sd.keys()):  # sorted: independent of parameter registration order
        v = sd[k]
        if v.dtype.is_floating_point and v.numel() > 1 and "lobe" not in k:
            s = float(v.float().std()) if v.numel() > 1 else 0.0
            out[k] = v + torch.randn(v.shape, generator=g) * (rel * s + 1e-3 * rel)
        else:
            out[k] = v.clone()
    return out


def stage2_input(h, w, n_lights, pose=None, all_surface=True, seed=3, mask_frac=0.5):
    """Model-input dict for PSNetwork.forward (stage2/model/renderer.py:112-124,154) on CPU."""
    n = h * w
    pts, nrm = shell_points(n, seed)
    if all_surface:
        smask = torch.ones(1, n, dtype=torch.bool)
    else:
        g = torch.Generator().manual_seed(seed + 100)
        smask = torch.rand(1, n, generator=g) < mask_frac
    return {
        "intrinsics": intrinsics(h, w), "uv": uv_grid_rowmajor(h, w),
        "pose": pose if pose is not None else look_at_pose(20.0, 10.0),
        "object_mask": smask.clone(), "surface_mask": smask, "points": pts, "normal": nrm,
        "light_direction": lights(n_lights),
    }


def latlong_light_grid(envmap_h, envmap_w=None, radius=1.0):
    """Directions (and solid-angle weights) of the centres of an envmap_h x envmap_w latitude-longitude environment map, poles
    excluded - the light set of the envmap relighting mode (stage2/utils/eval_utils.py:64-99 via stage2/eval.py:109-111,
    light_h = 16 -> 512 lights).  Returns float64 numpy-free torch tensors xyz [h*w, 3], areas [h*w]."""
    envmap_w = 2 * envmap_h if envmap_w is None else envmap_w
    lat_step = math.pi / (envmap_h + 2)
    lng_step = 2 * math.pi / (envmap_w + 2)
    lats = torch.linspace(math.pi / 2 - lat_step, -math.pi / 2 + lat_step, envmap_h, dtype=torch.float64)
    lngs = torch.linspace(math.pi - lng_step, -math.pi + lng_step, envmap_w, dtype=torch.float64)
    lat = lats[:, None].expand(envmap_h, envmap_w)
    lng = lngs[None, :].expand(envmap_h, envmap_w)
    xyz = torch.stack([radius * torch.cos(lat) * torch.cos(lng), radius * torch.cos(lat) * torch.sin(lng), radius * torch.sin(lat)], -1)
    sin_colat = torch.sin(math.pi / 2 - lat)
    areas = 4 * math.pi * sin_colat / sin_colat.sum()
    return xyz.reshape(-1, 3), areas.reshape(-1)
