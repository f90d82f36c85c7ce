"""Entry-point loops of the reference kept as thin host code over the CUDA path:

  render_stage1_view   <- stage1/eval.py:82-119          (one library call per view instead of 256 x 1024-ray chunks)
  extract_shape        <- stage1/shape_extract.py:103-165 (points / normals / mask / per-light visibility of a view)
  render_stage2_view   <- stage2/eval.py:314-417          (all light batches of a view)
  occupancy_grid_logits<- stage1/model/extracting.py:84-96,137-155 (dense -logit grid for mesh export)
  render_envmap_view   <- stage2/eval.py:173-231          (envmap relighting: RGB-intensity light grid, sum over lights)
  extract_and_shade    <- SURVEY.md §8f-1: stage-1 surface extraction feeding stage-2 shading directly, no .npy hand-off
  save_shape_view / load_shape_view <- shape_extract.py:144-162 / stage2/datasets/dataset.py:100-114: the .npy hand-off, kept as an option
  *_sharded            <- SURVEY.md §8e: rays of a view dealt over the ranks (stage 2: balanced by surface pixels), one all_gather at the end

All functions take / return torch tensors; `view` dicts carry the camera: stage 1 {camera_mat, world_mat}, stage 2
{intrinsics, pose}.  The pixel order conventions of the reference are preserved (stage 1: x-major, undone by to_hw;
stage 2: row-major uv)."""
import torch

from . import sharding
from .stage1.common import arange_pixels, to_hw


@torch.no_grad()
def render_stage1_view(renderer, h, w, camera_mat, world_mat, it=100000, pixels=None):
    """rgb [h,w,3], normal [h,w,3], acc [h,w], mask [h,w] exactly as stage1/eval.py assembles them."""
    dev = next(renderer.model.parameters()).device
    p_loc = arange_pixels((h, w))[0].to(dev) if pixels is None else pixels
    out = renderer(p_loc, camera_mat, world_mat, None, "unisurf", add_noise=False, eval_=True, it=it)
    if pixels is not None:
        return out
    return {"rgb": to_hw(out["rgb"][0], h, w), "normal": to_hw(out["normal_pred"][0], h, w),
            "acc": to_hw(out["acc_map"][0], h, w)[..., 0], "mask": to_hw(out["mask_pred"].float(), h, w)[..., 0] > 0.5}


@torch.no_grad()
def extract_shape(renderer, h, w, camera_mat, world_mat, light_dir=None):
    """points/normal/mask (+ visibility [L, h*w]) of one view in x-major pixel order (shape_extract.py:133-163)."""
    dev = next(renderer.model.parameters()).device
    p_loc = arange_pixels((h, w))[0].to(dev)
    return renderer(p_loc, camera_mat, world_mat, None, "shape_extract", visibility=light_dir is not None, light_dir=light_dir)


@torch.no_grad()
def render_stage2_view(model, model_input, light_dirs, light_batch=96, light_intensity=None):
    """Loop over light batches (stage2/eval.py:345-365) and concatenate along the light axis."""
    outs = []
    for s in range(0, light_dirs.shape[0], light_batch):
        inp = dict(model_input)
        inp["light_direction"] = light_dirs[s:s + light_batch]
        if light_intensity is not None:
            inp["light_intensity"] = light_intensity[s:s + light_batch]
        outs.append(model(inp))
    res = dict(outs[-1])
    for k in ("sg_rgb_values", "sg_specular_rgb_values", "visibility"):
        if k in res:
            res[k] = torch.cat([o[k] for o in outs], 0)
    return res


@torch.no_grad()
def occupancy_grid_logits(model, nx, padding=0.0, points=None):
    """Dense-grid query of the occupancy field for mesh export (stage1/model/extracting.py:84-96,137-155 with
    common.py:253-272): value_grid[nx,nx,nx] = model(p, return_logits=True) = -logit on the box_size * [-0.5, 0.5]^3 lattice,
    box_size = 2 + padding; `points` ([M,3], e.g. the MISE query points already mapped into the box) replaces the lattice.
    One library call on the device - the 100 000-point host round trips of Extractor3D.eval_points are gone; MISE / marching
    cubes stay CPU code outside this package."""
    dev = next(model.parameters()).device
    if points is None:
        box = 2.0 + padding
        ax = torch.linspace(-0.5, 0.5, nx)  # float32 on the host exactly as make_3d_grid, then scaled
        p = torch.stack([ax.view(-1, 1, 1).expand(nx, nx, nx), ax.view(1, -1, 1).expand(nx, nx, nx),
                         ax.view(1, 1, -1).expand(nx, nx, nx)], -1).reshape(-1, 3)
        vals = model((box * p).to(dev), None, return_logits=True).squeeze(-1)
        return vals.reshape(nx, nx, nx)
    return model(torch.as_tensor(points).float().to(dev), None, return_logits=True).squeeze(-1)


@torch.no_grad()
def render_envmap_view(model, model_input, env_light, light_xyz, light_batch=64):
    """Environment-map relighting of one view (stage2/eval.py:173-231): every texel of the (down-sampled) lat-long map is one
    directional light with an RGB intensity; batches of `light_batch` lights go through PSNetwork.forward and the per-light
    images are summed and clipped to [0, 1]; visibility is averaged over the lights.  env_light [Ld,3], light_xyz [Ld,3]
    (synth.latlong_light_grid).  The sums are accumulated on the device in float64 (the reference concatenates every batch on
    the host and sums there), so no [Ld, N, 3] stack ever exists."""
    dev = next(model.parameters()).device
    env_light = torch.as_tensor(env_light).float().reshape(-1, 3).to(dev)
    dirs = torch.nn.functional.normalize(torch.as_tensor(light_xyz).float().reshape(-1, 3).to(dev), p=2, dim=-1)
    Ld = dirs.shape[0]
    rgb_sum = vis_sum = None
    for s in range(0, Ld, light_batch):
        inp = dict(model_input)
        inp["light_direction"] = dirs[s:s + light_batch]
        inp["light_intensity"] = env_light[s:s + light_batch]
        out = model(inp)
        r = out["sg_rgb_values"].double().sum(0)
        rgb_sum = r if rgb_sum is None else rgb_sum + r
        if "visibility" in out:
            v = out["visibility"].double().sum(0)
            vis_sum = v if vis_sum is None else vis_sum + v
    res = {"rgb": rgb_sum.clamp(0, 1).float()}
    if vis_sum is not None:
        res["visibility"] = (vis_sum / Ld).float()
    return res


@torch.no_grad()
def extract_and_shade(renderer, ps_model, h, w, camera_mat, world_mat, light_dirs, light_batch=96):
    """Stage-1 surface search + analytic normals feeding stage-2 shading in one pass (no points/normal/mask .npy files).
    Stage-1 pixel order is x-major while stage 2 consumes uv pairs explicitly, so no re-ordering is needed."""
    dev = next(renderer.model.parameters()).device
    shp = extract_shape(renderer, h, w, camera_mat, world_mat)
    p_loc = arange_pixels((h, w))[0].to(dev)
    K = torch.eye(4).unsqueeze(0)
    cm = camera_mat.detach().float().cpu()
    K[0, 0, 0] = K[0, 1, 1] = cm[0, 0, 0]  # stage 1 uses fx for both axes (common.py:220)
    K[0, 0, 2], K[0, 1, 2] = cm[0, 0, 2], cm[0, 1, 2]
    inp = {"intrinsics": K, "uv": p_loc.float(), "pose": world_mat, "object_mask": shp["mask"], "surface_mask": shp["mask"],
           "points": shp["points"], "normal": shp["normal"]}
    return shp, render_stage2_view(ps_model, inp, light_dirs.to(dev), light_batch)


@torch.no_grad()
def render_stage1_view_sharded(renderer, h, w, camera_mat, world_mat, rank, world, it=100000):
    """Every rank renders its tiles of rays; ONE all_gather returns the full [h*w, 7] (rgb, normal, acc) image."""
    dev = next(renderer.model.parameters()).device
    p_all = arange_pixels((h, w))[0]
    idx = sharding.shard_indices(h * w, rank, world)
    out = renderer(p_all[:, idx].to(dev), camera_mat, world_mat, None, "unisurf", add_noise=False, eval_=True, it=it)
    local = torch.cat([out["rgb"][0], out["normal_pred"][0], out["acc_map"][0].unsqueeze(-1)], -1)
    return sharding.gather_pixels(local, h * w, rank, world)


PER_PIXEL_INPUTS = ("uv", "object_mask", "gt_normal", "normal", "depth", "points", "surface_mask", "visibility")


@torch.no_grad()
def render_stage2_view_sharded(model, model_input, light_dirs, rank, world, light_batch=96, light_intensity=None, group=None):
    """BASELINE config 4: the pixels of one stage-2 view dealt over the ranks with the SURFACE pixels balanced
    (sharding.shard_indices_by_mask), every rank shades its share under all lights, ONE all_gather returns the full view on every
    rank.  Output: the per-pixel entries of render_stage2_view in the reference's shapes - sg_rgb_values / visibility [L,N,3],
    normal_pred / sg_diffuse_albedo_values [1,N,3]."""
    smask = model_input["surface_mask"]
    N = smask.shape[1]
    dev = model_input["points"].device

    def indices_of(r):
        return sharding.shard_indices_by_mask(smask[0], r, world)

    idx = indices_of(rank).to(dev)
    sub = dict(model_input)
    for k in PER_PIXEL_INPUTS:
        if k in model_input and torch.is_tensor(model_input[k]) and model_input[k].dim() >= 2 and model_input[k].shape[1] == N:
            sub[k] = torch.index_select(model_input[k], 1, idx)
    out = render_stage2_view(model, sub, light_dirs, light_batch, light_intensity)
    L = light_dirs.shape[0]
    n_loc = idx.numel()
    rgb = out["sg_rgb_values"].reshape(L, n_loc, 3)
    vis = out["visibility"].reshape(L, n_loc, 3)
    local = torch.cat([rgb.permute(1, 0, 2).reshape(n_loc, L * 3), vis.permute(1, 0, 2).reshape(n_loc, L * 3),
                       out["normal_pred"].reshape(n_loc, 3), out["sg_diffuse_albedo_values"].reshape(n_loc, 3)], -1).contiguous()
    full = sharding.gather_rows(local, N, rank, world, indices_of, group=group)
    return {"sg_rgb_values": full[:, :L * 3].reshape(N, L, 3).permute(1, 0, 2).contiguous(),
            "visibility": full[:, L * 3:2 * L * 3].reshape(N, L, 3).permute(1, 0, 2).contiguous(),
            "normal_pred": full[:, 6 * L:6 * L + 3].reshape(1, N, 3), "sg_diffuse_albedo_values": full[:, 6 * L + 3:].reshape(1, N, 3)}


# ---- the on-disk hand-off between the stages (optional: extract_and_shade needs none of it) ------------------------------------
def save_shape_view(out_dir, view_number, shape, h, w):
    """Write one view of `extract_shape` in the layout stage1/shape_extract.py:144-162 produces and stage2/datasets/dataset.py:100-114
    reads: points/view_XX.npy, normal/view_XX.npy (float32 [h,w,3]), mask/view_XX.npy (bool [h,w]) and, when the shadow pass ran,
    visibility/view_XX.npy (float32; the reference reshapes the x-major [L, h*w] array as (L, h, w) and swaps the last two axes,
    which is only a transpose to row-major for square images - reproduced as is).  view_number is 1-based like the file names."""
    import os

    import numpy as np
    name = "view_%02d.npy" % view_number
    for sub in ("points", "normal", "mask") + (("visibility",) if "visibility" in shape else ()):
        os.makedirs(os.path.join(out_dir, sub), exist_ok=True)
    mask = to_hw(shape["mask"].reshape(1, -1, 1)[0], h, w).cpu().numpy()[..., 0]
    np.save(os.path.join(out_dir, "points", name), to_hw(shape["points"][0], h, w).cpu().numpy().astype(np.float32))
    np.save(os.path.join(out_dir, "normal", name), to_hw(shape["normal"][0], h, w).cpu().numpy().astype(np.float32))
    np.save(os.path.join(out_dir, "mask", name), mask.astype(bool))
    if "visibility" in shape:
        vis = shape["visibility"].cpu().numpy()
        L = vis.shape[0]
        np.save(os.path.join(out_dir, "visibility", name), vis.reshape(L, h, w).transpose(0, 2, 1).astype(np.float32))


def load_shape_view(shape_dir, view_number, with_visibility=False):
    """Read one view back the way the stage-2 dataset does (dataset.py:100-114): row-major pixel order, points / normal [1,N,3],
    surface_mask bool [1,N], visibility [L,N] - the model_input entries PSNetwork.forward consumes next to uv / pose / intrinsics."""
    import os

    import numpy as np
    name = "view_%02d.npy" % view_number
    def load(sub, dtype):  # C-ordered copy: a file written from a transposed view comes back Fortran-ordered
        return torch.from_numpy(np.ascontiguousarray(np.load(os.path.join(shape_dir, sub, name)).astype(dtype)))

    pts = load("points", np.float32)
    out = {"points": pts.view(1, -1, 3), "normal": load("normal", np.float32).view(1, -1, 3),
           "surface_mask": load("mask", bool).view(1, -1), "img_res": list(pts.shape[:2])}
    if with_visibility:
        vis = load("visibility", np.float32)
        out["visibility"] = vis.reshape(vis.shape[0], -1)
    return out
