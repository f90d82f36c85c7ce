"""Entry-point loops of the reference kept as thin host code over the CUDA path:

  render_stage1_view   <- stage1/eval.py:82-119          (one library call per view instead of 256 x 1024-ray chunks)
  extract_shape        <- stage1/shape_extract.py:103-165 (points / normals / mask / per-light visibility of a view)
  render_stage2_view   <- stage2/eval.py:314-417          (all light batches of a view)
  occupancy_grid_logits<- stage1/model/extracting.py:84-96,137-155 (dense -logit grid for mesh export)
  render_envmap_view   <- stage2/eval.py:173-231          (envmap relighting: RGB-intensity light grid, sum over lights)
  extract_and_shade    <- SURVEY.md §8f-1: stage-1 surface extraction + shadow-ray visibility feeding stage-2 shading directly, no .npy hand-off
  save_shape_view / load_shape_view <- shape_extract.py:144-162 / stage2/datasets/dataset.py:100-114: the .npy hand-off, kept as an option
  *_sharded            <- SURVEY.md §8e: rays of a view dealt over the ranks (stage 2: balanced by surface pixels), one all_gather at the end

All functions take / return torch tensors; `view` dicts carry the camera: stage 1 {camera_mat, world_mat}, stage 2
{intrinsics, pose}.  The pixel order conventions of the reference are preserved (stage 1: x-major, undone by to_hw;
stage 2: row-major uv)."""
import functools

import torch

from . import sharding
from .stage1.common import arange_pixels, to_hw


@functools.lru_cache(maxsize=64)
def _shard_pixels(h, w, rank, world, device):
    """[1, n_local, 2] integer pixel positions of one rank's ray tiles, resident on `device` (built once per view size)."""
    return arange_pixels((h, w))[0][:, sharding.shard_indices(h * w, rank, world)].to(device)


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


def _stage2_input_from_shape(shp, p_loc, camera_mat, world_mat):
    """The model_input PSNetwork.forward consumes, built from an extract_shape result instead of the .npy files of
    stage1/shape_extract.py:144-162 (stage2/datasets/dataset.py:100-114 reads them back).  Stage-1 pixel order is x-major while stage 2
    consumes explicit uv pairs, so no re-ordering is needed; stage 1 uses fx for both axes (common.py:220)."""
    K = torch.eye(4).unsqueeze(0)
    cm = camera_mat.detach().float().cpu()
    K[0, 0, 0] = K[0, 1, 1] = cm[0, 0, 0]
    K[0, 0, 2], K[0, 1, 2] = cm[0, 0, 2], cm[0, 1, 2]
    inp = {"intrinsics": K, "uv": p_loc.float(), "pose": world_mat, "object_mask": shp["mask"], "surface_mask": shp["mask"],
           "points": shp["points"], "normal": shp["normal"]}
    if "visibility" in shp:
        inp["visibility"] = shp["visibility"]  # [L, N]: the shadow-ray transmittances stage 2 trains its visibility net against
    return inp


@torch.no_grad()
def extract_and_shade(renderer, ps_model, h, w, camera_mat, world_mat, light_dirs, light_batch=96, shadows=True, pixels=None):
    """the whole per-view chain of the reference in one pass, nothing written to disk - stage-1 surface search
    (512-step ray march + secant), analytic normals, the shadow-ray visibility of every surface point towards every light
    (rendering.py:378-408; `shadows=False` skips it like shape_extract.py without --visibility) and stage-2 shading of the same
    points under the same lights.  Returns (shape dict incl. 'visibility' [L, N], stage-2 result dict).  `pixels` ([1,n,2] integer
    pixel positions) restricts the pass to a subset of the view (ray shards)."""
    dev = next(renderer.model.parameters()).device
    p_loc = (arange_pixels((h, w))[0] if pixels is None else pixels).to(dev)
    light_dirs = light_dirs.to(dev)
    shp = renderer(p_loc, camera_mat, world_mat, None, "shape_extract", visibility=bool(shadows), light_dir=light_dirs if shadows else None)
    inp = _stage2_input_from_shape(shp, p_loc, camera_mat, world_mat)
    return shp, render_stage2_view(ps_model, inp, light_dirs, light_batch)


def relit_row_layout(ps_model, L, shadows=True):
    """Column layout of one packed per-pixel row of a relit view: the stage-2 entries (see _s2_packed_layout) followed by the
    stage-1 mask (1), surface point (3), normal (3) and, with the shadow pass, the L shadow-ray transmittances."""
    layout, width = _s2_packed_layout(ps_model, L)
    return layout, width, width + 7 + (L if shadows else 0)


@torch.no_grad()
def extract_and_shade_rows(renderer, ps_model, h, w, camera_mat, world_mat, light_dirs, pixels, light_batch=96, shadows=True):
    """extract_and_shade over the given pixel positions ([1,n,2]), packed as one fp32 row per pixel (relit_row_layout): the unit a
    rank contributes to the pixel gather."""
    dev = next(renderer.model.parameters()).device
    shp, out = extract_and_shade(renderer, ps_model, h, w, camera_mat, world_mat, light_dirs, light_batch, shadows, pixels=pixels)
    L = light_dirs.shape[0]
    n_loc = int(pixels.shape[1])
    layout, width, total = relit_row_layout(ps_model, L, shadows)
    local = torch.empty(n_loc, total, dtype=torch.float32, device=dev)
    for key, per_light, chans, off, w_ in layout:
        v = out[key]
        local[:, off:off + w_] = (v.reshape(L, n_loc, chans).permute(1, 0, 2) if per_light else v).reshape(n_loc, w_)
    local[:, width] = shp["mask"].reshape(n_loc).float()
    local[:, width + 1:width + 4] = shp["points"].reshape(n_loc, 3)
    local[:, width + 4:width + 7] = shp["normal"].reshape(n_loc, 3)
    if shadows:
        local[:, width + 7:] = shp["visibility"].t()
    return local


def unpack_relit_rows(full, ps_model, L, shadows=True):
    """(shape dict, stage-2 result dict) views of packed rows [N, total] in the shapes extract_and_shade returns."""
    N = full.shape[0]
    layout, width, _ = relit_row_layout(ps_model, L, shadows)
    res = {}
    for key, per_light, chans, off, w_ in layout:
        v = full[:, off:off + w_]
        res[key] = v.reshape(N, L, chans).permute(1, 0, 2) if per_light else v.reshape(1, N, chans)
    shape = {"mask": full[:, width].reshape(1, N) > 0.5, "points": full[:, width + 1:width + 4].reshape(1, N, 3),
             "normal": full[:, width + 4:width + 7].reshape(1, N, 3)}
    if shadows:
        shape["visibility"] = full[:, width + 7:].t()
    return shape, res


@torch.no_grad()
def extract_and_shade_sharded(renderer, ps_model, h, w, camera_mat, world_mat, light_dirs, rank, world, light_batch=96, shadows=True,
                              group=None):
    """extract_and_shade with the rays of the view dealt over the ranks (128-ray tiles round-robin, sharding.shard_indices) and ONE
    all_gather of the packed per-pixel rows at the end: every rank returns the full view."""
    N = h * w
    dev = next(renderer.model.parameters()).device
    local = extract_and_shade_rows(renderer, ps_model, h, w, camera_mat, world_mat, light_dirs, _shard_pixels(h, w, rank, world, dev),
                                   light_batch, shadows)
    full = sharding.gather_pixels(local, N, rank, world, group=group)
    return unpack_relit_rows(full, ps_model, light_dirs.shape[0], shadows)


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


def split_input(model_input, total_pixels, n_pixels=None):
    """Pixel chunks of a stage-2 model input (the role of stage2/utils/general.py:23-37 in the eval loops, stage2/eval.py:345-365).
    The CUDA path shades a whole view per call, so the default is ONE chunk (n_pixels=None); a chunk size is honoured for callers
    that keep the reference's 1024-pixel loop.  Chunks are `narrow` views of the per-pixel entries (no index tensors, no copies);
    camera / light entries are shared by reference."""
    step = int(total_pixels) if not n_pixels else int(n_pixels)
    chunks = []
    for start in range(0, int(total_pixels), max(step, 1)):
        length = min(step, int(total_pixels) - start)
        chunk = {k: v for k, v in model_input.items() if k not in PER_PIXEL_INPUTS}
        chunk.update({k: model_input[k].narrow(1, start, length) for k in PER_PIXEL_INPUTS if k in model_input})
        chunks.append(chunk)
    return chunks


def merge_output(res, total_pixels, batch_size):
    """Stitch per-chunk model outputs back into whole-view tensors in the layout stage2/eval.py consumes (general.py:39-53):
    per-pixel scalars -> [batch * total_pixels]; [..., n, C] entries -> [prod(...) * total_pixels, C] with the pixel axis second
    to last.  Every output is allocated once and the chunks are copied into their pixel range; None entries are dropped."""
    merged = {}
    for key, first in res[0].items():
        if first is None:
            continue
        scalar = first.dim() < 3
        chans = 1 if scalar else first.shape[-1]
        lead = (batch_size,) if scalar else tuple(first.shape[:-2])
        buf = first.new_empty(*lead, total_pixels, chans)
        at = 0
        for r in res:
            part = r[key].reshape(*lead, -1, chans)
            buf.narrow(-2, at, part.shape[-2]).copy_(part)
            at += part.shape[-2]
        merged[key] = buf.reshape(-1) if scalar else buf.reshape(-1, chans)
    return merged


# per-pixel entries of a stage-2 result a sharded view returns: key -> (has a light axis, channels)
_S2_SHARDED_KEYS = (("sg_rgb_values", True, 3), ("sg_specular_rgb_values", True, 3), ("visibility", True, 3),
                    ("normal_pred", False, 3), ("sg_diffuse_albedo_values", False, 3), ("sg_weight", False, None))


def _s2_packed_layout(model, L):
    """Column layout of the packed per-pixel rows a stage-2 shard sends through the gather.  Derived from the MODEL's flags (which
    entries PSNetwork.forward emits: renderer.py:145-152,200-208), never from a rank's own output, so every rank - including one that
    was dealt no pixels - agrees on it."""
    micro = bool(getattr(model, "microfacet", False))  # microfacet: no SG weights, the "specular" entry is the [1,N,3] roughness image
    present = {"sg_rgb_values": True, "sg_specular_rgb_values": True, "sg_diffuse_albedo_values": True, "sg_weight": not micro,
               "visibility": bool(getattr(model, "visibility", False)), "normal_pred": bool(getattr(model, "normal_mlp", False))}
    layout, off = [], 0
    for key, per_light, chans in _S2_SHARDED_KEYS:
        if not present[key]:
            continue
        if key == "sg_specular_rgb_values" and micro:
            per_light = False
        if chans is None:
            chans = int(model.nbasis)
        width = (L if per_light else 1) * chans
        layout.append((key, per_light, chans, off, width))
        off += width
    return layout, off


@torch.no_grad()
def render_stage2_view_sharded(model, model_input, light_dirs, rank, world, light_batch=96, light_intensity=None, group=None):
    """BASELINE config 4: the pixels of one stage-2 view dealt over the ranks with the SURFACE pixels balanced
    (sharding.shard_plan_by_mask: one host copy of the mask per view), every rank shades its share under all lights, ONE all_gather
    returns the full view on every rank.  Output: the per-pixel entries of render_stage2_view in the reference's shapes -
    sg_rgb_values / sg_specular_rgb_values / visibility [L,N,3], normal_pred / sg_diffuse_albedo_values [1,N,3], sg_weight [1,N,27]
    (visibility / normal_pred only when the model emits them, like the unsharded path)."""
    smask = model_input["surface_mask"]
    N = smask.shape[1]
    first = next(model.parameters(), None)
    dev = first.device if first is not None else model_input["points"].device
    plan = sharding.shard_plan_by_mask(smask[0], world)
    idx = plan[rank]
    n_loc = int(idx.numel())
    L = light_dirs.shape[0]
    layout, width = _s2_packed_layout(model, L)
    local = torch.zeros(n_loc, width, dtype=torch.float32, device=dev)
    if n_loc > 0:
        sub = dict(model_input)
        for k in PER_PIXEL_INPUTS:
            v = model_input.get(k)
            if torch.is_tensor(v) and v.dim() >= 2 and v.shape[1] == N:
                sub[k] = torch.index_select(v, 1, idx.to(v.device))
        out = render_stage2_view(model, sub, light_dirs, light_batch, light_intensity)
        for key, per_light, chans, off, w_ in layout:
            v = out[key]
            if per_light:
                v = v.reshape(L, n_loc, chans).permute(1, 0, 2)
            local[:, off:off + w_] = v.reshape(n_loc, w_).to(dev)
    full = sharding.gather_rows(local, N, rank, world, lambda r: plan[r], group=group)
    res = {}
    for key, per_light, chans, off, w_ in layout:
        v = full[:, off:off + w_]
        res[key] = v.reshape(N, L, chans).permute(1, 0, 2).contiguous() if per_light else v.reshape(1, N, chans).contiguous()
    return res


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
