"""Ray sharding across the GPUs of one box + the single pixel gather that closes a render.

Every ray / pixel / (point, light) pair is independent and the weights (<= 3.2 MB) are replicated, so the data path
needs no collective; the only exchange is one all_gather of the rendered pixel shards.  Backend-agnostic
(`nccl` on GPUs, `gloo` in the CPU tests)."""
import functools

import torch
import torch.distributed as dist


@functools.lru_cache(maxsize=256)
def shard_indices(n_rays, rank, world, tile=128):
    """Ray ids owned by `rank`: tiles of `tile` consecutive rays, dealt in blocks of `world` tiles with the assignment rotated by a
    hash of the block index.  Every rank gets exactly one tile per block (equal ray counts), and the pseudo-random rotation keeps the
    deal from locking onto the image: in the x-major pixel order a 512-pixel column is 4 tiles, so a plain round-robin over 8 ranks
    hands rank r the same quarter of every column - rank 0 the top rows, where no ray hits the object and the shadow / shading passes
    have nothing to do (measured: 1/8 of the rays, 0 surface points); a fixed per-block rotation has the same problem at world = 2.
    Cached: the same (view size, rank, world) recurs on every view."""
    ids = torch.arange(n_rays)
    t = ids // tile
    block = t // world
    rot = ((block * 2654435761) >> 13) % world   # Knuth multiplicative hash of the block index
    owner = (t % world + rot) % world
    return ids[owner == rank]


@functools.lru_cache(maxsize=256)
def _shard_indices_on(n_rays, rank, world, tile, device):
    return shard_indices(n_rays, rank, world, tile).to(device)


@functools.lru_cache(maxsize=64)
def shard_counts(n_rays, world, tile=128):
    return [int(shard_indices(n_rays, r, world, tile).numel()) for r in range(world)]


def shard_plan_by_mask(mask, world, tile=128):
    """Pixel ids of EVERY rank when the work sits on the masked pixels (stage 2: cost grows with the surface points, not with the
    pixels): masked and unmasked pixels are dealt separately, tiles of `tile` round-robin, so every rank shades the
    same number of surface points (to within one tile) whatever the silhouette looks like.  ONE device-to-host copy of the mask per
    view; returns a list of sorted index tensors (host), identical on every rank."""
    m = mask.reshape(-1).bool().cpu()
    ids = torch.arange(m.numel())
    per_rank = [[] for _ in range(world)]
    for sel in (ids[m], ids[~m]):
        owner = (torch.arange(sel.numel()) // tile) % world
        for r in range(world):
            per_rank[r].append(sel[owner == r])
    return [torch.cat(p).sort().values for p in per_rank]


def shard_indices_by_mask(mask, rank, world, tile=128):
    """Pixel ids owned by `rank` (see shard_plan_by_mask; use the plan when more than one rank's list is needed)."""
    return shard_plan_by_mask(mask, world, tile)[rank]


def gather_rows(local, n_rows, rank, world, indices_of, group=None):
    """local: [n_local, C] values of the rows indices_of(rank) -> full [n_rows, C] on every rank.  ONE collective (all_gather on
    equally padded shards), then an index scatter; indices_of(r) must give every rank the same answer for every r."""
    if world == 1:
        return local
    idx = [indices_of(r) for r in range(world)]
    pad = max(int(i.numel()) for i in idx)
    C = local.shape[1]
    buf = local.new_zeros(pad, C)
    buf[: local.shape[0]] = local
    out = local.new_empty(world * pad, C)
    dist.all_gather_into_tensor(out, buf, group=group)
    full = local.new_empty(n_rows, C)
    for r in range(world):
        full[idx[r].to(local.device)] = out[r * pad: r * pad + idx[r].numel()]
    return full


def gather_pixels(local, n_rays, rank, world, tile=128, group=None):
    """local: [n_local, C] rendered values of this rank's rays -> full [n_rays, C] image on every rank.
    One collective (all_gather on equally padded shards), then an index scatter to undo the tile interleave."""
    if world == 1:
        return local
    counts = shard_counts(n_rays, world, tile)
    pad = max(counts)
    C = local.shape[1]
    buf = local.new_zeros(pad, C)
    buf[: local.shape[0]] = local
    out = local.new_empty(world * pad, C)
    dist.all_gather_into_tensor(out, buf, group=group)
    full = local.new_empty(n_rays, C)
    for r in range(world):
        full[_shard_indices_on(n_rays, r, world, tile, local.device)] = out[r * pad: r * pad + counts[r]]
    return full


def allreduce_gradients(module, world, group=None, average=True, extra_params=(), weight=None):
    """Data-parallel train step: ONE all_reduce of the flattened gradients per step
    (668 K floats for the stage-2 networks; latency bound, NVLS on NVSwitch).  Parameters without a gradient contribute zeros
    so that every rank reduces the same layout.
    extra_params: parameters that live outside `module` and must stay identical on every rank - the stage-2 light tables
    (light_para / light_inten_para, stepped by their own SparseAdam, stage2/trainer.py:165); sparse gradients are densified for the
    reduction and handed back dense.
    weight: this rank's share of the loss normalisation, e.g. its number of masked pixels.  The reference's losses are means over
    the masked pixels of a batch, so with unequal shards the single-GPU gradient is sum_r w_r g_r / sum_r w_r, not the plain average;
    the weight travels as one more element of the same all_reduce."""
    if world == 1:
        return
    params = [p for p in module.parameters() if p.requires_grad] + [p for p in extra_params if p.requires_grad]
    if not params:
        return

    def dense_grad(p):
        if p.grad is None:
            return torch.zeros_like(p, dtype=torch.float32)
        g = p.grad.to_dense() if p.grad.is_sparse else p.grad
        return g.float()
    parts = [dense_grad(p).reshape(-1) for p in params]
    w = None if weight is None else torch.as_tensor(float(weight), dtype=torch.float32, device=parts[0].device).reshape(1)
    flat = torch.cat(parts if w is None else [q * w for q in parts] + [w])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if w is not None:
        flat = flat[:-1] / flat[-1].clamp_min(1e-30)
    elif average:
        flat = flat / world
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None or p.grad.is_sparse:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
