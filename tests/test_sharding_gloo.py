"""world_size-2 gloo test of the multi-GPU host logic: ray sharding + the single pixel gather (CPU only)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from psnerf_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rays, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = sharding.shard_indices(n_rays, rank, world)
    # a fake "render": pixel value = f(ray id); every rank must end up with the identical full image
    local = torch.stack([idx.float() * 2.0, idx.float() + 0.5, -idx.float()], -1)
    full = sharding.gather_pixels(local, n_rays, rank, world)
    q.put((rank, full.numpy()))  # plain ndarray: a torch tensor travels as a shared-memory handle that dies with this process
    dist.destroy_process_group()


def test_shards_partition_the_rays():
    for n, w in [(1000, 2), (262144, 8), (129, 4), (5, 8)]:
        parts = [sharding.shard_indices(n, r, w) for r in range(w)]
        allidx = torch.cat(parts).sort().values
        assert torch.equal(allidx, torch.arange(n))
        assert sum(sharding.shard_counts(n, w)) == n
        if n >= 128 * w * 8:
            c = sharding.shard_counts(n, w)
            assert max(c) - min(c) <= 128


def test_gather_pixels_world2_gloo():
    n_rays, world = 1000, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_rays, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = torch.arange(n_rays).float()
    want = torch.stack([ids * 2.0, ids + 0.5, -ids], -1)
    for r in range(world):
        assert torch.equal(torch.from_numpy(res[r]), want)


def _grad_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 2))
    x = torch.arange(20, dtype=torch.float32).reshape(4, 5) / 10 + rank      # each rank sees a different ray shard
    m(x).sum().backward()
    if rank == 1:
        m[2].bias.grad = None                                                # a parameter that got no gradient on this rank
    sharding.allreduce_gradients(m, world)
    q.put((rank, [p.grad.clone().numpy() for p in m.parameters()]))
    dist.destroy_process_group()


def test_allreduce_gradients_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # reference: average of the two ranks' gradients computed serially
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 2))
    acc = None
    for rank in range(world):
        m.zero_grad()
        x = torch.arange(20, dtype=torch.float32).reshape(4, 5) / 10 + rank
        m(x).sum().backward()
        gs = [p.grad.clone() for p in m.parameters()]
        if rank == 1:
            gs[3] = torch.zeros_like(gs[3])
        acc = gs if acc is None else [a + b for a, b in zip(acc, gs)]
    want = [a / world for a in acc]
    for r in range(world):
        for got, w in zip(res[r], want):
            assert torch.allclose(torch.from_numpy(got), w, atol=1e-6)


def test_mask_balanced_shards():
    """Stage-2 sharding: every rank owns the same number of surface pixels (within one tile) for a lopsided silhouette."""
    g = torch.Generator().manual_seed(0)
    mask = torch.zeros(64, 64, dtype=torch.bool)
    mask[5:30, 40:62] = torch.rand(25, 22, generator=g) > 0.2   # the object sits in one corner of the image
    for world in (2, 4, 8):
        parts = [sharding.shard_indices_by_mask(mask, r, world, tile=16) for r in range(world)]
        assert torch.equal(torch.cat(parts).sort().values, torch.arange(64 * 64))
        on = [int(mask.reshape(-1)[p].sum()) for p in parts]
        assert max(on) - min(on) <= 16 and max(p.numel() for p in parts) - min(p.numel() for p in parts) <= 32
        assert all(torch.equal(p, p.sort().values) for p in parts)


class _StubPS(torch.nn.Module):
    """PSNetwork stand-in: per-pixel outputs in the reference's shapes as functions of uv, mask and light direction."""

    def forward(self, inp):
        uv, sm, l = inp["uv"][0], inp["surface_mask"][0], inp["light_direction"]
        L, n = l.shape[0], uv.shape[0]
        base = (uv[:, :1] * 0.01 + uv[:, 1:] * 0.02)[None] + l[:, None, :1]          # [L, n, 1]
        rgb = torch.where(sm[None, :, None], base.expand(L, n, 3) * torch.tensor([1.0, 2.0, 3.0]), torch.ones(L, n, 3))
        vis = torch.where(sm[None, :, None], (base * 0.5).expand(L, n, 3), torch.ones(L, n, 3))
        nrm = torch.where(sm[:, None], torch.stack([uv[:, 0], uv[:, 1], uv.sum(-1)], -1), torch.ones(n, 3))[None]
        return {"sg_rgb_values": rgb, "visibility": vis, "normal_pred": nrm, "sg_diffuse_albedo_values": nrm * 0.25}


def _s2_input(n_side=24):
    g = torch.Generator().manual_seed(1)
    ys, xs = torch.meshgrid(torch.arange(n_side), torch.arange(n_side), indexing="ij")
    uv = torch.stack([xs, ys], -1).reshape(1, -1, 2).float()
    n = n_side * n_side
    sm = (torch.rand(1, n, generator=g) > 0.6)
    return {"uv": uv, "surface_mask": sm, "object_mask": sm.clone(), "points": torch.randn(1, n, 3, generator=g),
            "normal": torch.randn(1, n, 3, generator=g), "intrinsics": torch.eye(4)[None], "pose": torch.eye(4)[None]}


def _s2_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from psnerf_b200 import pipeline
    lights = torch.nn.functional.normalize(torch.randn(7, 3, generator=torch.Generator().manual_seed(2)), dim=-1)
    out = pipeline.render_stage2_view_sharded(_StubPS(), _s2_input(), lights, rank, world, light_batch=3)
    q.put((rank, {k: v.numpy() for k, v in out.items()}))
    dist.destroy_process_group()


def test_stage2_view_sharded_world2_gloo():
    """BASELINE config 4 host logic: surface-balanced pixel shards, all light batches per rank, one all_gather - equal to the
    unsharded render of the same (stub) model on every rank."""
    from psnerf_b200 import pipeline
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_s2_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lights = torch.nn.functional.normalize(torch.randn(7, 3, generator=torch.Generator().manual_seed(2)), dim=-1)
    want = pipeline.render_stage2_view(_StubPS(), _s2_input(), lights, light_batch=3)
    single = pipeline.render_stage2_view_sharded(_StubPS(), _s2_input(), lights, 0, 1, light_batch=3)
    for k in ("sg_rgb_values", "visibility", "normal_pred", "sg_diffuse_albedo_values"):
        assert torch.equal(single[k], want[k].reshape(single[k].shape)), k
        for r in range(world):
            assert torch.equal(torch.from_numpy(res[r][k]), want[k].reshape(single[k].shape)), (k, r)
