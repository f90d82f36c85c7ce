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
