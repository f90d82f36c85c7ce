"""world_size-2 gloo test of the multi-GPU host logic: ray sharding + the single pixel gather (CPU only)."""
import os
import socket

import numpy as np
import pytest
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
    """PSNetwork stand-in: per-pixel outputs in the reference's shapes as functions of uv, mask and light direction.  Carries the
    flags the sharded path derives its packed layout from (visibility / normal_mlp / microfacet / nbasis)."""

    def __init__(self, visibility=True, normal_mlp=True):
        super().__init__()
        self.visibility, self.normal_mlp, self.microfacet, self.nbasis = visibility, normal_mlp, False, 9
        self.calls = 0

    def forward(self, inp):
        self.calls += 1
        uv, sm, l = inp["uv"][0], inp["surface_mask"][0], inp["light_direction"]
        L, n = l.shape[0], uv.shape[0]
        assert n > 0, "a rank without pixels must not call the model"
        base = (uv[:, :1] * 0.01 + uv[:, 1:] * 0.02)[None] + l[:, None, :1]          # [L, n, 1]
        rgb = torch.where(sm[None, :, None], base.expand(L, n, 3) * torch.tensor([1.0, 2.0, 3.0]), torch.ones(L, n, 3))
        vis = torch.where(sm[None, :, None], (base * 0.5).expand(L, n, 3), torch.ones(L, n, 3))
        nrm = torch.where(sm[:, None], torch.stack([uv[:, 0], uv[:, 1], uv.sum(-1)], -1), torch.ones(n, 3))[None]
        out = {"sg_rgb_values": rgb, "sg_specular_rgb_values": rgb * 0.125, "sg_diffuse_albedo_values": nrm * 0.25,
               "sg_weight": (uv.sum(-1, keepdim=True) * torch.arange(1.0, 10.0))[None]}
        if self.visibility:
            out["visibility"] = vis
        if self.normal_mlp:
            out["normal_pred"] = nrm
        return out


def _s2_input(n_side=24):
    g = torch.Generator().manual_seed(1)
    ys, xs = torch.meshgrid(torch.arange(n_side), torch.arange(n_side), indexing="ij")
    uv = torch.stack([xs, ys], -1).reshape(1, -1, 2).float()
    n = n_side * n_side
    sm = (torch.rand(1, n, generator=g) > 0.6)
    return {"uv": uv, "surface_mask": sm, "object_mask": sm.clone(), "points": torch.randn(1, n, 3, generator=g),
            "normal": torch.randn(1, n, 3, generator=g), "intrinsics": torch.eye(4)[None], "pose": torch.eye(4)[None]}


S2_CASES = {  # name: (image side, model flags) - "tiny": 64 pixels < world * tile, so rank 1 is dealt no pixels at all
    "full": (24, dict()),
    "no_vis_no_normal": (24, dict(visibility=False, normal_mlp=False)),
    "tiny": (8, dict()),
}


def _s2_worker(rank, world, port, q, case):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from psnerf_b200 import pipeline
    side, flags = S2_CASES[case]
    lights = torch.nn.functional.normalize(torch.randn(7, 3, generator=torch.Generator().manual_seed(2)), dim=-1)
    out = pipeline.render_stage2_view_sharded(_StubPS(**flags), _s2_input(side), lights, rank, world, light_batch=3)
    q.put((rank, {k: v.numpy() for k, v in out.items()}))
    dist.destroy_process_group()


@pytest.mark.parametrize("case", list(S2_CASES))
def test_stage2_view_sharded_world2_gloo(case):
    """BASELINE config 4 host logic: surface-balanced pixel shards, all light batches per rank, one all_gather - equal to the
    unsharded render of the same (stub) model on every rank; the entries follow the model's flags like the unsharded path, and a
    rank that was dealt no pixels takes part in the gather without calling the model."""
    from psnerf_b200 import pipeline
    world = 2
    side, flags = S2_CASES[case]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_s2_worker, args=(r, world, port, q, case)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lights = torch.nn.functional.normalize(torch.randn(7, 3, generator=torch.Generator().manual_seed(2)), dim=-1)
    want = pipeline.render_stage2_view(_StubPS(**flags), _s2_input(side), lights, light_batch=3)
    single = pipeline.render_stage2_view_sharded(_StubPS(**flags), _s2_input(side), lights, 0, 1, light_batch=3)
    keys = ["sg_rgb_values", "sg_specular_rgb_values", "sg_diffuse_albedo_values", "sg_weight"]
    keys += ["visibility"] if flags.get("visibility", True) else []
    keys += ["normal_pred"] if flags.get("normal_mlp", True) else []
    assert sorted(single.keys()) == sorted(keys)
    for k in keys:
        assert torch.equal(single[k], want[k].reshape(single[k].shape)), k
        for r in range(world):
            assert sorted(res[r].keys()) == sorted(keys)
            assert torch.equal(torch.from_numpy(res[r][k]), want[k].reshape(single[k].shape)), (k, r)


class _StubRenderer(torch.nn.Module):
    """Stage-1 Renderer stand-in for the relit-view chain: 'shape_extract' outputs as functions of the pixel positions."""

    def __init__(self):
        super().__init__()
        self.model = torch.nn.Linear(1, 1)

    def forward(self, pixels, camera_mat, world_mat, scale_mat, technique, visibility=False, light_dir=None):
        assert technique == "shape_extract"
        p = pixels[0].float()
        n = p.shape[0]
        mask = ((p[:, 0] * 7 + p[:, 1] * 3) % 5) < 2
        pts = torch.stack([p[:, 0] * 0.01, p[:, 1] * 0.02, p.sum(-1) * 0.005], -1)
        nrm = torch.nn.functional.normalize(pts + 0.1, dim=-1) * mask[:, None]
        out = {"mask": mask.reshape(1, n), "points": pts.reshape(1, n, 3), "normal": nrm.reshape(1, n, 3)}
        if visibility:
            out["visibility"] = torch.where(mask[None], (p[:, 0] * 0.001)[None] + light_dir[:, :1], torch.ones(light_dir.shape[0], n))
        return out


def _relit_worker(rank, world, port, q, shadows):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from psnerf_b200 import pipeline
    lights = torch.nn.functional.normalize(torch.randn(4, 3, generator=torch.Generator().manual_seed(5)), dim=-1)
    shp, out = pipeline.extract_and_shade_sharded(_StubRenderer(), _StubPS(), 20, 26, torch.eye(4)[None], torch.eye(4)[None], lights, rank, world,
                                                  light_batch=3, shadows=shadows)
    q.put((rank, {k: v.numpy() for k, v in shp.items()}, {k: v.numpy() for k, v in out.items()}))
    dist.destroy_process_group()


@pytest.mark.parametrize("shadows", [True, False])
def test_relit_view_sharded_world2_gloo(shadows):
    """The headline chain with the rays of one view dealt over two ranks (pipeline.extract_and_shade_sharded): packed per-pixel rows,
    ONE all_gather, every rank ends up with the unsharded result (stub renderer / stub PSNetwork: host logic only)."""
    from psnerf_b200 import pipeline
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_relit_worker, args=(r, world, port, q, shadows)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, shp, out = q.get(timeout=120)
        res[r] = (shp, out)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lights = torch.nn.functional.normalize(torch.randn(4, 3, generator=torch.Generator().manual_seed(5)), dim=-1)
    shp0, out0 = pipeline.extract_and_shade(_StubRenderer(), _StubPS(), 20, 26, torch.eye(4)[None], torch.eye(4)[None], lights, light_batch=3,
                                            shadows=shadows)
    assert ("visibility" in shp0) == shadows
    for r in range(world):
        shp, out = res[r]
        assert sorted(shp) == sorted(shp0)
        for k in shp0:
            assert torch.equal(torch.from_numpy(shp[k]), shp0[k].reshape(shp[k].shape)), (k, r)
        for k in ("sg_rgb_values", "sg_specular_rgb_values", "visibility", "normal_pred", "sg_diffuse_albedo_values", "sg_weight"):
            assert torch.equal(torch.from_numpy(out[k]), out0[k].reshape(out[k].shape)), (k, r)


def _weighted_grad_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    m = torch.nn.Linear(3, 2)
    table = torch.nn.Embedding(6, 3, sparse=True)          # a light table outside the model, sparse gradient (trainer.py:165)
    n = 3 if rank == 0 else 7                               # unequal shards of the masked pixels
    x = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3) / 10 + rank
    idx = torch.tensor([rank, 4])
    loss = m(x).pow(2).mean() + (table(idx).sum(0) * x).pow(2).mean()      # means over this rank's pixels, like the reference losses
    loss.backward()
    sharding.allreduce_gradients(m, world, extra_params=list(table.parameters()), weight=n)
    q.put((rank, [p.grad.clone().numpy() for p in m.parameters()] + [table.weight.grad.clone().numpy()]))
    dist.destroy_process_group()


def test_allreduce_gradients_weighted_with_light_tables_gloo():
    """Unequal shards + a sparse light table: the weighted reduction equals the gradient of ONE mean over all pixels of both ranks."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_weighted_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    m = torch.nn.Linear(3, 2)
    table = torch.nn.Embedding(6, 3)
    total = 0.0
    for rank, n in ((0, 3), (1, 7)):
        x = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3) / 10 + rank
        idx = torch.tensor([rank, 4])
        total = total + n * (m(x).pow(2).mean() + (table(idx).sum(0) * x).pow(2).mean())
    (total / 10).backward()
    want = [p.grad for p in m.parameters()] + [table.weight.grad]
    for r in range(world):
        assert not any(np.isnan(g).any() for g in res[r])
        for got, w in zip(res[r], want):
            assert torch.allclose(torch.from_numpy(got), w, atol=1e-6), r
