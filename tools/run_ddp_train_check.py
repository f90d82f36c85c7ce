"""torchrun check of the data-parallel stage-2 train step: each rank differentiates its shard of the in-mask pixels, ONE
all_reduce averages the gradients; the result must equal the single-GPU gradient of the whole batch.
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/run_ddp_train_check.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psnerf_b200 import sharding, synth  # noqa: E402
from psnerf_b200.stage2 import PSNetwork  # noqa: E402
from psnerf_b200.stage2.loss import MainLoss, NormalLoss  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
conf = synth.stage2_conf()
torch.manual_seed(0)
m = PSNetwork(conf).to(dev).train()
m.precision = "tc"
n_px, L = 2048, 12
full = synth.stage2_input(32, 64, L, all_surface=True, seed=4)
gen = torch.Generator().manual_seed(1)
gt_rgb = torch.rand(L, n_px, 3, generator=gen)
vt_gt = torch.rand(3, n_px, generator=gen)
noise = torch.randn(n_px, 3, generator=gen)
lm, ln = MainLoss(1.0, "L1", 0.05, 0.01, 1.0), NormalLoss(1.0, 0.05)


def grads(idx):
    inp = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in full.items()}
    for k in ("uv", "object_mask", "surface_mask", "points", "normal"):
        inp[k] = inp[k][:, idx.to(dev)]
    inp["light_vis_train"] = synth.lights(3, seed=5).to(dev)
    inp["vis_train_gt"] = vt_gt[:, idx].to(dev)
    inp["visibility"] = torch.zeros(L, idx.numel(), device=dev)
    m.zero_grad(set_to_none=True)
    out = m(inp, noise={"xyz": noise[idx]})
    loss = lm(out, {"rgb": gt_rgb[:, idx]}, inp)["loss"] + ln(out)["loss"]
    loss.backward()
    return float(loss)


idx = sharding.shard_indices(n_px, rank, world)
loss_local = grads(idx)
sharding.allreduce_gradients(m, world)
dp = [p.grad.clone() for p in m.parameters() if p.requires_grad]
grads(torch.arange(n_px))
ref = [p.grad.clone() for p in m.parameters() if p.requires_grad]
err = max(float((a - b).abs().max() / (b.abs().max() + 1e-12)) for a, b in zip(dp, ref))
t = torch.tensor([err], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("data-parallel vs single-GPU gradients, max relative-to-max error over %d ranks: %.3e" % (world, float(t)))
    assert float(t) < 2e-3
    print("OK")
dist.destroy_process_group()
