"""A/B of the optimizer step on cuda:0: psnerf_b200.optim.Adam (one psn_adam_step launch) against torch.optim.Adam (foreach,
the default on CUDA, and fused=True) on the parameter sets of the two train loops, and SparseAdam on a [1920,3] light table.
   python tools/time_optim.py [--out gpurun_out/optim_ab.json]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psnerf_b200 import optim, synth  # noqa: E402
from psnerf_b200.stage1 import NeuralNetwork  # noqa: E402
from psnerf_b200.stage2 import PSNetwork  # noqa: E402


def time_steps(opt, reps=200):
    for _ in range(10):
        opt.step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        opt.step()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3  # us per step, device time of a back-to-back loop (host-bound when launches dominate)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "optim_ab.json"))
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    res = {}
    torch.manual_seed(0)
    models = {"stage1_field_802490": NeuralNetwork(synth.stage1_cfg()).to(dev), "stage2_psnetwork": PSNetwork(synth.stage2_conf()).to(dev)}
    for name, m in models.items():
        params = [p for p in m.parameters() if p.requires_grad]
        for p in params:
            p.grad = torch.randn_like(p) * 1e-3
        row = {"tensors": len(params), "values": sum(p.numel() for p in params)}
        row["psn_adam_us"] = time_steps(optim.Adam(params, lr=1e-7))
        row["torch_foreach_us"] = time_steps(torch.optim.Adam(params, lr=1e-7))
        row["torch_fused_us"] = time_steps(torch.optim.Adam(params, lr=1e-7, fused=True))
        row["torch_single_tensor_us"] = time_steps(torch.optim.Adam(params, lr=1e-7, foreach=False), reps=50)
        res[name] = row
    tab = torch.nn.Embedding(1920, 3, sparse=True).to(dev)
    idx = torch.arange(96, device=dev) + 96 * 7
    torch.nn.functional.normalize(tab(idx), dim=-1).sum().backward()
    row = {"rows": 1920, "touched": 96}
    row["psn_sparse_adam_us"] = time_steps(optim.SparseAdam(list(tab.parameters()), lr=1e-7))
    row["torch_sparse_adam_us"] = time_steps(torch.optim.SparseAdam(list(tab.parameters()), lr=1e-7))
    res["light_table_1920x3"] = row
    res["gpu"] = torch.cuda.get_device_name(0)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
