"""Fused optimizers of the two train loops: drop-ins for torch.optim.Adam (stage1/train.py:62, stage2/trainer.py:116) and
torch.optim.SparseAdam (stage2/trainer.py:165), same constructor arguments, param_groups (so MultiStepLR drives `lr` as it
does in the reference, train.py:75-77 / trainer.py:121-123) and state_dict layout (state[p] = {step, exp_avg, exp_avg_sq}:
OptimizerParameters/*.pth of a reference run load, and a state saved here loads into the torch classes).

Adam.step() is ONE psn_adam_step launch over every parameter that has a gradient (torch's foreach path issues ~10 kernels per
step, its single-tensor path ~8 per parameter); SparseAdam.step() is one psn_sparse_adam_step launch per embedding, fed with the
un-coalesced indices / values of the sparse gradient.  CUDA fp32 parameters only: there is no CPU fallback."""
import ctypes as C

import torch

from . import _binding as B
from . import engine


def _step_int(s):
    return int(s.item()) if torch.is_tensor(s) else int(s)


def _bump(tensors):
    """The kernels write through raw pointers: tell autograd (and everything keyed on Tensor._version, e.g. the packed-weight cache
    of engine.packed_handles and saved-tensor checks) that these tensors were modified in place."""
    torch.autograd.graph.increment_version(tensors)


def _check_param(p, who):
    if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
        raise RuntimeError("psnerf_b200.optim.%s: parameters must be contiguous fp32 CUDA tensors (got %s %s)"
                           % (who, p.device, p.dtype))


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
        if amsgrad:
            raise RuntimeError("psnerf_b200.optim.Adam: amsgrad is not supported (the reference never enables it)")
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("psnerf_b200.optim.Adam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = B.load()
        for group in self.param_groups:
            if group.get("amsgrad") or group.get("maximize"):
                raise RuntimeError("psnerf_b200.optim.Adam: amsgrad / maximize are not supported")
            by_step = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                _check_param(p, "Adam")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] = _step_int(st["step"]) + 1
                by_step.setdefault((st["step"], p.device.index), []).append((p, engine.f32c(p.grad), st))
            for (step, dev), items in by_step.items():
                arr = (B.AdamTensor * len(items))()
                for i, (p, g, st) in enumerate(items):
                    arr[i] = B.AdamTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel())
                hyper = B.AdamHyper(group["lr"], group["betas"][0], group["betas"][1], group["eps"], group["weight_decay"], step)
                with torch.cuda.device(dev):
                    B.check(lib.psn_adam_step(arr, len(items), C.byref(hyper), engine._stream()), "psn_adam_step")
                _bump([t for p, _, st in items for t in (p, st["exp_avg"], st["exp_avg_sq"])])
        return loss


class SparseAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("psnerf_b200.optim.SparseAdam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = B.load()
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.grad.is_sparse:
                    raise RuntimeError("SparseAdam does not support dense gradients, please consider Adam instead")
                _check_param(p, "SparseAdam")
                if p.dim() != 2:
                    raise RuntimeError("psnerf_b200.optim.SparseAdam: [rows, dim] embedding tables only")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] = _step_int(st["step"]) + 1
                g = p.grad
                rows = g._indices()[0].contiguous()
                vals = engine.f32c(g._values()).reshape(rows.numel(), -1)
                hyper = B.AdamHyper(group["lr"], group["betas"][0], group["betas"][1], group["eps"], 0.0, st["step"])
                with torch.cuda.device(p.device):
                    B.check(lib.psn_sparse_adam_step(p.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.shape[0],
                                                     p.shape[1], rows.data_ptr(), vals.data_ptr(), rows.numel(), C.byref(hyper),
                                                     engine._stream()), "psn_sparse_adam_step")
                _bump([p, st["exp_avg"], st["exp_avg_sq"]])
        return loss
