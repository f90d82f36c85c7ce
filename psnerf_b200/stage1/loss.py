"""Host-side (PyTorch, device-agnostic) restatement of the stage-1 training loss, kept in PyTorch like the reference
(stage1/model/losses.py:6-70): L1 rgb (sum / n_rays), normal-consistency mean, masked L1 normal (sum / n_masked), BCE mask."""
import torch
from torch import nn


class Loss(nn.Module):
    def __init__(self, full_weight, grad_weight, norm_weight=1.0, mask_weight=1.0, device=None):
        super().__init__()
        self.full_weight, self.grad_weight = full_weight, grad_weight
        self.norm_weight, self.mask_weight = norm_weight, mask_weight
        self.l1_loss = nn.L1Loss(reduction="sum")
        self.device = device

    def forward(self, out_dict, rgb_gt, normal_gt=None, norm_mask=None, mask=None, mask_gt=None, mask_valid=None):
        rgb_pred, diff_norm, normal = out_dict["rgb"], out_dict["diff_norm"], out_dict.get("normal_pred", None)
        dev = rgb_pred.device
        rgb_gt = rgb_gt.to(dev)
        zero = torch.zeros((), device=dev)
        rgb_full = self.l1_loss(rgb_pred, rgb_gt) / float(rgb_pred.shape[1]) if self.full_weight != 0.0 else zero
        grad_loss = diff_norm.mean() if (diff_norm is not None and diff_norm.shape[0] > 0 and self.grad_weight != 0.0) else zero
        loss = self.full_weight * rgb_full + self.grad_weight * grad_loss
        terms = {"fullrgb_loss": rgb_full, "grad_loss": grad_loss}
        # Masked terms as where-sums over their counts instead of boolean gathers: no device -> host synchronisation per step (the
        # reference's `norm_mask.sum() > 0` and `x[mask]` cost one each); an empty mask contributes 0 with zero gradients.
        if normal is not None and normal_gt is not None:
            nm = norm_mask.to(dev).unsqueeze(-1)
            nl = torch.where(nm, (normal - normal_gt.to(dev)).abs(), zero).sum() / nm.sum().clamp_min(1).to(normal.dtype)
            loss = loss + self.norm_weight * nl
            terms["normal_loss"] = nl
        if mask is not None and mask_gt is not None:
            mv = mask_valid.to(dev)
            bce = torch.nn.functional.binary_cross_entropy(mask.clamp(0, 1), mask_gt.to(dev).to(mask.dtype), reduction="none")
            lm = torch.where(mv, bce, zero).sum() / mv.sum().clamp_min(1).to(bce.dtype)
            loss = loss + self.mask_weight * lm
            terms["mask_loss"] = lm
        terms["loss"] = loss
        return terms
