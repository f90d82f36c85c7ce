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
        self.mask_loss = nn.BCELoss(reduction="mean")
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
        if normal is not None and normal_gt is not None and norm_mask.sum() > 0:
            nl = self.l1_loss(normal[norm_mask], normal_gt.to(dev)[norm_mask]) / float(normal[norm_mask].shape[0])
            loss = loss + self.norm_weight * nl
            terms["normal_loss"] = nl
        if mask is not None and mask_gt is not None:
            lm = self.mask_loss(mask[mask_valid].clamp(0, 1), mask_gt.to(dev)[mask_valid])
            loss = loss + self.mask_weight * lm
            terms["mask_loss"] = lm
        terms["loss"] = loss
        return terms
