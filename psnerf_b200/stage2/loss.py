"""Host-side (PyTorch, device-agnostic) restatement of the stage-2 losses, kept in PyTorch like the reference
(stage2/model/loss.py:6-141; the reference hard-codes .cuda() at :30,61)."""
import torch
import torch.nn.functional as F
from torch import nn


class MainLoss(nn.Module):
    def __init__(self, sg_rgb_weight, loss_type="L1", albedo_smooth_weight=0, rough_smooth_weight=0, vis_weight=1.0):
        super().__init__()
        self.sg_rgb_weight, self.albedo_smooth_weight = sg_rgb_weight, albedo_smooth_weight
        self.rough_smooth_weight, self.vis_weight = rough_smooth_weight, vis_weight
        if loss_type not in ("L1", "L2"):
            raise Exception("Unknown loss_type!")
        self.img_loss = nn.L1Loss(reduction="mean") if loss_type == "L1" else nn.MSELoss(reduction="mean")
        self.smooth_loss = nn.L1Loss(reduction="mean")

    def _masked(self, fn, a, b, mask, last=None):
        """fn(a[m], b[m]) (mean reduction) without the boolean gather: sum of the masked element losses / their count.  No
        device -> host synchronisation (the reference's `mask.sum() == 0` test and `a[m]` each cost one per term and step); an empty
        mask gives 0 with zero gradients, like the reference's early return."""
        m = mask.expand(a.shape[0], -1)
        d = a - b.to(a.dtype)
        e = d.abs() if isinstance(fn, nn.L1Loss) else d * d
        per = 1
        if e.dim() == m.dim() + 1:
            m, per = m.unsqueeze(-1), e.shape[-1]
        return torch.where(m, e, torch.zeros((), device=e.device, dtype=e.dtype)).sum() / (m.sum() * per).clamp_min(1).to(e.dtype)

    def forward(self, model_outputs, ground_truth, model_input=None):
        dev = model_outputs["sg_rgb_values"].device
        rgb_gt = ground_truth["rgb"].to(dev)
        mask = model_outputs["network_object_mask"].to(dev) & model_outputs["object_mask"].to(dev)
        sg_rgb_loss = self._masked(self.img_loss, model_outputs["sg_rgb_values"], rgb_gt, mask, (-1, 3))
        loss = self.sg_rgb_weight * sg_rgb_loss
        albedo_loss = rough_loss = None
        if "albedo_jitter" in model_outputs and self.albedo_smooth_weight > 0:
            albedo_loss = self._masked(self.smooth_loss, model_outputs["albedo_values"], model_outputs["albedo_jitter"], mask)
            loss = loss + self.albedo_smooth_weight * albedo_loss
        if "rough_jitter" in model_outputs and self.rough_smooth_weight > 0:
            rough_loss = self._masked(self.smooth_loss, model_outputs["rough_values"], model_outputs["rough_jitter"], mask)
            loss = loss + self.rough_smooth_weight * rough_loss
        lterm = {"sg_rgb_loss": sg_rgb_loss, "albedo_smooth_loss": albedo_loss, "rough_smooth_loss": rough_loss}
        if model_input is not None and "visibility" in model_input and "visibility" in model_outputs:  # loss.py:81
            if "vis_train_gt" in model_input and "vis_train" in model_outputs:
                vis_loss = self._masked(self.img_loss, model_outputs["vis_train"][..., 0], model_input["vis_train_gt"].to(dev), mask, (-1,))
            elif "light_vis_train" in model_input and "vis_train" in model_outputs:
                vis_loss = self._masked(self.img_loss, model_outputs["vis_train"][..., 0], model_input["visibility"].to(dev), mask, (-1,))
            else:
                vis_loss = self._masked(self.img_loss, model_outputs["visibility"][..., 0], model_input["visibility"].to(dev), mask, (-1,))
            loss = loss + self.vis_weight * vis_loss
            lterm["vis_loss"] = vis_loss
        lterm["loss"] = loss
        return lterm


class NormalLoss(nn.Module):
    def __init__(self, normal_weight, normal_smooth_weight=0):
        super().__init__()
        self.normal_weight, self.normal_smooth_weight = normal_weight, normal_smooth_weight

    def forward(self, model_outputs):
        dev = model_outputs["normal_pred"].device
        gt = F.normalize(model_outputs["normal_values"].to(dev), dim=-1)
        mask = model_outputs["network_object_mask"].to(dev) & model_outputs["object_mask"].to(dev)
        m3 = mask.unsqueeze(-1)
        cnt = (mask.sum() * 3).clamp_min(1).to(gt.dtype)
        zero = torch.zeros((), device=dev, dtype=gt.dtype)
        d = model_outputs["normal_pred"] - gt
        norm_loss = torch.where(m3, d * d, zero).sum() / cnt   # = F.mse_loss(pred[mask], gt[mask]) without the gather / host sync
        loss = self.normal_weight * norm_loss
        smooth = None
        if "normal_jitter" in model_outputs and self.normal_smooth_weight > 0:
            smooth = torch.where(m3, (model_outputs["normal_pred"] - model_outputs["normal_jitter"]).abs(), zero).sum() / cnt
            loss = loss + self.normal_smooth_weight * smooth
        return {"loss": loss, "normal_loss": norm_loss, "normal_smooth_loss": smooth}
