"""Drop-in for the train-step half of stage1/model/training.py (Trainer.__init__, train_step, process_data_dict, compute_loss:
training.py:20-60,122-198): pixel sampling, ground-truth lookup, the call into Renderer.forward (which runs the differentiable CUDA
path of stage1/train.py in train() mode) and the Loss.  The optimizer is whatever the caller passes - psnerf_b200.optim.Adam for a
fused step.  The visualisation half (render_visdata: PIL / matplotlib image dumps) is outside the hot path and not provided."""
import numpy as np
import torch

from .common import arange_pixels, get_tensor_values, sample_patch_points
from .loss import Loss


class Trainer(object):
    def __init__(self, model, optimizer, cfg_all, device=None, **kwargs):
        cfg = cfg_all["training"]
        self.model = model  # the Renderer
        self.optimizer = optimizer
        self.device = device
        self.n_training_points = cfg["n_training_points"]
        self.n_eval_points = cfg["n_training_points"]
        self.overwrite_visualization = True
        self.normal_loss = cfg.get("normal_loss", False)
        self.normal_after = cfg.get("normal_after", -1)
        self.cfg = cfg
        self.angle = cfg.get("normal_angle", None)
        self.normal_weight_decay = cfg.get("normal_weight_decay", False)
        self.mask_loss = cfg.get("mask_loss", False)
        self.mask_loss_type = cfg.get("mask_loss_type", "acc")
        self.rendering_technique = cfg["type"]
        self.loss = Loss(cfg["lambda_l1_rgb"], cfg["lambda_normals"], cfg.get("lambda_normloss", 1.0), cfg.get("lambda_mask", 1.0),
                         device=device)

    def train_step(self, data, it=None):
        """zero_grad -> compute_loss -> backward -> optimizer.step (training.py:46-60); returns the loss dictionary."""
        self.model.train()
        self.optimizer.zero_grad()
        loss_dict = self.compute_loss(data, it=it)
        loss_dict["loss"].backward()
        self.optimizer.step()
        return loss_dict

    def render_visdata(self, data_loader, it, out_render_path):
        raise NotImplementedError("psnerf_b200: Trainer.render_visdata (image dumps through PIL / matplotlib) is outside the hot path; "
                                  "render with psnerf_b200.pipeline.render_stage1_view")

    def process_data_dict(self, data):
        """(img, mask, world_mat, camera_mat, scale_mat, img_idx, normal, norm_mask, mask_valid) on the device (training.py:122-139)."""
        device = self.device
        img = data.get("img").to(device)
        img_idx = data.get("img.idx")
        batch_size, _, h, w = img.shape
        mask_img = data.get("img.mask", torch.ones(batch_size, h, w)).unsqueeze(1).to(device)
        world_mat = data.get("img.world_mat").to(device)
        camera_mat = data.get("img.camera_mat").to(device)
        scale_mat = data.get("img.scale_mat").to(device)
        normal = data.get("img.normal").to(device) if self.normal_loss else None
        norm_mask = data.get("img.norm_mask").unsqueeze(1).to(device) if self.normal_loss else None
        mask_valid = data.get("img.mask_valid", torch.ones(batch_size, h, w)).unsqueeze(1).to(device)
        return (img, mask_img, world_mat, camera_mat, scale_mat, img_idx, normal, norm_mask, mask_valid)

    def compute_loss(self, data, eval_mode=False, it=None):
        """Sample n_training_points pixels, render them, look the ground truth up at the same pixels, evaluate the Loss
        (training.py:141-198)."""
        n_points = self.n_eval_points if eval_mode else self.n_training_points
        (img, mask_img, world_mat, camera_mat, scale_mat, img_idx, normal, norm_mask, mask_valid) = self.process_data_dict(data)
        device = self.device
        batch_size, _, h, w = img.shape
        assert ((h, w) == mask_img.shape[2:4]) and (n_points > 0)
        if n_points >= h * w:  # every pixel (the reference passes the integer grid on; the lookups below need it as float)
            p = arange_pixels((h, w), batch_size)[0].to(device)
            mask_gt = mask_img.bool().reshape(batch_size, -1).to(torch.float32)
            mask_valid = mask_valid.bool().reshape(batch_size, -1)
            norm_mask_gt = norm_mask.bool() if self.normal_loss else None
            pix = p.float()
        else:
            p, pix = sample_patch_points(batch_size, n_points, patch_size=1.0, image_resolution=(h, w), continuous=False)
            p = pix.to(device)
            pix = pix.to(device)
            mask_gt = get_tensor_values(mask_img, pix.clone()).bool().reshape(batch_size, -1).to(torch.float32)
            mask_valid = get_tensor_values(mask_valid * 1.0, pix.clone()).bool().reshape(batch_size, -1)
            norm_mask_gt = get_tensor_values(norm_mask, pix.clone()).bool().squeeze(-1) if self.normal_loss else None
        out_dict = self.model(p, camera_mat, world_mat, scale_mat, self.rendering_technique, it=it, eval_=eval_mode)
        rgb_gt = get_tensor_values(img, pix.clone())
        normal_gt = None
        if self.normal_loss and it >= self.normal_after:
            normal_gt = get_tensor_values(normal, pix.clone())
            if self.angle is not None:
                norm_mask_gt[normal_gt[..., -1] < np.cos(np.deg2rad(self.angle))] = False
            flip = torch.tensor([[[1, -1, -1]]], dtype=torch.float32).to(device)  # camera-frame normals -> world frame
            normal_gt = torch.einsum("bij,bnj->bni", world_mat[:, :3, :3] * flip, normal_gt)
        mask_pred = out_dict.get("acc_map", None)
        if not self.mask_loss:
            mask_gt = None
        return self.loss(out_dict, rgb_gt, normal_gt, norm_mask_gt, mask_pred, mask_gt, mask_valid)
