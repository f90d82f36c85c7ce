"""Host helpers of the stage-1 entry points (stage1/model/common.py:55-93)."""
import torch


def arange_pixels(resolution=(128, 128), batch_size=1, image_range=(-1.0, 1.0)):
    """Integer pixel grid in x-major order and its copy scaled to image_range (common.py:55-93).
    Returns (pixel_locations long[B,N,2], pixel_scaled float[B,N,2])."""
    h, w = resolution
    gx, gy = torch.meshgrid(torch.arange(0, w), torch.arange(0, h), indexing="ij")
    loc = torch.stack([gx, gy], dim=-1).long().view(1, -1, 2).repeat(batch_size, 1, 1)
    sc = loc.clone().float()
    span = image_range[1] - image_range[0]
    sc[:, :, 0] = span * sc[:, :, 0] / (w - 1) - span / 2
    sc[:, :, 1] = span * sc[:, :, 1] / (h - 1) - span / 2
    return loc, sc


def to_hw(x, h, w):
    """Undo the x-major pixel order of arange_pixels (stage1/eval.py:22)."""
    return x.reshape(w, h, -1).permute(1, 0, 2)
