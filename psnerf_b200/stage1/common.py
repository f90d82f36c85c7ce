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


def sample_patch_points(batch_size, n_points, patch_size=1, image_resolution=(128, 128), sensor_size=((-1, 1), (-1, 1)),
                        continuous=True):
    """Random pixel positions of a training batch (common.py:9-53): returns (p scaled to the sensor range, pix = raw positions),
    both [B, n_points, 2] as (x, y).  continuous=False draws integer pixels: x from randint(0, W), then y from randint(0, H) -
    the same two generator calls in the same order as the reference, so a seeded run samples the same pixels."""
    assert patch_size > 0
    n = int(n_points)
    h, w = image_resolution
    if continuous:
        p = torch.rand(batch_size, n, 2)
    else:
        px = torch.randint(0, w, size=(batch_size, n, 1)).float()
        py = torch.randint(0, h, size=(batch_size, n, 1)).float()
        p = torch.cat([px, py], dim=-1)
    p = p.view(batch_size, -1, 2)
    pix = p.clone()
    (y0, y1), (x0, x1) = sensor_size
    p[:, :, 0] *= (x1 - x0) / (w - 1)
    p[:, :, 1] *= (y1 - y0) / (h - 1)
    p[:, :, 0] += x0
    p[:, :, 1] += y0
    lo, hi = min(x0, y0), max(x1, y1)
    assert p.max() <= hi and p.min() >= lo and pix.max() < max(image_resolution) and pix.min() >= 0
    return p, pix


def get_tensor_values(tensor, pe, grid_sample=True, mode="nearest", with_mask=False, squeeze_channel_dim=False):
    """Values of tensor [B,C,H,W] at the pixel positions pe [B,N,2] -> [B,N,C] (common.py:172-203).  The reference normalises
    x by W and y by H (not W-1 / H-1) before an align_corners=True nearest lookup, i.e. it reads pixel round(x (W-1)/W): kept."""
    _, _, h, w = tensor.shape
    p = pe.clone().detach().to(tensor.dtype if tensor.dtype.is_floating_point else torch.float32)
    p[:, :, 0] = 2.0 * p[:, :, 0] / w - 1
    p[:, :, 1] = 2.0 * p[:, :, 1] / h - 1
    values = torch.nn.functional.grid_sample(tensor, p.unsqueeze(1), mode=mode, align_corners=True)
    values = values.squeeze(2).detach().permute(0, 2, 1)
    return values.squeeze(-1) if squeeze_channel_dim else values
