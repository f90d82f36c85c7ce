"""Chunking helpers kept for entry-point compatibility (stage2/utils/general.py:23-53).  The CUDA path renders a
whole view per call, so eval loops may pass n_pixels=total_pixels; the 1024-pixel default is the reference's."""
import torch


def split_input(model_input, total_pixels, n_pixels=1024):
    dev = model_input["uv"].device
    out = []
    for idx in torch.split(torch.arange(total_pixels, device=dev), n_pixels, dim=0):
        d = dict(model_input)
        for k in ["uv", "object_mask", "gt_normal", "normal", "depth", "points", "surface_mask", "visibility"]:
            if k in model_input:
                d[k] = torch.index_select(model_input[k], 1, idx)
        out.append(d)
    return out


def merge_output(res, total_pixels, batch_size):
    merged = {}
    for k in res[0]:
        if res[0][k] is None:
            continue
        if len(res[0][k].shape) < 3:
            merged[k] = torch.cat([r[k].reshape(batch_size, -1, 1) for r in res], 1).reshape(batch_size * total_pixels)
        else:
            merged[k] = torch.cat([r[k].reshape(*r[k].shape[:-2], -1, r[k].shape[-1]) for r in res], -2
                                  ).reshape(-1, res[0][k].shape[-1])
    return merged
