"""Import the REAL reference hot-path modules (read-only, /root/reference) for fixture generation.

Only usable in the authoring container (the GPU box has no /root/reference).  No reference
source is modified or copied: stage-1 modules are loaded by file path under a synthetic package
(their own ``model/__init__`` pulls matplotlib via training.py), stage-2 ``utils.rend_util`` is
replaced by a device-agnostic stand-in because the original downloads a plugin at import time
and hard-codes ``.cuda()``.
"""
import importlib.util
import os
import sys
import types
import warnings

import torch
import torch.nn.functional as F

REF = os.environ.get("PSNERF_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "stage1", "model"))


def _load(modname, path, package=None):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    if package:
        mod.__package__ = package
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load_stage1():
    """Returns (network_module, rendering_module, common_module) of the reference stage 1."""
    warnings.filterwarnings("ignore")
    pkg = "psnerf_ref_stage1"
    if pkg + ".rendering" in sys.modules:
        return (sys.modules[pkg + ".network"], sys.modules[pkg + ".rendering"], sys.modules[pkg + ".common"])
    p = types.ModuleType(pkg)
    p.__path__ = [os.path.join(REF, "stage1", "model")]
    sys.modules[pkg] = p
    common = _load(pkg + ".common", os.path.join(REF, "stage1/model/common.py"), pkg)
    network = _load(pkg + ".network", os.path.join(REF, "stage1/model/network.py"), pkg)
    rendering = _load(pkg + ".rendering", os.path.join(REF, "stage1/model/rendering.py"), pkg)
    return network, rendering, common


def _camera_params_standin(uv, pose, intrinsics):
    # device-agnostic equivalent of stage2/utils/rend_util.py:90-147 (pose-matrix branch)
    cam_loc = pose[:, :3, 3]
    fx = intrinsics[:, 0, 0].unsqueeze(-1)
    fy = intrinsics[:, 1, 1].unsqueeze(-1)
    cx = intrinsics[:, 0, 2].unsqueeze(-1)
    cy = intrinsics[:, 1, 2].unsqueeze(-1)
    z = torch.ones_like(uv[:, :, 0])
    x = (uv[:, :, 0] - cx) / fx * z
    y = (uv[:, :, 1] - cy) / fy * z
    pc = torch.stack((x, y, z), dim=-1)
    ray_dirs = torch.einsum('bij,bnj->bni', pose[:, :3, :3], pc)
    return F.normalize(ray_dirs, dim=2), cam_loc


def load_stage2():
    """Returns the reference stage2 ``model.renderer`` module (PSNetwork etc.)."""
    warnings.filterwarnings("ignore")
    if "psnerf_ref_stage2_loaded" in sys.modules:
        return sys.modules["model.renderer"]
    utils = types.ModuleType("utils")
    utils.__path__ = []
    ru = types.ModuleType("utils.rend_util")
    ru.get_camera_params = _camera_params_standin
    utils.rend_util = ru
    saved = {k: sys.modules.get(k) for k in ("utils", "utils.rend_util", "model")}
    sys.modules["utils"] = utils
    sys.modules["utils.rend_util"] = ru
    model = types.ModuleType("model")
    model.__path__ = [os.path.join(REF, "stage2", "model")]
    sys.modules["model"] = model
    for name in ("embedder", "microfacet", "sgbasis", "renderer"):
        _load("model." + name, os.path.join(REF, "stage2/model/%s.py" % name), "model")
    sys.modules["psnerf_ref_stage2_loaded"] = types.ModuleType("x")
    return sys.modules["model.renderer"]


class DictConf:
    """Minimal pyhocon-like getter object over a flat {'a.b.c': value} dict."""

    def __init__(self, d):
        self.d = dict(d)

    def _get(self, key, default=None, **kw):
        if key in self.d:
            return self.d[key]
        if "default" in kw:
            return kw["default"]
        if default is not None:
            return default
        raise KeyError(key)

    def get_string(self, key, default=None, **kw):
        return str(self._get(key, default, **kw))

    def get_int(self, key, default=None, **kw):
        return int(self._get(key, default, **kw))

    def get_float(self, key, default=None, **kw):
        return float(self._get(key, default, **kw))

    def get_bool(self, key, default=None, **kw):
        if key in self.d:
            return bool(self.d[key])
        if "default" in kw:
            return bool(kw["default"])
        return bool(default) if default is not None else False
