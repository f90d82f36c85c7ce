"""Drop-in replacement of stage1/model/network.py:NeuralNetwork.

Same constructor argument (cfg_all['model']), same parameter names (lin{l}.weight_g / weight_v / bias,
lina{l}.*: reference checkpoints load unchanged), same initialisation draws under a given torch seed, same
forward / gradient / infer_occ / infer_app signatures.  The arithmetic runs in the CUDA library
(psn_occupancy, psn_infer_occ, psn_gradient, psn_radiance); there is no CPU implementation.
In eval() mode (or under no_grad) the outputs carry no autograd graph (tensor-core or fp32 inference kernels); in train()
mode with gradients enabled forward(p, ray_d) and gradient(p) run the differentiable fp32 path of stage1/train.py
(psn_s1_train_forward / _backward), including the create_graph normals the appearance MLP consumes.
"""
import math

import torch
import torch.nn as nn

from .. import _binding as B
from .. import engine


class WNLinear(nn.Module):
    """Weight-normalised Linear with the parameter names torch.nn.utils.weight_norm produces
    (weight_g [out,1], weight_v [out,in], bias [out]); W = g * v / ||v||_row (network.py:64,77)."""

    def __init__(self, linear):
        super().__init__()
        w = linear.weight.detach()
        self.weight_g = nn.Parameter(w.norm(2, dim=1, keepdim=True).clone())
        self.weight_v = nn.Parameter(w.clone())
        self.bias = nn.Parameter(linear.bias.detach().clone())

    def effective_weight(self):
        return torch._weight_norm(self.weight_v, self.weight_g, 0)


def _fresh_linear(n_in, n_out):
    return nn.Linear(n_in, n_out)  # same RNG draws as the reference's nn.Linear(...) call


class NeuralNetwork(nn.Module):
    def __init__(self, cfg_all, **kwargs):
        super().__init__()
        cfg = cfg_all["model"]
        self.octaves_pe = cfg["octaves_pe"]
        self.octaves_pe_views = cfg["octaves_pe_views"]
        self.skips = list(cfg["skips"])
        self.rescale = cfg["rescale"]
        self.feat_size = cfg["feat_size"]
        hidden = cfg["hidden_dim"]
        d_pe = 3 + 6 * self.octaves_pe
        d_app_in = 3 + (3 + 6 * self.octaves_pe_views) + 3 + self.feat_size
        # geo stack: pe -> hidden x num_layers -> feat+1; the layer feeding a skip emits hidden - d_pe (network.py:37-45)
        widths = [d_pe] + [hidden] * cfg["num_layers"] + [self.feat_size + 1]
        self.num_layers = len(widths)
        last = self.num_layers - 2
        for l in range(self.num_layers - 1):
            n_out = widths[l + 1] - widths[0] if (l + 1) in self.skips else widths[l + 1]
            lin = _fresh_linear(widths[l], n_out)
            if cfg["geometric_init"]:  # sphere-like initial occupancy (network.py:47-61)
                std = math.sqrt(2) / math.sqrt(n_out)
                if l == last:
                    nn.init.normal_(lin.weight, mean=math.sqrt(math.pi) / math.sqrt(widths[l]), std=0.0001)
                    nn.init.constant_(lin.bias, -0.6)
                elif self.octaves_pe > 0 and l == 0:
                    nn.init.constant_(lin.bias, 0.0)
                    nn.init.constant_(lin.weight[:, 3:], 0.0)
                    nn.init.normal_(lin.weight[:, :3], 0.0, std)
                elif self.octaves_pe > 0 and l in self.skips:
                    nn.init.constant_(lin.bias, 0.0)
                    nn.init.normal_(lin.weight, 0.0, std)
                    nn.init.constant_(lin.weight[:, -(widths[0] - 3):], 0.0)
                else:
                    nn.init.constant_(lin.bias, 0.0)
                    nn.init.normal_(lin.weight, 0.0, std)
            setattr(self, "lin%d" % l, WNLinear(lin))
        app_widths = [d_app_in] + [hidden] * 4 + [3]
        self.num_layers_app = len(app_widths)
        for l in range(self.num_layers_app - 1):
            setattr(self, "lina%d" % l, WNLinear(_fresh_linear(app_widths[l], app_widths[l + 1])))
        self.precision = None  # None = engine default; 'fp32' | 'tc' | 'tc_mixed' per module

    # ---- packing -------------------------------------------------------------------------------------------
    def _prec(self):
        return None if self.precision is None else B.PRECISIONS[self.precision]

    def _packed(self):
        def build():
            with torch.no_grad():
                geo = [getattr(self, "lin%d" % l) for l in range(self.num_layers - 1)]
                app = [getattr(self, "lina%d" % l) for l in range(self.num_layers_app - 1)]
                if len(self.skips) > 1:
                    raise RuntimeError("psnerf_b200: more than one skip layer is unsupported")
                g = engine.PackedMLP(B.NET_GEO, [m.effective_weight() for m in geo], [m.bias for m in geo],
                                     octaves=self.octaves_pe, skip=self.skips[0] if self.skips else -1,
                                     rescale=self.rescale)
                a = engine.PackedMLP(B.NET_APP, [m.effective_weight() for m in app], [m.bias for m in app],
                                     octaves=self.octaves_pe_views)
            return g, a
        if next(self.parameters()).device.type != "cuda":
            raise RuntimeError("psnerf_b200: NeuralNetwork must live on a CUDA device (no CPU fallback)")
        return engine.cached_pack(self, build)

    # ---- reference API ---------------------------------------------------------------------------------------
    def infer_occ(self, p):
        g, _ = self._packed()
        shp = p.shape[:-1]
        out = engine.infer_occ(g, p.reshape(-1, 3), self.feat_size + 1, self._prec())
        return out.reshape(*shp, self.feat_size + 1)

    def _differentiable(self):
        return self.training and torch.is_grad_enabled() and any(q.requires_grad for q in self.parameters())

    def gradient(self, p, tflag=True):
        if tflag and self._differentiable():  # network.py:108-120 with create_graph=True
            from . import train as T
            return T.field(self, p.detach().reshape(-1, 3), None)[2].unsqueeze(1)
        g, _ = self._packed()
        return engine.gradient(g, p.detach().reshape(-1, 3), self._prec()).unsqueeze(1)

    def infer_app(self, points, normals, view_dirs, feature_vectors):
        raise NotImplementedError("infer_app is fused into forward(p, ray_d) in psnerf_b200 (psn_radiance)")

    def forward(self, p, ray_d=None, only_occupancy=False, return_logits=False, return_addocc=False, noise=False, **kwargs):
        g, a = self._packed()
        shp = p.shape[:-1]
        flat = p.detach().reshape(-1, 3)
        if only_occupancy:
            return engine.occupancy(g, flat, B.OUT_ALPHA, self._prec()).reshape(*shp, 1)
        if ray_d is not None and self._differentiable():
            from . import train as T
            rgb, logit, _ = T.field(self, flat, ray_d.detach().reshape(-1, 3))
            rgb = rgb.reshape(*shp, 3)
            return (rgb, torch.sigmoid(-10.0 * logit).reshape(*shp, 1)) if return_addocc else rgb
        if ray_d is not None:
            rgb, alpha = engine.radiance(g, a, flat, ray_d.detach().reshape(-1, 3), self._prec())
            rgb = rgb.reshape(*shp, 3)
            return (rgb, alpha.reshape(*shp, 1)) if return_addocc else rgb
        if return_logits:
            return engine.occupancy(g, flat, B.OUT_NEG_LOGIT, self._prec()).reshape(*shp, 1)
        return None
