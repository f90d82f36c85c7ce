"""Drop-in replacement of stage2/model/renderer.py:PSNetwork (+ Network / Normal_Network / SGBasis containers).

Same constructor (a conf object with get_string/get_int/get_float/get_bool, or a flat dict), same state-dict
keys ({albedo_net,normal_net,rough_net,visibility_net}.linears.{i}.{weight,bias}, sgbasis.lobe), same input /
output dict keys.  forward() runs psn_shade_stage2 (per-point MLPs, per-(light,point) visibility MLP, SG
shading, image-shaped writes) on the CUDA library; inference only (no autograd graph).
Supported configuration family: render_model=sgbasis, shape_pregen=True (all 7 shipped confs).
"""
import ctypes as C
import math

import torch
import torch.nn as nn

from .. import _binding as B
from .. import engine


class _DictConf:
    def __init__(self, d):
        self.d = d

    def _g(self, k, default=None, **kw):
        if k in self.d:
            return self.d[k]
        if "default" in kw:
            return kw["default"]
        if default is not None:
            return default
        raise KeyError(k)

    def get_string(self, k, default=None, **kw):
        return str(self._g(k, default, **kw))

    def get_int(self, k, default=None, **kw):
        return int(self._g(k, default, **kw))

    def get_float(self, k, default=None, **kw):
        return float(self._g(k, default, **kw))

    def get_bool(self, k, default=None, **kw):
        if k in self.d:
            return bool(self.d[k])
        return bool(kw.get("default", default if default is not None else False))


class _MLP(nn.Module):
    """Parameter container shaped like renderer.py:17-49 (linears ModuleList, cat[y, x] after skip_at layers)."""
    final_act = 0

    def __init__(self, din, dout, W, depth, skip_at=()):
        super().__init__()
        self.skip_at = list(skip_at)
        self.linears = nn.ModuleList([nn.Linear(din, W)] +
                                     [nn.Linear(W + din if i in self.skip_at else W, W) for i in range(depth - 1)] +
                                     [nn.Linear(W, dout)])

    def packed(self):
        def build():
            n = len(self.linears)
            skips = [s for s in self.skip_at if 0 <= s < n]
            if len(skips) > 1:
                raise RuntimeError("psnerf_b200: more than one skip is unsupported")
            return engine.PackedMLP(B.NET_S2, [l.weight for l in self.linears], [l.bias for l in self.linears],
                                    skip=skips[0] if skips else -1, final_act=self.final_act)
        return engine.cached_pack(self, build)


class Normal_Network(_MLP):
    final_act = 0


class Network(_MLP):
    final_act = 1


class SGBasis(nn.Module):
    def __init__(self, nbasis=9, specular_rgb=False):
        super().__init__()
        self.nbasis = nbasis
        self.specular_rgb = specular_rgb
        self.lobe = nn.Parameter(torch.tensor([math.exp(i) for i in range(2, 11)], dtype=torch.float32))
        self.lobe.requires_grad_(False)


class PSNetwork(nn.Module):
    def __init__(self, conf):
        super().__init__()
        if isinstance(conf, dict):
            conf = _DictConf(conf)
        self.conf = conf
        self.render_model = conf.get_string("train.render_model", default="sgbasis")
        if self.render_model not in ("sgbasis", "microfacet"):
            raise NotImplementedError("psnerf_b200: render_model must be sgbasis or microfacet (renderer.py:62-67)")
        self.microfacet = self.render_model == "microfacet"
        self.fresnel_f0 = conf.get_float("brdf.fresnel_f0", default=0.05)
        nbasis = conf.get_int("train.nbasis", default=9)
        self.specular_rgb = conf.get_bool("train.specular_rgb", default=False) and not self.microfacet
        if not self.microfacet:
            self.sgbasis = SGBasis(nbasis=nbasis, specular_rgb=self.specular_rgb)
        self.n_freqs = conf.get_int("brdf.net.n_freqs_xyz")
        dim_emb = 3 + 6 * self.n_freqs if self.n_freqs > 0 else 3
        W, depth = conf.get_int("brdf.net.mlp_width"), conf.get_int("brdf.net.mlp_depth")
        self.albedo_net = Network(dim_emb, 3, W, depth, skip_at=[conf.get_int("brdf.net.mlp_skip_at")])
        self.nbasis_lobes = nbasis
        if self.microfacet:  # renderer.py:73-74: one sigmoid roughness per point, same trunk shape as the albedo net
            self.rough_net = Network(dim_emb, 1, W, depth, skip_at=[conf.get_int("brdf.net.mlp_skip_at")])
            nbasis = 1
        else:
            if self.specular_rgb:
                nbasis *= 3
            self.rough_net = Normal_Network(dim_emb, nbasis, conf.get_int("brdf.sgnet.mlp_width", 128),
                                            conf.get_int("brdf.sgnet.mlp_depth", 4),
                                            skip_at=[conf.get_int("brdf.sgnet.mlp_skip_at", 2)])
        self.nbasis = nbasis
        self.light_int = conf.get_float("brdf.light_intensity", default=4.0)
        self.shape_pregen = conf.get_bool("train.shape_pregen", default=False)
        self.xyz_jitter_std = conf.get_float("brdf.net.xyz_jitter_std", default=0)
        self.normal_mlp = conf.get_bool("train.normal_mlp", default=False)
        self.n_freqs_n = self.n_freqs
        if self.normal_mlp:
            self.n_freqs_n = conf.get_int("normal.net.n_freqs_xyz")
            dn = 3 + 6 * self.n_freqs_n
            self.normal_net = Normal_Network(dn, 3, conf.get_int("normal.net.mlp_width"), conf.get_int("normal.net.mlp_depth"),
                                             skip_at=[conf.get_int("normal.net.mlp_skip_at")])
            self.normal_joint = conf.get_bool("train.normal_joint", default=False)
            self.normal_jitter_std = conf.get_float("normal.net.xyz_jitter_std", default=0)
            if not self.normal_joint:
                self.normal_net = self.normal_net.eval().requires_grad_(False)
                self.normal_jitter_std = 0
        self.visibility = conf.get_bool("train.visibility", default=False)
        self.light_vis_detach = conf.get_bool("train.light_vis_detach", default=False)
        if self.visibility:
            self.visibility_net = Normal_Network(dim_emb * 2, 1, conf.get_int("visibility.net.mlp_width"),
                                                 conf.get_int("visibility.net.mlp_depth"),
                                                 skip_at=[conf.get_int("visibility.net.mlp_skip_at")])
        self.precision = None

    def _prec(self):
        return engine.default_precision() if self.precision is None else B.PRECISIONS[self.precision]

    def forward(self, input, albedo_new=None, basis_new=None, noise=None):
        """Inference (no autograd graph) unless the module is in train() mode, gradients are enabled and a parameter / the
        lights require them, in which case the differentiable train-step path (psn_s2_train_forward / _backward,
        stage2/train.py) runs.  eval() always takes the inference kernels (stage2/eval.py calls model.eval())."""
        if next(self.parameters()).device.type != "cuda":
            raise RuntimeError("psnerf_b200: PSNetwork must live on a CUDA device (no CPU fallback)")
        wants_grad = self.training and torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or
                                                  (torch.is_tensor(input.get("light_direction")) and input["light_direction"].requires_grad))
        if wants_grad:
            return self._forward_train(input, noise)
        with torch.no_grad():
            return self._forward_eval(input, albedo_new, basis_new, noise)

    def _surface(self, input, dev):
        """Camera rays + gather of the surface pixels (renderer.py:118-125,162)."""
        uv, pose, K = input["uv"], input["pose"], input["intrinsics"]
        assert uv.shape[0] == 1
        Kc, Pc = K.detach().float().cpu(), pose.detach().float().cpu()
        dirs = engine.rays_from_pixels(uv[0].to(dev), Pc[0, :3, :3].reshape(-1).tolist(), Pc[0, :3, 3].tolist(),
                                       float(Kc[0, 0, 0]), float(Kc[0, 1, 1]), float(Kc[0, 0, 2]), float(Kc[0, 1, 2]), stage2=True)
        smask = input["surface_mask"].to(dev)
        pix = torch.nonzero(smask[0], as_tuple=False).squeeze(-1).to(torch.int32).contiguous()
        pixl = pix.long()
        surf = engine.f32c(input["points"].to(dev).float()[0][pixl])
        view = engine.f32c(-dirs[pixl])
        return surf, view, pix, pixl, uv.shape[1]

    def _forward_train(self, input, noise=None):
        from .train import S2TrainStep, flat_params
        if self.microfacet:
            raise NotImplementedError("psnerf_b200 train step: render_model=microfacet is inference-only (no shipped conf trains it)")
        if not (self.shape_pregen and self.normal_mlp and self.visibility):
            raise NotImplementedError("psnerf_b200 train step: needs shape_pregen, normal_mlp and visibility (all shipped confs)")
        if not (self.light_vis_detach and self.conf.get_bool("train.vis_rgb_detach", default=False)):
            raise NotImplementedError("psnerf_b200 train step: implemented for light_vis_detach = vis_rgb_detach = True (all shipped confs)")
        dev = next(self.parameters()).device
        with torch.no_grad():
            surf, view, pix, pixl, N = self._surface(input, dev)
        Ns = int(pix.shape[0])
        lights = input["light_direction"].to(dev)
        inten = input.get("light_intensity", None)
        if torch.is_tensor(inten):
            inten = inten.to(dev)
        jp = None
        if self.xyz_jitter_std > 0 and Ns > 0:
            z = noise["xyz"].to(dev) if (noise is not None and "xyz" in noise) else torch.randn(Ns, 3, device=dev)
            jp = surf + z * self.xyz_jitter_std
        lvt = input["light_vis_train"].to(dev) if ("light_vis_train" in input and Ns > 0) else None
        if "light_vis_train" not in input and "visibility" in input:
            # The reference's MainLoss then trains visibility_net through model_outputs['visibility'] (stage2/model/loss.py:86-87);
            # here the L-light pass is detached (light_vis_detach), so that loss term would silently carry no gradient.
            raise NotImplementedError("psnerf_b200 train step: a ground-truth 'visibility' input without 'light_vis_train' would train "
                                      "visibility_net through the detached L-light pass; supply light_vis_train (trainer.py:312-330 does)")
        rgb, spec, vis, normal, albedo, sgw, aj, wj, vt = S2TrainStep.apply(self, (surf, view, pix, N), lights, inten, jp, lvt,
                                                                            *flat_params(self))
        out = {"points": input["points"], "object_mask": input["object_mask"], "network_object_mask": input["surface_mask"],
               "sg_rgb_values": rgb, "normal_values": input["normal"], "sg_diffuse_albedo_values": albedo,
               "sg_specular_rgb_values": spec, "normal_pred": normal, "visibility": vis, "sg_weight": sgw}
        if jp is not None:
            albedo_jitter = torch.ones(1, N, 3, device=dev).index_put((torch.zeros_like(pixl), pixl), aj)
            rough_jitter = torch.ones(1, N, self.nbasis, device=dev).index_put((torch.zeros_like(pixl), pixl), wj)
            out.update({"albedo_values": albedo, "albedo_jitter": albedo_jitter, "rough_values": sgw, "rough_jitter": rough_jitter})
        if lvt is not None:
            Lt = lvt.shape[0]
            base = torch.ones(Lt, N, 3, device=dev)
            idx_l = torch.arange(Lt, device=dev).unsqueeze(1).expand(Lt, Ns)
            out["vis_train"] = base.index_put((idx_l, pixl.unsqueeze(0).expand(Lt, Ns)), vt.unsqueeze(-1).expand(Lt, Ns, 3))
        return out

    def _forward_eval(self, input, albedo_new=None, basis_new=None, noise=None):
        if not self.shape_pregen:
            raise NotImplementedError("psnerf_b200: train.shape_pregen=False is not a shipped configuration")
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("psnerf_b200: PSNetwork must live on a CUDA device (no CPU fallback)")
        lib = B.load()
        uv, pose, K = input["uv"], input["pose"], input["intrinsics"]
        assert uv.shape[0] == 1
        N = uv.shape[1]
        Kc, Pc = K.detach().float().cpu(), pose.detach().float().cpu()
        dirs = engine.rays_from_pixels(uv[0].to(dev), Pc[0, :3, :3].reshape(-1).tolist(), Pc[0, :3, 3].tolist(),
                                       float(Kc[0, 0, 0]), float(Kc[0, 1, 1]), float(Kc[0, 0, 2]), float(Kc[0, 1, 2]),
                                       stage2=True)
        smask = input["surface_mask"].to(dev)
        points = input["points"].to(dev).float()
        normals_in = input["normal"].to(dev).float()
        pix = torch.nonzero(smask[0], as_tuple=False).squeeze(-1).to(torch.int32).contiguous()
        Ns = int(pix.shape[0])
        pixl = pix.long()
        surf = engine.f32c(points[0][pixl])
        view = engine.f32c(-dirs[pixl])
        nin = engine.f32c(normals_in[0][pixl])
        lights = engine.f32c(input["light_direction"].to(dev)).reshape(-1, 3)
        L = lights.shape[0]
        inten = input.get("light_intensity", self.light_int)
        ikind, iscalar, iptr = 0, float(self.light_int), None
        if torch.is_tensor(inten):
            t = engine.f32c(inten.to(dev))
            if t.numel() == 1:
                iscalar = float(t)
            elif t.shape[0] > 1 and t.shape[-1] == 3 and t.dim() == 2:
                ikind, iptr = 2, t
            elif t.shape[0] > 1:
                ikind, iptr = 1, t.reshape(-1)
            else:  # [1,3]: one RGB intensity shared by all lights
                ikind, iptr = 2, t.reshape(1, 3).expand(L, 3).contiguous()
        else:
            iscalar = float(inten)
        prm = B.ShadeParams(self.n_freqs, self.n_freqs_n, self.nbasis_lobes, 1 if self.specular_rgb else 0, ikind, iscalar,
                            1 if self.microfacet else 0, self.fresnel_f0)
        rgb = torch.empty(L, N, 3, device=dev)
        spec = torch.empty(1 if self.microfacet else L, N, 3, device=dev)  # microfacet: the per-pixel roughness image
        vis = torch.empty(L, N, 3, device=dev) if self.visibility else None
        normal = torch.empty(1, N, 3, device=dev) if self.normal_mlp else None
        albedo = torch.empty(1, N, 3, device=dev)
        sgw = torch.empty(1, N, self.nbasis, device=dev)
        ws = engine.workspace(dev, "shade", max(N, Ns), 1, L)
        P = engine._ptr
        # material editing (stage2/eval.py:116-132 -> renderer.py:167-168,175-181): one albedo / one SG weight vector for all points
        a_new = w_new = None
        if albedo_new is not None:
            a_new = engine.f32c(torch.as_tensor(albedo_new).to(dev).reshape(3))
        if basis_new is not None and not self.microfacet:  # renderer.py:175 sits in the sgbasis branch
            nb = self.nbasis_lobes
            wn = torch.zeros(3 if self.specular_rgb else 1, nb)
            wn[:, basis_new] = torch.as_tensor(2.0 ** torch.as_tensor(basis_new, dtype=torch.float64) / 100).float()
            w_new = engine.f32c(wn.reshape(-1).to(dev))
        with torch.cuda.device(dev):
            B.check(lib.psn_shade_stage2_edit(
                P(self.normal_net.packed().handle) if self.normal_mlp else C.c_void_p(0),
                P(self.albedo_net.packed().handle), P(self.rough_net.packed().handle),
                P(self.visibility_net.packed().handle) if self.visibility else C.c_void_p(0),
                P(None if self.microfacet else engine.f32c(self.sgbasis.lobe.detach())), C.byref(prm), P(surf), P(view), P(nin),
                P(pix), Ns, N, P(lights), L, P(iptr), P(a_new), P(w_new), P(rgb), P(spec), P(vis), P(normal), P(albedo),
                P(None if self.microfacet else sgw), P(ws),
                ws.numel(), self._prec(), engine._stream()), "psn_shade_stage2_edit")
        out = {"points": input["points"], "object_mask": input["object_mask"], "network_object_mask": input["surface_mask"],
               "sg_rgb_values": rgb, "normal_values": input["normal"], "sg_diffuse_albedo_values": albedo,
               "sg_specular_rgb_values": spec}
        if self.xyz_jitter_std > 0 and Ns > 0:  # renderer.py:211-231 (smoothness-loss inputs)
            z = noise["xyz"].to(dev) if (noise is not None and "xyz" in noise) else torch.randn(Ns, 3, device=dev)
            pj = engine.f32c(surf + z * self.xyz_jitter_std)
            aj = torch.empty(Ns, 3, device=dev)
            wj = torch.empty(Ns, self.nbasis, device=dev)
            with torch.cuda.device(dev):
                B.check(lib.psn_s2_point_nets(P(self.albedo_net.packed().handle), P(self.rough_net.packed().handle),
                                              self.n_freqs, P(pj), Ns, P(aj), P(wj), self.nbasis, self._prec(),
                                              engine._stream()), "psn_s2_point_nets")
            albedo_jitter = torch.ones(1, N, 3, device=dev)
            albedo_jitter[0, pixl] = aj
            if self.microfacet:  # renderer.py:224-226: [1,N,3] roughness images
                rough_jitter = torch.ones(1, N, 3, device=dev)
                rough_jitter[0, pixl] = wj.expand(-1, 3)
                rough_ori = spec
            else:
                rough_jitter = torch.ones(1, N, self.nbasis, device=dev)
                rough_jitter[0, pixl] = wj
                rough_ori = sgw
            out.update({"albedo_values": albedo, "albedo_jitter": albedo_jitter, "rough_values": rough_ori,
                        "rough_jitter": rough_jitter})
        if self.normal_mlp:
            out["normal_pred"] = normal
        if self.visibility:
            out["visibility"] = vis
            if "vis_train_gt" in input or "light_vis_train" in input:  # renderer.py:251-262
                lt = engine.f32c(input["light_vis_train"].to(dev)).reshape(-1, 3)
                Lt = lt.shape[0]
                vt = torch.ones(Lt, N, 3, device=dev)
                if Ns > 0:
                    raw = torch.empty(Lt, Ns, device=dev)
                    ws2 = engine.workspace(dev, "s2_vis", Ns, 1, Lt)
                    with torch.cuda.device(dev):
                        B.check(lib.psn_s2_visibility(P(self.visibility_net.packed().handle), self.n_freqs, P(surf), Ns, P(lt),
                                                      Lt, P(raw), P(ws2), ws2.numel(), self._prec(), engine._stream()),
                                "psn_s2_visibility")
                    vt[:, pixl, :] = raw.unsqueeze(-1).expand(-1, -1, 3)
                out["vis_train"] = vt
        if not self.microfacet:
            out["sg_weight"] = sgw
        return out
