"""Drop-in replacement of stage1/model/rendering.py:Renderer.

Same constructor (model, cfg_all, device) and forward(pixels, camera_mat, world_mat, scale_mat,
rendering_technique, add_noise, eval_, it, visibility, light_dir) -> dict with the reference's keys.
One library call renders the whole pixel batch (no 1024-ray / 64000-point chunk loops are needed: the
kernels never materialise per-sample tensors beyond a few floats per sample).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _binding as B
from .. import engine


class Renderer(nn.Module):
    def __init__(self, model, cfg_all, device=None, **kwargs):
        super().__init__()
        cfg = cfg_all["rendering"]
        self._device = device
        self.depth_range = [cfg["near"], cfg["far"]]
        self.n_max_network_queries = cfg["n_max_network_queries"]  # kept for config compatibility; unused
        self.white_background = cfg["white_background"]
        self.cfg = cfg
        self.model = model.to(device)

    def to(self, device):
        m = super().to(device)
        m._device = device
        return m

    # ---- helpers ---------------------------------------------------------------------------------------------
    def _rays(self, pixels, camera_mat, world_mat):
        """origin (3 host floats) and unit directions [N,3] (common.py:205-226, rendering.py:67-71)."""
        assert pixels.shape[0] == 1, "batch size must be 1 (rendering.py:100)"
        dev = next(self.model.parameters()).device
        cm = camera_mat.detach().float().cpu()
        wm = world_mat.detach().float().cpu()
        fx, cx, cy = float(cm[0, 0, 0]), float(cm[0, 0, 2]), float(cm[0, 1, 2])
        R = wm[0, :3, :3].reshape(-1).tolist()
        origin = wm[0, :3, 3].tolist()
        pix = pixels[0].to(dev).float()
        dirs = engine.rays_from_pixels(pix, R, origin, fx, fx, cx, cy, stage2=False)
        return origin, dirs

    def _geo_app(self):
        return self.model._packed()

    def _normal_prec(self):
        """Precision of the surface-normal OUTPUT launches (one gradient evaluation per hit ray): fp32 under every tensor-core
        precision, like the library's own unisurf normals (csrc/api_stage1.cu: normal_precision)."""
        p = self.model._prec()
        p = engine.default_precision() if p is None else p
        return B.PREC_FP32 if p != B.PREC_FP32 else p

    def _surface(self, depth, origin, dirs):
        """mask / surface points exactly as rendering.py:88-108."""
        zero_occ = depth == 0
        finite = (depth.abs() != np.inf) & ~torch.isnan(depth)
        dists = torch.ones_like(depth)
        dists[finite] = depth[finite]
        dists[zero_occ] = 0.0
        obj = finite & ~zero_occ
        o = torch.tensor(origin, dtype=torch.float32, device=dirs.device)
        pts = o[None, :] + dirs * dists[:, None]
        return obj, pts

    # ---- reference API ---------------------------------------------------------------------------------------
    def forward(self, pixels, camera_mat, world_mat, scale_mat, rendering_technique, add_noise=True, eval_=False, it=0,
                visibility=False, light_dir=None):
        if rendering_technique == "unisurf":
            if self.model.training and torch.is_grad_enabled() and not eval_ and any(p.requires_grad for p in self.model.parameters()):
                return self.unisurf_train(pixels, camera_mat, world_mat, it=it, add_noise=add_noise)
            return self.unisurf(pixels, camera_mat, world_mat, scale_mat, it=it, add_noise=add_noise, eval_=eval_)
        if rendering_technique == "phong_renderer":
            return self.phong_renderer(pixels, camera_mat, world_mat, scale_mat)
        if rendering_technique == "shape_extract":
            return self.shape_extract(pixels, camera_mat, world_mat, scale_mat, it=it, visibility=visibility,
                                      light_dir=light_dir)
        raise ValueError("Choose unisurf, phong_renderer or shape_extract")

    @torch.no_grad()
    def ray_marching(self, ray0, ray_direction, model=None, c=None, tau=0.5, n_steps=(128, 129), n_secant_steps=8,
                     depth_range=(25, 40), max_points=3500000, rad=1.0, clip=False):
        """rendering.py:410-523: [1,N] depths, inf = miss, 0 = first proposal occupied."""
        if clip:
            raise NotImplementedError("clip=True is unused by the reference entry points")
        g, _ = self._geo_app()
        origin = ray0[0, 0].detach().float().cpu().tolist()
        d = engine.raymarch(g, origin, ray_direction[0], depth_range[0], rad, int(n_steps[0]), n_secant_steps, 0.5,
                            self.model._prec())
        return d.unsqueeze(0)

    def _unisurf_plan(self, g, origin, dirs, it):
        """UnisurfParams of a call: interval half-width delta(it) and the full_steps switch (rendering.py:116-127)."""
        cfg = self.cfg
        near = float(cfg["near"])
        steps, steps_out = int(cfg["num_points_in"]), int(cfg["num_points_out"])
        delta = float(torch.max(cfg["interval_start"] * torch.exp(-1 * cfg["interval_decay"] * it * torch.ones(1)),
                                cfg["interval_end"] * torch.ones(1)))
        full_ok = it > 5000
        if full_ok and not near > 0:
            d0 = engine.raymarch(g, origin, dirs, near, cfg["radius"], int(cfg["ray_marching_steps"]), 8, 0.5, self.model._prec())
            hit = (d0.abs() != np.inf) & (d0 != 0)
            dnp = torch.clamp(d0[hit] - delta, min=near)
            full_ok = bool((dnp != 0.0).all())
        return B.UnisurfParams(near, float(cfg["radius"]), delta, 0.5, int(cfg["ray_marching_steps"]), 8, steps,
                               steps_out if full_ok else 0, 1 if self.white_background else 0)

    def unisurf_train(self, pixels, camera_mat, world_mat, it=100000, add_noise=True, noise=None):
        """Training forward of Renderer.unisurf (rendering.py:50-226 with eval_=False, called from training.py:180): the outputs
        carry an autograd graph to every parameter of the field.  The surface search and the sample depths are computed by the
        inference kernels without gradient (rendering.py:79-87 is under no_grad in the reference too); the radiance samples with
        their create_graph normals, the compositing and the surface normals run through the differentiable CUDA path
        (stage1/train.py: psn_s1_train_forward / _backward, psn_composite / _bwd).  ``noise`` optionally supplies the random
        draws: 'depth' [N,S] uniform samples (rendering.py:139,163) and 'neigh' [Ns,3] (rendering.py:204)."""
        from . import train as T
        noise = noise or {}
        with torch.no_grad():
            g, a = self._geo_app()
            origin, dirs = self._rays(pixels, camera_mat, world_mat)
            N = dirs.shape[0]
            prm = self._unisurf_plan(g, origin, dirs, it)
            S = prm.steps_in + prm.steps_out
            u = None
            if add_noise:
                u = noise["depth"].to(dirs.device) if "depth" in noise else torch.rand(N, S, device=dirs.device)
            info = engine.render_unisurf(g, a, origin, dirs, prm, noise=u, want_sample_depth=True, precision=self.model._prec())
            obj, pts = self._surface(info["depth"], origin, dirs)
            o = torch.tensor(origin, dtype=torch.float32, device=dirs.device)
            p_fg = (o[None, None, :] + dirs[:, None, :] * info["sample_depth"][:, :, None]).reshape(-1, 3)
            v_fg = (-dirs)[:, None, :].expand(N, S, 3).reshape(-1, 3)
            sp = pts[obj]
            Ns = sp.shape[0]
            nu = noise["neigh"].to(dirs.device) if "neigh" in noise else torch.rand_like(sp)
            pp = torch.cat([sp, sp + (nu - 0.5) * 0.01], 0)
        params = T.effective_params(self.model)
        rgb_s, logit, _ = T.field(self.model, p_fg, v_fg, params)
        alpha = torch.sigmoid(-10.0 * logit).view(N, S)  # network.py:134
        rgb, acc = T.composite(rgb_s.view(N, S, 3), alpha, bool(self.white_background))
        norm_pred = torch.zeros(N, 3, device=dirs.device)
        if Ns > 0:
            _, _, gg = T.field(self.model, pp, None, params)
            normals_ = gg / (gg.norm(2, dim=1).unsqueeze(-1) + 10 ** (-5))
            norm_pred = norm_pred.index_put((torch.nonzero(obj).squeeze(-1),), normals_[:Ns])
            diff_norm = torch.norm(normals_[:Ns] - normals_[Ns:], dim=-1)
        else:
            diff_norm = torch.zeros(0, device=dirs.device)
        return {"rgb": rgb.reshape(1, -1, 3), "mask_pred": obj, "diff_norm": diff_norm,
                "normal_pred": norm_pred.reshape(1, -1, 3), "acc_map": acc.reshape(1, -1)}

    @torch.no_grad()
    def unisurf(self, pixels, camera_mat, world_mat, scale_mat, add_noise=False, it=100000, eval_=False):
        cfg = self.cfg
        g, a = self._geo_app()
        origin, dirs = self._rays(pixels, camera_mat, world_mat)
        N = dirs.shape[0]
        near = float(cfg["near"])
        steps, steps_out = int(cfg["num_points_in"]), int(cfg["num_points_out"])
        # delta = max(start*exp(-decay*it), end) in float32 like rendering.py:116
        delta = float(torch.max(cfg["interval_start"] * torch.exp(-1 * cfg["interval_decay"] * it * torch.ones(1)),
                                cfg["interval_end"] * torch.ones(1)))
        full_ok = it > 5000
        if full_ok and not near > 0:
            # (dnp != 0).all() can only fail when near <= 0 (rendering.py:124): decide it from a surface search
            d0 = engine.raymarch(g, origin, dirs, near, cfg["radius"], int(cfg["ray_marching_steps"]), 8, 0.5,
                                 self.model._prec())
            hit = (d0.abs() != np.inf) & (d0 != 0)
            dnp = torch.clamp(d0[hit] - delta, min=near)
            full_ok = bool((dnp != 0.0).all())
        prm = B.UnisurfParams(near, float(cfg["radius"]), delta, 0.5, int(cfg["ray_marching_steps"]), 8, steps,
                              steps_out if full_ok else 0, 1 if self.white_background else 0)
        S = prm.steps_in + prm.steps_out
        noise = torch.rand(N, S, device=dirs.device) if add_noise else None
        out = engine.render_unisurf(g, a, origin, dirs, prm, noise=noise, precision=self.model._prec())
        diff_norm = None
        if not eval_:  # rendering.py:203-211: normal consistency between surface points and jittered neighbours
            obj, pts = self._surface(out["depth"], origin, dirs)
            sp = pts[obj]
            if sp.shape[0] > 0:
                nb = sp + (torch.rand_like(sp) - 0.5) * 0.01
                gn = engine.gradient(g, nb, self.model._prec())
                n2 = gn / (gn.norm(2, dim=1, keepdim=True) + 10 ** (-5))
                diff_norm = torch.norm(out["normal"][obj] - n2, dim=-1)
            else:
                diff_norm = torch.zeros(0, device=dirs.device)
        return {"rgb": out["rgb"].reshape(1, -1, 3), "mask_pred": out["mask"], "diff_norm": diff_norm,
                "normal_pred": out["normal"].reshape(1, -1, 3), "acc_map": out["acc"].reshape(1, -1)}

    @torch.no_grad()
    def phong_renderer(self, pixels, camera_mat, world_mat, scale_mat):
        """rendering.py:228-293 (debug visualisation of the marched surface, 512 steps)."""
        g, _ = self._geo_app()
        origin, dirs = self._rays(pixels, camera_mat, world_mat)
        d = engine.raymarch(g, origin, dirs, float(self.cfg["near"]), self.cfg["radius"], 512, 8, 0.5, self.model._prec())
        obj, pts = self._surface(d, origin, dirs)
        rgb = torch.ones_like(pts)
        if int(obj.sum()) > 0:
            grad = engine.gradient(g, pts[obj], self._normal_prec())
            nrm = grad / grad.norm(2, 1, keepdim=True)
            o = torch.tensor(origin, dtype=torch.float32, device=dirs.device)
            light = (o / o.norm(2)).unsqueeze(1)
            diffuse = torch.mm(nrm, light).clamp_min(0).repeat(1, 3) * 0.7
            rgb[obj] = (0.3 + diffuse).clamp_max(1.0)
        return {"rgb": rgb.reshape(1, -1, 3)}

    @torch.no_grad()
    def shape_extract(self, pixels, camera_mat, world_mat, scale_mat, it=100000, visibility=False, light_dir=None):
        """rendering.py:297-376: surface points / normals / mask (+ per-light shadow-ray visibility)."""
        g, _ = self._geo_app()
        origin, dirs = self._rays(pixels, camera_mat, world_mat)
        N = dirs.shape[0]
        d = engine.raymarch(g, origin, dirs, float(self.cfg["near"]), self.cfg["radius"], 512, 8, 0.5, self.model._prec())
        obj, pts = self._surface(d, origin, dirs)
        surf = pts[obj]
        normal = torch.zeros(N, 3, device=dirs.device)
        if surf.shape[0] > 0:
            normal[obj] = F.normalize(engine.gradient(g, surf, self._normal_prec()), dim=-1)
        out = {"mask": obj.reshape(1, -1), "normal": normal.reshape(1, -1, 3), "points": pts.reshape(1, -1, 3)}
        if visibility and light_dir is not None:
            light_dir = light_dir.to(dirs.device).float()
            vis = torch.ones(light_dir.shape[0], N, device=dirs.device)
            if surf.shape[0] > 0:
                vis[:, obj] = self.light_visibility(surf=surf, light_dir=light_dir).view(light_dir.shape[0], -1)
            out["visibility"] = vis
        return out

    @torch.no_grad()
    def light_visibility(self, surf=None, light_dir=None, lnear=0.1, lfar=3.5, tau=0.5, n_steps=128, max_points=3500000):
        """rendering.py:378-408: flat [L*Ns] light-major transmittances."""
        g, _ = self._geo_app()
        return engine.shadow_visibility(g, surf, light_dir, lnear, lfar, n_steps, 1.1, self.model._prec()).reshape(-1)
