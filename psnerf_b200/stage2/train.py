"""Autograd bridge of the stage-2 train step (BASELINE config 5): psn_s2_train_forward / psn_s2_train_backward.

The differentiable outputs are rgb [L,N,3], normal_pred [1,N,3], albedo [1,N,3], sg weights [1,N,nbt], the jittered
albedo / weights (per surface point) and the vis-train visibilities [Lt,Ns]; gradients flow to every parameter of
normal_net / albedo_net / rough_net / visibility_net, to the light directions and to the light intensities, with the
reference's detach semantics (renderer.py:193-199: the L-light visibility pass carries no gradient)."""
import ctypes as C

import torch

from .. import _binding as B
from .. import engine


def _net_view(mlp, grads=None):
    """psn_train_net view of a Network / Normal_Network container over its live parameters (and gradient buffers)."""
    n = len(mlp.linears)
    ws = [engine.f32c(l.weight.detach()) for l in mlp.linears]
    bs = [engine.f32c(l.bias.detach()) for l in mlp.linears]
    keep = [ws, bs]
    in_dims = (C.c_int * n)(*[w.shape[1] for w in ws])
    out_dims = (C.c_int * n)(*[w.shape[0] for w in ws])
    W = (C.c_void_p * n)(*[w.data_ptr() for w in ws])
    Bv = (C.c_void_p * n)(*[b.data_ptr() for b in bs])
    if grads is not None:
        dW = (C.c_void_p * n)(*[g.data_ptr() for g in grads[0]])
        dB = (C.c_void_p * n)(*[g.data_ptr() for g in grads[1]])
    else:
        dW = (C.c_void_p * n)()
        dB = (C.c_void_p * n)()
    skips = [s for s in mlp.skip_at if 0 <= s < n]
    tn = B.TrainNet(n, skips[0] if skips else -1, mlp.final_act, in_dims, out_dims, W, Bv, dW, dB)
    keep += [in_dims, out_dims, W, Bv, dW, dB, grads]
    return tn, keep


class S2TrainStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, geom, lights, intensity, jitter_pts, lights_vt, *params):
        """geom = (surf [Ns,3], view [Ns,3], pix int32 [Ns], N).  params: flat list of the four nets' parameters (only used so
        that autograd tracks them); the kernels read the live parameter storage through `model`."""
        lib = B.load()
        surf, view, pix, N = geom
        dev = surf.device
        Ns, L = surf.shape[0], lights.shape[0]
        Lt = 0 if lights_vt is None else lights_vt.shape[0]
        lights = engine.f32c(lights.detach())
        ikind, iscalar, iptr = 0, float(model.light_int), None
        if torch.is_tensor(intensity):
            t = engine.f32c(intensity.detach())
            if t.numel() == 1:
                iscalar = float(t)
            elif t.dim() == 2 and t.shape[-1] == 3:
                # [1,3] is one RGB intensity shared by all lights (renderer.py:188-190 broadcasts it): expanded like the eval path
                ikind, iptr = 2, (t.expand(L, 3).contiguous() if t.shape[0] == 1 else t)
            else:
                ikind, iptr = 1, t.reshape(-1)
        elif intensity is not None:
            iscalar = float(intensity)
        prm = B.ShadeParams(model.n_freqs, model.n_freqs_n, model.nbasis_lobes, 1 if model.specular_rgb else 0, ikind, iscalar)
        nets = [model.normal_net, model.albedo_net, model.rough_net, model.visibility_net]
        views = [_net_view(m) for m in nets]
        tape_bytes = lib.psn_s2_train_tape_bytes(C.byref(views[0][0]), C.byref(views[1][0]), C.byref(views[2][0]), C.byref(views[3][0]),
                                                 Ns, L, Lt)
        tape = torch.empty(max(int(tape_bytes), 16), dtype=torch.uint8, device=dev)
        ws = engine.workspace(dev, "s2_train", max(N, Ns), max(Lt, 1), L)
        nbt = model.nbasis
        rgb = torch.empty(L, N, 3, device=dev)
        spec = torch.empty(L, N, 3, device=dev)
        vis = torch.empty(L, N, 3, device=dev)
        normal = torch.empty(1, N, 3, device=dev)
        albedo = torch.empty(1, N, 3, device=dev)
        sgw = torch.empty(1, N, nbt, device=dev)
        aj = torch.empty(Ns, 3, device=dev) if jitter_pts is not None else None
        wj = torch.empty(Ns, nbt, device=dev) if jitter_pts is not None else None
        vt = torch.empty(Lt, Ns, device=dev) if Lt > 0 else None
        jp = engine.f32c(jitter_pts.detach()) if jitter_pts is not None else None
        lvt = engine.f32c(lights_vt.detach()) if Lt > 0 else None
        P = engine._ptr
        lobe = engine.f32c(model.sgbasis.lobe.detach())
        with torch.cuda.device(dev):
            B.check(lib.psn_s2_train_forward(C.byref(views[0][0]), C.byref(views[1][0]), C.byref(views[2][0]), C.byref(views[3][0]),
                                             P(model.visibility_net.packed().handle), P(lobe), C.byref(prm), P(surf), P(view), P(pix), Ns, N,
                                             P(lights), L, P(iptr), P(jp), P(lvt), Lt, P(rgb), P(spec), P(vis), P(normal), P(albedo), P(sgw),
                                             P(aj), P(wj), P(vt), P(tape), tape.numel(), P(ws), ws.numel(), model._prec(), engine._stream()),
                    "psn_s2_train_forward")
        ctx.model, ctx.prm, ctx.tape, ctx.geom = model, prm, tape, (surf, view, pix, N)
        ctx.lights, ctx.iptr, ctx.ikind, ctx.Lt, ctx.lobe = lights, iptr, ikind, Lt, lobe
        ctx.int_shape = intensity.shape if torch.is_tensor(intensity) else None
        ctx.has_jit, ctx.nparams = jitter_pts is not None, len(params)
        ctx.mark_non_differentiable(spec, vis)
        outs = [rgb, spec, vis, normal, albedo, sgw]
        outs += [aj if aj is not None else torch.zeros(0, device=dev), wj if wj is not None else torch.zeros(0, device=dev),
                 vt if vt is not None else torch.zeros(0, device=dev)]
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_rgb, g_spec, g_vis, g_normal, g_albedo, g_sgw, g_aj, g_wj, g_vt):
        lib = B.load()
        model, prm = ctx.model, ctx.prm
        surf, view, pix, N = ctx.geom
        dev = surf.device
        Ns, L, Lt = surf.shape[0], ctx.lights.shape[0], ctx.Lt
        nets = [model.normal_net, model.albedo_net, model.rough_net, model.visibility_net]
        grads = [([torch.zeros_like(l.weight, dtype=torch.float32) for l in m.linears],
                  [torch.zeros_like(l.bias, dtype=torch.float32) for l in m.linears]) for m in nets]
        views = [_net_view(m, g) for m, g in zip(nets, grads)]
        d_l = torch.zeros(L, 3, device=dev)
        d_i = torch.zeros(max(ctx.lights.shape[0] * 3, 4), device=dev)
        ws = engine.workspace(dev, "s2_train", max(N, Ns), max(Lt, 1), L)
        P = engine._ptr

        def c(t):
            return None if t is None else engine.f32c(t)
        g_aj = c(g_aj) if ctx.has_jit else None
        g_wj = c(g_wj) if ctx.has_jit else None
        g_vt = c(g_vt) if Lt > 0 else None
        with torch.cuda.device(dev):
            B.check(lib.psn_s2_train_backward(C.byref(views[0][0]), C.byref(views[1][0]), C.byref(views[2][0]), C.byref(views[3][0]),
                                              P(ctx.lobe), C.byref(prm), P(view), P(pix), Ns, N, P(ctx.lights), L, P(ctx.iptr), Lt,
                                              P(c(g_rgb)), None, P(c(g_normal)), P(c(g_albedo)), P(c(g_sgw)), P(g_aj), P(g_wj), P(g_vt),
                                              P(d_l), P(d_i), P(ctx.tape), ctx.tape.numel(), P(ws), ws.numel(), engine._stream()),
                    "psn_s2_train_backward")
        if ctx.int_shape is None:
            g_int = None
        elif ctx.ikind == 0:
            g_int = d_i[:1].reshape(ctx.int_shape)
        elif ctx.ikind == 1:
            g_int = d_i[:L].reshape(ctx.int_shape)
        else:
            g_int = d_i[:L * 3].reshape(L, 3)
            g_int = g_int.sum(0, keepdim=True) if ctx.int_shape[0] == 1 else g_int.reshape(ctx.int_shape)
        flat = []
        for gw, gb in grads:
            for a, b in zip(gw, gb):
                flat += [a, b]
        return (None, None, d_l, g_int, None, None) + tuple(flat)


def flat_params(model):
    out = []
    for m in (model.normal_net, model.albedo_net, model.rough_net, model.visibility_net):
        for l in m.linears:
            out += [l.weight, l.bias]
    return out


def train_step(model, loss, model_input, ground_truth, sg_optimizer, loss_n=None, light_optimizer=None, light_para=None):
    """One iteration of TrainRunner.run (stage2/trainer.py:394-410): forward, MainLoss (+ NormalLoss when normal_train), zero_grad,
    backward, optimizer steps.  The light optimizer (SparseAdam over the light tables, trainer.py:165) is zeroed and stepped only while
    the light-direction table still requires grad (train_fix may freeze it, trainer.py:359-360).  Returns (loss_output, loss_normal)."""
    model_outputs = model(model_input)
    loss_output = loss(model_outputs, ground_truth, model_input)
    total = loss_output["loss"]
    loss_normal = None
    if loss_n is not None:
        loss_normal = loss_n(model_outputs)
        total = total + loss_normal["loss"]
    step_lights = light_optimizer is not None and (light_para is None or light_para.weight.requires_grad)
    sg_optimizer.zero_grad()
    if step_lights:
        light_optimizer.zero_grad()
    total.backward()
    sg_optimizer.step()
    if step_lights:
        light_optimizer.step()
    loss_output["loss"] = total
    return loss_output, loss_normal
