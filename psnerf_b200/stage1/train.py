"""Autograd bridge of the stage-1 train step: psn_s1_train_forward / psn_s1_train_backward / psn_composite(_bwd).

`field(model, pts, views)` evaluates NeuralNetwork.forward(p, ray_d, return_addocc=True) under autograd
(stage1/model/network.py:122-136 with the create_graph normals of :108-120): per sample rgb [M,3], logit [M] and
grad = d logit / d p [M,3], differentiable w.r.t. every parameter of the geo and appearance nets (first order through
logit / feature, second order through grad).  `views=None` is the gradient-only evaluation of the surface normals
(rendering.py:203-211).  The kernels work on the EFFECTIVE weights W = g v / |v|; torch._weight_norm (a [out,in]
element-wise op on the parameters) carries the gradients on to weight_g / weight_v.
`composite(rgb_s, alpha, white)` is rendering.py:196-197,214-216 with its hand-derived backward."""
import ctypes as C

import torch

from .. import _binding as B
from .. import engine


def _net_view(Ws, bs, skip, grads=None):
    n = len(Ws)
    keep = [Ws, bs, grads]
    in_dims = (C.c_int * n)(*[w.shape[1] for w in Ws])
    out_dims = (C.c_int * n)(*[w.shape[0] for w in Ws])
    W = (C.c_void_p * n)(*[w.data_ptr() for w in Ws])
    Bv = (C.c_void_p * n)(*[b.data_ptr() for b in bs])
    if grads is not None:
        dW = (C.c_void_p * n)(*[g.data_ptr() for g in grads[0]])
        dB = (C.c_void_p * n)(*[g.data_ptr() for g in grads[1]])
    else:
        dW = (C.c_void_p * n)()
        dB = (C.c_void_p * n)()
    tn = B.TrainNet(n, skip, 0, in_dims, out_dims, W, Bv, dW, dB)
    keep += [in_dims, out_dims, W, Bv, dW, dB]
    return tn, keep


class S1Field(torch.autograd.Function):
    @staticmethod
    def forward(ctx, meta, pts, views, *wb):
        """meta = (n_geo, n_app, octaves, octaves_view, skip, rescale); wb = geo W.., geo b.., app W.., app b.. (effective)."""
        n_geo, n_app, octaves, octaves_view, skip, rescale = meta
        lib = B.load()
        dev = pts.device
        pts = engine.f32c(pts.detach()).reshape(-1, 3)
        M = pts.shape[0]
        with_app = views is not None
        if with_app:
            views = engine.f32c(views.detach()).reshape(-1, 3)
        gW = [engine.f32c(t.detach()) for t in wb[:n_geo]]
        gb = [engine.f32c(t.detach()) for t in wb[n_geo:2 * n_geo]]
        aW = [engine.f32c(t.detach()) for t in wb[2 * n_geo:2 * n_geo + n_app]]
        ab = [engine.f32c(t.detach()) for t in wb[2 * n_geo + n_app:]]
        geo, k1 = _net_view(gW, gb, skip)
        app, k2 = _net_view(aW, ab, -1) if with_app else (None, None)
        app_p = C.byref(app) if with_app else None
        tape_bytes = int(lib.psn_s1_train_tape_bytes(C.byref(geo), app_p, octaves, octaves_view, M))
        ws_bytes = int(lib.psn_s1_train_ws_bytes(C.byref(geo), app_p, octaves, octaves_view, M))
        if tape_bytes < 0 or ws_bytes < 0:
            raise RuntimeError("psnerf_b200 stage-1 train: " + (lib.psn_last_error() or b"?").decode())
        tape = torch.empty(max(tape_bytes, 16), dtype=torch.uint8, device=dev)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        rgb = torch.empty(M, 3, device=dev) if with_app else torch.zeros(M, 3, device=dev)
        logit = torch.empty(M, device=dev)
        grad = torch.empty(M, 3, device=dev)
        P = engine._ptr
        with torch.cuda.device(dev):
            B.check(lib.psn_s1_train_forward(C.byref(geo), app_p, octaves, octaves_view, float(rescale), P(pts),
                                             P(views) if with_app else None, M, P(rgb) if with_app else None, P(logit), P(grad),
                                             P(tape), tape.numel(), P(ws), ws.numel(), engine._stream()), "psn_s1_train_forward")
        ctx.meta, ctx.with_app, ctx.M = meta, with_app, M
        ctx.tensors = (gW, gb, aW, ab, tape, ws)
        if not with_app:
            ctx.mark_non_differentiable(rgb)
        return rgb, logit, grad

    @staticmethod
    def backward(ctx, g_rgb, g_logit, g_grad):
        n_geo, n_app, octaves, octaves_view, skip, rescale = ctx.meta
        gW, gb, aW, ab, tape, ws = ctx.tensors
        lib = B.load()
        dev = tape.device
        M = ctx.M
        dgW = [torch.zeros_like(t) for t in gW]
        dgb = [torch.zeros_like(t) for t in gb]
        daW = [torch.zeros_like(t) for t in aW]
        dab = [torch.zeros_like(t) for t in ab]
        if M > 0:
            geo, k1 = _net_view(gW, gb, skip, (dgW, dgb))
            app, k2 = _net_view(aW, ab, -1, (daW, dab)) if ctx.with_app else (None, None)
            P = engine._ptr

            def cot(t):
                return None if t is None else engine.f32c(t)
            g_rgb, g_logit, g_grad = (cot(g_rgb) if ctx.with_app else None), cot(g_logit), cot(g_grad)
            with torch.cuda.device(dev):
                B.check(lib.psn_s1_train_backward(C.byref(geo), C.byref(app) if ctx.with_app else None, octaves, octaves_view,
                                                  float(rescale), M, P(g_rgb), P(g_logit), P(g_grad), P(tape), tape.numel(), P(ws),
                                                  ws.numel(), engine._stream()), "psn_s1_train_backward")
        ctx.tensors = None
        return (None, None, None, *dgW, *dgb, *daW, *dab)


def effective_params(model):
    """Effective (weight-norm folded) weights and biases of the two nets, attached to the autograd graph of weight_g / weight_v."""
    geo = [getattr(model, "lin%d" % l) for l in range(model.num_layers - 1)]
    app = [getattr(model, "lina%d" % l) for l in range(model.num_layers_app - 1)]
    return ([m.effective_weight() for m in geo], [m.bias for m in geo], [m.effective_weight() for m in app], [m.bias for m in app])


def field(model, pts, views=None, params=None):
    """(rgb [M,3] | None, logit [M], grad [M,3]) of the stage-1 field at pts (and view directions), differentiable."""
    gW, gb, aW, ab = effective_params(model) if params is None else params
    if len(model.skips) > 1:
        raise RuntimeError("psnerf_b200: more than one skip layer is unsupported")
    meta = (len(gW), len(aW), int(model.octaves_pe), int(model.octaves_pe_views), model.skips[0] if model.skips else -1,
            float(model.rescale))
    rgb, logit, grad = S1Field.apply(meta, pts, views, *gW, *gb, *aW, *ab)
    return (rgb if views is not None else None), logit, grad


class Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb_s, alpha, white):
        rgb_s, alpha = engine.f32c(rgb_s.detach()), engine.f32c(alpha.detach())
        rgb, acc = engine.composite(rgb_s, alpha, white)
        ctx.save_for_backward(rgb_s, alpha)
        ctx.white = white
        return rgb, acc

    @staticmethod
    def backward(ctx, g_rgb, g_acc):
        rgb_s, alpha = ctx.saved_tensors
        N, S = alpha.shape
        d_rgb_s = torch.empty_like(rgb_s)
        d_alpha = torch.empty_like(alpha)
        P = engine._ptr
        g_rgb = None if g_rgb is None else engine.f32c(g_rgb)
        g_acc = None if g_acc is None else engine.f32c(g_acc)
        with torch.cuda.device(alpha.device):
            B.check(B.load().psn_composite_bwd(P(rgb_s), P(alpha), N, S, 1 if ctx.white else 0, P(g_rgb), P(g_acc), P(d_rgb_s),
                                               P(d_alpha), engine._stream()), "psn_composite_bwd")
        return d_rgb_s, d_alpha, None


def composite(rgb_s, alpha, white=True):
    """rgb [N,3], acc [N] from per-sample rgb_s [N,S,3], alpha [N,S] (rendering.py:196-197,214-216), differentiable."""
    return Composite.apply(rgb_s, alpha, white)
