"""Host-side plumbing between the drop-in nn.Modules and the C ABI: packed-weight handles (cached per
parameter version), a grow-only device workspace, and thin typed wrappers around each entry point.
PyTorch is used only for device memory and the current stream."""
import ctypes as C

import torch

from . import _binding as B

# The default is the fastest program that holds the north star's 1e-4 gate (tests/test_gpu_parity_at_size.py): the tcgen05 kernels
# with the mixed radiance program and the two-level surface march.  'fp32' (FFMA kernels) is the debug / cross-check path.
_DEFAULT_PRECISION = [B.PREC_TC_TWOLEVEL]


def set_default_precision(name):
    """'tc_two_level' (default: 'tc_mixed' + the two-level surface march: single-pass occupancy over all march proposals, full
    re-evaluation where the sign scan can tell the difference - depths / masks identical to 'tc_mixed'), 'tc_mixed' ('tc' with the
    appearance side of the radiance program in single fp16 passes: rgb within 1e-5 of 'tc', everything else bit-identical), 'tc'
    (tcgen05, fp16 hi/lo split operands everywhere, fp32 accumulate) or 'fp32' (FFMA kernels: debug / cross-check path)."""
    _DEFAULT_PRECISION[0] = B.PRECISIONS[name]


def tc_available():
    return bool(B.load().psn_has_tensor_path())


def default_precision():
    return _DEFAULT_PRECISION[0]


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, C.c_void_p):
        return t
    return C.c_void_p(t.data_ptr())


def f32c(t, device=None):
    """contiguous fp32 CUDA tensor (copying only when needed)."""
    if device is not None and t.device != device:
        t = t.to(device)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class PackedMLP:
    """Owns one psn_mlp handle."""

    def __init__(self, kind, weights, biases, octaves=0, skip=-1, final_act=0, rescale=1.0):
        B.require_device()
        lib = B.load()
        n = len(weights)
        self._keep = [f32c(w.detach()) for w in weights] + [f32c(b.detach()) for b in biases]
        ws, bs = self._keep[:n], self._keep[n:]
        desc = B.MlpDesc(kind, n, int(octaves), int(skip), int(final_act), float(rescale))
        in_dims = (C.c_int * n)(*[w.shape[1] for w in ws])
        out_dims = (C.c_int * n)(*[w.shape[0] for w in ws])
        wp = (C.c_void_p * n)(*[w.data_ptr() for w in ws])
        bp = (C.c_void_p * n)(*[b.data_ptr() for b in bs])
        h = C.c_void_p()
        with torch.cuda.device(ws[0].device):
            B.check(lib.psn_mlp_create(C.byref(desc), in_dims, out_dims, wp, bp, _stream(), C.byref(h)), "psn_mlp_create")
        # No host synchronisation: the pack kernels read the staging copies on the current stream, and torch's caching allocator only
        # hands their memory to later work of the same stream, so dropping the references here is safe (train steps re-pack every step).
        self.handle = h
        self.device = ws[0].device
        self._keep = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                B.load().psn_mlp_free(self.handle)
                self.handle = None
        except Exception:
            pass


class _Workspace:
    def __init__(self):
        self.buf = {}

    def get(self, device, nbytes):
        # one buffer per (device, stream): calls on different streams may run concurrently and must not share scratch
        key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
        cur = self.buf.get(key)
        if cur is None or cur.numel() < nbytes:
            self.buf[key] = None
            cur = torch.empty(int(nbytes * 1.1) + 4096, dtype=torch.uint8, device=device)
            self.buf[key] = cur
        return cur


_WS = _Workspace()


def workspace(device, op, n_rays=1, n_samples=1, n_lights=1):
    need = B.load().psn_workspace_bytes(op.encode(), int(n_rays), int(n_samples), int(n_lights))
    if need < 0:
        B.check(-1, "psn_workspace_bytes")
    return _WS.get(device, need)


def _params_version(module):
    return tuple((p.data_ptr(), p._version) for p in module.parameters()) + \
        tuple((b.data_ptr(), b._version) for b in module.buffers())


def cached_pack(module, builder):
    """Re-pack only when a parameter was modified in place or replaced (version / pointer change)."""
    key = _params_version(module)
    cache = module.__dict__.get("_psn_pack")
    if cache is None or cache[0] != key:
        cache = (key, builder())
        module.__dict__["_psn_pack"] = cache
    return cache[1]


# ---- typed wrappers --------------------------------------------------------------------------------------------
def occupancy(geo, pts, out_kind=B.OUT_ALPHA, precision=None):
    pts = f32c(pts).reshape(-1, 3)
    M = pts.shape[0]
    out = torch.empty(M, dtype=torch.float32, device=pts.device)
    with torch.cuda.device(pts.device):
        B.check(B.load().psn_occupancy(geo.handle, _ptr(pts), M, out_kind, _ptr(out),
                                       default_precision() if precision is None else precision, _stream()), "psn_occupancy")
    return out


def infer_occ(geo, pts, width, precision=None):
    pts = f32c(pts).reshape(-1, 3)
    M = pts.shape[0]
    out = torch.empty(M, width, dtype=torch.float32, device=pts.device)
    with torch.cuda.device(pts.device):
        B.check(B.load().psn_infer_occ(geo.handle, _ptr(pts), M, _ptr(out),
                                       default_precision() if precision is None else precision, _stream()), "psn_infer_occ")
    return out


def gradient(geo, pts, precision=None):
    pts = f32c(pts).reshape(-1, 3)
    M = pts.shape[0]
    out = torch.empty(M, 3, dtype=torch.float32, device=pts.device)
    ws = workspace(pts.device, "gradient", M)
    with torch.cuda.device(pts.device):
        B.check(B.load().psn_gradient(geo.handle, _ptr(pts), M, _ptr(out), _ptr(ws), ws.numel(),
                                      default_precision() if precision is None else precision, _stream()), "psn_gradient")
    return out


def radiance(geo, app, pts, views, precision=None):
    pts = f32c(pts).reshape(-1, 3)
    views = f32c(views).reshape(-1, 3)
    M = pts.shape[0]
    rgb = torch.empty(M, 3, dtype=torch.float32, device=pts.device)
    alpha = torch.empty(M, dtype=torch.float32, device=pts.device)
    ws = workspace(pts.device, "radiance", M)
    with torch.cuda.device(pts.device):
        B.check(B.load().psn_radiance(geo.handle, app.handle, _ptr(pts), _ptr(views), M, _ptr(rgb), _ptr(alpha), _ptr(ws),
                                      ws.numel(), default_precision() if precision is None else precision, _stream()),
                "psn_radiance")
    return rgb, alpha


def rays_from_pixels(pixels, R, origin, fx, fy, cx, cy, stage2=False):
    """pixels [N,2] float CUDA; R 3x3 / origin 3 host floats."""
    pixels = f32c(pixels).reshape(-1, 2)
    N = pixels.shape[0]
    cam = (C.c_float * 16)(*([float(x) for x in R] + [float(x) for x in origin] + [float(fx), float(fy), float(cx), float(cy)]))
    dirs = torch.empty(N, 3, dtype=torch.float32, device=pixels.device)
    with torch.cuda.device(pixels.device):
        B.check(B.load().psn_rays_from_pixels(_ptr(pixels), N, cam, 1 if stage2 else 0, _ptr(dirs), _stream()),
                "psn_rays_from_pixels")
    return dirs


def raymarch(geo, origin, dirs, near, radius, n_steps, n_secant=8, tau=0.5, precision=None):
    dirs = f32c(dirs).reshape(-1, 3)
    N = dirs.shape[0]
    depth = torch.empty(N, dtype=torch.float32, device=dirs.device)
    ws = workspace(dirs.device, "raymarch", N, n_steps)
    o = (C.c_float * 3)(*[float(x) for x in origin])
    with torch.cuda.device(dirs.device):
        B.check(B.load().psn_raymarch(geo.handle, o, _ptr(dirs), N, float(near), float(radius), int(n_steps), int(n_secant),
                                      float(tau), _ptr(depth), _ptr(ws), ws.numel(),
                                      default_precision() if precision is None else precision, _stream()), "psn_raymarch")
    return depth


def render_unisurf(geo, app, origin, dirs, prm, noise=None, want_sample_depth=False, precision=None):
    dirs = f32c(dirs).reshape(-1, 3)
    N = dirs.shape[0]
    dev = dirs.device
    S = prm.steps_in + prm.steps_out
    rgb = torch.empty(N, 3, dtype=torch.float32, device=dev)
    acc = torch.empty(N, dtype=torch.float32, device=dev)
    normal = torch.empty(N, 3, dtype=torch.float32, device=dev)
    mask = torch.empty(N, dtype=torch.uint8, device=dev)
    depth = torch.empty(N, dtype=torch.float32, device=dev)
    sd = torch.empty(N, S, dtype=torch.float32, device=dev) if want_sample_depth else None
    if noise is not None:
        noise = f32c(noise, dev).reshape(N, S)
    ws = workspace(dev, "unisurf", N, max(S, prm.march_steps))
    o = (C.c_float * 3)(*[float(x) for x in origin])
    with torch.cuda.device(dev):
        B.check(B.load().psn_render_unisurf(geo.handle, app.handle, o, _ptr(dirs), N, C.byref(prm), _ptr(noise), _ptr(rgb),
                                            _ptr(acc), _ptr(normal), _ptr(mask), _ptr(depth), _ptr(sd), _ptr(ws), ws.numel(),
                                            default_precision() if precision is None else precision, _stream()),
                "psn_render_unisurf")
    return {"rgb": rgb, "acc": acc, "normal": normal, "mask": mask.bool(), "depth": depth, "sample_depth": sd}


def shadow_visibility(geo, surf, lights, lnear=0.1, lfar=3.5, n_steps=128, box=1.1, precision=None, return_stats=False):
    """vis [L, Ns] (rendering.py:378-408).  return_stats=True also returns {'evaluated': in-box samples the MLP kernel ran on,
    'nominal': L * Ns * n_steps} (one device-to-host read: measurements only)."""
    surf = f32c(surf).reshape(-1, 3)
    lights = f32c(lights, surf.device).reshape(-1, 3)
    Ns, L = surf.shape[0], lights.shape[0]
    vis = torch.empty(L, Ns, dtype=torch.float32, device=surf.device)
    ws = workspace(surf.device, "shadow", Ns, n_steps, L)
    with torch.cuda.device(surf.device):
        B.check(B.load().psn_shadow_visibility(geo.handle, _ptr(surf), _ptr(lights), Ns, L, float(lnear), float(lfar),
                                               int(n_steps), float(box), _ptr(vis), _ptr(ws), ws.numel(),
                                               default_precision() if precision is None else precision, _stream()),
                "psn_shadow_visibility")
    if return_stats:
        nominal = Ns * L * int(n_steps)
        head = ws[:24].view(torch.int64).tolist() if nominal > 0 else [0, 0, 0]  # counters the library leaves at the head of ws
        culled = bool(head[2] & 0xffffffff)
        return vis, {"evaluated": int(head[1]) if culled else nominal, "nominal": nominal, "culled": culled}
    return vis


def composite(rgb_s, alpha, white_background=True):
    alpha = f32c(alpha)
    N, S = alpha.shape
    rgb_s = f32c(rgb_s).reshape(N, S, 3)
    rgb = torch.empty(N, 3, dtype=torch.float32, device=alpha.device)
    acc = torch.empty(N, dtype=torch.float32, device=alpha.device)
    with torch.cuda.device(alpha.device):
        B.check(B.load().psn_composite(_ptr(rgb_s), _ptr(alpha), N, S, 1 if white_background else 0, _ptr(rgb), _ptr(acc),
                                       _stream()), "psn_composite")
    return rgb, acc
