#!/usr/bin/env python
"""bench.py — headline benchmark of the PS-NeRF hot path on B200 (contract: DESIGN.md "Measurement").

Headline workload = the metric's own configuration, 512 x 512 rays x 128 samples x 96 lights: ONE relit view through
psnerf_b200.pipeline.extract_and_shade, i.e. the reference's per-view chain  stage1/shape_extract.py --visibility  ->
stage2/eval.py --light_batch 96  without the .npy hand-off:
    ray generation, 512-step surface march + 8 secant refinements, analytic normals   (stage1/model/rendering.py:297-361)
    shadow-ray visibility: surface points x 96 lights x 128 march samples              (rendering.py:378-408, the rays x samples x lights loop)
    stage-2 shading of the same points under the same 96 lights                        (stage2/model/renderer.py:110-266)
metric: Msamples/s = surface points x 128 samples x 96 lights / second  (the unit of the fused shadow-march + shade pass).
The stage-1 volume render of the same view (BASELINE configs[1], 512 x 512 x 128 spp, lights = 1) is timed in the same run and
reported under "secondary", with its own roofline and parity block.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision auto|tc_two_level|tc_mixed|tc|fp32] [--impl reference]

N > 1 (torchrun): weak scaling - N views per step, every rank runs the chain on its 128-ray tiles of every view and ONE NCCL
all_gather of the packed per-pixel rows closes the step.  The same line carries "multi_gpu": one view split N ways (strong scaling),
BASELINE config 4 (stage-2 view sharded by surface pixels) and config 5 (data-parallel train step, strong + weak).
--impl reference: the CPU oracle port of the same chain on a bounded strided sample of the same view, all host threads.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 512
L_LIGHTS, S_SHADOW, MARCH_EXTRACT = 96, 128, 512   # shape_extract: 512 march steps (rendering.py:302); 128 shadow steps (rendering.py:380)
S_IN, S_OUT, MARCH = 96, 32, 256                   # secondary: 128 samples/ray, bear.yaml ray_marching_steps
MFLOP_OCC = 0.918016      # occupancy sample, logit row only (2 * 459,008 MAC): what the occupancy kernels compute
MFLOP_RAD = 2.509824      # radiance sample: fwd 524,544 + reverse 459,008 + app 271,360 MAC
MFLOP_PAIR = 1.04704      # stage-2 (point, light) pair: visibility MLP
MFLOP_POINT = 0.282368    # stage-2 per-point nets
PROF_TAGS = ["occ_march", "occ_secant", "radiance", "gradient", "shadow", "s2_vis", "s2_point", "occ_other"]
SAMPLE_GRID = 24          # CPU legs / parity: a SAMPLE_GRID^2 strided sub-grid of the full view (same hit fraction as the view)
METRIC = "Msamples/sec (rays x samples x lights)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"tflops": d.get("bf16_tflops_sustained", 1400.0), "tflops_burst": d.get("bf16_tflops", 1590.0),
                "hbm_gbs": d.get("hbm_gbs", 6650.0), "src": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(r[0])) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = sorted(float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w": pw[len(pw) // 2] if pw else None, "samples": len(sm)}


# ---- the synthetic scene -----------------------------------------------------------------------------------------
def scene(view):
    from psnerf_b200 import synth
    pose = synth.look_at_pose(20.0 + 37.0 * view, 10.0 + 3.0 * (view % 5))
    return synth.intrinsics(H, W), pose


def scene_lights(pose):
    from psnerf_b200 import synth
    return synth.lights(L_LIGHTS, axis=tuple((-pose[0, :3, 2]).tolist()))  # upper hemisphere about the view axis


def stage1_cfg():
    from psnerf_b200 import synth
    return synth.stage1_cfg(num_points_in=S_IN, num_points_out=S_OUT, ray_marching_steps=MARCH)


def state_dicts():
    """Reference-constructor weights under torch.manual_seed(0) for both stages (CPU tensors; the same on every arm)."""
    from psnerf_b200 import synth
    from psnerf_b200.stage1 import NeuralNetwork
    from psnerf_b200.stage2 import PSNetwork
    cfg = stage1_cfg()
    torch.manual_seed(0)
    sd1 = {k: v.detach().clone() for k, v in NeuralNetwork(cfg).state_dict().items()}
    conf = synth.stage2_conf()
    torch.manual_seed(0)
    sd2 = {k: v.detach().clone() for k, v in PSNetwork(conf).state_dict().items()}
    return cfg, sd1, conf, sd2


def build_models(dev, precision):
    from psnerf_b200.stage1 import NeuralNetwork, Renderer
    from psnerf_b200.stage2 import PSNetwork
    cfg, sd1, conf, sd2 = state_dicts()
    net = NeuralNetwork(cfg)
    net.load_state_dict(sd1)
    net = net.eval()  # geometric init: sphere-like occupancy, ~18 % of the rays hit the surface
    net.precision = precision
    rend = Renderer(net, cfg, device=dev)
    ps = PSNetwork(conf)
    ps.load_state_dict(sd2)
    ps = ps.to(dev).eval()
    ps.precision = precision
    return cfg, net, rend, conf, ps


def sample_pixels():
    """[1, SAMPLE_GRID^2, 2] integer pixels on a regular sub-grid of the whole view and their indices in the x-major pixel order."""
    step = W // SAMPLE_GRID
    xs = torch.arange(SAMPLE_GRID) * step + step // 2
    gx, gy = torch.meshgrid(xs, xs, indexing="ij")
    pix = torch.stack([gx, gy], -1).long().view(1, -1, 2)
    return pix, (pix[0, :, 0] * H + pix[0, :, 1])


# ---- CPU legs: the oracle port of the same chain ----------------------------------------------------------------------------------
def oracle_relit_sample(O, cfg, sd1, conf, sd2, pix, K, pose, lights):
    """The reference chain on the sample pixels: shape_extract(visibility) -> PSNetwork.forward, both through the oracle port."""
    shp = O.shape_extract(sd1, cfg, pix, K, pose, visibility=True, light_dir=lights, ray_steps=MARCH_EXTRACT)
    Ks = torch.eye(4).unsqueeze(0)
    Ks[0, 0, 0] = Ks[0, 1, 1] = K[0, 0, 0]
    Ks[0, 0, 2], Ks[0, 1, 2] = K[0, 0, 2], K[0, 1, 2]
    inp = {"intrinsics": Ks, "uv": pix.float(), "pose": pose, "object_mask": shp["mask"], "surface_mask": shp["mask"],
           "points": shp["points"], "normal": shp["normal"], "light_direction": lights}
    with torch.no_grad():
        out = O.psnetwork_forward(sd2, conf, inp)
    return shp, out


def sample_text(n_surf):
    return ("%dx%d strided sub-grid of the 512x512 view (every %dth pixel in x and y: the view's own hit fraction, %d surface points) "
            "x 128 shadow samples x 96 lights, oracle port of shape_extract(visibility) -> PSNetwork.forward"
            % (SAMPLE_GRID, SAMPLE_GRID, W // SAMPLE_GRID, n_surf))


def run_reference(args):
    """CPU arm: the oracle port of the reference chain (oracle/psnerf_oracle.py) on the bounded sample, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import psnerf_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, sd1, conf, sd2 = state_dicts()
    K, pose = scene(0)
    lights = scene_lights(pose)
    pix, _ = sample_pixels()
    times, n_surf = [], 0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        shp, _ = oracle_relit_sample(O, cfg, sd1, conf, sd2, pix, K, pose, lights)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
        n_surf = int(shp["mask"].sum())
    ms = 1e3 * sum(times) / len(times)
    val = n_surf * S_SHADOW * L_LIGHTS / (ms / 1e3) / 1e6
    sample = sample_text(n_surf)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_text() + ", CPU oracle port of the reference path", "sample": sample},
            "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_text():
    return ("relit view 512x512 x 128spp x 96L: shape_extract (512 march steps + 8 secant, analytic normals) + shadow-ray visibility "
            "(surface points x 96 lights x 128 samples) + stage-2 shading x 96 lights, one pipeline.extract_and_shade call per view")


def _stats(a, b):
    """max-abs / relative-L2 / PSNR of GPU values a against oracle values b (CPU tensors of equal shape)."""
    if a.numel() == 0:
        return {"max_abs": 0.0, "rel_l2": 0.0, "psnr": 100.0}
    d = (a.double() - b.double())
    mse = float((d ** 2).mean())
    return {"max_abs": float(d.abs().max()), "rel_l2": float(d.norm() / b.double().norm().clamp_min(1e-30)),
            "psnr": 100.0 if mse == 0 else float(20.0 * torch.log10(torch.tensor(1.0)) - 10.0 * torch.log10(torch.tensor(mse)))}


def cpu_baseline_and_parity(gpu_shape, gpu_out, gpu_render, K, pose, lights):
    """Runs the oracle port once on the strided sample (about 10-20 s of CPU work): its time is the cpu_baseline of the headline, its
    outputs are the parity reference for the GPU results AT THE SAME PIXELS of the full-view run (rays are independent).  The same is
    done for the secondary workload (stage-1 unisurf render at 96+32 samples, 256 march steps)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import psnerf_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, sd1, conf, sd2 = state_dicts()
    pix, idx = sample_pixels()
    lights = lights.cpu()
    O.shape_extract(sd1, cfg, pix[:, :32], K, pose, ray_steps=64)  # warm-up of the CPU kernels
    t0 = time.perf_counter()
    shp, out = oracle_relit_sample(O, cfg, sd1, conf, sd2, pix, K, pose, lights)
    dt = time.perf_counter() - t0
    n_surf = int(shp["mask"].sum())
    cpu = {"value": n_surf * S_SHADOW * L_LIGHTS / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
           "sample": sample_text(n_surf) + " (torch CPU fp32, %d threads, 1 pass)" % cores}
    # ---- parity of the headline chain
    g_mask = gpu_shape["mask"][0][idx].cpu()
    agree = g_mask == shp["mask"][0]
    both = agree & shp["mask"][0]
    par = {"pixels": int(idx.numel()), "surface_points_oracle": n_surf, "mask_agree": float(agree.float().mean()),
           "gate": "north star: 1e-4 relative; compared on the pixels whose hit / miss decision agrees (flips sit at |occupancy - 0.5| < rounding)"}
    par["points"] = _stats(gpu_shape["points"][0][idx].cpu()[both], shp["points"][0][both])
    par["normal"] = _stats(gpu_shape["normal"][0][idx].cpu()[both], shp["normal"][0][both])
    par["shadow_visibility"] = _stats(gpu_shape["visibility"][:, idx].cpu()[:, both], shp["visibility"][:, both])
    for k_gpu, name in (("sg_rgb_values", "rgb"), ("sg_diffuse_albedo_values", "albedo"), ("normal_pred", "normal_pred"),
                        ("visibility", "s2_visibility"), ("sg_specular_rgb_values", "specular")):
        a = gpu_out[k_gpu][:, idx].cpu()[:, agree]
        b = out[k_gpu].reshape(a.shape[0], -1, a.shape[-1])[:, agree]
        par[name] = _stats(a, b)
    # ---- stage 2 on IDENTICAL inputs: the oracle's PSNetwork.forward fed with the GPU's own surface (points / normals / mask) at the
    # sample pixels, so that only the shading kernels are compared (the chain above also carries the 2^9-frequency encoding of the
    # ~1e-5 surface-point differences between the two stage-1 results)
    Ks = torch.eye(4).unsqueeze(0)
    Ks[0, 0, 0] = Ks[0, 1, 1] = K[0, 0, 0]
    Ks[0, 0, 2], Ks[0, 1, 2] = K[0, 0, 2], K[0, 1, 2]
    gm = gpu_shape["mask"][:, idx].cpu()
    inp_same = {"intrinsics": Ks, "uv": pix.float(), "pose": pose, "object_mask": gm, "surface_mask": gm,
                "points": gpu_shape["points"][:, idx].cpu(), "normal": gpu_shape["normal"][:, idx].cpu(), "light_direction": lights}
    with torch.no_grad():
        same = O.psnetwork_forward(sd2, conf, inp_same)
    par["stage2_identical_inputs"] = {}
    for k_gpu, name in (("sg_rgb_values", "rgb"), ("sg_diffuse_albedo_values", "albedo"), ("normal_pred", "normal_pred"),
                        ("visibility", "s2_visibility"), ("sg_specular_rgb_values", "specular")):
        a = gpu_out[k_gpu][:, idx].cpu()
        par["stage2_identical_inputs"][name] = _stats(a, same[k_gpu].reshape(a.shape))
    # ---- secondary: stage-1 volume render (configs[1]) on the same sample pixels
    t0 = time.perf_counter()
    ref = O.unisurf_render(sd1, cfg, pix, K, pose, it=100000)
    dt2 = time.perf_counter() - t0
    m2 = gpu_render["mask_pred"][idx].cpu()
    ag2 = m2 == ref["mask_pred"]
    par2 = {"pixels": int(idx.numel()), "mask_agree": float(ag2.float().mean()),
            "rgb": _stats(gpu_render["rgb"][0][idx].cpu()[ag2], ref["rgb"][0][ag2]),
            "normal": _stats(gpu_render["normal_pred"][0][idx].cpu()[ag2], ref["normal_pred"][0][ag2]),
            "acc": _stats(gpu_render["acc_map"][0][idx].cpu()[ag2], ref["acc_map"][0][ag2])}
    cpu2 = {"value": idx.numel() * (S_IN + S_OUT) / dt2 / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
            "sample": "the same %d pixels, unisurf render at 96+32 samples / 256 march steps" % idx.numel()}
    return cpu, par, cpu2, par2


def ncu_capture(kernel_substr, workload):
    """{dram_bytes_total, file} of ONE launch of a kernel on this bench's workload from the newest committed ncu summary under
    profiles/ whose capture carries the matching "workload" tag, or (None, None)."""
    pdir = os.path.join(ROOT, "profiles")
    try:
        cands = sorted(f for f in os.listdir(pdir) if "_ncu_" in f and f.endswith(".json"))
    except OSError:
        return None, None
    for name in reversed(cands):
        try:
            with open(os.path.join(pdir, name)) as f:
                caps = [c for c in json.load(f) if kernel_substr in c.get("kernel", "") and c.get("workload") == workload]
            if caps:
                big = max(caps, key=lambda c: c["metrics"]["gpu__time_duration.sum"]["value"])
                return big["dram_bytes_total"], name
        except Exception:
            continue
    return None, None


def _time_cuda(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def collect_kernels(lib, steps):
    nt = len(PROF_TAGS)
    pl, pms, prow = (C.c_int64 * nt)(), (C.c_double * nt)(), (C.c_double * nt)()
    lib.psn_profile_collect(nt, pl, pms, prow)
    kern = {}
    for i, t in enumerate(PROF_TAGS):
        if pl[i]:
            kern[t] = {"launches_per_step": pl[i] / steps, "ms_per_launch": pms[i] / pl[i], "ms_per_step": pms[i] / steps,
                       "rows_per_launch": prow[i] / pl[i]}
    return kern


# ---- secondary / other BASELINE configurations --------------------------------------------------------------------------------------
def secondary_stage1_render(lib, dev, rend, precision, view, peaks, steps):
    """BASELINE configs[1]: Renderer.unisurf over the 512x512 view at 96+32 samples/ray, 256 march steps (the round-1 headline)."""
    from psnerf_b200 import synth
    K, pose = view
    pix = synth.pixel_grid_xmajor(H, W).to(dev)

    def step():
        return rend(pix, K, pose, None, "unisurf", add_noise=False, eval_=True, it=100000)
    for _ in range(2):
        step()
    lib.psn_profile_enable(1)
    ms = _time_cuda(step, reps=steps)
    kern = collect_kernels(lib, steps + 1)
    lib.psn_profile_enable(0)
    units = H * W * (S_IN + S_OUT)
    sec = {"workload": "stage1-unisurf-render 512x512x128spp (BASELINE configs[1]): %d march steps + 8 secant, %d+%d samples/ray, lights=1"
                       % (MARCH, S_IN, S_OUT),
           "ms_per_step": ms, "value": units / (ms / 1e3) / 1e6, "unit": "Msamples/s", "kernels": kern}
    rad = kern.get("radiance")
    if rad:
        tf = rad["rows_per_launch"] * MFLOP_RAD * 1e6 / (rad["ms_per_launch"] * 1e-3) / 1e12
        traffic, src = ncu_capture("k_tc_rad", "stage1_render")
        issued = {"tc": 3.0, "fp32": None}.get(precision, (3.0 * 0.918 + 1.0 * (MFLOP_RAD - 0.918)) / MFLOP_RAD)
        sec["roofline"] = {"bound": "tensor", "kernel": "k_tc_rad (geo fwd + analytic normal + app MLP)", "achieved": tf, "peak": peaks["tflops"],
                           "unit": "TFLOP/s", "frac": tf / peaks["tflops"], "issued_frac": None if issued is None else issued * tf / peaks["tflops"],
                           "traffic": traffic, "traffic_source": src, "flops_per_row": MFLOP_RAD * 1e6}
    occ = kern.get("occ_march")
    if occ:
        tfo = occ["rows_per_launch"] * MFLOP_OCC * 1e6 / (occ["ms_per_launch"] * 1e-3) / 1e12
        sec["march"] = {"ms": occ["ms_per_launch"], "achieved_TFLOPs": tfo, "frac": tfo / peaks["tflops"],
                        "note": "all N x 256 proposals; under tc_two_level the first level runs in ONE fp16 pass, so algorithmic and issued FLOPs coincide there"}
    return sec, step()


def other_workloads(dev, precision, rend, ps, view, lights, peaks):
    """The remaining BASELINE configurations at N = 1, timed once each with device-resident inputs (CUDA events)."""
    from psnerf_b200 import synth, engine
    out = {}
    # configs[2]: stage-2 shading 512x512 x 96 lights on the all-surface synthetic points
    inp = synth.stage2_input(H, W, L_LIGHTS, all_surface=True)
    inp = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inp.items()}
    ms = _time_cuda(lambda: ps(inp))
    pairs = H * W * L_LIGHTS
    tf = (pairs * MFLOP_PAIR + H * W * MFLOP_POINT) * 1e6 / (ms * 1e-3) / 1e12
    out["stage2_shade_512x512x96L_all_surface"] = {"ms": ms, "Mpairs_per_s": pairs / ms / 1e3, "algorithmic_TFLOPs": tf,
                                                   "frac_of_peak": tf / peaks["tflops"],
                                                   "hbm_output_GBps": 3 * pairs * 3 * 4 / (ms * 1e-3) / 1e9}
    del inp
    # the shadow pass alone, box-culled (default) against the unculled evaluation of every step (what the reference executes)
    K, pose = view
    g, _ = rend._geo_app()
    origin, dirs = rend._rays(synth.pixel_grid_xmajor(H, W).to(dev), K, pose)
    d = engine.raymarch(g, origin, dirs, 2.0, 2.0, MARCH_EXTRACT, 8, 0.5, rend.model._prec())
    obj, pts = rend._surface(d, origin, dirs)
    surf = pts[obj].contiguous()
    lt = lights.to(dev)
    ms_c = _time_cuda(lambda: engine.shadow_visibility(g, surf, lt, precision=rend.model._prec()), reps=2)
    vis_c, st = engine.shadow_visibility(g, surf, lt, precision=rend.model._prec(), return_stats=True)
    os.environ["PSNERF_B200_SHADOW_UNCULLED"] = "1"
    try:
        ms_u = _time_cuda(lambda: engine.shadow_visibility(g, surf, lt, precision=rend.model._prec()), reps=1)
        vis_u = engine.shadow_visibility(g, surf, lt, precision=rend.model._prec())
    finally:
        os.environ.pop("PSNERF_B200_SHADOW_UNCULLED", None)
    out["shadow_visibility_96L_x128"] = {
        "surface_points": int(surf.shape[0]), "nominal_samples": st["nominal"], "evaluated_samples": st["evaluated"],
        "ms_box_culled": ms_c, "ms_every_step": ms_u, "max_abs_diff_culled_vs_every_step": float((vis_c - vis_u).abs().max()),
        "Msamples_per_s_nominal": st["nominal"] / ms_c / 1e3,
        "tensor_TFLOPs_box_culled": st["evaluated"] * MFLOP_OCC * 1e6 / (ms_c * 1e-3) / 1e12,
        "tensor_TFLOPs_every_step": st["nominal"] * MFLOP_OCC * 1e6 / (ms_u * 1e-3) / 1e12}
    del vis_c, vis_u
    out.update(train_steps(dev, rend, ps, view, world=1, rank=0))
    return out


def train_steps(dev, rend, ps, view, world, rank, n_px=4096, tag="", dist=None):
    """BASELINE configs[4]: stage-2 train step over n_px in-mask pixels x 96 lights (+ 8 vis-train lights, xyz jitter), forward +
    losses + backward (+ ONE all_reduce of the gradients when world > 1) + Adam; and its stage-1 analogue over n_px rays x 128 samples."""
    from psnerf_b200 import synth, sharding
    from psnerf_b200.stage2.loss import MainLoss, NormalLoss
    out = {}
    ps.train()
    side = int(round(n_px ** 0.5))
    n_px = side * side
    tin = synth.stage2_input(side, side, L_LIGHTS, all_surface=True, seed=11 + rank)
    tin = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in tin.items()}
    gen = torch.Generator().manual_seed(1 + rank)
    lraw = torch.nn.Parameter(synth.lights(L_LIGHTS).to(dev) + 0.01)
    linten = torch.nn.Parameter(torch.full((L_LIGHTS, 1), 2.0, device=dev))
    tin["light_vis_train"] = synth.lights(8, seed=5).to(dev)
    tin["vis_train_gt"] = torch.rand(8, n_px, generator=gen).to(dev)
    tin["visibility"] = torch.rand(L_LIGHTS, n_px, generator=gen).to(dev)
    gt = {"rgb": torch.rand(L_LIGHTS, n_px, 3, generator=gen).to(dev)}
    lm, ln = MainLoss(1.0, "L1", 0.05, 0.01, 1.0), NormalLoss(1.0, 0.05)
    opt = torch.optim.Adam(list(ps.parameters()) + [lraw, linten], lr=1e-6)

    def s2_step():
        tin["light_direction"] = torch.nn.functional.normalize(lraw, p=2, dim=-1)
        tin["light_intensity"] = linten
        o = ps(tin)
        loss = lm(o, gt, tin)["loss"] + ln(o)["loss"]
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if world > 1:
            sharding.allreduce_gradients(ps, world, extra_params=(lraw, linten))  # the light tables are reduced with the model
        opt.step()
    ms3 = _time_dist(s2_step, 7, dev, dist)
    out["stage2_train_step%s" % tag] = {"pixels_per_rank": n_px, "lights": L_LIGHTS, "vis_train_lights": 8, "ms_fwd_bwd_adam": ms3,
                                        "Mpairs_per_s": world * n_px * L_LIGHTS / ms3 / 1e3}
    ps.eval()
    del opt
    try:
        from psnerf_b200.stage1 import Loss
        K, pose = view
        net = rend.model
        net.train()
        gen = torch.Generator().manual_seed(2 + rank)
        pix_all = synth.pixel_grid_xmajor(H, W)
        sel = torch.randperm(H * W, generator=gen)[:n_px]
        pix_t = pix_all[:, sel].to(dev)
        rgb_gt = torch.rand(1, n_px, 3, generator=gen).to(dev)
        n_gt = torch.nn.functional.normalize(torch.randn(1, n_px, 3, generator=gen), dim=-1).to(dev)
        n_mask = torch.rand(1, n_px, generator=gen).to(dev) > 0.5
        m_gt = (torch.rand(1, n_px, generator=gen) > 0.5).float().to(dev)
        m_valid = torch.ones(1, n_px, dtype=torch.bool, device=dev)
        crit = Loss(1.0, 0.01, 0.05, 0.1, device=dev)
        opt1 = torch.optim.Adam(net.parameters(), lr=1e-7)  # tiny rate: the timing loop must not walk the field away

        def s1_step():
            o = rend(pix_t, K, pose, None, "unisurf", add_noise=True, eval_=False, it=100000)
            loss = crit(o, rgb_gt, n_gt, n_mask, o["acc_map"], m_gt, m_valid)["loss"]
            opt1.zero_grad(set_to_none=True)
            loss.backward()
            if world > 1:
                sharding.allreduce_gradients(net, world)
            opt1.step()
        ms4 = _time_dist(s1_step, 5, dev, dist)
        out["stage1_train_step%s" % tag] = {"rays_per_rank": n_px, "samples_per_ray": S_IN + S_OUT, "ms_fwd_bwd_adam": ms4,
                                            "Msamples_per_s": world * n_px * (S_IN + S_OUT) / ms4 / 1e3}
        net.eval()
        del opt1
        torch.cuda.empty_cache()
    except Exception as e:  # an extra: never let it take the headline line down
        out["stage1_train_step%s" % tag] = {"error": repr(e)[:300]}
    return out


def _time_dist(fn, reps, dev, dist):
    """fn timed with CUDA events after two warm-up calls (the first use of a kernel pays its lazy module load).  Every repetition has
    its own event pair and the MEDIAN is reported: the train steps allocate a multi-GB activation tape per step, and a one-off
    allocator stall (cudaMalloc / cudaFree of that block, 200-300 ms) inside one of three repetitions otherwise triples the mean
    (r2: 83 / 93 / 373 ms in one process).  With a process group: barrier on both sides, max over ranks."""
    fn()
    fn()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    reps = max(int(reps), 3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    per = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    ms = per[len(per) // 2]
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms


def multi_gpu_block(dev, rend, ps, view, lights, rank, world, td, ms_single_view):
    """What only exists at N > 1: one view split N ways (strong scaling of the headline chain), BASELINE config 4
    (one stage-2 view sharded by surface pixels, gathered) and config 5 (data-parallel train steps, strong and weak)."""
    from psnerf_b200 import pipeline, synth
    K, pose = view
    lt = lights.to(dev)
    out = {}
    ms = _time_dist(lambda: pipeline.extract_and_shade_sharded(rend, ps, H, W, K, pose, lt, rank, world), 3, dev, td)
    out["strong_one_view"] = {"ms_per_view": ms, "note": "the headline chain on ONE 512x512x96L view, rays dealt over the ranks, one "
                                                         "all_gather of the packed rows; compare with ms_per_step of the N=1 run"}
    inp = synth.stage2_input(H, W, L_LIGHTS, all_surface=False, seed=9, mask_frac=0.25)
    inp = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inp.items()}
    ms4 = _time_dist(lambda: pipeline.render_stage2_view_sharded(ps, inp, lt, rank, world), 3, dev, td)
    n_surf = int(inp["surface_mask"].sum())
    out["config4_stage2_view_sharded_96L"] = {"ms_per_view": ms4, "surface_points": n_surf, "Mpairs_per_s": n_surf * L_LIGHTS / ms4 / 1e3,
                                              "note": "render_stage2_view_sharded: surface pixels balanced over the ranks, 96 lights, one "
                                                      "all_gather of [N, 6L+..] rows"}
    del inp
    out.update(train_steps(dev, rend, ps, view, world, rank, n_px=4096, tag="_ddp_weak_4096_per_rank", dist=td))
    out.update(train_steps(dev, rend, ps, view, world, rank, n_px=max(4096 // world, 64), tag="_ddp_strong_4096_total", dist=td))
    return out


_JSON_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to stdout when
    NCCL_DEBUG=VERSION is set, as it is on the GPU boxes), so file descriptor 1 is pointed at stderr for the whole run and the
    JSON line goes to a private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


DTYPES = {"tc": "f16x3-split operands, f32 accumulate", "fp32": "f32",
          "tc_mixed": "f16x3-split operands (softplus stacks) + f16 single pass (appearance side), f32 accumulate",
          "tc_two_level": "f16x3-split operands (softplus stacks; march proposals away from the threshold in one f16 pass) + f16 single pass "
                          "(appearance side), f32 accumulate"}


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default="auto", choices=["auto", "tc", "tc_mixed", "tc_two_level", "fp32"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sample-grid", type=int, default=SAMPLE_GRID, help="side of the strided pixel sub-grid the CPU legs run on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="headline steps only (launch lists / profiles of the headline)")
    args = ap.parse_args()
    globals()["SAMPLE_GRID"] = args.sample_grid
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)

    from psnerf_b200 import _binding as B, engine, pipeline, sharding, synth
    if not os.path.exists(B.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = world > 1
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    td = None
    if dist:
        import torch.distributed as td
        td.init_process_group("nccl", device_id=dev)
    lib = B.load()
    precision = args.precision
    if precision == "auto":
        precision = "tc_two_level"  # the engine default: fastest program that holds the parity gate (tests/test_gpu_parity_at_size.py)
    cfg, net, rend, conf, ps = build_models(dev, precision)
    N = H * W
    n_views = world                                            # weak scaling: one view's worth of rays per GPU
    # N copies of the SAME bench view: surface coverage and in-box shadow fractions differ between poses by more than 10 %, which
    # would turn the weak-scaling ratio into a statement about the scenes; nothing is cached between the copies
    views = [scene(0) for _ in range(n_views)]
    lights = [scene_lights(pose).to(dev) for (_, pose) in views]
    pix_all = synth.pixel_grid_xmajor(H, W)                    # long [1,N,2]
    shard = sharding.shard_indices(N, rank, world)             # this rank's 128-ray tiles of every view (round-robin)
    pix_host = pix_all[:, shard].contiguous().pin_memory()
    pix_dev = pix_host.to(dev)
    n_local = pix_dev.shape[1]
    layout, width, row_w = pipeline.relit_row_layout(ps, L_LIGHTS, True)
    gather_buf = torch.empty(world, n_views * n_local, row_w, device=dev) if dist else None
    # e2e result read-back: the relit images [L,n,3], albedo, normal, mask and the shadow visibilities [L,n] (what eval.py / shape_extract.py save)
    col_rgb = next(off for key, _, _, off, _ in layout if key == "sg_rgb_values")
    col_alb = next(off for key, _, _, off, _ in layout if key == "sg_diffuse_albedo_values")
    host_cols = 3 * L_LIGHTS + 3 + (row_w - width)
    out_host = torch.empty(n_views * n_local, host_cols, dtype=torch.float32).pin_memory()

    def step(e2e=False):
        rows = []
        for (K, pose), lt in zip(views, lights):
            p = pix_host.to(dev, non_blocking=True) if e2e else pix_dev
            rows.append(pipeline.extract_and_shade_rows(rend, ps, H, W, K, pose, lt, p))
        res = rows[0] if len(rows) == 1 else torch.cat(rows, 0)
        if dist:
            td.all_gather_into_tensor(gather_buf.view(-1), res.view(-1))  # the single NCCL pixel gather
        if e2e:  # one contiguous device buffer -> one DMA into pinned host memory
            packed = torch.cat([res[:, col_rgb:col_rgb + 3 * L_LIGHTS], res[:, col_alb:col_alb + 3], res[:, width:]], 1)
            out_host.copy_(packed, non_blocking=True)
        return res

    def timed(e2e, steps):
        if dist:
            td.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            step(e2e)
        b.record()
        torch.cuda.synchronize()
        if dist:
            td.barrier()
        ms = a.elapsed_time(b) / steps
        if dist:
            t = torch.tensor([ms], device=dev)
            td.all_reduce(t, op=td.ReduceOp.MAX)
            ms = float(t)
        return ms

    for _ in range(args.warmup):
        res = step(False)
    torch.cuda.synchronize()
    # surface points of this rank's rays over all views -> units of the whole job (counted once, outside the timed region), and the
    # number of in-box shadow samples the MLP kernel evaluates for them (device-side list length, read back here for the roofline)
    n_surf = (res[:, width] > 0.5).sum().to(torch.int64)
    evaluated_rank = 0
    g_geo, _ = rend._geo_app()
    for v in range(n_views):
        rows_v = res[v * n_local:(v + 1) * n_local]
        surf_v = rows_v[:, width + 1:width + 4][rows_v[:, width] > 0.5].contiguous()
        if surf_v.shape[0] > 0:
            _, st = engine.shadow_visibility(g_geo, surf_v, lights[v], precision=rend.model._prec(), return_stats=True)
            evaluated_rank += st["evaluated"]
    if dist:
        td.all_reduce(n_surf)
    n_surf = int(n_surf)
    del res
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # device-resident timing, with per-kernel CUDA-event profiling of the same steps
    lib.psn_profile_enable(1)
    l0 = lib.psn_launch_count()
    ms = timed(False, args.steps)
    launches = (lib.psn_launch_count() - l0) / args.steps
    kern = collect_kernels(lib, args.steps)
    lib.psn_profile_enable(0)
    step(True)
    ms_e2e = timed(True, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    units = n_surf * S_SHADOW * L_LIGHTS  # whole job per step: every surface point's 128 shadow samples towards every light
    value = units / (ms / 1e3) / 1e6
    mg = None
    if dist:
        mg = multi_gpu_block(dev, rend, ps, views[0], lights[0], rank, world, td, ms)
    if rank == 0:
        peaks = load_peaks()
        roof = None
        sh = kern.get("shadow")
        if sh:
            # dominant kernel: k_tc_occ over the box-culled shadow list; its rows are the in-box samples (device-side count)
            rows_per_launch = evaluated_rank / max(sh["launches_per_step"], 1.0)
            tf = rows_per_launch * MFLOP_OCC * 1e6 / (sh["ms_per_launch"] * 1e-3) / 1e12
            traffic, src = ncu_capture("k_tc_occ", "relit_view_shadow")
            roof = {"bound": "tensor", "kernel": "k_tc_occ<MODE_OUT> over the box-culled shadow list (GEN_SHADOW_LIST), %s" % precision,
                    "achieved": tf, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": tf / peaks["tflops"],
                    "issued_frac": None if precision == "fp32" else 3.0 * tf / peaks["tflops"], "traffic": traffic, "traffic_source": src,
                    "peak_source": peaks["src"], "flops_per_row": MFLOP_OCC * 1e6, "rows_per_launch": rows_per_launch,
                    "share_of_step": sh["ms_per_step"] / ms,
                    "note": "rows = in-box shadow samples actually evaluated (the reference also evaluates the out-of-box steps and then "
                            "zeroes them, rendering.py:402-404); nominal rows would be %d per view" % (n_surf // n_views * S_SHADOW * L_LIGHTS)}
        line = {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPES[precision],
                "data": "synthetic",
                "config": {"workload": workload_text(), "views_per_step": n_views, "rays_per_gpu_per_step": n_views * n_local,
                           "surface_points_per_step": n_surf, "units_per_step": units,
                           "units": "surface points x 128 shadow samples x 96 lights (SURVEY.md 8d); the march proposals (rays x 512) and the "
                                    "stage-2 pairs (points x 96) of the same step are not counted",
                           "parallelism": "ray-shard x%d" % world, "precision": precision,
                           "l2": "per-step working set (268 MB of march occupancies, >= 600 MB of shadow lists, 1 GB of outputs) exceeds the 126 MB L2",
                           "weights": "reference constructors, torch.manual_seed(0) (stage 1: geometric init)"},
                "gpu_launches": launches, "clocks": clocks,
                "e2e": {"value": units / (ms_e2e / 1e3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": n_views * n_local * 2 * 8, "d2h_bytes_per_step": n_views * n_local * host_cols * 4},
                "roofline": roof, "kernels": kern}
        if mg is not None:
            line["multi_gpu"] = mg
        if world == 1 and not args.no_secondary:
            K, pose = views[0]
            sec, gpu_render = secondary_stage1_render(lib, dev, rend, precision, views[0], peaks, 3)
            line["secondary"] = {"stage1_unisurf_512x512x128spp": sec}
            if not args.no_cpu_baseline:
                full = step(False)
                gshape, gout = pipeline.unpack_relit_rows(full, ps, L_LIGHTS, True)
                cpu, par, cpu2, par2 = cpu_baseline_and_parity(gshape, gout, gpu_render, K, pose, lights[0])
                line["cpu_baseline"] = cpu
                line["parity"] = par
                sec["cpu_baseline"] = cpu2
                sec["parity"] = par2
                del full, gshape, gout
            del gpu_render
            if not args.no_extras:
                try:
                    line["other_workloads"] = other_workloads(dev, precision, rend, ps, views[0], lights[0], peaks)
                except Exception as e:
                    line["other_workloads"] = {"error": repr(e)[:400]}
        emit(line)
    if dist:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
