"""Generate golden fixtures by running the REAL reference (/root/reference) on CPU.

Run in the authoring container only:   python tests/golden/make_golden.py
Writes tests/golden/{stage1_net,stage1_render,stage2_shade}.npz.  Weights are NOT stored: both the reference
constructors and the drop-in constructors are deterministic under torch.manual_seed, so fixtures carry weight
checksums instead and the tests rebuild the weights (tests/test_host.py checks the checksums).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_loader  # noqa: E402
from psnerf_b200 import synth  # noqa: E402

torch.set_num_threads(8)


def checksum(sd):
    return np.array([[float(v.double().sum()), float(v.double().abs().sum())] for _, v in sorted(sd.items())])


def np_(t):
    return t.detach().cpu().numpy()


def stage1_variants(net_mod):
    cfg = synth.stage1_cfg()
    torch.manual_seed(0)
    model = net_mod.NeuralNetwork(cfg)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    sd1 = synth.perturb_state_dict(sd0, rel=0.1, seed=1)
    return cfg, model, {"init": sd0, "trained": sd1}


def make_stage1_net():
    net_mod, rend_mod, _ = ref_loader.load_stage1()
    cfg, model, variants = stage1_variants(net_mod)
    g = torch.Generator().manual_seed(11)
    pts = (torch.rand(96, 3, generator=g) * 2.4 - 1.2)
    views = torch.randn(96, 3, generator=g)
    views = views / views.norm(dim=-1, keepdim=True)
    out = {"pts": np_(pts), "views": np_(views)}
    for name, sd in variants.items():
        model.load_state_dict(sd)
        with torch.no_grad():
            out[name + "_infer_occ"] = np_(model.infer_occ(pts.clone()))
            out[name + "_alpha"] = np_(model(pts.clone(), only_occupancy=True))
            out[name + "_neg_logit"] = np_(model(pts.clone(), return_logits=True))
        out[name + "_grad"] = np_(model.gradient(pts.clone(), tflag=False))
        rgb, a = model(pts.clone(), views.clone(), return_addocc=True)
        out[name + "_rgb"] = np_(rgb)
        out[name + "_rgb_alpha"] = np_(a)
        out[name + "_checksum"] = checksum(sd)
    np.savez_compressed(os.path.join(HERE, "stage1_net.npz"), **out)
    print("stage1_net", {k: v.shape for k, v in out.items()})


def make_stage1_render():
    net_mod, rend_mod, common = ref_loader.load_stage1()
    _, model, variants = stage1_variants(net_mod)
    out = {}
    cases = {
        # name: (h, w, num_points_in, num_points_out, march steps, it)
        "cfg1small": (24, 24, 32, 0, 256, 100000),     # BASELINE config 1 shape at 24x24
        "inout": (20, 20, 16, 8, 128, 100000),         # interval + outside samples, sorted (rendering.py:150-155)
        "early": (16, 16, 16, 8, 64, 1000),            # it <= 5000: full_steps == steps, wide delta
    }
    pose = synth.look_at_pose(25.0, 15.0)
    for vname, sd in variants.items():
        model.load_state_dict(sd)
        for cname, (h, w, s_in, s_out, msteps, it) in cases.items():
            cfg = synth.stage1_cfg(num_points_in=s_in, num_points_out=s_out, ray_marching_steps=msteps)
            rend = rend_mod.Renderer(model, cfg, device=torch.device("cpu"))
            pix = synth.pixel_grid_xmajor(h, w)
            K = synth.intrinsics(h, w)
            with torch.no_grad():
                res = rend(pix, K, pose, torch.eye(4)[None], "unisurf", add_noise=False, eval_=True, it=it)
                ray0 = common.origin_to_world(pix.shape[1], K, pose, None)
                rayd = common.image_points_to_ray(pix, K, pose)
                rayd = rayd / rayd.norm(2, 2).unsqueeze(-1)
                d_i = rend.ray_marching(ray0, rayd, model, n_secant_steps=8, n_steps=[msteps, msteps + 1], rad=cfg["rendering"]["radius"],
                                        depth_range=rend.depth_range)
            key = "%s_%s_" % (vname, cname)
            out[key + "rgb"] = np_(res["rgb"])
            out[key + "normal"] = np_(res["normal_pred"])
            out[key + "acc"] = np_(res["acc_map"])
            out[key + "mask"] = np_(res["mask_pred"])
            out[key + "d_i"] = np_(d_i)
            out[key + "dirs"] = np_(rayd)
            print(key, "hit rays", int(res["mask_pred"].sum()), "/", h * w)
        # shape_extract + shadow visibility (rendering.py:297-408)
        cfg = synth.stage1_cfg()
        rend = rend_mod.Renderer(model, cfg, device=torch.device("cpu"))
        h = w = 14
        pix = synth.pixel_grid_xmajor(h, w)
        lights = synth.lights(3, seed=5, axis=tuple((-pose[0, :3, 2]).tolist()))
        res = rend(pix, synth.intrinsics(h, w), pose, torch.eye(4)[None], "shape_extract", visibility=True, light_dir=lights)
        key = "%s_extract_" % vname
        for k in ("mask", "normal", "points", "visibility"):
            out[key + k] = np_(res[k])
        out[key + "lights"] = np_(lights)
        print(key, "surface", int(res["mask"].sum()), "vis range", float(res["visibility"].min()), float(res["visibility"].max()))
    out["pose"] = np_(pose)
    np.savez_compressed(os.path.join(HERE, "stage1_render.npz"), **out)


def make_stage2():
    m2 = ref_loader.load_stage2()
    conf = synth.stage2_conf()
    torch.manual_seed(0)
    model = m2.PSNetwork(ref_loader.DictConf(conf))
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    sd1 = synth.perturb_state_dict(sd0, rel=0.5, seed=1)
    out = {}
    for vname, sd in {"init": sd0, "trained": sd1}.items():
        model.load_state_dict(sd)
        for cname, (h, w, L, all_surf) in {"multi": (16, 16, 4, False), "single": (12, 12, 1, False), "full": (8, 8, 3, True)}.items():
            inp = synth.stage2_input(h, w, L, all_surface=all_surf)
            if cname == "multi":
                inp["light_intensity"] = torch.tensor([[1.0], [2.0], [0.5], [3.0]])
                inp["light_vis_train"] = synth.lights(2, seed=9)
            if cname == "full":
                inp["light_intensity"] = torch.tensor([[1.0, 2.0, 0.5], [3.0, 1.0, 1.0], [0.3, 0.6, 2.0]])
            torch.manual_seed(123)  # the jitter branch draws torch.normal (renderer.py:212); its outputs are not compared
            with torch.no_grad():
                res = model(inp)
            key = "%s_%s_" % (vname, cname)
            for k in ("sg_rgb_values", "sg_specular_rgb_values", "visibility", "normal_pred", "sg_diffuse_albedo_values",
                      "sg_weight", "vis_train"):
                if k in res:
                    out[key + k] = np_(res[k])
            print(key, {k: tuple(v.shape) for k, v in res.items() if torch.is_tensor(v)})
        out[vname + "_checksum"] = checksum(sd)
    np.savez_compressed(os.path.join(HERE, "stage2_shade.npz"), **out)


def make_stage2_edit():
    """Material editing (albedo_new / basis_new, stage2/eval.py:116-132) and an envmap-style batch of RGB intensities
    (stage2/eval.py:199-203) through the REAL reference PSNetwork."""
    m2 = ref_loader.load_stage2()
    conf = synth.stage2_conf()
    torch.manual_seed(0)
    model = m2.PSNetwork(ref_loader.DictConf(conf))
    sd1 = synth.perturb_state_dict({k: v.clone() for k, v in model.state_dict().items()}, rel=0.5, seed=1)
    model.load_state_dict(sd1)
    out = {}
    for cname, (albedo_new, basis_new) in EDIT_CASES.items():
        inp = edit_case_input()
        torch.manual_seed(123)
        with torch.no_grad():
            res = model(inp, albedo_new=None if albedo_new is None else np.asarray(albedo_new, np.float32), basis_new=basis_new)
        for k in ("sg_rgb_values", "sg_specular_rgb_values", "visibility", "sg_diffuse_albedo_values", "sg_weight"):
            out["%s_%s" % (cname, k)] = np_(res[k])
    out["checksum"] = checksum(sd1)
    # render_model = microfacet (GGX; no shipped conf uses it): REAL reference PSNetwork under the same seed protocol
    confm = synth.stage2_conf(**{"train.render_model": "microfacet"})
    torch.manual_seed(0)
    mm = m2.PSNetwork(ref_loader.DictConf(confm))
    sdm = synth.perturb_state_dict({k: v.clone() for k, v in mm.state_dict().items()}, rel=0.5, seed=1)
    mm.load_state_dict(sdm)
    for cname in ("multi", "single"):
        inp = micro_case_input(cname)
        torch.manual_seed(123)
        with torch.no_grad():
            res = mm(inp)
        for k in ("sg_rgb_values", "sg_specular_rgb_values", "visibility", "sg_diffuse_albedo_values", "normal_pred"):
            out["micro_%s_%s" % (cname, k)] = np_(res[k])
        assert "sg_weight" not in res
    out["micro_checksum"] = checksum(sdm)
    # light grid of the envmap mode: gen_light_xyz / sph2cart taken from the reference file by name (the module itself imports
    # cv2 / imageio, absent here), light_h = 16 as in stage2/eval.py:100
    import ast
    src = open("/root/reference/stage2/utils/eval_utils.py").read()
    tree = ast.parse(src)
    want = {"gen_light_xyz", "sph2cart", "_warn_degree", "_convert_sph_conventions"}
    code = "\n\n".join(ast.get_source_segment(src, n) for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want)
    ns = {"np": np}
    exec(code, ns)
    xyz, areas = ns["gen_light_xyz"](16, 32, envmap_radius=1)
    out["envgrid_xyz"] = xyz.reshape(-1, 3)
    out["envgrid_areas"] = areas.reshape(-1)
    np.savez_compressed(os.path.join(HERE, "stage2_edit.npz"), **out)


# name: (albedo_new, basis_new); shared with tests/util.py
EDIT_CASES = {"albedo": ([0.2, 0.55, 0.31], None), "basis": (None, 4), "both": ([0.05, 0.1, 0.4], 7)}


def edit_case_input():
    h, w, L = 12, 14, 5
    inp = synth.stage2_input(h, w, L, all_surface=False, seed=33, mask_frac=0.5)
    g = torch.Generator().manual_seed(8)
    inp["light_intensity"] = torch.rand(L, 3, generator=g) * 2.0  # env_light[lstart:lend] is [L,3]
    return inp


def micro_case_input(cname):
    if cname == "multi":
        inp = synth.stage2_input(12, 14, 5, all_surface=False, seed=33, mask_frac=0.5)
        inp["light_intensity"] = torch.rand(5, 3, generator=torch.Generator().manual_seed(8)) * 2.0
    else:
        inp = synth.stage2_input(9, 11, 1, all_surface=True, seed=34)
    return inp


S1_TRAIN_CASE = dict(h=12, w=10, s_in=12, s_out=6, msteps=64, it=100000, pose=(25.0, 15.0))
S1_TRAIN_KEYS = ["rgb", "normal_pred", "acc_map", "diff_norm"]


def s2_loss_ground_truth(L, n, Lt):
    g = torch.Generator().manual_seed(12)
    return {"rgb": torch.rand(L, n, 3, generator=g), "visibility": torch.rand(L, n, generator=g), "vis_train_gt": torch.rand(Lt, n, generator=g)}


def make_stage2_losses():
    """MainLoss / NormalLoss of the REAL reference (stage2/model/loss.py) on the outputs of the real PSNetwork for the train-like
    case of make_stage2_grads.  The reference hard-codes .cuda() (loss.py:30,61): torch.Tensor.cuda is replaced by the identity for
    the duration of the call (no source edit)."""
    m2 = ref_loader.load_stage2()
    lossmod = ref_loader._load("model.loss", os.path.join(ref_loader.REF, "stage2/model/loss.py"), "model")
    conf = synth.stage2_conf()
    torch.manual_seed(0)
    model = m2.PSNetwork(ref_loader.DictConf(conf))
    sd1 = synth.perturb_state_dict({k: v.clone() for k, v in model.state_dict().items()}, rel=0.5, seed=1)
    model.load_state_dict(sd1)
    h, w, L, Lt = 12, 10, 5, 2
    inp = synth.stage2_input(h, w, L, all_surface=False, seed=21, mask_frac=0.6)
    inp["light_vis_train"] = synth.lights(Lt, seed=9)
    gt = s2_loss_ground_truth(L, h * w, Lt)
    inp["visibility"], inp["vis_train_gt"] = gt["visibility"], gt["vis_train_gt"]
    ns = int(inp["surface_mask"].sum())
    std = conf["brdf.net.xyz_jitter_std"]
    torch.manual_seed(4242)
    z = torch.normal(0, torch.ones(ns, 3) * std) / std
    torch.manual_seed(4242)
    with torch.no_grad():
        out = model(inp)
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        with torch.no_grad():
            lm = lossmod.MainLoss(1.0, "L1", 0.05, 0.01, 1.0)(out, {"rgb": gt["rgb"]}, inp)
            ln = lossmod.NormalLoss(1.0, 0.05)(out)
    finally:
        torch.Tensor.cuda = orig
    res = {"xyz_noise": np_(z)}
    for k, v in lm.items():
        if v is not None:
            res["main_" + k] = np.array(float(v))
    for k, v in ln.items():
        if v is not None:
            res["normal_" + k] = np.array(float(v))
    np.savez_compressed(os.path.join(HERE, "stage2_loss.npz"), **res)
    print("stage2_loss", {k: float(v) for k, v in res.items() if k != "xyz_noise"})


def make_stage1_phong():
    """Renderer.phong_renderer (rendering.py:228-293; 512 march steps, headlight Phong shading) of the REAL reference."""
    net_mod, rend_mod, common = ref_loader.load_stage1()
    _, model, variants = stage1_variants(net_mod)
    cfg = synth.stage1_cfg()
    out = {}
    for vname, sd in variants.items():
        model.load_state_dict(sd)
        rend = rend_mod.Renderer(model, cfg, device=torch.device("cpu"))
        with torch.no_grad():
            res = rend(synth.pixel_grid_xmajor(12, 12), synth.intrinsics(12, 12), synth.look_at_pose(10.0, 5.0), torch.eye(4)[None],
                       "phong_renderer")
        out[vname + "_rgb"] = np_(res["rgb"])
        print("phong", vname, "shaded pixels", int((res["rgb"] != 1).any(-1).sum()))
    np.savez_compressed(os.path.join(HERE, "stage1_phong.npz"), **out)


def s1_loss_ground_truth(n):
    g = torch.Generator().manual_seed(6)
    return {"rgb": torch.rand(1, n, 3, generator=g), "normal": torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g), dim=-1),
            "norm_mask": torch.rand(1, n, generator=g) > 0.4, "mask": (torch.rand(1, n, generator=g) > 0.5).float(),
            "mask_valid": torch.rand(1, n, generator=g) > 0.1}


def make_stage1_grads():
    """Gradients of the REAL reference's stage-1 training forward (Renderer.unisurf, eval_=False, add_noise=False) w.r.t. every
    parameter of the field: scalar = sum of the differentiable outputs against fixed cotangents.  torch.rand_like (neighbour
    points of the normal-consistency term, rendering.py:204) is recorded through a wrapper so that the restatement can be fed
    the same draw."""
    net_mod, rend_mod, common = ref_loader.load_stage1()
    _, model, variants = stage1_variants(net_mod)
    c = S1_TRAIN_CASE
    cfg = synth.stage1_cfg(num_points_in=c["s_in"], num_points_out=c["s_out"], ray_marching_steps=c["msteps"])
    model.load_state_dict(variants["trained"])
    model.train()
    rend = rend_mod.Renderer(model, cfg, device=torch.device("cpu"))
    pix = synth.pixel_grid_xmajor(c["h"], c["w"])
    K, pose = synth.intrinsics(c["h"], c["w"]), synth.look_at_pose(*c["pose"])
    drawn = []
    orig = torch.rand_like

    def recording(t, *a, **k):
        r = orig(t, *a, **k)
        drawn.append(r.clone())
        return r
    torch.manual_seed(99)
    torch.rand_like = recording
    try:
        out = rend(pix, K, pose, torch.eye(4)[None], "unisurf", add_noise=False, eval_=False, it=c["it"])
    finally:
        torch.rand_like = orig
    assert len(drawn) == 1
    g = torch.Generator().manual_seed(5)
    cot = {k: torch.randn(out[k].shape, generator=g) for k in S1_TRAIN_KEYS}
    scalar = sum((out[k] * cot[k]).sum() for k in S1_TRAIN_KEYS)
    params = dict(model.named_parameters())
    names = sorted(params)
    grads = torch.autograd.grad(scalar, [params[n] for n in names], allow_unused=True)
    res = {"neigh_u": np_(drawn[0]), "scalar": np.array(float(scalar)), "mask": np_(out["mask_pred"])}
    for k in S1_TRAIN_KEYS:
        res["cot_" + k] = np_(cot[k])
        res["out_" + k] = np_(out[k])
    for n, gr in zip(names, grads):
        gr = torch.zeros_like(params[n]) if gr is None else gr
        res["gsum_" + n] = np.array([float(gr.double().sum()), float(gr.double().abs().sum())])
        if gr.numel() <= 300:
            res["g_" + n] = np_(gr)
    res["checksum"] = checksum(variants["trained"])
    # loss terms of the REAL reference Loss (stage1/model/losses.py) on these outputs with seeded ground truth
    losses = ref_loader._load("psnerf_ref_stage1.losses", os.path.join(ref_loader.REF, "stage1/model/losses.py"), "psnerf_ref_stage1")
    gt = s1_loss_ground_truth(out["rgb"].shape[1])
    crit = losses.Loss(1.0, 0.01, 0.05, 0.1, device=torch.device("cpu"))
    with torch.no_grad():
        terms = crit({k: out[k].detach() for k in out if torch.is_tensor(out[k])}, gt["rgb"], gt["normal"], gt["norm_mask"],
                     out["acc_map"].detach(), gt["mask"], gt["mask_valid"])
    for k, v in terms.items():
        res["loss_" + k] = np.array(float(v))
    np.savez_compressed(os.path.join(HERE, "stage1_grads.npz"), **res)
    print("stage1_grads: scalar", float(scalar), "params", len(names), "surface", int(out["mask_pred"].sum()))


def make_stage2_grads():
    """Gradients of the REAL reference's PSNetwork (autograd) for a train-like step: trainable light directions and per-light
    intensities, 2 vis-train lights, xyz jitter; scalar = sum of outputs against fixed cotangents (the reference's loss module
    hard-codes .cuda(), stage2/model/loss.py:30,61, so it cannot run here)."""
    m2 = ref_loader.load_stage2()
    conf = synth.stage2_conf()
    torch.manual_seed(0)
    model = m2.PSNetwork(ref_loader.DictConf(conf))
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    sd1 = synth.perturb_state_dict(sd0, rel=0.5, seed=1)
    model.load_state_dict(sd1)
    model.train()
    h, w, L = 12, 10, 5
    inp = synth.stage2_input(h, w, L, all_surface=False, seed=21, mask_frac=0.6)
    g = torch.Generator().manual_seed(77)
    lraw = torch.randn(L, 3, generator=g)
    lraw.requires_grad_(True)
    inten = (1.0 + torch.rand(L, 1, generator=g)).requires_grad_(True)
    inp["light_direction"] = torch.nn.functional.normalize(lraw, p=2, dim=-1)
    inp["light_intensity"] = inten
    inp["light_vis_train"] = synth.lights(2, seed=9)
    ns = int(inp["surface_mask"].sum())
    std = conf["brdf.net.xyz_jitter_std"]
    torch.manual_seed(4242)
    z = torch.normal(0, torch.ones(ns, 3) * std) / std        # what the reference will draw (renderer.py:212)
    torch.manual_seed(4242)
    out = model(inp)
    keys = ["sg_rgb_values", "normal_pred", "albedo_values", "rough_values", "albedo_jitter", "rough_jitter", "vis_train"]
    cot = {k: torch.randn(out[k].shape, generator=g) for k in keys}
    scalar = sum((out[k] * cot[k]).sum() for k in keys)
    params = dict(model.named_parameters())
    names = [n for n, p_ in params.items() if p_.requires_grad]
    grads = torch.autograd.grad(scalar, [params[n] for n in names] + [lraw, inten], allow_unused=True)
    res = {"xyz_noise": np_(z), "light_raw": np_(lraw), "light_intensity": np_(inten), "scalar": np.array(float(scalar))}
    for k in keys:
        res["cot_" + k] = np_(cot[k])
        res["out_" + k] = np_(out[k])
    for n, gr in zip(names, grads[:-2]):
        gr = torch.zeros_like(params[n]) if gr is None else gr
        res["gsum_" + n] = np.array([float(gr.double().sum()), float(gr.double().abs().sum())])
        if gr.numel() <= 256:
            res["g_" + n] = np_(gr)
    res["g_light_raw"] = np_(grads[-2])
    res["g_light_intensity"] = np_(grads[-1])
    np.savez_compressed(os.path.join(HERE, "stage2_grads.npz"), **res)
    print("stage2_grads: scalar", float(scalar), "params", len(names), "surface", ns)


def general_case():
    """Inputs of the split/merge fixture: a 2500-pixel model_input (3 chunks at the reference's 1024) and per-chunk outputs
    of rank 1, 2, 3 and 4 (the ranks PSNetwork.forward returns: [N] masks, [1,N] flags, [1,N,3] colours, [L,1,N,3] per-light)."""
    g = torch.Generator().manual_seed(41)
    n = 2500
    inp = {"uv": torch.rand(1, n, 2, generator=g), "object_mask": torch.rand(1, n, generator=g) > 0.3,
           "points": torch.randn(1, n, 3, generator=g), "normal": torch.randn(1, n, 3, generator=g),
           "visibility": torch.rand(1, n, 2, generator=g), "intrinsics": torch.eye(4)[None], "pose": torch.eye(4)[None],
           "light_direction": torch.randn(2, 3, generator=g)}
    return inp, n


def general_chunk_outputs(chunk):
    m = chunk["uv"].shape[1]
    return {"mask": chunk["object_mask"].reshape(-1), "flag": chunk["object_mask"].float(), "rgb": chunk["points"] * 2,
            "per_light": chunk["visibility"].permute(2, 0, 1)[..., None].expand(2, 1, m, 3).contiguous(), "none": None}


def make_stage2_general():
    """split_input / merge_output of the REAL stage2/utils/general.py:23-53 on CPU (its .cuda() on the index is made a no-op)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("psnerf_ref_general", os.path.join(ref_loader.REF, "stage2/utils/general.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    inp, n = general_case()
    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        chunks = gen.split_input(inp, n)
    finally:
        torch.Tensor.cuda = saved
    out = {"n_chunks": np.array(len(chunks)), "chunk_sizes": np.array([c["uv"].shape[1] for c in chunks]),
           "keys": np.array(sorted(chunks[0].keys()))}
    for i, c in enumerate(chunks):
        for k in ("uv", "object_mask", "points", "normal", "visibility"):
            out["c%d_%s" % (i, k)] = np.array([float(c[k].double().sum()), float(c[k][0, 0].double().sum()), float(c[k][0, -1].double().sum())])
    merged = gen.merge_output([general_chunk_outputs(c) for c in chunks], n, 1)
    out["merged_keys"] = np.array(sorted(merged.keys()))
    for k, v in merged.items():
        out["m_" + k] = np_(v)
    np.savez_compressed(os.path.join(HERE, "stage2_general.npz"), **out)


def metrics_case():
    g = torch.Generator().manual_seed(77)
    a = torch.rand(9, 11, 3, generator=g).numpy()
    b = (torch.rand(9, 11, 3, generator=g) * 0.1).numpy() + a * 0.9
    m = (torch.rand(9, 11, generator=g) > 0.4).numpy()
    n1 = torch.randn(9, 11, 3, generator=g).numpy()
    n2 = n1 + 0.2 * torch.randn(9, 11, 3, generator=g).numpy()
    n1[0, 0] = 0  # a zero normal (background pixel)
    return a, b, m, n1, n2


def make_metrics():
    """PSNR / MAE of the REAL stage2/utils/metrics.py:16-51 (its skimage / lpips / trimesh imports, unused by these two functions,
    are shimmed with empty modules)."""
    import importlib.util
    import types
    for name in ("skimage", "skimage.metrics", "lpips", "trimesh", "trimesh.proximity", "trimesh.sample"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["skimage.metrics"].structural_similarity = None
    spec = importlib.util.spec_from_file_location("psnerf_ref_metrics", os.path.join(ref_loader.REF, "stage2/utils/metrics.py"))
    met = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(met)
    a, b, m, n1, n2 = metrics_case()
    mae_all, ang_all = met.MAE(n1, n2)
    mae_m, ang_m = met.MAE(n1, n2, m)
    mae_raw, _ = met.MAE(n1 / 3.0, n2, m, normalize=False)
    out = {"psnr": np.array(met.PSNR(a, b)), "psnr_masked": np.array(met.PSNR(a, b, m)), "psnr_same": np.array(met.PSNR(a, a)),
           "mae": np.array(mae_all), "ang": ang_all, "mae_masked": np.array(mae_m), "ang_masked": ang_m, "mae_raw": np.array(mae_raw)}
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)


if __name__ == "__main__":
    make_stage1_net()
    make_stage1_render()
    make_stage2()
    make_stage2_grads()
    make_stage2_edit()
    make_stage1_grads()
    make_stage1_phong()
    make_stage2_losses()
    make_stage2_general()
    make_metrics()
    for f in ("stage1_net", "stage1_render", "stage2_shade", "stage2_grads", "stage2_edit", "stage1_grads", "stage1_phong", "stage2_loss", "stage2_general", "metrics"):
        print(f, os.path.getsize(os.path.join(HERE, f + ".npz")) // 1024, "KB")
