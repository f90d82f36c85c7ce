"""CPU-only checks of the host side: drop-in constructors / state-dict layout, C-ABI symbol export, loud failure
without a GPU."""
import ctypes

import numpy as np
import pytest
import torch

import util
from psnerf_b200 import synth


def test_stage1_state_dict_layout():
    from psnerf_b200.stage1 import NeuralNetwork
    m = NeuralNetwork(synth.stage1_cfg())
    sd = m.state_dict()
    assert sum(v.numel() for v in sd.values()) == 802490
    for l in range(9):
        assert sd["lin%d.weight_g" % l].shape[1] == 1
        assert sd["lin%d.weight_v" % l].shape[0] == sd["lin%d.bias" % l].shape[0]
    assert sd["lin3.weight_v"].shape == (217, 256) and sd["lin4.weight_v"].shape == (256, 256)
    assert sd["lin8.weight_v"].shape == (257, 256) and sd["lina0.weight_v"].shape == (256, 289)
    # weight_g initialised to the row norms of weight_v, like torch.nn.utils.weight_norm
    assert torch.allclose(sd["lin2.weight_g"], sd["lin2.weight_v"].norm(2, dim=1, keepdim=True))


def test_stage2_state_dict_layout():
    from psnerf_b200.stage2 import PSNetwork
    m = PSNetwork(synth.stage2_conf())
    sd = m.state_dict()
    assert sum(v.numel() for v in sd.values()) == 667947
    assert sd["visibility_net.linears.5.weight"].shape == (256, 382)
    assert sd["albedo_net.linears.3.weight"].shape == (128, 191)
    assert sd["rough_net.linears.2.weight"].shape == (27, 64)
    assert not m.sgbasis.lobe.requires_grad
    assert np.allclose(sd["sgbasis.lobe"].numpy(), np.exp(np.arange(2, 11)), rtol=1e-6)


def test_constructor_weights_equal_reference():
    g1, g2 = util.golden("stage1_net"), util.golden("stage2_shade")
    _, s1 = util.stage1_state_dicts()
    _, s2 = util.stage2_state_dicts()
    np.testing.assert_array_equal(util.checksum(s1["init"]), g1["init_checksum"])
    np.testing.assert_array_equal(util.checksum(s2["init"]), g2["init_checksum"])


def test_stage1_loss_equals_reference_loss():
    """psnerf_b200.stage1.Loss on the reference's own training outputs vs the loss terms the REAL reference Loss produced."""
    from psnerf_b200.stage1 import Loss
    g = util.golden("stage1_grads")
    out = {k: torch.from_numpy(g["out_" + k]) for k in util.S1_TRAIN_KEYS}
    gt = util.s1_loss_ground_truth(out["rgb"].shape[1])
    terms = Loss(1.0, 0.01, 0.05, 0.1)(out, gt["rgb"], gt["normal"], gt["norm_mask"], out["acc_map"], gt["mask"], gt["mask_valid"])
    for k in ("fullrgb_loss", "grad_loss", "normal_loss", "mask_loss", "loss"):
        assert abs(float(terms[k]) - float(g["loss_" + k])) < 1e-5 * max(1.0, abs(float(g["loss_" + k]))), k


def test_abi_exports_every_declared_symbol(lib_built):
    from psnerf_b200 import _binding
    names = _binding.declared_symbols()
    assert len(names) >= 17 and "psn_render_unisurf" in names and "psn_shade_stage2" in names
    for n in names:
        assert hasattr(lib_built, n), n
    assert lib_built.psn_version() >= 100


def test_abi_rejects_bad_arguments_without_gpu(lib_built):
    assert lib_built.psn_workspace_bytes(b"no-such-op", 1, 1, 1) < 0
    assert b"unknown op" in lib_built.psn_last_error()
    assert lib_built.psn_workspace_bytes(b"unisurf", 1024, 256, 1) > 1024 * 256 * 4
    rc = lib_built.psn_occupancy(None, None, 0, 0, None, 0, None)
    assert rc == -1 and b"null" in lib_built.psn_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    from psnerf_b200.stage1 import NeuralNetwork
    from psnerf_b200.stage2 import PSNetwork
    m = NeuralNetwork(synth.stage1_cfg())
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        m(torch.zeros(4, 3), only_occupancy=True)
    p = PSNetwork(synth.stage2_conf())
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        p(synth.stage2_input(4, 4, 2))


def test_fused_optimizers_host_contract(lib_built):
    """psnerf_b200.optim.{Adam,SparseAdam} without a GPU: argument validation of the C entry points, torch-compatible state_dict
    layout (a torch.optim.Adam checkpoint resumes and the state goes back), MultiStepLR drives param_groups, and no CPU fallback."""
    import ctypes as C
    from psnerf_b200 import _binding as B
    from psnerf_b200 import optim
    h = B.AdamHyper(1e-3, 0.9, 0.999, 1e-8, 0.0, 0)
    assert lib_built.psn_adam_step(None, 0, C.byref(h), None) < 0 and b"step" in lib_built.psn_last_error()
    h.step = 1
    assert lib_built.psn_adam_step(None, 0, C.byref(h), None) == 0  # empty list: nothing to launch
    assert lib_built.psn_adam_step(None, 2, C.byref(h), None) < 0
    h.beta1 = 1.0
    assert lib_built.psn_sparse_adam_step(None, None, None, 4, 3, None, None, 2, C.byref(h), None) < 0
    h.beta1, h.weight_decay = 0.9, 0.1
    assert lib_built.psn_sparse_adam_step(None, None, None, 4, 3, None, None, 2, C.byref(h), None) < 0
    assert b"weight decay" in lib_built.psn_last_error()
    h.weight_decay = 0.0
    assert lib_built.psn_sparse_adam_step(None, None, None, 4, 3, None, None, 0, C.byref(h), None) == 0  # K = 0
    assert lib_built.psn_sparse_adam_step(None, None, None, 4, 3, None, None, 2, C.byref(h), None) < 0  # null pointers

    p = torch.nn.Parameter(torch.ones(4, 3))
    ot = torch.optim.Adam([p], lr=1e-2)
    p.grad = torch.ones(4, 3)
    ot.step()
    ot.step()
    om = optim.Adam([p], lr=1e-2)
    om.load_state_dict(ot.state_dict())
    assert optim._step_int(om.state[p]["step"]) == 2 and set(om.state[p]) == {"step", "exp_avg", "exp_avg_sq"}
    back = torch.optim.Adam([p], lr=1e-2)
    back.load_state_dict(om.state_dict())
    assert torch.equal(back.state[p]["exp_avg"], ot.state[p]["exp_avg"])
    sched = torch.optim.lr_scheduler.MultiStepLR(om, [1], gamma=0.5)
    om.param_groups[0]["params"][0].grad = None
    om.step()  # nothing has a gradient: a no-op even without a GPU
    sched.step()
    assert om.param_groups[0]["lr"] == pytest.approx(5e-3)
    p.grad = torch.ones(4, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        om.step()
    with pytest.raises(RuntimeError, match="amsgrad"):
        optim.Adam([p], amsgrad=True)
    with pytest.raises(RuntimeError, match="dense"):
        optim.SparseAdam([p]).step()
    e = torch.nn.Embedding(5, 3, sparse=True)
    e(torch.tensor([1, 1, 3])).sum().backward()
    with pytest.raises(RuntimeError, match="CUDA"):
        optim.SparseAdam(list(e.parameters())).step()
    with pytest.raises(RuntimeError, match="sparse"):
        optim.Adam(list(e.parameters())).step()


def test_stage2_train_step_order_of_operations():
    """psnerf_b200.stage2.train_step = trainer.py:394-410: losses summed, both optimizers zeroed before and stepped after ONE backward,
    the light optimizer left alone once the light table is frozen."""
    from psnerf_b200.stage2 import train_step
    w = torch.nn.Parameter(torch.tensor([1.0, 2.0]))
    table = torch.nn.Embedding(4, 3)
    log = []

    class Opt(torch.optim.SGD):
        def __init__(self, params, name):
            super().__init__(params, lr=0.1)
            self.name = name

        def zero_grad(self, set_to_none=True):
            log.append(self.name + ".zero")
            super().zero_grad(set_to_none)

        def step(self, closure=None):
            log.append(self.name + ".step")
            return super().step(closure)

    def model(inp):
        log.append("forward")
        return {"y": (w * inp["x"]).sum() + table.weight.sum() * 0.0 + table.weight[1, 0]}

    def loss(out, gt, inp):
        return {"loss": (out["y"] - gt["y"]) ** 2, "sg_rgb_loss": out["y"].detach()}

    def loss_n(out):
        return {"loss": out["y"] * 0.5}

    sg, lo = Opt([w], "sg"), Opt(table.parameters(), "light")
    y = 3.0 + float(table.weight[1, 0])  # before the step
    lo_out, ln_out = train_step(model, loss, {"x": torch.tensor([1.0, 1.0])}, {"y": torch.tensor(0.0)}, sg, loss_n, lo, table)
    assert log == ["forward", "sg.zero", "light.zero", "sg.step", "light.step"]
    assert float(w[0]) != 1.0 and ln_out is not None and abs(float(lo_out["loss"]) - ((y - 0.0) ** 2 + 0.5 * y)) < 1e-3 * max(1.0, abs(y))
    table.weight.requires_grad_(False)
    del log[:]
    train_step(model, loss, {"x": torch.tensor([1.0, 1.0])}, {"y": torch.tensor(0.0)}, sg, None, lo, table)
    assert log == ["forward", "sg.zero", "sg.step"]


def test_shape_files_round_trip_in_the_reference_layout(tmp_path):
    """points / normal / mask / visibility .npy files of one view: written from the x-major stage-1 arrays as shape_extract.py:144-162
    does, read back row-major as dataset.py:100-114 does - pixel (y, x) of the file is ray x*h + y of the renderer."""
    from psnerf_b200 import pipeline
    h, w, L = 6, 6, 3
    g = torch.Generator().manual_seed(4)
    n = h * w
    shape = {"points": torch.randn(1, n, 3, generator=g), "normal": torch.randn(1, n, 3, generator=g),
             "mask": torch.rand(1, n, generator=g) > 0.5, "visibility": torch.rand(L, n, generator=g)}
    pipeline.save_shape_view(str(tmp_path), 7, shape, h, w)
    pts = np.load(tmp_path / "points" / "view_07.npy")
    msk = np.load(tmp_path / "mask" / "view_07.npy")
    vis = np.load(tmp_path / "visibility" / "view_07.npy")
    assert pts.shape == (h, w, 3) and pts.dtype == np.float32 and msk.shape == (h, w) and msk.dtype == bool and vis.shape == (L, h, w)
    for (y, x) in [(0, 0), (2, 5), (5, 1)]:
        ray = x * h + y
        assert np.array_equal(pts[y, x], shape["points"][0, ray].numpy()) and msk[y, x] == bool(shape["mask"][0, ray])
        assert np.array_equal(vis[:, y, x], shape["visibility"][:, ray].numpy())
    back = pipeline.load_shape_view(str(tmp_path), 7, with_visibility=True)
    assert back["points"].shape == (1, n, 3) and back["surface_mask"].dtype == torch.bool and back["visibility"].shape == (L, n)
    assert back["img_res"] == [h, w]
    row_major = torch.arange(n).view(w, h).t().reshape(-1)  # ray index of row-major pixel k
    assert torch.equal(back["points"][0], shape["points"][0][row_major])
    assert torch.equal(back["normal"][0], shape["normal"][0][row_major])
    assert torch.equal(back["surface_mask"][0], shape["mask"][0][row_major])
    assert torch.equal(back["visibility"], shape["visibility"][:, row_major])
    pipeline.save_shape_view(str(tmp_path), 8, {k: v for k, v in shape.items() if k != "visibility"}, h, w)
    assert not (tmp_path / "visibility" / "view_08.npy").exists()


@pytest.mark.skipif(not __import__("os").path.exists("/root/reference/stage1/model/checkpoints.py"),
                    reason="needs the reference checkout (authoring container only)")
def test_reference_checkpoint_io_round_trips_the_drop_in_modules(tmp_path):
    """The reference's OWN CheckpointIO (stage1/model/checkpoints.py, loaded by path, unmodified) saves and restores the drop-in
    NeuralNetwork and the fused Adam exactly as it does its own classes (stage1/train.py:66-72): same 'model' / 'optimizer' entries,
    scalars passed through."""
    import importlib.util
    from psnerf_b200 import optim
    from psnerf_b200.stage1 import NeuralNetwork
    spec = importlib.util.spec_from_file_location("psnerf_ref_checkpoints", "/root/reference/stage1/model/checkpoints.py")
    ck = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ck)
    torch.manual_seed(3)
    net = NeuralNetwork(synth.stage1_cfg())
    opt = optim.Adam(net.parameters(), lr=1e-4)
    for p in list(net.parameters())[:2]:  # give the optimizer some state without running a CUDA step
        opt.state[p] = {"step": 5, "exp_avg": torch.full_like(p, 0.25), "exp_avg_sq": torch.full_like(p, 0.5)}
    io = ck.CheckpointIO(str(tmp_path), model=net, optimizer=opt)
    io.save("model.pt", epoch_it=7, it=1234, loss_val_best=0.5)
    raw = torch.load(tmp_path / "model.pt", weights_only=False)
    assert set(raw) == {"model", "optimizer", "epoch_it", "it", "loss_val_best"}
    assert set(raw["model"]) == set(net.state_dict()) and len(raw["optimizer"]["state"]) == 2
    torch.manual_seed(4)
    net2 = NeuralNetwork(synth.stage1_cfg())
    opt2 = optim.Adam(net2.parameters(), lr=1e-4)
    scalars = ck.CheckpointIO(str(tmp_path), model=net2, optimizer=opt2).load("model.pt")
    assert scalars == {"epoch_it": 7, "it": 1234, "loss_val_best": 0.5}
    for (k, a), b in zip(net.state_dict().items(), net2.state_dict().values()):
        assert torch.equal(a, b), k
    p0 = list(net2.parameters())[0]
    assert opt2.state[p0]["step"] == 5 and float(opt2.state[p0]["exp_avg"].mean()) == 0.25


def test_arange_pixels_is_xmajor():
    from psnerf_b200.stage1 import arange_pixels
    import psnerf_oracle as O
    loc, sc = arange_pixels((3, 5))
    assert loc.shape == (1, 15, 2) and loc[0, 1].tolist() == [0, 1] and loc[0, 3].tolist() == [1, 0]
    lo, so = O.arange_pixels((3, 5))
    assert torch.equal(loc, lo) and torch.equal(sc, so)


def test_bench_traffic_lookup_matches_workload():
    """roofline.traffic comes from a committed ncu capture of the named kernel ON THE NAMED WORKLOAD (captures carry a tag)."""
    import bench
    t_rad, src = bench.ncu_capture("k_tc_rad", "stage1_render")
    assert src and 1e9 < t_rad < 3e11
    assert bench.ncu_capture("k_tc_rad", "no_such_workload") == (None, None)
    assert bench.ncu_capture("k_no_such_kernel", "stage1_render") == (None, None)


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm the driver times next to the GPU arm): exactly one JSON line on stdout with the
    contract keys; library chatter must not reach stdout."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--sample-grid", "10"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_sync_free_losses_equal_the_boolean_gather_formulation():
    """psnerf_b200.stage{1,2} losses compute their masked means as where-sums over counts (no `mask.sum() == 0` / `x[mask]` host
    synchronisation per term).  Against the reference's own formulation - boolean gathers + nn.L1Loss / mse / BCELoss (stage1/model/
    losses.py:24-66, stage2/model/loss.py:24-104) - on random data: same values, same gradients, and an EMPTY mask gives a zero term with
    zero gradients instead of a NaN."""
    import torch.nn.functional as F
    from psnerf_b200.stage1 import Loss
    from psnerf_b200.stage2.loss import MainLoss, NormalLoss
    gen = torch.Generator().manual_seed(21)
    L, N = 5, 97
    # ---- stage 2
    def s2_case(mask):
        out = {"sg_rgb_values": torch.rand(L, N, 3, generator=gen).requires_grad_(True), "network_object_mask": mask, "object_mask": torch.ones_like(mask),
               "albedo_values": torch.rand(1, N, 3, generator=gen).requires_grad_(True), "albedo_jitter": torch.rand(1, N, 3, generator=gen),
               "rough_values": torch.rand(1, N, 9, generator=gen).requires_grad_(True), "rough_jitter": torch.rand(1, N, 9, generator=gen),
               "vis_train": torch.rand(2, N, 3, generator=gen).requires_grad_(True), "visibility": torch.rand(L, N, 3, generator=gen),
               "normal_pred": torch.randn(1, N, 3, generator=gen).requires_grad_(True), "normal_values": torch.randn(1, N, 3, generator=gen)}
        inp = {"visibility": torch.rand(L, N, generator=gen), "vis_train_gt": torch.rand(2, N, generator=gen), "light_vis_train": torch.rand(2, 3, generator=gen)}
        gt = {"rgb": torch.rand(L, N, 3, generator=gen)}
        return out, inp, gt
    mask = torch.rand(1, N, generator=gen) > 0.4
    out, inp, gt = s2_case(mask)
    terms = MainLoss(1.0, "L1", 0.05, 0.01, 1.0)(out, gt, inp)
    nterms = NormalLoss(1.0, 0.0)(out)
    m = mask.expand(L, -1)
    ref_rgb = F.l1_loss(out["sg_rgb_values"][m].reshape(-1, 3), gt["rgb"][m].reshape(-1, 3))
    ref_alb = F.l1_loss(out["albedo_values"][mask], out["albedo_jitter"][mask])
    ref_rough = F.l1_loss(out["rough_values"][mask], out["rough_jitter"][mask])
    ref_vis = F.l1_loss(out["vis_train"][..., 0][mask.expand(2, -1)].reshape(-1), inp["vis_train_gt"][mask.expand(2, -1)].reshape(-1))
    ref_n = F.mse_loss(out["normal_pred"][mask].reshape(-1, 3), F.normalize(out["normal_values"], dim=-1)[mask].reshape(-1, 3))
    for got, ref in ((terms["sg_rgb_loss"], ref_rgb), (terms["albedo_smooth_loss"], ref_alb), (terms["rough_smooth_loss"], ref_rough),
                     (terms["vis_loss"], ref_vis), (nterms["normal_loss"], ref_n)):
        assert abs(float(got) - float(ref)) < 1e-6 * max(1.0, abs(float(ref)))
    total = terms["loss"] + nterms["loss"]
    ref_total = ref_rgb + 0.05 * ref_alb + 0.01 * ref_rough + ref_vis + ref_n
    leaves = [out["sg_rgb_values"], out["albedo_values"], out["rough_values"], out["vis_train"], out["normal_pred"]]
    g1 = torch.autograd.grad(total, leaves, retain_graph=True)
    g2 = torch.autograd.grad(ref_total, leaves)
    for a, b in zip(g1, g2):
        assert float((a - b).abs().max()) < 1e-7
    out, inp, gt = s2_case(torch.zeros(1, N, dtype=torch.bool))
    t0 = MainLoss(1.0, "L1", 0.05, 0.01, 1.0)(out, gt, inp)["loss"] + NormalLoss(1.0, 0.0)(out)["loss"]
    assert float(t0) == 0.0
    for gz in torch.autograd.grad(t0, [out["sg_rgb_values"], out["normal_pred"]], allow_unused=True):
        assert gz is None or float(gz.abs().max()) == 0.0
    # ---- stage 1
    n = 61
    o1 = {"rgb": torch.rand(1, n, 3, generator=gen).requires_grad_(True), "diff_norm": torch.rand(n, generator=gen).requires_grad_(True),
          "normal_pred": torch.randn(1, n, 3, generator=gen).requires_grad_(True)}
    acc = torch.rand(1, n, generator=gen).requires_grad_(True)
    rgb_gt, n_gt = torch.rand(1, n, 3, generator=gen), torch.randn(1, n, 3, generator=gen)
    n_mask, m_valid = torch.rand(1, n, generator=gen) > 0.5, torch.rand(1, n, generator=gen) > 0.2
    m_gt = (torch.rand(1, n, generator=gen) > 0.5).float()
    t1 = Loss(1.0, 0.01, 0.05, 0.1)(o1, rgb_gt, n_gt, n_mask, acc, m_gt, m_valid)
    ref1 = (F.l1_loss(o1["rgb"], rgb_gt, reduction="sum") / n + 0.01 * o1["diff_norm"].mean()
            + 0.05 * F.l1_loss(o1["normal_pred"][n_mask], n_gt[n_mask], reduction="sum") / float(n_mask.sum())
            + 0.1 * F.binary_cross_entropy(acc[m_valid].clamp(0, 1), m_gt[m_valid]))
    assert abs(float(t1["loss"]) - float(ref1)) < 1e-6 * max(1.0, abs(float(ref1)))
    for a, b in zip(torch.autograd.grad(t1["loss"], [o1["rgb"], o1["normal_pred"], acc], retain_graph=True),
                    torch.autograd.grad(ref1, [o1["rgb"], o1["normal_pred"], acc])):
        assert float((a - b).abs().max()) < 1e-7
    t2 = Loss(1.0, 0.01, 0.05, 0.1)(o1, rgb_gt, n_gt, torch.zeros(1, n, dtype=torch.bool), acc, m_gt, torch.zeros(1, n, dtype=torch.bool))
    assert torch.isfinite(t2["loss"]) and float(t2["normal_loss"]) == 0.0 and float(t2["mask_loss"]) == 0.0
