"""Fused Adam / SparseAdam (psn_adam_step, psn_sparse_adam_step through psnerf_b200.optim) against the CPU oracle
restatement (O.adam_step / O.sparse_adam_step, pinned to torch.optim in test_oracle_golden.py) and against torch.optim itself
on the same device.  fp32 tolerance: max |a - b| <= 2e-6 * max |b| per tensor per step."""
import numpy as np
import pytest
import torch

import psnerf_oracle as O
from psnerf_b200 import optim, synth

pytestmark = pytest.mark.gpu
TOL = 2e-6


def _close(a, b, tol=TOL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    assert float(np.abs(a - b).max()) <= tol * max(float(np.abs(b).max()), 1e-30)


# ragged sizes: scalars, sizes that are not multiples of 4, a tensor spanning many blocks, and a mis-aligned view (offset 1 float)
SHAPES = [(1,), (3,), (257, 39), (256, 1), (5, 7), (1024, 256), (2,), (513,)]


@pytest.mark.parametrize("weight_decay", [0.0, 0.01])
def test_adam_vs_oracle_and_torch(weight_decay):
    g = torch.Generator().manual_seed(3)
    base = [torch.randn(s, generator=g) for s in SHAPES]
    flat = torch.zeros(1 + 77, device="cuda")  # a parameter living at a 4-byte offset of its storage: scalar path
    odd0 = torch.randn(77, generator=g)
    flat[1:].copy_(odd0)
    mine = [torch.nn.Parameter(t.clone().cuda()) for t in base] + [torch.nn.Parameter(flat[1:])]
    theirs = [torch.nn.Parameter(t.clone().cuda()) for t in base] + [torch.nn.Parameter(odd0.clone().cuda())]
    ora = [(t.numpy().copy(), np.zeros(t.shape, np.float32), np.zeros(t.shape, np.float32)) for t in base + [odd0]]
    kw = dict(lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay)
    o_mine, o_theirs = optim.Adam(mine, **kw), torch.optim.Adam(theirs, **kw)
    for step in range(1, 6):
        lr = 3e-3 if step < 4 else 3e-4
        o_mine.param_groups[0]["lr"] = o_theirs.param_groups[0]["lr"] = lr
        grads = [torch.randn(p.shape, generator=g) * (0.01 if i % 2 else 30.0) for i, p in enumerate(mine)]
        for a, b, gr in zip(mine, theirs, grads):
            a.grad, b.grad = gr.cuda(), gr.cuda()
        o_mine.step()
        o_theirs.step()
        ora = [O.adam_step(p, gr.numpy(), m, v, step, lr=lr, weight_decay=weight_decay) for (p, m, v), gr in zip(ora, grads)]
        for a, b, (p, m, v) in zip(mine, theirs, ora):
            _close(a.detach().cpu().numpy(), p)
            _close(o_mine.state[a]["exp_avg"].cpu().numpy(), m)
            _close(o_mine.state[a]["exp_avg_sq"].cpu().numpy(), v)
            _close(a.detach().cpu().numpy(), b.detach().cpu().numpy())
    assert flat[0].item() == 0.0  # the neighbour of the mis-aligned view is untouched


def test_adam_skips_params_without_grad_and_counts_steps_per_param():
    a = torch.nn.Parameter(torch.ones(10, device="cuda"))
    b = torch.nn.Parameter(torch.ones(10, device="cuda"))
    a2, b2 = torch.nn.Parameter(a.detach().clone()), torch.nn.Parameter(b.detach().clone())
    o1, o2 = optim.Adam([a, b], lr=0.1), torch.optim.Adam([a2, b2], lr=0.1)
    for step in range(4):
        for p in (a, a2):
            p.grad = torch.full((10,), 0.5 + step, device="cuda")
        for p in (b, b2):
            p.grad = None if step % 2 == 0 else torch.full((10,), -1.0 - step, device="cuda")
        o1.step()
        o2.step()
    assert o1.state[a]["step"] == 4 and o1.state[b]["step"] == 2
    assert a._version >= 4 and b._version >= 2  # in-place updates are visible to autograd and to the packed-weight cache
    _close(a.detach().cpu().numpy(), a2.detach().cpu().numpy())
    _close(b.detach().cpu().numpy(), b2.detach().cpu().numpy())


def test_adam_state_dict_is_interchangeable_with_torch():
    """A torch.optim.Adam state (the reference's OptimizerParameters/*.pth layout) resumes in the fused optimizer and back."""
    g = torch.Generator().manual_seed(9)
    w0 = torch.randn(33, 5, generator=g)
    grads = [torch.randn(33, 5, generator=g).cuda() for _ in range(4)]
    ref = torch.nn.Parameter(w0.clone().cuda())
    o_ref = torch.optim.Adam([ref], lr=1e-2)
    for gr in grads:
        ref.grad = gr
        o_ref.step()
    p = torch.nn.Parameter(w0.clone().cuda())
    o_t = torch.optim.Adam([p], lr=1e-2)
    for gr in grads[:2]:
        p.grad = gr
        o_t.step()
    o_m = optim.Adam([p], lr=1e-2)
    o_m.load_state_dict(o_t.state_dict())
    p.grad = grads[2]
    o_m.step()
    o_t2 = torch.optim.Adam([p], lr=1e-2)
    o_t2.load_state_dict(o_m.state_dict())
    p.grad = grads[3]
    o_t2.step()
    _close(p.detach().cpu().numpy(), ref.detach().cpu().numpy())


def test_adam_whole_model_is_one_launch(lib_built):
    """Every parameter of the stage-1 field (42 tensors, 802 490 values) in ONE kernel launch, equal to torch.optim.Adam."""
    from psnerf_b200.stage1 import NeuralNetwork
    torch.manual_seed(0)
    net = NeuralNetwork(synth.stage1_cfg()).cuda()
    twin = NeuralNetwork(synth.stage1_cfg()).cuda()
    twin.load_state_dict(net.state_dict())
    o1, o2 = optim.Adam(net.parameters(), lr=1e-4), torch.optim.Adam(twin.parameters(), lr=1e-4)
    g = torch.Generator(device="cuda").manual_seed(1)
    for _ in range(3):
        for a, b in zip(net.parameters(), twin.parameters()):
            a.grad = torch.randn(a.shape, generator=g, device="cuda")
            b.grad = a.grad.clone()
        n0 = lib_built.psn_launch_count()
        o1.step()
        assert lib_built.psn_launch_count() - n0 == 1
        o2.step()
    assert sum(p.numel() for p in net.parameters()) == 802490
    for a, b in zip(net.parameters(), twin.parameters()):
        _close(a.detach().cpu().numpy(), b.detach().cpu().numpy())


def test_adam_rejects_what_it_cannot_run():
    p = torch.nn.Parameter(torch.ones(4))  # CPU parameter: no fallback
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        optim.Adam([p]).step()
    with pytest.raises(RuntimeError):
        optim.Adam([torch.nn.Parameter(torch.ones(4, device="cuda"))], amsgrad=True)
    q = torch.nn.Parameter(torch.ones(4, 2, device="cuda"))
    q.grad = torch.ones(4, 2, device="cuda")
    with pytest.raises(RuntimeError):
        optim.SparseAdam([q]).step()  # dense gradient, as torch.optim.SparseAdam refuses


@pytest.mark.parametrize("dim", [3, 1])
def test_sparse_adam_vs_oracle_and_torch(dim):
    """The light-direction [llen,3] and light-intensity [llen,1] tables of stage2/trainer.py:136-165: a batch looks up the lights of
    one view (here also with repeated rows), SparseAdam moves only those rows."""
    g = torch.Generator().manual_seed(17)
    R = 1920  # 20 views x 96 lights
    w0 = torch.randn(R, dim, generator=g)
    e1 = torch.nn.Embedding(R, dim, sparse=True).cuda()
    e2 = torch.nn.Embedding(R, dim, sparse=True).cuda()
    e1.weight.data.copy_(w0)
    e2.weight.data.copy_(w0)
    o1, o2 = optim.SparseAdam(list(e1.parameters()), lr=1e-3), torch.optim.SparseAdam(list(e2.parameters()), lr=1e-3)
    p, m, v = w0.numpy().copy(), np.zeros((R, dim), np.float32), np.zeros((R, dim), np.float32)
    batches = [torch.arange(96) + 96 * 3, torch.arange(96) + 96 * 3, torch.tensor([5, 5, 700, 5, 1919, 0, 700]),
               torch.randint(0, R, (300,), generator=g), torch.arange(96) + 96 * 19]
    for step, idx in enumerate(batches, start=1):
        w = torch.randn(idx.numel(), dim, generator=g)
        for e, o in ((e1, o1), (e2, o2)):
            o.zero_grad()
            out = e(idx.cuda())
            out = torch.nn.functional.normalize(out, dim=-1) if dim == 3 else out * out
            (out * w.cuda()).sum().backward()
        gr = e1.weight.grad
        gi, gv = gr._indices()[0].cpu().numpy(), gr._values().cpu().numpy()
        o1.step()
        o2.step()
        p, m, v = O.sparse_adam_step(p, gi, gv, m, v, step, lr=1e-3)
        _close(e1.weight.detach().cpu().numpy(), p)
        _close(o1.state[e1.weight]["exp_avg"].cpu().numpy(), m)
        _close(o1.state[e1.weight]["exp_avg_sq"].cpu().numpy(), v)
        _close(e1.weight.detach().cpu().numpy(), e2.weight.detach().cpu().numpy())
    untouched = np.setdiff1d(np.arange(R), np.concatenate([b.numpy() for b in batches]))
    assert np.array_equal(e1.weight.detach().cpu().numpy()[untouched], w0.numpy()[untouched])


def test_stage2_train_loop_with_fused_optimizers_reduces_loss():
    """The reference's optimizer arrangement (trainer.py:116-165,394-410): Adam on the PSNetwork, SparseAdam on the light tables,
    MultiStepLR on both; a few steps on a fixed batch must reduce the loss, and track the torch optimizers closely."""
    import util
    from psnerf_b200.stage2 import PSNetwork
    from psnerf_b200.stage2.loss import MainLoss, NormalLoss
    conf, sds = util.stage2_state_dicts()
    inp = synth.stage2_input(48, 48, 12, all_surface=False, seed=5, mask_frac=0.7)
    ci = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in inp.items()}
    ci["light_vis_train"] = synth.lights(4, seed=3).cuda()
    gen = torch.Generator().manual_seed(0)
    gt = {"rgb": torch.rand(12, 48 * 48, 3, generator=gen).cuda()}
    ci["vis_train_gt"] = torch.rand(4, 48 * 48, generator=gen).cuda()
    ci["visibility"] = torch.rand(12, 48 * 48, generator=gen).cuda()
    idx = torch.arange(12, device="cuda") + 24
    losses = {}
    for kind in ("fused", "torch"):
        m = PSNetwork(conf)
        m.load_state_dict(sds["init"])
        m = m.cuda().train()
        m.precision = "fp32"
        table = torch.nn.Embedding(60, 3, sparse=True).cuda()
        table.weight.data.copy_(synth.lights(60).cuda())
        A, S = (optim.Adam, optim.SparseAdam) if kind == "fused" else (torch.optim.Adam, torch.optim.SparseAdam)
        oa, os_ = A(m.parameters(), lr=5e-4), S(list(table.parameters()), lr=1e-3)
        scheds = [torch.optim.lr_scheduler.MultiStepLR(o, [3], gamma=0.5) for o in (oa, os_)]
        lm, ln = MainLoss(1.0, "L1", 0.05, 0.01, 1.0), NormalLoss(1.0, 0.05)
        hist = []
        for _ in range(6):
            ci["light_direction"] = torch.nn.functional.normalize(table(idx), p=2, dim=-1)
            out = m(ci)
            loss = lm(out, gt, ci)["loss"] + ln(out)["loss"]
            oa.zero_grad()
            os_.zero_grad()
            loss.backward()
            oa.step()
            os_.step()
            for sc in scheds:
                sc.step()
            hist.append(float(loss))
        assert table.weight.grad.is_sparse
        assert oa.param_groups[0]["lr"] == pytest.approx(2.5e-4) and os_.param_groups[0]["lr"] == pytest.approx(5e-4)
        losses[kind] = hist
    assert all(np.isfinite(losses["fused"])) and losses["fused"][-1] < losses["fused"][0]
    np.testing.assert_allclose(losses["fused"], losses["torch"], rtol=1e-2)
