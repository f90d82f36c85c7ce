"""The GEMM of the train steps on tensor cores (csrc/tc_gemm.cu: tcgen05 kind::tf32, x = hi + lo split, three passes, fp32
accumulate) against torch's float64 matmul: the three operand forms of train_gemm.cuh, ragged sizes, sub-matrix views (leading
dimension > width), unaligned base pointers, wide dynamic range (gradients), the fused epilogues and the split-K accumulation.

Gate: 1e-5 of max |C|.  Measured on the B200: <= 2.3e-6 (K = 256) - the tensor core adds every MMA's K = 8 partial sum into the fp32
accumulator with truncation, so the error grows with the number of accumulating instructions (3 passes x K / 8), not with the
2^-21 of the operand split; an fp32 FFMA GEMM sits at ~3e-7 on the same data."""
import ctypes as C

import pytest
import torch

from psnerf_b200 import _binding as B

pytestmark = pytest.mark.gpu
GATE = 1e-5


def _gemm(form, A, B_, Cm, bias, M, N, K, epi):
    lib = B.load()
    B.require_device()
    P = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
    B.check(lib.psn_tc_gemm_debug(form, P(A), A.stride(0), P(B_), B_.stride(0), P(Cm), Cm.stride(0), P(bias), M, N, K, epi,
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)), "psn_tc_gemm_debug")


def _rand(shape, gen, spread=0.0):
    """values with magnitudes spread over `spread` decades (gradient-like dynamic range when spread > 0)"""
    x = torch.randn(shape, generator=gen)
    if spread > 0:
        x = x * torch.pow(10.0, -spread * torch.rand(shape, generator=gen))
    return x.cuda()


def _view(rows, cols, gen, pad=0, offset=0, spread=0.0):
    """[rows, cols] view with leading dimension cols + pad, starting `offset` floats into its storage"""
    buf = torch.zeros(rows * (cols + pad) + offset + 8, device="cuda")
    v = buf[offset:offset + rows * (cols + pad)].view(rows, cols + pad)[:, :cols]
    v.copy_(_rand((rows, cols), gen, spread))
    return v


SHAPES = [(1, 16, 39), (127, 256, 256), (129, 217, 256), (1000, 3, 256), (4133, 256, 289), (300, 289, 256), (2000, 304, 64), (64, 27, 63)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_forward_form_nt(M, N, K, epi):
    gen = torch.Generator().manual_seed(M * 7 + N * 3 + K + epi)
    A = _view(M, K, gen, pad=(M + K) % 5, offset=(M + N) % 4)
    W = _view(N, K, gen, pad=K % 3)
    bias = _rand((N,), gen)
    out = _view(M, N, gen, pad=N % 7)
    out.fill_(7.0)
    _gemm(0, A, W, out, bias if epi else None, M, N, K, epi)
    ref = A.double() @ W.double().t()
    if epi >= 1:
        ref = ref + bias.double()
    if epi == 2:
        ref = ref.clamp_min(0)
    if epi == 3:
        ref = torch.sigmoid(ref)
    err = float((out.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    assert err < (4 * GATE if epi == 3 else GATE), err  # sigmoid: relative to max |C| = 1, and expf adds its own 2 ulp


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_input_gradient_form_nn(M, N, K):
    gen = torch.Generator().manual_seed(M + N * 5 + K * 11)
    dZ = _view(M, K, gen, pad=K % 4, spread=6.0)   # gradients: six decades of dynamic range
    W = _view(K, N, gen, pad=(N + 1) % 4, offset=N % 3)
    out = _view(M, N, gen)
    _gemm(1, dZ, W, out, None, M, N, K, 0)
    ref = dZ.double() @ W.double()
    err = float((out.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    assert err < GATE, err
    # small entries keep their relative accuracy too (an fp16 split would flush them)
    rows = ref.abs().max(dim=1).values
    small = rows < rows.median()
    if int(small.sum()) > 0:
        rel = ((out.double() - ref)[small].abs().max(dim=1).values / rows[small].clamp_min(1e-300)).max()
        assert float(rel) < 1e-4, float(rel)


@pytest.mark.parametrize("Kc,M,N", [(1, 16, 16), (100, 256, 256), (5000, 217, 39), (40000, 256, 289), (33, 3, 256), (70001, 289, 64)])
def test_weight_gradient_form_tn_accumulates(Kc, M, N):
    gen = torch.Generator().manual_seed(Kc + M + N)
    dZ = _view(Kc, M, gen, pad=M % 3, spread=4.0)
    X = _view(Kc, N, gen, pad=N % 5, offset=1)
    out = _view(M, N, gen)
    before = out.clone()
    _gemm(2, dZ, X, out, None, M, N, Kc, 0)
    ref = before.double() + dZ.double().t() @ X.double()
    err = float((out.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    assert err < 3 * GATE, err   # K = samples: thousands of accumulating MMAs per tile; split-K partial sums are combined with fp32 atomics


def _gemm_fused(form, A, B_, Cm, bias, M, N, K, epi, C2=None, E1=None, E2=None, lde=0, scale=1.0):
    lib = B.load()
    B.require_device()
    P = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
    B.check(lib.psn_tc_gemm_debug_fused(form, P(A), A.stride(0), P(B_), B_.stride(0), P(Cm), Cm.stride(0), P(bias), M, N, K, epi, P(C2), P(E1),
                                        P(E2), lde, scale, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "psn_tc_gemm_debug_fused")


@pytest.mark.parametrize("M,N,K", [(127, 256, 256), (4133, 217, 256), (300, 256, 39), (1000, 257, 256)])
def test_fused_softplus_epilogue(M, N, K):
    """EPI 4 (forward layer of the stage-1 train step): h = softplus_100(x W^T + b), s = sigmoid(100 (x W^T + b))."""
    gen = torch.Generator().manual_seed(M + N + K)
    A = _view(M, K, gen, pad=(4 - K % 4) % 4)
    W = _view(N, K, gen) * 0.05
    bias = _rand((N,), gen) * 0.01
    h = torch.zeros(M, N, device="cuda")
    sg = torch.zeros(M, N, device="cuda")
    _gemm_fused(0, A, W, h, bias, M, N, K, 4, C2=sg, lde=N)
    z = A.double() @ W.double().t() + bias.double()
    assert float((h.double() - torch.nn.functional.softplus(z, beta=100)).abs().max()) < GATE * float(z.abs().max())
    assert float((sg.double() - torch.sigmoid(100 * z)).abs().max()) < 2e-3  # 100 x the 1e-5 of z


@pytest.mark.parametrize("form", [0, 1])
@pytest.mark.parametrize("M,N,K", [(129, 256, 256), (2000, 217, 256), (4133, 256, 217)])
def test_fused_elementwise_epilogues(form, M, N, K):
    """EPI 5 / 6 / 7 of the stage-1 backward: products of the accumulator with saved [M, N] matrices, one of them updated in place."""
    gen = torch.Generator().manual_seed(7 * M + N + K + form)
    A = _view(M, K, gen, spread=3.0)
    Bm = _view(N, K, gen) if form == 0 else _view(K, N, gen)
    ref = A.double() @ (Bm.double().t() if form == 0 else Bm.double())
    mx = float(ref.abs().max())
    E1 = torch.rand(M, N, generator=gen).cuda()
    E2 = _rand((M, N), gen)
    # 7: C = acc, C2 = acc E1
    Cm, C2 = torch.zeros(M, N, device="cuda"), torch.zeros(M, N, device="cuda")
    _gemm_fused(form, A, Bm, Cm, None, M, N, K, 7, C2=C2, E1=E1, lde=N)
    assert float((Cm.double() - ref).abs().max()) < GATE * mx
    assert float((C2.double() - ref * E1.double()).abs().max()) < GATE * mx
    # 6: C = acc scale E1 + E2
    _gemm_fused(form, A, Bm, Cm, None, M, N, K, 6, E1=E1, E2=E2, lde=N, scale=0.70710678)
    assert float((Cm.double() - (ref * 0.70710678 * E1.double() + E2.double())).abs().max()) < GATE * max(mx, 1.0)
    # 5: C = acc E1 ; E2 := acc E2 100 E1 (1 - E1)
    E2w = E2.clone()
    _gemm_fused(form, A, Bm, Cm, None, M, N, K, 5, E1=E1, E2=E2w, lde=N)
    assert float((Cm.double() - ref * E1.double()).abs().max()) < GATE * mx
    want = ref * E2.double() * (100 * E1.double() * (1 - E1.double()))
    assert float((E2w.double() - want).abs().max()) < GATE * float(want.abs().max())


def test_train_steps_use_the_tensor_gemm(monkeypatch):
    """The switch the train steps read: tensor-core GEMM unless PSNERF_B200_TRAIN_GEMM=ffma (evaluated once per process)."""
    import os
    assert os.environ.get("PSNERF_B200_TRAIN_GEMM") != "ffma"
