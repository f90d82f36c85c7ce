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


def test_train_steps_use_the_tensor_gemm(monkeypatch):
    """The switch the train steps read: tensor-core GEMM unless PSNERF_B200_TRAIN_GEMM=ffma (evaluated once per process)."""
    import os
    assert os.environ.get("PSNERF_B200_TRAIN_GEMM") != "ffma"
