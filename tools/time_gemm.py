"""Times the tcgen05 train-step GEMM (csrc/tc_gemm.cu) alone at the shapes of the stage-1 4096 x 128 train step:
   python tools/time_gemm.py [rows]       -> one line per operand form: ms, algorithmic TFLOP/s, HBM GB/s of the fp32 matrices."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psnerf_b200 import _binding as B  # noqa: E402

lib = B.load()
B.require_device()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * 128
P = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(form, A, Bm, Cm, M, N, K, reps=10):
    f = lambda: B.check(lib.psn_tc_gemm_debug(form, P(A), A.stride(0), P(Bm), Bm.stride(0), P(Cm), Cm.stride(0), P(None), M, N, K, 0, st), "gemm")
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
for n, k in ((256, 256), (256, 64), (217, 256)):
    X = torch.randn(rows, k, device="cuda")
    W = torch.randn(n, k, device="cuda")
    Y = torch.empty(rows, n, device="cuda")
    ms = run(0, X, W, Y, rows, n, k)
    out["nt_%dx%dx%d" % (rows, n, k)] = {"ms": ms, "TFLOPs": 2.0 * rows * n * k / ms * 1e-9, "GBs": 4.0 * rows * (n + k) / ms * 1e-6}
    dZ = torch.randn(rows, n, device="cuda")
    dX = torch.empty(rows, k, device="cuda")
    ms = run(1, dZ, W, dX, rows, k, n)
    out["nn_%dx%dx%d" % (rows, k, n)] = {"ms": ms, "TFLOPs": 2.0 * rows * n * k / ms * 1e-9, "GBs": 4.0 * rows * (n + k) / ms * 1e-6}
    dW = torch.zeros(n, k, device="cuda")
    ms = run(2, dZ, X, dW, n, k, rows)
    out["tn_%dx%dx%d" % (n, k, rows)] = {"ms": ms, "TFLOPs": 2.0 * rows * n * k / ms * 1e-9, "GBs": 4.0 * rows * (n + k) / ms * 1e-6}
    del X, Y, dZ, dX
print(json.dumps(out, indent=1))
