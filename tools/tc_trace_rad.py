"""Timeline (SM cycles) of one 128-row tile of the tcgen05 radiance kernel, step by step: when the MMA lane started the step,
how long it waited for activations / weight stages, when the epilogue (row 0) finished it.  python tools/tc_trace_rad.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psnerf_b200 import _binding as B, engine, synth  # noqa: E402
from psnerf_b200.stage1 import NeuralNetwork  # noqa: E402

NAMES = (["geo fwd %d" % l for l in range(8)] + ["feature head", "app L0 feat (park)"] + ["reverse L%d" % l for l in range(7, 0, -1)] +
         ["reverse L0 -> grad", "app L0 rest", "app L1", "app L2", "app L3", "app L4 -> rgb"])
cfg = synth.stage1_cfg()
torch.manual_seed(0)
m = NeuralNetwork(cfg).cuda().eval()
g, a = m._packed()
M = 148 * 128 * 8
MIXED = 1 if "--mixed" in sys.argv else 0  # the PSN_PREC_TC_MIXED program (s8.. single-pass)
pts = (torch.rand(M, 3, device="cuda") * 2.4 - 1.2).contiguous()
views = torch.nn.functional.normalize(torch.randn(M, 3, device="cuda"), dim=-1).contiguous()
rgb = torch.empty(M, 3, device="cuda")
alpha = torch.empty(M, device="cuda")
stash = torch.empty(148 * 640 * 1024 + 4096, dtype=torch.uint8, device="cuda")
trace = torch.zeros(256, dtype=torch.int64, device="cuda")
lib = B.load()
for _ in range(2):
    B.check(lib.psn_tc_debug_trace_rad(g.handle, a.handle, C.c_void_p(pts.data_ptr()), C.c_void_p(views.data_ptr()), M,
                                       C.c_void_p(rgb.data_ptr()), C.c_void_p(alpha.data_ptr()), C.c_void_p(stash.data_ptr()),
                                       C.c_void_p(trace.data_ptr()), MIXED, engine._stream()), "trace")
torch.cuda.synchronize()
t = trace.cpu().tolist()
t0 = t[0]
print("step                 | MMA start | wait_a wait_w | last commit | epilogue done | step period")
prev = None
for st in range(23):
    start = t[st * 8] - t0
    commit = t[st * 8 + 7] - t0
    epi = t[192 + st] - t0 if t[192 + st] else -1
    per = (start - prev) if prev is not None else 0
    prev = start
    print("%2d %-18s| %9d | %6d %6d | %11d | %13d | %6d" % (st, NAMES[st], start, t[st * 8 + 4], t[st * 8 + 5], commit, epi, per))
print("tile total (MMA start of step 0 -> epilogue of step 22):", t[192 + 22] - t0)
