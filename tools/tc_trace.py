"""Timeline (SM cycles) of one 128-row tile of the tcgen05 occupancy kernel: when the MMA lane got each K block, when it
committed each layer, when the epilogue warps woke up / finished their chunk passes.  python tools/tc_trace.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psnerf_b200 import _binding as B, engine, synth  # noqa: E402
from psnerf_b200.stage1 import NeuralNetwork  # noqa: E402

cfg = synth.stage1_cfg()
torch.manual_seed(0)
m = NeuralNetwork(cfg).cuda().eval()
g, _ = m._packed()
M = 148 * 128 * 8
pts = (torch.rand(M, 3, device="cuda") * 2.4 - 1.2).contiguous()
out = torch.empty(M, device="cuda")
trace = torch.zeros(256, dtype=torch.int64, device="cuda")
lib = B.load()
for _ in range(2):
    B.check(lib.psn_tc_debug_trace(g.handle, C.c_void_p(pts.data_ptr()), M, C.c_void_p(out.data_ptr()),
                                   C.c_void_p(trace.data_ptr()), engine._stream()), "trace")
torch.cuda.synchronize()
t = trace.cpu().tolist()
t0 = min(t[l * 8] for l in range(8) if t[l * 8] > 0)  # time stamps only (slots +4 / +5 hold cycle counts)
rel = lambda x: (x - t0) if x > 0 else -1
print("layer | MMA: a_ready kb0 kb1 kb2 kb3 | commit | wait_a wait_w || epilogue sub0: wake p0 p1 p2 p3 | sub3: wake p0 p1 p2 p3")
for l in range(8):
    mm = [rel(t[l * 8 + k]) for k in range(4)]
    cm = rel(t[l * 8 + 7])
    e0 = [rel(t[64 + 0 * 40 + l * 5 + k]) for k in range(5)]
    e3 = [rel(t[64 + 3 * 40 + l * 5 + k]) for k in range(5)]
    print("%5d | %6d %6d %6d %6d | %6d | %6d %6d || %6d %6d %6d %6d %6d | %6d %6d %6d %6d %6d" %
          (l, *mm, cm, t[l * 8 + 4], t[l * 8 + 5], *e0, *e3))
print("per K block (layers 1..4): wait for activations kb0..3 | wait for weight stages kb0..3")
for l in range(1, 5):
    print("%5d | %s | %s" % (l, " ".join("%6d" % t[224 + (l - 1) * 8 + k] for k in range(4)),
                           " ".join("%6d" % t[224 + (l - 1) * 8 + 4 + k] for k in range(4))))
# pass end times of layers 1..3 for the four TMEM lane quadrants (one scheduler each; quadrant 1 shares its scheduler with the MMA warp)
for q in range(4):
    trace.zero_()
    for _ in range(2):
        B.check(lib.psn_tc_debug_trace_q(g.handle, C.c_void_p(pts.data_ptr()), M, C.c_void_p(out.data_ptr()),
                                         C.c_void_p(trace.data_ptr()), q, engine._stream()), "trace")
    torch.cuda.synchronize()
    tq = trace.cpu().tolist()
    t0q = min(tq[l * 8] for l in range(8) if tq[l * 8] > 0)
    rows = []
    for l in range(1, 4):
        rows.append(" ".join("%6d" % (tq[64 + s_ * 40 + l * 5 + 4] - t0q) for s_ in range(4)))
    print("quadrant %d: end of pass 3 of layers 1..3 for subs 0..3 | %s" % (q, " | ".join(rows)))
