"""Bring-up report of the tcgen05 geo stack: per-layer activation error vs the CPU oracle (prints, no asserts).
Run on the GPU box:  python tests/tc_bringup.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import psnerf_oracle as O  # noqa: E402
import util  # noqa: E402
from psnerf_b200 import _binding as B, engine  # noqa: E402
from psnerf_b200.stage1 import NeuralNetwork  # noqa: E402


def main():
    cfg, sds = util.stage1_state_dicts()
    lib = B.load()
    for variant in ("init", "trained"):
        sd = sds[variant]
        m = NeuralNetwork(cfg).eval()
        m.load_state_dict(sd)
        m = m.cuda()
        g, _ = m._packed()
        M = 300
        gen = torch.Generator().manual_seed(5)
        pts = (torch.rand(M, 3, generator=gen) * 2.4 - 1.2)
        with torch.no_grad():
            out, pre = O.geo_forward(sd, pts, cfg["model"], return_pre=True)
        pe = O.positional_encoding(pts, 6)
        pc = pts.cuda().contiguous()
        for l in range(8):
            ncmp = 256
            act = torch.nn.functional.softplus(pre[l], beta=100)
            if l == 3:
                act = torch.cat([act, pe], -1) / (2 ** 0.5)
            dump = torch.full((M, 256), float("nan"), device="cuda")
            logits = torch.empty(M, device="cuda")
            rc = lib.psn_tc_debug_layer(g.handle, C.c_void_p(pc.data_ptr()), M, l, C.c_void_p(dump.data_ptr()),
                                        C.c_void_p(logits.data_ptr()), engine._stream())
            torch.cuda.synchronize()
            if rc != 0:
                print("rc", rc, lib.psn_last_error())
                return
            d = dump.cpu()[:, :ncmp]
            act = act[:, :ncmp]
            err = (d - act).abs()
            print("%s layer %d: max abs err %.3e  rel-l2 %.3e  nan %d   (act max %.3f)" %
                  (variant, l, float(err.nan_to_num(1e9).max()), float((d - act).norm() / act.norm()),
                   int(torch.isnan(d).sum()), float(act.abs().max())))
            if l == 7:
                print("   logit max abs err %.3e (logit range %.3f..%.3f)" %
                      (float((logits.cpu() - out[:, 0]).abs().max()), float(out[:, 0].min()), float(out[:, 0].max())))
        a_tc = engine.occupancy(g, pc, B.OUT_ALPHA, B.PREC_TC).cpu()
        a_32 = engine.occupancy(g, pc, B.OUT_ALPHA, B.PREC_FP32).cpu()
        a_ref = torch.sigmoid(-10 * out[:, 0])
        print("%s alpha: tc-vs-oracle %.3e   fp32-vs-oracle %.3e" % (variant, float((a_tc - a_ref).abs().max()),
                                                                      float((a_32 - a_ref).abs().max())))


if __name__ == "__main__":
    main()
