"""Summarise an ncu report (`ncu --set full ... -o X`) into the small JSON committed under profiles/.

  python tools/ncu_summary.py gpurun_out/prof_s1.ncu-rep profiles/r1_ncu_v2.json

Per captured launch of the tensor-core kernels: duration, DRAM bytes, tensor / XU / issue utilisation, shared-memory
pipe shares, L2 -> SM bytes and the top warp-stall reasons.  Runs here (no GPU needed: ncu only reads the report)."""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        if d.get("gpu__time_duration.sum", "") in ("", "-nan") or d.get("dram__bytes_read.sum", "-nan") == "-nan":
            continue
        item = {"id": int(d["ID"]), "kernel": d["Kernel Name"].split("(")[0], "metrics": {}}
        for m in METRICS:
            if m in d and d[m] not in ("", "-nan"):
                item["metrics"][m] = {"value": float(d[m].replace(",", "")), "unit": u[m]}
        mm = item["metrics"]
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if k in mm:
                tot += mm[k]["value"] * SCALE.get(mm[k]["unit"], 1.0)
        item["dram_bytes_total"] = tot
        st = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(v.replace(",", ""))) for h, v in d.items()
              if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and v not in ("", "-nan")]
        s = sum(v for _, v in st) or 1.0
        item["stall_pct"] = {n: round(100 * v / s, 1) for n, v in sorted(st, key=lambda x: -x[1])[:8]}
        res.append(item)
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    for it in res:
        m = it["metrics"]
        print(it["id"], it["kernel"], "%.3f %s" % (m["gpu__time_duration.sum"]["value"], m["gpu__time_duration.sum"]["unit"]),
              "tensor %.1f%%" % m["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]["value"],
              "dram %.3f GB" % (it["dram_bytes_total"] / 1e9))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
