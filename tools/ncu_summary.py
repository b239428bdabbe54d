#!/usr/bin/env python
"""Summarise an `ncu --set full` report (.ncu-rep) into a small JSON: one entry per captured kernel with the metrics
DESIGN.md / profiles/README.md quote.  Usage: python tools/ncu_summary.py gpurun_out/final_full.ncu-rep out.json"""
import csv
import json
import re
import subprocess
import sys

WANT = {
    "duration_us": ("gpu__time_duration.sum", 1e-3),
    "dram_read_bytes": ("dram__bytes_read.sum", None),
    "dram_write_bytes": ("dram__bytes_write.sum", None),
    "grid": ("launch__grid_size", 1), "block": ("launch__block_size", 1), "cluster": ("launch__cluster_size", 1),
    "clusters_max_active": ("launch__cluster_max_active", 1),
    "regs_per_thread": ("launch__registers_per_thread", 1),
    "dyn_smem_kb": ("launch__shared_mem_per_block_dynamic", 1),
    "tensor_pipe_active_pct_of_active": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
    "tensor_pipe_active_pct_of_elapsed": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
    "utchmma_tf32_pct_of_peak": ("sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", 1),
    "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
    "sm_throughput_pct": ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "dram_throughput_pct": ("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", 1),
}
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "ns": 1, "us": 1e3, "ms": 1e6}


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in data:
        e = {"kernel": re.sub(r"\(.*\)$", "", r[col["Kernel Name"]].replace("void <unnamed>::", ""))}
        for k, (name, scale) in WANT.items():
            if name not in col or r[col[name]] == "":
                continue
            try:
                v = float(r[col[name]].replace(",", ""))
            except ValueError:          # "no data" (a metric the pass did not collect for this launch)
                continue
            u = units[col[name]]
            if scale is None:
                v *= UNIT_SCALE.get(u, 1)
            elif k == "duration_us":
                v = v * UNIT_SCALE.get(u, 1) * 1e-3
            e[k] = round(v, 3)
        stalls = []
        for h, i in col.items():
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio", h)
            if m and r[i] != "":
                try:
                    stalls.append((float(r[i].replace(",", "")), m.group(1) or m.group(2)))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        e["top_stalls"] = [{"reason": n, "ratio": round(v, 2)} for v, n in stalls[:5]]
        res.append(e)
    json.dump(res, open(out, "w"), indent=1)
    for e in res:
        print(json.dumps(e))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
