#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small CSV that can be committed under profiles/.

  python tools/ncu_summary.py gpurun_out/r2_prof.ncu-rep profiles/r1_full_summary.csv [profiles/traffic.json]

One row per profiled launch: duration, achieved DRAM traffic, issue-slot utilisation, occupancy, SIMT
efficiency, cache hit rates, pipe utilisation and the top stall reasons.
"""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_inst"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "pipe_alu_pct"),
    ("sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "pipe_fma_pct"),
    ("sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "pipe_fp64_pct"),
    ("sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "pipe_xu_pct"),
    ("sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "pipe_lsu_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"
STALL_SUFFIX = "_per_issue_active.ratio"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(head)}
    stall_cols = [n for n in head if n.startswith(STALL_PREFIX) and n.endswith(STALL_SUFFIX) and "not_issued" not in n]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel"] + [f"{short} [{units[col[m]]}]" if units[col[m]] else short for m, short in METRICS if m in col] +
                   ["top_stalls (warps stalled per issue)"])
        for r in data:
            name = r[col["Kernel Name"]].replace("<unnamed>::", "").split("(")[0].replace("void ", "")
            vals = [r[col[m]] for m, _ in METRICS if m in col]
            st = []
            for n in stall_cols:
                try:
                    st.append((float(r[col[n]]), n[len(STALL_PREFIX):-len(STALL_SUFFIX)]))
                except ValueError:
                    pass
            st.sort(reverse=True)
            w.writerow([r[col["ID"]], name] + vals + [" ".join(f"{n}={v:.2f}" for v, n in st[:4])])
    print("wrote", out, len(data), "launches")
    # per-kernel DRAM traffic per launch (bytes), for bench.py's roofline.traffic
    if len(sys.argv) > 3:
        import json
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        agg = {}
        for r in data:
            name = r[col["Kernel Name"]].replace("<unnamed>::", "").split("(")[0].replace("void ", "")
            rd = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
            wr = float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
            a = agg.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "time_ms": 0.0})
            a["launches"] += 1
            a["dram_bytes"] += rd + wr
            tu = units[col["gpu__time_duration.sum"]]
            a["time_ms"] += float(r[col["gpu__time_duration.sum"]]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(tu, 1.0)
        for a in agg.values():
            a["dram_bytes_per_launch"] = a.pop("dram_bytes") / a["launches"]
            a["time_ms_per_launch"] = a.pop("time_ms") / a["launches"]
        json.dump({"source": rep.split("/")[-1], "note": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch",
                   "kernels": agg}, open(sys.argv[3], "w"), indent=1)
        print("wrote", sys.argv[3])


if __name__ == "__main__":
    main()
