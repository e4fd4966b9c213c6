#!/usr/bin/env python
"""Hot source lines of one kernel: joins the SASS-level stall samples / instruction counts of an .ncu-rep
(ncu --set full --import-source on) with the line table of the library it was taken from (nvdisasm -g).

  python tools/ncu_hot_lines.py REPORT.ncu-rep LIB.so KERNEL_SUBSTRING [launch_index=0] [top=40]

The innermost line (after inlining) is what nvdisasm reports, so header lines (rb_passes.cuh, det_math.h ...)
show up directly.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def line_table(lib, kernel_sub):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
    cubin = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True, check=True).stdout
    table, cur, on = {}, ("?", 0), False
    for ln in txt.splitlines():
        if ln.startswith("\t.section\t.text."):
            on = kernel_sub in ln
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    rep, lib, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
    launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    table = line_table(lib, ksub)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + re.match(r"[a-z_0-9]+", ksub).group(0)],
                         capture_output=True, text=True).stdout
    segs, cur = [], None
    for r in csv.reader(txt.splitlines()):
        if not r:
            continue
        if r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            segs.append(cur)
        elif r[0] == "Address":
            cur["head"] = r
        elif cur is not None:
            cur["rows"].append(r)
    segs = [s for s in segs if len(s["rows"]) == len(table)] or segs
    s = segs[launch]
    ix = {n: i for i, n in enumerate(s["head"])}
    base = int(s["rows"][0][ix["Address"]], 16)
    by_line = collections.defaultdict(lambda: [0, 0, 0])
    tot_s = tot_i = 0
    for r in s["rows"]:
        off = int(r[ix["Address"]], 16) - base
        (f, l), _ = table.get(off, (("?", 0), ""))
        smp, ins, thr = int(r[ix["Warp Stall Sampling (All Samples)"]]), int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]])
        a = by_line[(f, l)]
        a[0] += smp
        a[1] += ins
        a[2] += thr
        tot_s += smp
        tot_i += ins
    print(f"{s['name'][:80]}: {len(s['rows'])} SASS instructions, {tot_i} warp instructions executed, {tot_s} samples")
    src_cache = {}

    def src(f, l):
        for root in ("restir_embree_b200/csrc", "include"):
            p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), root, f)
            if os.path.exists(p):
                if p not in src_cache:
                    src_cache[p] = open(p).read().splitlines()
                return src_cache[p][l - 1].strip()[:100] if 0 < l <= len(src_cache[p]) else ""
        return ""

    print(" samples   insts  thr/inst  file:line  source")
    for (f, l), (smp, ins, thr) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{smp / max(tot_s, 1) * 100:7.2f}% {ins / max(tot_i, 1) * 100:6.2f}% {thr / max(ins, 1):8.1f}  {f}:{l}  {src(f, l)}")


if __name__ == "__main__":
    main()
