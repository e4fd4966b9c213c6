#!/usr/bin/env python
"""Compile-time resource table of every kernel (registers, stack frame, spill bytes) from `nvcc -Xptxas -v`, for
sm_100a with the library's own flags. No GPU needed.   python tools/ptxas_table.py > profiles/<round>_ptxas_kernels.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    csrc = ge.CSRC
    cmd = [ge.NVCC] + ge.NVCC_FLAGS + ["-Xptxas", "-v", "-o", "/tmp/_ptxas_table.so", os.path.join(csrc, "restir_b200.cu")]
    out = subprocess.run(cmd, cwd=csrc, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, check=True).stdout
    demangle = subprocess.run(["c++filt"], input=out, stdout=subprocess.PIPE, text=True).stdout
    rows, cur = {}, None
    for line in demangle.splitlines():
        m = re.search(r"Function properties for (.+)$", line)
        if m:
            cur = re.sub(r"\(anonymous namespace\)::", "", m.group(1)).strip()
            cur = re.sub(r"\(rb::(FrameCtx|BuildCtx)\)", "", cur)
            rows.setdefault(cur, {})
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and cur:
            rows[cur].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
        m = re.search(r"Used (\d+) registers", line)
        if m and cur:
            rows[cur]["regs"] = int(m.group(1))
    print("# nvcc " + " ".join(ge.NVCC_FLAGS) + " -Xptxas -v   (restir_b200.cu, CUDA 12.9)")
    print(f"{'kernel / out-of-line device function':<86} {'regs':>5} {'stack B':>8} {'spill st/ld B':>14}")
    for k in sorted(rows, key=lambda s: (("regs" not in rows[s]), s)):
        r = rows[k]
        print(f"{k[:86]:<86} {str(r.get('regs', '-')):>5} {r.get('stack', 0):>8} {str(r.get('spill_st', 0)) + '/' + str(r.get('spill_ld', 0)):>14}")


if __name__ == "__main__":
    main()
