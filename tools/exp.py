#!/usr/bin/env python
"""A/B runner for kernel experiments on the GPU box: runs bench.py (no CPU leg) once per variant and prints one line each.

    python tools/exp.py name[:LIB][:ENV=V,ENV=V] ...      LIB = file under restir_embree_b200/variants/ (an alternative build)

Variant builds are made in the build container with tools/build_variant.sh NAME -DFLAG...; they travel with the snapshot."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    steps = os.environ.get("EXP_STEPS", "20")
    out = []
    for spec in sys.argv[1:]:
        parts = spec.split(":")
        name = parts[0]
        env = dict(os.environ)
        for p in parts[1:]:
            if "=" in p:
                for kv in p.split(","):
                    k, v = kv.split("=")
                    env[k] = v
            elif p:
                env["RB_LIB"] = os.path.join(ROOT, "restir_embree_b200", "variants", p)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", steps, "--warmup", "5", "--no-cpu"],
                           env=env, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            ro = d["roofline"]
            row = dict(name=name, fps=round(d["value"], 2), e2e=round(d["e2e"]["value"], 2),
                       e2e_blocking=round(d["e2e"].get("blocking_call_value", 0), 2),
                       per_pass={k: round(v, 3) for k, v in ro["per_pass_ms"].items()},
                       stream={k: round(v, 3) for k, v in ro["stream_ms"].items()},
                       trace={k: round(v, 3) for k, v in ro["trace_ms"].items()})
        except Exception as e:  # noqa: BLE001
            row = dict(name=name, error=str(e), stderr=r.stderr[-600:], stdout=r.stdout[-300:])
        print(json.dumps(row), flush=True)
        out.append(row)
    return 0


if __name__ == "__main__":
    sys.exit(main())
