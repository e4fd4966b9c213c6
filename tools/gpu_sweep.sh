#!/bin/bash
# GPU experiment helper: parity tests (unless SKIP_TESTS), then bench.py once per environment setting in "$@"
# (each argument is a space-separated list of VAR=value; "-" = defaults). One summary line per run.
[ -n "$SKIP_TESTS" ] || python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for cfg in "$@"; do
  [ "$cfg" = "-" ] && cfg=""
  env $cfg python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('[$cfg]', round(d['value'],1),'fps | pass', {k:round(v,2) for k,v in r['per_pass_ms'].items()}, '| trace', {k:round(v,2) for k,v in r['trace_ms'].items() if v>0}, '| stream', {k:round(v,2) for k,v in r['stream_ms'].items()}, '| build_ms', round(d['config']['scene']['build_ms'],1), 'nodes', d['config']['scene']['n_bvh_nodes'])"
done
