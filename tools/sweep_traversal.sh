#!/bin/bash
# GPU experiment helper: parity tests, then a sweep of the traversal scheduling thresholds
# (RB_REFILL = refill_lanes, RB_POSTPONE = tri_lanes, RB_TRACE_BLOCKS = resident blocks per SM).
[ -n "$SKIP_TESTS" ] || python -m pytest tests -m gpu -x -q 2>&1 | tail -2
SWEEP=${SWEEP:-22,8,8 22,1,8 22,4,8 22,12,8 22,16,8 26,8,8 18,8,8 22,8,12}
for cfg in $SWEEP; do
  IFS=, read a b c <<< "$cfg"
  RB_REFILL=$a RB_POSTPONE=$b RB_TRACE_BLOCKS=$c python bench.py --steps 8 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('refill,tri_lanes,blocks $cfg', round(d['value'],1),'fps trace', {k:round(v,2) for k,v in r['trace_ms'].items() if v>0}, 'stream', {k:round(v,2) for k,v in r['stream_ms'].items()})"
done
