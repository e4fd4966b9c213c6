#!/bin/sh
# compute-sanitizer passes over the hot path (SURVEY §5 "race detection / sanitizers"):
#   sh tools/sanitize.sh            -> gpurun_out/r2_sanitizer_*.log
# NOTE (round 2): compute-sanitizer is CLOSED on the GPU pool this repo is developed on ("runs under it have left GPUs
# needing a reset"); the script is kept for pools where it is available. What replaces it here is listed in DESIGN.md
# ("Race and bounds evidence without a sanitizer").
OUT=gpurun_out
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  $CS --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2_sanitizer_smoke_$tool.log 2>&1
  echo "$tool smoke rc=$?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|closed' $OUT/r2_sanitizer_smoke_$tool.log | tr '\n' ' ')"
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  CHECK_FRAMES=4 RB_BAL_PERIOD=1 $CS --tool memcheck --target-processes all --print-limit 20 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_nccl_bands.py \
    > $OUT/r2_sanitizer_bands_memcheck.log 2>&1
  echo "memcheck 2 bands rc=$?: $(grep -E 'ERROR SUMMARY|bit-identical|DIFFERENT|closed' $OUT/r2_sanitizer_bands_memcheck.log | tail -8 | tr '\n' ' ')"
fi
