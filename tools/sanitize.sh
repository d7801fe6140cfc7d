#!/bin/bash
# compute-sanitizer over one small pass of every CUDA path: memcheck on the CTA-pair build, the
# single-CTA build and the cluster-multicast build, racecheck on the pair and single builds.  Run on the GPU box:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# Logs land in gpurun_out/sanitizer_*.log (copied to profiles/ by hand).  Every run is bounded by `timeout`.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {   # tag, tool, env..., -- args
  local tag=$1 tool=$2; shift 2
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  local log=gpurun_out/sanitizer_${tool}_${tag}.log
  env "${envs[@]}" timeout ${SAN_TIMEOUT:-420} $CS --tool $tool --error-exitcode 7 --print-limit 4000 --log-file $log \
    python tools/sanitize_run.py "$@" > gpurun_out/sanitizer_${tool}_${tag}.out 2>&1
  local rc=$?
  echo "[$tool/$tag] rc=$rc $(tail -1 gpurun_out/sanitizer_${tool}_${tag}.out) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
}
[ -n "$SAN_FUSE128_ONLY" ] && { run fuse128 memcheck DTTS_TC_PAIR=1 -- --skip-acoustic --fuse-all; run fuse128 racecheck DTTS_TC_PAIR=1 -- --skip-acoustic --fuse-all --frames 8; exit 0; }
run pair   memcheck DTTS_TC_PAIR=1 --
[ -n "$SAN_FIRST_ONLY" ] && { run pair racecheck DTTS_TC_PAIR=1 -- --frames 8; exit 0; }
run single memcheck DTTS_TC_PAIR=0 --
run cluster2 memcheck DTTS_TC_PAIR=0 DTTS_TC_CLUSTER=2 -- --skip-acoustic --vocoder-precision 3
run p1     memcheck DTTS_TC_PAIR=1 -- --skip-acoustic --vocoder-precision 1
run fuse128 memcheck DTTS_TC_PAIR=1 -- --skip-acoustic --fuse-all
run pair   racecheck DTTS_TC_PAIR=1 -- --frames 8
run fuse128 racecheck DTTS_TC_PAIR=1 -- --skip-acoustic --fuse-all --frames 8
run single racecheck DTTS_TC_PAIR=0 -- --frames 8
