#!/bin/bash
# compute-sanitizer over the acoustic model only (the fused kernels of round 2's second session: flow_fused_kernel, the
# LayerNorm / gate epilogues of tc_conv_kernel<0>, fvae_pre_net_planes_kernel, l2_prefetch), fused and per-layer builds:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/sanitize_ac.sh'
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  log=gpurun_out/sanitizer_${tool}_acoustic.log
  timeout ${SAN_TIMEOUT:-600} $CS --tool $tool --error-exitcode 7 --print-limit 4000 --log-file $log \
    python tools/sanitize_run.py --skip-vocoder --long > gpurun_out/sanitizer_${tool}_acoustic.out 2>&1
  echo "[$tool/acoustic] rc=$? $(tail -1 gpurun_out/sanitizer_${tool}_acoustic.out) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
done
