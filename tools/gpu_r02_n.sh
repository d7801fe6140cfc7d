#!/bin/bash
# round 2, call N: pipeline timeline of rb_pair128_kernel (trace build)
mkdir -p gpurun_out
for kd in "3 1" "7 3" "11 1"; do
  set -- $kd
  DTTS_TC_P128_TG=2 python tools/p128_trace.py --k $1 --dil $2 --tiles 8 > gpurun_out/r02n_trace_k$1.txt 2>&1
  tail -16 gpurun_out/r02n_trace_k$1.txt
done
