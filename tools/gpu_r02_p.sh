#!/bin/bash
# round 2, call P: weight-stage refill latency inside rb_pair128_kernel
mkdir -p gpurun_out
for tg in 2 1; do
  DTTS_TC_P128_TG=$tg python tools/p128_trace.py --k 7 --dil 3 --tiles 8 > gpurun_out/r02p_trace_k7_tg$tg.txt 2>&1
  tail -17 gpurun_out/r02p_trace_k7_tg$tg.txt
done
