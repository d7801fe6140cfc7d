#!/bin/bash
# round 2, call O: does the epilogue's HBM traffic inflate the weight / input latency of rb_pair128_kernel?
mkdir -p gpurun_out
for kd in "7 3" "11 1"; do
  set -- $kd
  DTTS_P128_NOEPI=1 DTTS_TC_P128_TG=2 python tools/p128_trace.py --k $1 --dil $2 --tiles 8 > gpurun_out/r02o_trace_noepi_k$1.txt 2>&1
  tail -14 gpurun_out/r02o_trace_noepi_k$1.txt
done
