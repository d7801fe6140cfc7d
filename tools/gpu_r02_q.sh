#!/bin/bash
# round 2, call Q: latency of one weight-stage copy inside rb_pair128_kernel
mkdir -p gpurun_out
for tg in 2 1 4; do
  DTTS_P128_COPYLAT=1 DTTS_TC_P128_TG=$tg python tools/p128_trace.py --k 7 --dil 3 --tiles 8 > gpurun_out/r02q_copylat_k7_tg$tg.txt 2>&1
  echo "== TG=$tg"; grep -E "tile period|issue span|COPYLAT|Error" gpurun_out/r02q_copylat_k7_tg$tg.txt | cut -c1-140
done
