#!/bin/bash
# gpurun call 7: full capture of a short-tile layer (stage 4, k = 3, conv1) to see what bounds it.
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 59 -c 2 -o gpurun_out/tc_conv_s4_k3 -f \
  python tools/prof_vocoder.py --precision 3 --iters 0 > gpurun_out/ncu_s4.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
