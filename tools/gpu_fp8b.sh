#!/bin/bash
mkdir -p gpurun_out
for prec in 3 6; do
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  -k regex:tc_conv --log-file gpurun_out/vocoder_full_p${prec}.csv python tools/prof_vocoder.py --precision $prec --iters 0 > /dev/null 2>&1; echo "ncu p$prec rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 33 -c 1 -o gpurun_out/tc_conv_s2_k11_p6 -f \
  python tools/prof_vocoder.py --precision 6 --iters 0 > /dev/null 2>&1; echo "ncu s2 p6 rc=$?"
