#!/bin/bash
# round 2, call T: FP8 lo plane also in the k = 3 layers of the C >= 128 stages (DTTS_TC_LO8_MINTAPS=3)
mkdir -p gpurun_out
for v in 7 3 7 3; do
  DTTS_TC_LO8_MINTAPS=$v python tools/prof_vocoder.py --precision 6 --iters 4 --lens 2>&1 | tail -1 | sed "s/^/mintaps=$v /" | tee -a gpurun_out/r02t_ab.log
done
DTTS_TC_LO8_MINTAPS=3 timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q > gpurun_out/r02t_pytest_tc_mintaps3.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02t_pytest_tc_mintaps3.log
DTTS_TC_LO8_MINTAPS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  -k regex:"tc_conv_kernel|rb_pair128" -c 40 --log-file gpurun_out/r02t_mintaps3.csv python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1
echo "mintaps=3: $(grep "time_duration" gpurun_out/r02t_mintaps3.csv | awk -F'","' '{print int($NF/1000)}' | tr -d '"' | tr '\n' ' ')"
