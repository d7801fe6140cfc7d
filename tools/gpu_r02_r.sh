#!/bin/bash
# round 2, call R: rb_pair128_kernel with the L2 prefetch of the next tile's input and input-first shared-memory plan
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q -k "fused or lengths" > gpurun_out/r02r_pytest_fused.log 2>&1; echo "pytest fused rc=$?" | tee -a gpurun_out/r02r_pytest_fused.log
tail -3 gpurun_out/r02r_pytest_fused.log
for v in "A=1" "DTTS_TC_P128_PREFETCH=0" "DTTS_TC_P128_ASTAGES=2" "DTTS_TC_P128_ASTAGES=2 DTTS_TC_P128_PREFETCH=0"; do
  env $v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    -k regex:"rb_pair128" --log-file gpurun_out/r02r_pair128.csv python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1
  echo "$v: $(grep "time_duration" gpurun_out/r02r_pair128.csv | awk -F'","' '{print $NF}' | tr '\n' ' ')" | tee -a gpurun_out/r02r_variants.log
done
python tools/p128_trace.py --k 7 --dil 3 --tiles 8 > gpurun_out/r02r_trace_k7.txt 2>&1
grep -E "tile period|issue span|of which|acc2empty|tempty" gpurun_out/r02r_trace_k7.txt | cut -c1-120
python tools/p128_trace.py --k 3 --dil 1 --tiles 8 > gpurun_out/r02r_trace_k3.txt 2>&1
grep -E "tile period|issue span|of which|acc2empty|tempty|e2:|e1" gpurun_out/r02r_trace_k3.txt | cut -c1-120
