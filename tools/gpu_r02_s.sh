#!/bin/bash
# round 2, call S: rb_pair128_kernel, E2 with two residual chunks in flight
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q -k "fused or lengths" > gpurun_out/r02s_pytest_fused.log 2>&1; echo "pytest fused rc=$?" | tee -a gpurun_out/r02s_pytest_fused.log
tail -3 gpurun_out/r02s_pytest_fused.log
for v in "A=1" "DTTS_TC_P128_PREFETCH=0"; do
  env $v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    -k regex:"rb_pair128" --log-file gpurun_out/r02s_pair128.csv python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1
  echo "$v: $(grep "time_duration" gpurun_out/r02s_pair128.csv | awk -F'","' '{print $NF}' | tr '\n' ' ')" | tee -a gpurun_out/r02s_variants.log
done
python tools/p128_trace.py --k 3 --dil 1 --tiles 8 > gpurun_out/r02s_trace_k3.txt 2>&1
grep -E "tile period|issue span|of which|acc2empty|tempty|e2:|e1" gpurun_out/r02s_trace_k3.txt | cut -c1-120
