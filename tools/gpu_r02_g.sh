#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02g_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02g_pytest_gpu.log
tail -12 gpurun_out/r02g_pytest_gpu.log
grep -h "worst wav" gpurun_out/r02g_pytest_gpu.log
python tools/prof_acoustic.py --iters 3 2>&1 | tail -2 | tee gpurun_out/r02g_acoustic.log
python tools/prof_acoustic.py --iters 3 --alias 2>&1 | tail -2 | tee -a gpurun_out/r02g_acoustic.log
for al in "" "--alias"; do
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none --csv \
  -k regex:"s2pa_stream|pointwise_small" --log-file gpurun_out/r02g_s2pa$al.csv python tools/prof_acoustic.py --iters 0 $al > /dev/null 2>&1
grep -E "s2pa_stream|pointwise" gpurun_out/r02g_s2pa$al.csv | grep -E "time_duration|s2pa_stream" | awk -F'","' '{print substr($5,1,40), $(NF-2), $NF}' | head -6
done
