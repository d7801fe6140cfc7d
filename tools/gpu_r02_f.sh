#!/bin/bash
# round 2, call F: single-pass S2PA kernel, pointwise flow projections, one-launch space-to-depth staging
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02f_pytest_gpu.log
tail -12 gpurun_out/r02f_pytest_gpu.log
python tools/prof_acoustic.py --iters 3 2>&1 | tail -3 | tee gpurun_out/r02f_acoustic.log
python tools/prof_acoustic.py --iters 3 --alias 2>&1 | tail -2 | tee -a gpurun_out/r02f_acoustic.log
for al in "" "--alias"; do
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  -k regex:"s2pa_stream|pointwise_small" --log-file gpurun_out/r02f_s2pa$al.csv python tools/prof_acoustic.py --iters 0 $al > /dev/null 2>&1
grep -E "s2pa_stream|pointwise" gpurun_out/r02f_s2pa$al.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | head -12
done
timeout 300 python bench.py --quick --steps 20 --warmup 5 > gpurun_out/r02f_bench_quick.json 2> gpurun_out/r02f_bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench_quick.json'))
print(round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stages_ms'].items()}, round(d['e2e']['ms_per_step'],3), d['gpu_launches']/d['steps'])
PY
