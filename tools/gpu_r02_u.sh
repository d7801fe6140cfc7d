#!/bin/bash
# round 2, call U: DTTS_TC_LO8_MINTAPS=3 in the real bench (sustained, power-capped clocks) + its error margins
mkdir -p gpurun_out
for v in 7 3 7 3; do
  DTTS_TC_LO8_MINTAPS=$v python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('mintaps=$v', round(d['ms_per_step'],3), d['stages_ms']['vocode'], d['clocks']['sm_mhz'], round(d['roofline']['frac'],4))" | tee -a gpurun_out/r02u_bench_ab.log
done
for v in 7 3; do
  DTTS_TC_LO8_MINTAPS=$v python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -s -k "ten_weight_seeds or golden or hot" 2>&1 | grep -E "worst|RMS|passed|failed" | sed "s/^/mintaps=$v /" | tee -a gpurun_out/r02u_errors.log
done
