#!/bin/bash
# round 2, call E: programmatic dependent launch on the tcgen05 kernels: full parity suite + A/B timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02e_pytest_gpu.log
tail -6 gpurun_out/r02e_pytest_gpu.log
for pdl in 0 1 0 1; do
  DTTS_TC_PDL=$pdl timeout 300 python bench.py --quick --steps 20 --warmup 5 > gpurun_out/r02e_bench_pdl$pdl.json 2> gpurun_out/r02e_bench_pdl$pdl.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02e_bench_pdl$pdl.json'))
print('pdl=$pdl', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stages_ms'].items()}, round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'])
PY
done | tee gpurun_out/r02e_pdl_ab.log
