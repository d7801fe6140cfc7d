#!/bin/bash
# gpurun call 3: full GPU suite (incl. full-size parity), bench (default and hybrid precision 5), A-stage sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for as in 2 3 4; do for nacc in 0 2; do
  echo "ASTAGES=$as NACC=$nacc: $(DTTS_TC_ASTAGES=$as DTTS_TC_NACC=$nacc python tools/prof_vocoder.py --precision 3 --iters 3 2>&1 | tail -1)" | tee -a gpurun_out/astages_sweep.log
done; done
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json
timeout 300 python bench.py --vocoder-precision 5 --steps 5 --no-cpu-baseline > gpurun_out/bench_p5.json 2> gpurun_out/bench_p5.err; echo "bench p5 rc=$?"
cat gpurun_out/bench_p5.json
