#!/bin/bash
# round 2, call A: GPU parity suite after the advisor fixes, smoke, the new multi-config bench line, both arms,
# and a first bounded compute-sanitizer pass.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02a_pytest_gpu.log
tail -5 gpurun_out/r02a_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02a_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err ) 2>&1 | grep real; echo "bench rc=$?"
head -c 1500 gpurun_out/r02a_bench_n1.json; echo
tail -3 gpurun_out/r02a_bench_n1.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err ) 2>&1 | grep real
head -c 600 gpurun_out/r02a_bench_ref.json; echo
SAN_FIRST_ONLY=1 SAN_TIMEOUT=300 bash tools/sanitize.sh 2>&1 | head -3
