#!/bin/bash
# round 2, call B: PortaSpeech sibling parity on the GPU + the remaining compute-sanitizer variants
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_portaspeech.py -m gpu -x -q > gpurun_out/r02b_pytest_ps.log 2>&1; echo "pytest ps rc=$?" | tee -a gpurun_out/r02b_pytest_ps.log
tail -25 gpurun_out/r02b_pytest_ps.log
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_portaspeech.py > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rest rc=$?" | tee -a gpurun_out/r02b_pytest_gpu.log
tail -4 gpurun_out/r02b_pytest_gpu.log
SAN_TIMEOUT=300 bash tools/sanitize.sh 2>&1 | tail -8
