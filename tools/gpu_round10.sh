#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python tools/prof_vocoder.py --precision 3 --iters 3 2>&1 | tail -1 | tee gpurun_out/vocoder_times10.log
python tools/prof_vocoder.py --precision 3 --iters 3 --lens 2>&1 | tail -1 | tee -a gpurun_out/vocoder_times10.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  -k regex:tc_conv --log-file gpurun_out/vocoder_full_times10.csv python tools/prof_vocoder.py --precision 3 --iters 0 > /dev/null 2>&1; echo "ncu rc=$?"
