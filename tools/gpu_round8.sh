#!/bin/bash
# gpurun call 8: packed-fp32 epilogue + no residual shuffle without a residual: tests, vocoder timings, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python tools/prof_vocoder.py --precision 3 --iters 3 2>&1 | tail -1 | tee gpurun_out/vocoder_times8.log
python tools/prof_vocoder.py --precision 3 --iters 3 --lens 2>&1 | tail -1 | tee -a gpurun_out/vocoder_times8.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  -k regex:tc_conv --log-file gpurun_out/vocoder_full_times8.csv python tools/prof_vocoder.py --precision 3 --iters 0 > gpurun_out/ncu_voc_full8.log 2>&1; echo "ncu full rc=$?"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "bench rc=$?"
cut -c1-1200 gpurun_out/bench8.json
