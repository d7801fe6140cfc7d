#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q -s -k "fp8_lo" > gpurun_out/pytest_fp8.log 2>&1; echo "fp8 tests rc=$?" | tee -a gpurun_out/pytest_fp8.log
grep -E "wav rms|passed|failed|Error|error" gpurun_out/pytest_fp8.log | head -30
python tools/prof_vocoder.py --precision 3 --iters 3 --lens 2>&1 | tail -1 | tee gpurun_out/vocoder_fp8_times.log
python tools/prof_vocoder.py --precision 6 --iters 3 --lens 2>&1 | tail -1 | tee -a gpurun_out/vocoder_fp8_times.log
DTTS_TC_LO8_MINTAPS=3 python tools/prof_vocoder.py --precision 6 --iters 3 --lens 2>&1 | tail -1 | tee -a gpurun_out/vocoder_fp8_times.log
DTTS_TC_LO8_MINTAPS=11 python tools/prof_vocoder.py --precision 6 --iters 3 --lens 2>&1 | tail -1 | tee -a gpurun_out/vocoder_fp8_times.log
DTTS_TC_ASTAGES=2 python tools/prof_vocoder.py --precision 6 --iters 3 --lens 2>&1 | tail -1 | tee -a gpurun_out/vocoder_fp8_times.log
python tools/prof_vocoder.py --precision 3 --iters 3 --lens 2>&1 | tail -1 | tee -a gpurun_out/vocoder_fp8_times.log
DTTS_TC_LO8_MINTAPS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  -k regex:tc_conv --log-file gpurun_out/vocoder_full_p6.csv python tools/prof_vocoder.py --precision 6 --iters 0 > /dev/null 2>&1; echo "ncu p6 rc=$?"
