#!/bin/bash
# gpurun call 4: A-stage sweep on the acoustic model (deep-K small GEMMs), printed waveform errors of the fp16 modes, smoke.
mkdir -p gpurun_out
for as in 2 3 4; do
  echo "ASTAGES=$as: $(DTTS_TC_ASTAGES=$as python tools/prof_acoustic.py --iters 3 2>&1 | tail -2 | tr '\n' ' ')" | tee -a gpurun_out/astages_acoustic.log
done
timeout 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -s -k "fp16_vocoder" 2>&1 | grep -E "precision|passed|failed" | tee gpurun_out/fp16_mode_errors.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
