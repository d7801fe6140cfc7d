#!/bin/bash
# round 2, call V: more loads in flight in the small acoustic kernels (LayerNorm, attention, pointwise, duration head)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_task.py tests/test_gpu_portaspeech.py -m gpu -x -q > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02v_pytest.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r02v_launches_step.csv python tools/prof_step.py --bank > /dev/null 2>&1
python tools/agg_launches.py gpurun_out/r02v_launches_step.csv | grep -E "channel_ln|self_attn|pointwise|dur_head|embed|launches" 
for i in 1 2; do python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), d['stages_ms'], d['clocks']['sm_mhz'])"; done
