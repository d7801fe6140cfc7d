#!/bin/bash
# Short closing run of a round (after tools/gpu_round.sh produced the captures): both bench arms, the launch list of one
# step and of decode_mel with the final build.   gpurun --timeout 1500 -- 'bash tools/gpu_final.sh r02h'
TAG=${1:-r02h}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference_arm.json 2> /dev/null; echo "reference arm rc=$?"
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_step.csv python tools/prof_step.py --bank > gpurun_out/${TAG}_ncu_step.log 2>&1; echo "ncu step rc=$?"
python tools/agg_launches.py gpurun_out/${TAG}_launches_step.csv > gpurun_out/${TAG}_launches_step_agg.txt 2>&1; head -16 gpurun_out/${TAG}_launches_step_agg.txt
python tools/prof_acoustic.py --iters 3 --alias 2>&1 | tail -2 | tee gpurun_out/${TAG}_acoustic_times.log
python tools/prof_decode.py --iters 20 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_acoustic_times.log
DTTS_AC_FUSE=0 python bench.py --quick --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1_ac_unfused.json 2> /dev/null; echo "unfused bench rc=$?"
