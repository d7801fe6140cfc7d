#!/bin/bash
# One gpurun call that produces everything profiles/ holds for a round (run from the repo root on the GPU box):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round.sh'
# GPU parity tests, the bench line, the ncu launch list of one step, per-launch DRAM traffic of the vocoder
# (full-length and valid-length pass) and two full captures of the top kernel.  ncu numbers are never bench values.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
# the single-CTA build of the FP8 lo-plane path and of the wide layers (the switch is read once per process)
DTTS_TC_PAIR=0 timeout 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q -k "fp8_lo or lengths or baseline_shapes" > gpurun_out/pytest_gpu_nopair.log 2>&1; echo "pytest (DTTS_TC_PAIR=0) rc=$?" | tee -a gpurun_out/pytest_gpu_nopair.log
tail -3 gpurun_out/pytest_gpu_nopair.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_step.csv python tools/prof_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu step rc=$?"
python tools/agg_launches.py gpurun_out/launches_step.csv > gpurun_out/launches_step_agg.txt 2>&1; head -12 gpurun_out/launches_step_agg.txt
for mode in "" "--lens"; do
  tag=full; [ -n "$mode" ] && tag=lens
  python tools/prof_vocoder.py --precision 6 --iters 3 $mode 2>&1 | tail -1 | tee -a gpurun_out/vocoder_times.log
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    -k regex:tc_conv --log-file gpurun_out/vocoder_${tag}_dram.csv python tools/prof_vocoder.py --precision 6 --iters 0 $mode > /dev/null 2>&1
  python tools/agg_launches.py gpurun_out/vocoder_${tag}_dram.csv ALL > gpurun_out/vocoder_${tag}_dram_agg.txt 2>&1; head -5 gpurun_out/vocoder_${tag}_dram_agg.txt
done
# stage-2 k=11 ResBlock convolution (CTA pairs) and a stage-4 k=3 one (short tiles): launch indices 33 and 59
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 33 -c 1 -o gpurun_out/tc_conv_s2_k11 -f \
  python tools/prof_vocoder.py --precision 6 --iters 0 > /dev/null 2>&1; echo "ncu s2 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 59 -c 2 -o gpurun_out/tc_conv_s4_k3 -f \
  python tools/prof_vocoder.py --precision 6 --iters 0 > /dev/null 2>&1; echo "ncu s4 rc=$?"
