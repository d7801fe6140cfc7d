#!/bin/bash
# One gpurun call that produces everything profiles/ holds for a round (run from the repo root on the GPU box):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r02'
# GPU parity tests, smoke, both bench arms, the ncu launch list of one step, per-launch DRAM traffic of the vocoder
# (valid-length pass) and full captures of the top kernels.  ncu numbers are never bench values.
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
# the opt-in builds of the tcgen05 kernels (switches are read once per process)
DTTS_TC_PAIR64_CLUSTER=1 timeout 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -k "fused or lengths" > gpurun_out/${TAG}_pytest_gpu_cluster64.log 2>&1; echo "pytest (DTTS_TC_PAIR64_CLUSTER=1) rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu_cluster64.log
DTTS_TC_PAIR=0 DTTS_TC_PDL=0 timeout 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -k "fp8_lo or lengths or baseline_shapes or fused" > gpurun_out/${TAG}_pytest_gpu_nopair.log 2>&1; echo "pytest (DTTS_TC_PAIR=0 DTTS_TC_PDL=0) rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu_nopair.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"
head -c 600 gpurun_out/${TAG}_bench_n1.json; echo
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference_arm.json 2> /dev/null; echo "reference arm rc=$?"
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_step.csv python tools/prof_step.py --bank > gpurun_out/${TAG}_ncu_step.log 2>&1; echo "ncu step rc=$?"
python tools/agg_launches.py gpurun_out/${TAG}_launches_step.csv > gpurun_out/${TAG}_launches_step_agg.txt 2>&1; head -14 gpurun_out/${TAG}_launches_step_agg.txt
python tools/prof_vocoder.py --precision 6 --iters 3 --lens 2>&1 | tail -1 | tee gpurun_out/${TAG}_vocoder_times.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  -k regex:"tc_conv|rb_pair|rb_block" --log-file gpurun_out/${TAG}_vocoder_lens_dram_per_launch.csv python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1
python tools/agg_launches.py gpurun_out/${TAG}_vocoder_lens_dram_per_launch.csv ALL > gpurun_out/${TAG}_vocoder_lens_dram_agg.txt 2>&1; head -8 gpurun_out/${TAG}_vocoder_lens_dram_agg.txt
# full captures: the fused pair kernels (C = 32: launch 2 of its 6 = k 7; C = 64: launch 5 = k 11), the two block-fused k = 3
# ResBlocks, a stage-2 k=11 pair-mode convolution, the fused prior flow and the S2PA stream kernel
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rb_pair32 -s 1 -c 1 -o gpurun_out/${TAG}_rb_pair32_k7 -f \
  python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1; echo "ncu pair32 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rb_pair64 -s 4 -c 1 -o gpurun_out/${TAG}_rb_pair64_k11 -f \
  python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1; echo "ncu pair64 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rb_block -c 2 -o gpurun_out/${TAG}_rb_block -f \
  python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1; echo "ncu block rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 27 -c 1 -o gpurun_out/${TAG}_tc_conv_s2_k11 -f \
  python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1; echo "ncu s2 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rb_pair128 -s 0 -c 1 -o gpurun_out/${TAG}_rb_pair128_k3 -f \
  python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1; echo "ncu pair128 rc=$?"
mkdir -p tools/_build
[ -x tools/_build/mma_rate ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dict_tts_b200/csrc -o tools/_build/mma_rate tools/mma_rate.cu
./tools/_build/mma_rate > gpurun_out/${TAG}_mma_rate.log 2>&1; echo "mma_rate rc=$?"
python tools/prof_decode.py --iters 20 2>&1 | tail -4 | tee gpurun_out/${TAG}_decode_times.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_decode.csv \
  python tools/prof_decode.py --ncu > /dev/null 2>&1; echo "ncu decode rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flow_fused -c 1 -o gpurun_out/${TAG}_flow_fused -f \
  python tools/prof_decode.py --ncu > /dev/null 2>&1; echo "ncu flow rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:s2pa_stream -c 1 -o gpurun_out/${TAG}_s2pa_stream -f \
  python tools/prof_acoustic.py --iters 0 --alias > /dev/null 2>&1; echo "ncu s2pa rc=$?"
