#!/bin/bash
# round 2, call K: rb_pair128_kernel (fused ResBlock pairs of the C = 128 stage on CTA pairs) + the MMA issue-rate microbenchmark
mkdir -p gpurun_out
timeout 120 ./tools/_build/mma_rate > gpurun_out/r02k_mma_rate.log 2>&1; echo "mma_rate rc=$?"; cat gpurun_out/r02k_mma_rate.log
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q -k "fused" > gpurun_out/r02k_pytest_fused.log 2>&1; echo "pytest fused rc=$?" | tee -a gpurun_out/r02k_pytest_fused.log
tail -15 gpurun_out/r02k_pytest_fused.log
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_fullsize.py -m gpu -q > gpurun_out/r02k_pytest_tc.log 2>&1; echo "pytest tc rc=$?" | tee -a gpurun_out/r02k_pytest_tc.log
tail -5 gpurun_out/r02k_pytest_tc.log
for f in 0 1 0 1; do
  DTTS_TC_FUSE128=$f python tools/prof_vocoder.py --precision 6 --iters 4 --lens 2>&1 | tail -1 | sed "s/^/fuse128=$f /" | tee -a gpurun_out/r02k_fuse128_ab.log
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  -k regex:"tc_conv|rb_pair" --log-file gpurun_out/r02k_vocoder_lens_dram_per_launch.csv python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1
python tools/agg_launches.py gpurun_out/r02k_vocoder_lens_dram_per_launch.csv ALL > gpurun_out/r02k_vocoder_lens_dram_agg.txt 2>&1; head -8 gpurun_out/r02k_vocoder_lens_dram_agg.txt
grep "rb_pair128" gpurun_out/r02k_vocoder_lens_dram_per_launch.csv | grep "time_duration" | awk -F'","' '{print $NF}' | tr '\n' ' '; echo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rb_pair128 -s 4 -c 1 -o gpurun_out/r02k_rb_pair128_k7 -f \
  python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1; echo "ncu pair128 rc=$?"
