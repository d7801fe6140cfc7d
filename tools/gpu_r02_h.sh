#!/bin/bash
# round 2, call H: fused pair kernel for C = 64 (weights streamed): parity + timing A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r02h_pytest_tc.log 2>&1; echo "pytest tc rc=$?" | tee -a gpurun_out/r02h_pytest_tc.log
tail -12 gpurun_out/r02h_pytest_tc.log
for f in 0 1 0 1; do
  DTTS_TC_FUSE64=$f python tools/prof_vocoder.py --precision 6 --iters 4 --lens 2>&1 | tail -1 | sed "s/^/fuse64=$f /" | tee -a gpurun_out/r02h_fuse64_ab.log
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  -k regex:"tc_conv|rb_pair" --log-file gpurun_out/r02h_vocoder_lens_dram.csv python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1
python tools/agg_launches.py gpurun_out/r02h_vocoder_lens_dram.csv ALL > gpurun_out/r02h_vocoder_lens_dram_agg.txt 2>&1; head -8 gpurun_out/r02h_vocoder_lens_dram_agg.txt
