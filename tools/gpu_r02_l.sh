#!/bin/bash
# round 2, call L: rb_pair128_kernel with the deep one-tap weight ring; lean MMA issue-rate microbenchmark
mkdir -p gpurun_out
timeout 120 ./tools/_build/mma_rate > gpurun_out/r02l_mma_rate.log 2>&1; echo "mma_rate rc=$?"; cat gpurun_out/r02l_mma_rate.log
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q -k "fused or lengths" > gpurun_out/r02l_pytest_fused.log 2>&1; echo "pytest fused rc=$?" | tee -a gpurun_out/r02l_pytest_fused.log
tail -5 gpurun_out/r02l_pytest_fused.log
run() { env "$@" python tools/prof_vocoder.py --precision 6 --iters 4 --lens 2>&1 | tail -1 | sed "s/^/$* /" | tee -a gpurun_out/r02l_ab.log; }
run DTTS_TC_FUSE128=0
run DTTS_TC_FUSE128=1
run DTTS_TC_P128_ASTAGES=3
run DTTS_TC_P128_TG=2
run DTTS_TC_P128_WSTAGES=6
run DTTS_TC_FUSE128=0
run DTTS_TC_FUSE128=1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  -k regex:"rb_pair128" --log-file gpurun_out/r02l_pair128.csv python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1
grep "time_duration" gpurun_out/r02l_pair128.csv | awk -F'","' '{print $NF}' | tr '\n' ' '; echo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rb_pair128 -s 4 -c 1 -o gpurun_out/r02l_rb_pair128_k7 -f \
  python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1; echo "ncu pair128 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rb_pair128 -s 0 -c 1 -o gpurun_out/r02l_rb_pair128_k3 -f \
  python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1; echo "ncu pair128 k3 rc=$?"
