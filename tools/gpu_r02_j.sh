#!/bin/bash
# round 2, call J: 2-CTA weight multicast in rb_pair64_kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r02j_pytest_tc.log 2>&1; echo "pytest tc rc=$?" | tee -a gpurun_out/r02j_pytest_tc.log
tail -8 gpurun_out/r02j_pytest_tc.log
for f in 0 1 0 1; do
  DTTS_TC_PAIR64_CLUSTER=$f python tools/prof_vocoder.py --precision 6 --iters 4 --lens 2>&1 | tail -1 | sed "s/^/cluster=$f /" | tee -a gpurun_out/r02j_cluster_ab.log
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  -k regex:"rb_pair64" --log-file gpurun_out/r02j_pair64.csv python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1
grep "time_duration" gpurun_out/r02j_pair64.csv | awk -F'","' '{print $NF}' | tr '\n' ' '; echo
DTTS_TC_PAIR64_CLUSTER=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  -k regex:"rb_pair64" --log-file gpurun_out/r02j_pair64_nocluster.csv python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1
grep "time_duration" gpurun_out/r02j_pair64_nocluster.csv | awk -F'","' '{print $NF}' | tr '\n' ' '; echo
