#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_fullsize.py tests/test_gpu_portaspeech.py -m gpu -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02d_pytest.log
tail -15 gpurun_out/r02d_pytest.log
for f in 0 1; do
  for i in 1 2; do DTTS_TC_FUSE=$f python tools/prof_vocoder.py --precision 6 --iters 4 --lens 2>&1 | tail -1 | sed "s/^/fuse=$f /" | tee -a gpurun_out/r02d_fuse_ab.log; done
done
