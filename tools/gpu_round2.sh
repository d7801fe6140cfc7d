#!/bin/bash
# gpurun call 2: GPU tests, launch list of one step, DRAM traffic of the vocoder convolutions, two full captures.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python tools/prof_acoustic.py --s2pa-route 0 > gpurun_out/acoustic_routes.log 2>&1
python tools/prof_acoustic.py --s2pa-route 1 >> gpurun_out/acoustic_routes.log 2>&1
cat gpurun_out/acoustic_routes.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_step.csv python tools/prof_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu step rc=$?"
python tools/agg_launches.py gpurun_out/launches_step.csv > gpurun_out/launches_step_agg.txt 2>&1; head -40 gpurun_out/launches_step_agg.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  -k regex:tc_conv --log-file gpurun_out/vocoder_dram.csv python tools/prof_vocoder.py --precision 3 --iters 0 > gpurun_out/ncu_voc.log 2>&1; echo "ncu dram rc=$?"
python tools/agg_launches.py gpurun_out/vocoder_dram.csv ALL > gpurun_out/vocoder_dram_agg.txt 2>&1; head -12 gpurun_out/vocoder_dram_agg.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 17 -c 1 -o gpurun_out/s2pa_gemm -f \
  python tools/prof_acoustic.py --s2pa-route 1 --iters 0 > gpurun_out/ncu_s2pa.log 2>&1; echo "ncu s2pa rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 33 -c 1 -o gpurun_out/tc_conv_s2_k11 -f \
  python tools/prof_vocoder.py --precision 3 --iters 0 > gpurun_out/ncu_tc.log 2>&1; echo "ncu tc rc=$?"
ls -la gpurun_out
