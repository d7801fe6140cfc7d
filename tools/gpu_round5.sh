#!/bin/bash
# gpurun call 5: valid-length vocoding (ragged tile schedule): tests first (bounded), then timings, traffic, bench.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q -k "lengths" > gpurun_out/pytest_lens.log 2>&1; echo "lens tests rc=$?" | tee -a gpurun_out/pytest_lens.log
tail -25 gpurun_out/pytest_lens.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python tools/prof_vocoder.py --precision 3 --iters 3 2>&1 | tail -1 | tee gpurun_out/vocoder_lens_times.log
python tools/prof_vocoder.py --precision 3 --iters 3 --lens 2>&1 | tail -1 | tee -a gpurun_out/vocoder_lens_times.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  -k regex:tc_conv --log-file gpurun_out/vocoder_lens_dram.csv python tools/prof_vocoder.py --precision 3 --iters 0 --lens > gpurun_out/ncu_voc_lens.log 2>&1; echo "ncu dram rc=$?"
python tools/agg_launches.py gpurun_out/vocoder_lens_dram.csv ALL > gpurun_out/vocoder_lens_dram_agg.txt 2>&1; head -6 gpurun_out/vocoder_lens_dram_agg.txt
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json
