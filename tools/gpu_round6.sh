#!/bin/bash
# gpurun call 6: cursor-based tile decode: tests, vocoder timings (full / valid lengths), per-launch list, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python tools/prof_vocoder.py --precision 3 --iters 3 2>&1 | tail -1 | tee gpurun_out/vocoder_lens_times.log
python tools/prof_vocoder.py --precision 3 --iters 3 --lens 2>&1 | tail -1 | tee -a gpurun_out/vocoder_lens_times.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  -k regex:tc_conv --log-file gpurun_out/vocoder_lens_dram.csv python tools/prof_vocoder.py --precision 3 --iters 0 --lens > gpurun_out/ncu_voc_lens.log 2>&1; echo "ncu dram rc=$?"
python tools/agg_launches.py gpurun_out/vocoder_lens_dram.csv ALL > gpurun_out/vocoder_lens_dram_agg.txt 2>&1; head -6 gpurun_out/vocoder_lens_dram_agg.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  -k regex:tc_conv --log-file gpurun_out/vocoder_full_times.csv python tools/prof_vocoder.py --precision 3 --iters 0 > gpurun_out/ncu_voc_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json
