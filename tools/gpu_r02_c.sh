#!/bin/bash
# round 2, call C: fused ResBlock pair kernel (stage 4) + staged conv_post: parity, timing A/B, per-launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r02c_pytest_tc.log 2>&1; echo "pytest tc rc=$?" | tee -a gpurun_out/r02c_pytest_tc.log
tail -15 gpurun_out/r02c_pytest_tc.log
for f in 0 1; do
  for i in 1 2 3; do DTTS_TC_FUSE=$f python tools/prof_vocoder.py --precision 6 --iters 4 --lens 2>&1 | tail -1 | sed "s/^/fuse=$f /" | tee -a gpurun_out/r02c_fuse_ab.log; done
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  -k regex:"tc_conv|rb_pair" --log-file gpurun_out/r02c_vocoder_lens_dram.csv python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1
python tools/agg_launches.py gpurun_out/r02c_vocoder_lens_dram.csv ALL > gpurun_out/r02c_vocoder_lens_dram_agg.txt 2>&1; head -8 gpurun_out/r02c_vocoder_lens_dram_agg.txt
timeout 600 python bench.py --quick --steps 20 --warmup 5 > gpurun_out/r02c_bench_quick.json 2> gpurun_out/r02c_bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c_bench_quick.json'))
print(d['ms_per_step'], d['stages_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'])
PY
