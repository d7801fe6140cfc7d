#!/bin/bash
# round 2, call I: conv_post folded into the last fused pair
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02i_pytest_gpu.log
tail -8 gpurun_out/r02i_pytest_gpu.log
for f in 0 1 0 1; do
  DTTS_TC_FOLD_POST=$f python tools/prof_vocoder.py --precision 6 --iters 4 --lens 2>&1 | tail -1 | sed "s/^/fold=$f /" | tee -a gpurun_out/r02i_fold_ab.log
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  -k regex:"rb_pair32|conv_post" --log-file gpurun_out/r02i_post.csv python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1
python tools/agg_launches.py gpurun_out/r02i_post.csv ALL 2>&1 | head -6
timeout 300 python bench.py --quick --steps 20 --warmup 5 > gpurun_out/r02i_bench_quick.json 2> gpurun_out/r02i_bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02i_bench_quick.json'))
print(round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stages_ms'].items()}, round(d['e2e']['ms_per_step'],3), d['roofline']['frac'])
PY
