#!/bin/bash
# round 2, call M: rb_pair128_kernel with per-CTA weight streams (one bulk copy per weight stage)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q -k "fused or lengths" > gpurun_out/r02m_pytest_fused.log 2>&1; echo "pytest fused rc=$?" | tee -a gpurun_out/r02m_pytest_fused.log
tail -5 gpurun_out/r02m_pytest_fused.log
run() { env "$@" python tools/prof_vocoder.py --precision 6 --iters 4 --lens 2>&1 | tail -1 | sed "s/^/$* /" | tee -a gpurun_out/r02m_ab.log; }
run DTTS_TC_FUSE128=0
run DTTS_TC_P128_TG=1
run DTTS_TC_P128_TG=2
run DTTS_TC_P128_TG=3
run DTTS_TC_P128_TG=4
run DTTS_TC_P128_TG=1 DTTS_TC_P128_ASTAGES=3
run DTTS_TC_P128_TG=2 DTTS_TC_P128_ASTAGES=3
run DTTS_TC_FUSE128=0
for tg in 1 2 3; do
DTTS_TC_P128_TG=$tg timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  -k regex:"rb_pair128" --log-file gpurun_out/r02m_pair128_tg$tg.csv python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1
echo "TG=$tg: $(grep "time_duration" gpurun_out/r02m_pair128_tg$tg.csv | awk -F'","' '{print $NF}' | tr '\n' ' ')"
done
DTTS_TC_P128_TG=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:rb_pair128 -s 4 -c 1 -o gpurun_out/r02m_rb_pair128_k7 -f \
  python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1; echo "ncu pair128 rc=$?"
DTTS_TC_P128_TG=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:rb_pair128 -s 0 -c 1 -o gpurun_out/r02m_rb_pair128_k3 -f \
  python tools/prof_vocoder.py --precision 6 --iters 0 --lens > /dev/null 2>&1; echo "ncu pair128 k3 rc=$?"
