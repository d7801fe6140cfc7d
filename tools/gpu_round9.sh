#!/bin/bash
# gpurun call 9: A/B of the packed-fp32 epilogue on one box (scalar build = libdtts_nopack.so), alternating runs.
mkdir -p gpurun_out
cp dict_tts_b200/libdtts.so /tmp/libdtts_packed.so
for rep in 1 2; do
  for v in packed nopack; do
    if [ $v = packed ]; then cp /tmp/libdtts_packed.so dict_tts_b200/libdtts.so; else cp dict_tts_b200/libdtts_nopack.so dict_tts_b200/libdtts.so; fi
    echo "$v full: $(python tools/prof_vocoder.py --precision 3 --iters 3 2>&1 | tail -1)" | tee -a gpurun_out/ab_packed.log
    echo "$v lens: $(python tools/prof_vocoder.py --precision 3 --iters 3 --lens 2>&1 | tail -1)" | tee -a gpurun_out/ab_packed.log
  done
done
cp dict_tts_b200/libdtts_nopack.so dict_tts_b200/libdtts.so
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  -k regex:tc_conv --log-file gpurun_out/vocoder_full_times_nopack.csv python tools/prof_vocoder.py --precision 3 --iters 0 > /dev/null 2>&1
cp /tmp/libdtts_packed.so dict_tts_b200/libdtts.so
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  -k regex:tc_conv --log-file gpurun_out/vocoder_full_times_packed.csv python tools/prof_vocoder.py --precision 3 --iters 0 > /dev/null 2>&1
echo done
