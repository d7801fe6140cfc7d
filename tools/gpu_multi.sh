#!/bin/bash
# N-GPU records of a round (N = 2, 4 or 8):  /usr/local/graft/bin/gpurun --gpus N --timeout 1200 -- 'bash tools/gpu_multi.sh N r02'
# (1) bench.py under torchrun, both arms, as the driver launches them; (2) the --infer task seam on N ranks (tools/multi_infer.py)
N=${1:-2}
TAG=${2:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench N=$N rc=$?"
head -c 400 gpurun_out/${TAG}_bench_n$N.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
  bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm_n$N.json 2> /dev/null; echo "reference arm N=$N rc=$?"
timeout 900 python tools/multi_infer.py --gpus $N | tee gpurun_out/${TAG}_multi_infer_n$N.json
