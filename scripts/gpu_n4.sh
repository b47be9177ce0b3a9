#!/bin/bash
# Four-GPU bench line (dam break, 8M particles per GPU).  usage (gpurun --gpus 4): scripts/gpu_n4.sh
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 4 --steps 100 --warmup 10 > gpurun_out/n4_bench_dam.json 2> gpurun_out/n4_bench_dam.err; echo "dam rc=$?"
tail -1 gpurun_out/n4_bench_dam.json | cut -c1-300
