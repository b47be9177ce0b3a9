#!/bin/bash
# Eight-GPU session of the final build: configs[3] / configs[4] at full size on 8 ranks against ONE GPU
# (tests/mg_full_check.py: per-field hashes + oracle strip around a cut), then the N = 8 bench lines.
# usage (gpurun --gpus 8): scripts/gpu_r2_n8.sh
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
N=$(nvidia-smi -L | wc -l)
tr() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 "$@"; }
{ tr tests/mg_full_check.py dam64m 5; tr tests/mg_full_check.py slosh16m 20; } 2>&1 | grep -v "^W\|^\*\|OMP_NUM" > $OUT/r02_mg_full_check_${N}gpu.log
echo "full check rc=$?"; cat $OUT/r02_mg_full_check_${N}gpu.log | tail -8
bash scripts/gpu_scale.sh $N 100 dam64m
bash scripts/gpu_scale.sh $N 200 slosh16m --workload slosh16m
