#!/bin/bash
# Eight-GPU session: the two BASELINE multi-GPU configurations as bench lines.
#   configs[3] dam break 64M particles (8M per GPU)    configs[4] sloshing tank 16M, tilt trace
# usage (gpurun --gpus 8): scripts/gpu_n8.sh [N]   -> gpurun_out/n<N>_*
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
N=${1:-8}
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@"; }
run --steps 100 --warmup 10 > $OUT/n${N}_bench_dam.json 2> $OUT/n${N}_bench_dam.err; echo "dam rc=$?"; tail -1 $OUT/n${N}_bench_dam.json | cut -c1-400
run --steps 100 --warmup 10 --workload slosh16m > $OUT/n${N}_bench_slosh.json 2> $OUT/n${N}_bench_slosh.err; echo "slosh rc=$?"; tail -1 $OUT/n${N}_bench_slosh.json | cut -c1-400
run --steps 20 --warmup 3 --impl reference > $OUT/n${N}_bench_ref.json 2> $OUT/n${N}_bench_ref.err; echo "ref rc=$?"; tail -1 $OUT/n${N}_bench_ref.json | cut -c1-600
