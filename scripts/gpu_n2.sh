#!/bin/bash
# Two-GPU session: NCCL transport test, slab bench lines (dam break 8M/GPU, sloshing tank 8M/GPU).
# usage (gpurun --gpus 2): scripts/gpu_n2.sh [N]   -> gpurun_out/n<N>_*
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
N=${1:-2}
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@"; }
timeout 600 python -m pytest tests/test_gpu_slabs.py -m gpu -q --timeout=500 -p no:cacheprovider > $OUT/n${N}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/n${N}_pytest.log
run --steps 100 --warmup 10 > $OUT/n${N}_bench_dam.json 2> $OUT/n${N}_bench_dam.err; echo "dam rc=$?"; tail -1 $OUT/n${N}_bench_dam.json
run --steps 100 --warmup 10 --workload slosh16m > $OUT/n${N}_bench_slosh.json 2> $OUT/n${N}_bench_slosh.err; echo "slosh rc=$?"; tail -1 $OUT/n${N}_bench_slosh.json; tail -5 $OUT/n${N}_bench_slosh.err
