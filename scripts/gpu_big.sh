#!/bin/bash
# Large-N single-GPU session: bench + ncu of the pair kernels on an 8M-particle dam break.
# usage: scripts/gpu_big.sh [workload]   -> gpurun_out/big_*
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
WL=${1:-dam8m}
timeout 600 python bench.py --workload $WL --steps 100 --warmup 10 --no-cpu-baseline --no-e2e > $OUT/big_bench_$WL.json 2> $OUT/big_bench_$WL.err; echo "bench rc=$?"; cat $OUT/big_bench_$WL.json
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_force|k_density' -s 20 -c 2 -f -o $OUT/big_pair_$WL \
  python bench.py --workload $WL --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/big_ncu_$WL.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_advect|k_scan|k_reorder|k_scatter' -s 40 -c 4 -f -o $OUT/big_build_$WL \
  python bench.py --workload $WL --steps 12 --warmup 3 --no-cpu-baseline --no-e2e >> $OUT/big_ncu_$WL.log 2>&1; echo "ncu2 rc=$?"
ls -la $OUT
