#!/bin/bash
# One GPU-box session: smoke (under compute-sanitizer), parity tests, bench, ncu launch list,
# ncu full capture of the two pair kernels.  Everything lands in gpurun_out/.
# usage: scripts/gpu_check.sh [stage ...]   stages: smoke san tests bench launches ncu
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
STAGES=${*:-"smoke tests bench launches ncu"}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
for st in $STAGES; do
  echo "=== stage $st ($(date +%T))"
  case $st in
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log ;;
    san)
      timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -5 $OUT/sanitizer.log ;;
    tests)
      timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 $OUT/pytest_gpu.log ;;
    bench)
      timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 360 --csv --log-file $OUT/launches.csv \
        python bench.py --steps 60 --warmup 5 --no-cpu-baseline > $OUT/launches_bench.log 2>&1; echo "launches rc=$?" ;;
    ncu)
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_force|k_density' -s 40 -c 4 -f -o $OUT/prof_pair \
        python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1; echo "ncu rc=$?"
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_advect|k_scan|k_reorder|k_scatter' -s 40 -c 4 -f -o $OUT/prof_build \
        python bench.py --steps 30 --warmup 5 --no-cpu-baseline >> $OUT/ncu_bench.log 2>&1; echo "ncu2 rc=$?" ;;
  esac
done
ls -la $OUT
