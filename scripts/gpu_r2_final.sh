#!/bin/bash
# Round-2 measurement session on ONE GPU: bench line (as the driver runs it), ncu launch list of the same
# command, ncu --set full of the pair and build kernels (configs[3] on one GPU and the 8M-per-GPU slab load).
# usage: scripts/gpu_r2_final.sh [stage ...]   stages: bench launches ncu64 ncu8   -> gpurun_out/r02_*
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
STAGES=${*:-"bench launches ncu64 ncu8"}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/r02_gpu.csv 2>&1
for st in $STAGES; do
  echo "=== stage $st ($(date +%T))"
  case $st in
    bench)
      timeout 1500 python bench.py > $OUT/r02_bench_n1_dam64m.json 2> $OUT/r02_bench_n1.err; echo "bench rc=$?"; tail -3 $OUT/r02_bench_n1.err
      python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02_bench_n1_dam64m.json").read().strip().splitlines()[-1])
print("value=%.4e ms/step=%.4f e2e=%.4e (%.3f ms/step)" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"]), j["clocks"])
print({k: v["ms"] for k, v in j["roofline"]["kernels"].items()})
for k, v in j.get("secondary", {}).items():
    print(" ", k, "value=%.4e ms/step=%.5f" % (v["value"], v["ms_per_step"]), v.get("e2e", {}).get("ms_per_step"), v.get("cpu_baseline", {}).get("value"), v.get("force_ms"))
print("cpu_baseline", j.get("cpu_baseline"))
PY
      ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $OUT/r02_launches_dam64m.csv \
        python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/r02_launches_bench.log 2>&1; echo "launches rc=$?" ;;
    ncu64)
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_force|k_density|k_advect|k_scan|k_reorder|k_scatter' -s 48 -c 6 -f -o $OUT/r02_dam64m \
        python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/r02_ncu64.log 2>&1; echo "ncu64 rc=$?"
      python scripts/ncu_summary.py $OUT/r02_dam64m.ncu-rep > $OUT/r02_ncu_dam64m.txt 2>&1
      ncu -i $OUT/r02_dam64m.ncu-rep --page source --csv -k regex:k_force > $OUT/r02_force_src_dam64m.csv 2>/dev/null
      ncu -i $OUT/r02_dam64m.ncu-rep --page source --csv -k regex:k_density > $OUT/r02_density_src_dam64m.csv 2>/dev/null ;;
    ncu8)
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_force|k_density|k_advect|k_scan|k_reorder|k_scatter' -s 48 -c 6 -f -o $OUT/r02_dam8m \
        python bench.py --workload dam8m --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/r02_ncu8.log 2>&1; echo "ncu8 rc=$?"
      python scripts/ncu_summary.py $OUT/r02_dam8m.ncu-rep > $OUT/r02_ncu_dam8m.txt 2>&1 ;;
  esac
done
ls -la $OUT | grep r02_
