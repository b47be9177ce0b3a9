#!/bin/bash
# Full bench lines (value + e2e + roofline) of single-GPU workloads other than the default:
# usage: scripts/gpu_lines.sh [workload ...]   (default: dam8m dam4m)  -> gpurun_out/line_<workload>.json
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
for w in ${*:-dam8m dam4m}; do
  timeout 200 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline > $OUT/line_$w.json 2> $OUT/line_$w.err; echo "$w rc=$?"
  python - "$OUT/line_$w.json" <<'PY' || tail -3 $OUT/line_$w.err
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(j["config"]["workload"], "value=%.4e ms/step=%.4f e2e=%.4e" % (j["value"], j["ms_per_step"], j["e2e"]["value"]),
      {k: v["ms"] for k, v in j["roofline"]["kernels"].items()}, "hbm frac", j["roofline"]["frac"], j["roofline"]["fp32"])
PY
done
