#!/bin/bash
# N-GPU bench line (strong scaling on configs[3] unless a workload is given) -> gpurun_out/n<N>_bench_<tag>.json
# usage (gpurun --gpus N): scripts/gpu_scale.sh N [steps] [tag] [extra bench args...]
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
N=${1:-8}; STEPS=${2:-100}; TAG=${3:-dam64m}; shift 3 2>/dev/null || true
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps $STEPS --warmup 10 "$@" > $OUT/n${N}_bench_$TAG.json 2> $OUT/n${N}_bench_$TAG.err; echo "bench rc=$?"
python - "$OUT/n${N}_bench_$TAG.json" <<'PY' || tail -5 $OUT/n${N}_bench_$TAG.err
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(j["config"].get("transport"), "value=%.4e ms/step=%.4f e2e=%.4e" % (j["value"], j["ms_per_step"], j["e2e"]["value"]))
print("particles/gpu", j["config"]["particles_per_gpu"])
for k, v in j["roofline"]["kernels_per_rank_ms"].items():
    print("  %-10s" % k, v)
print(j["config"].get("transport_fallback", ""), j["config"]["merged_stats"])
PY
