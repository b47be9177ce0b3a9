#!/bin/bash
# Eight-GPU session: BASELINE configs[4] (sloshing tank 16M, tilt trace) on the peer-store transport.
# usage (gpurun --gpus 8): scripts/gpu_n8_slosh_ipc.sh [N] [steps]   -> gpurun_out/n<N>_bench_slosh_ipc.*
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
N=${1:-8}; STEPS=${2:-100}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps $STEPS --warmup 10 --workload slosh16m > $OUT/n${N}_bench_slosh_ipc.json 2> $OUT/n${N}_bench_slosh_ipc.err; echo "slosh ipc rc=$?"
python - "$OUT/n${N}_bench_slosh_ipc.json" <<'PY' || tail -5 $OUT/n${N}_bench_slosh_ipc.err
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(j["config"].get("transport"), "value=%.4e ms/step=%.4f e2e=%.4e" % (j["value"], j["ms_per_step"], j["e2e"]["value"]),
      {k: v["ms"] for k, v in j["roofline"]["kernels"].items()}, j["config"].get("transport_fallback", ""), j["config"]["merged_stats"])
print(j["roofline"].get("kernels_per_rank_ms"))
PY
