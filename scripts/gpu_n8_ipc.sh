#!/bin/bash
# Eight-GPU session for the peer-store (CUDA IPC) transport: bit-identity check with interior ranks
# (two neighbours each), then BASELINE configs[3] (dam break 64M) as a bench line.
# usage (gpurun --gpus 8): scripts/gpu_n8_ipc.sh [N] [steps]   -> gpurun_out/n<N>_ipc_*
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
N=${1:-8}; STEPS=${2:-100}
tr() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 "$@"; }
tr tests/mg_nccl_check.py ipc > $OUT/n${N}_ipc_check.log 2>&1; echo "check rc=$?"; grep -v "^W\|^\*" $OUT/n${N}_ipc_check.log | tail -4
tr bench.py --gpus $N --steps $STEPS --warmup 10 > $OUT/n${N}_bench_dam_ipc.json 2> $OUT/n${N}_bench_dam_ipc.err; echo "dam ipc rc=$?"
python - "$OUT/n${N}_bench_dam_ipc.json" <<'PY' || tail -5 $OUT/n${N}_bench_dam_ipc.err
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(j["config"].get("transport"), "value=%.4e ms/step=%.4f e2e=%.4e" % (j["value"], j["ms_per_step"], j["e2e"]["value"]),
      {k: v["ms"] for k, v in j["roofline"]["kernels"].items()}, j["config"].get("transport_fallback", ""), j["config"]["merged_stats"])
PY
