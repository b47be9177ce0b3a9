#!/bin/bash
# Two-GPU session for the peer-store (CUDA IPC) transport: slab tests (both transports), then the
# dam-break slab bench with each transport back to back.
# usage (gpurun --gpus 2): scripts/gpu_n2_ipc.sh [N] [steps]   -> gpurun_out/n<N>_ipc_*
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
N=${1:-2}; STEPS=${2:-100}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@"; }
timeout 400 python -m pytest tests/test_gpu_slabs.py -m gpu -q --timeout=300 -p no:cacheprovider -k "torchrun or loopback" > $OUT/n${N}_ipc_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/n${N}_ipc_pytest.log
for t in ipc nccl; do
  SPHB_BENCH_TRANSPORT=$t run --steps $STEPS --warmup 10 > $OUT/n${N}_bench_dam_$t.json 2> $OUT/n${N}_bench_dam_$t.err; echo "dam $t rc=$?"
  python - "$OUT/n${N}_bench_dam_$t.json" <<'PY' || tail -5 $OUT/n${N}_bench_dam_$t.err
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(j["config"].get("transport"), "value=%.4e ms/step=%.4f e2e=%.4e" % (j["value"], j["ms_per_step"], j["e2e"]["value"]),
      {k: v["ms"] for k, v in j["roofline"]["kernels"].items()}, j["config"].get("transport_fallback", ""), j["config"]["merged_stats"])
PY
done
