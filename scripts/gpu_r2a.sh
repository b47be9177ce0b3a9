#!/bin/bash
# round 2, first GPU session: smoke + parity tests with the reference-arithmetic force pass, then its cost
# against fast_force at 262k and 8M particles.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=900 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 $OUT/pytest_gpu.log
for w in drop256k dam8m; do for ff in "" "--fast-force"; do
  timeout 600 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline --no-e2e $ff 2> $OUT/ab.err | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); k = j['roofline']['kernels']
print('$w'.ljust(9), '$ff'.ljust(12), 'value=%.4e ms/step=%.4f' % (j['value'], j['ms_per_step']), ' '.join('%s=%.4f' % (n, d['ms']) for n, d in k.items()))
" || tail -3 $OUT/ab.err
done; done 2>&1 | tee $OUT/r2a_ab.log
