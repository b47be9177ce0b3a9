#!/bin/bash
# A/B timing of library variants on the GPU box: scripts/ab.sh "<variants>" "<workloads>" [steps]
# (variant "" = the default libsphb200.so; others are libsphb200_<tag>.so built by build_variant)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VARS=${1:-"old new"}; WLS=${2:-"drop256k dam4m"}; STEPS=${3:-200}
for w in $WLS; do for v in $VARS; do
  if [ "$v" = "new" ]; then unset SPHB_LIB_VARIANT; else export SPHB_LIB_VARIANT=$v; fi
  timeout 600 python bench.py --workload $w --steps $STEPS --warmup 10 --no-cpu-baseline --no-e2e 2> gpurun_out/ab_$v.err | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); k = j['roofline']['kernels']
print('$v'.ljust(14), '$w'.ljust(9), 'value=%.4e ms/step=%.4f' % (j['value'], j['ms_per_step']), ' '.join('%s=%.4f' % (n, d['ms']) for n, d in k.items()))
" || tail -3 gpurun_out/ab_$v.err
done; done 2>&1 | tee -a gpurun_out/ab.log
