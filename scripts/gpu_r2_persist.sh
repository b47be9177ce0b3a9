#!/bin/bash
# Pair kernels with static-stride persistent CTAs (SPHB_PERSISTENT=2, grid = resident slots x 1/2/4/8) against one
# CTA per chunk: parity subset on one variant, then timing.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SPHB_LIB_VARIANT=p2x2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py -x -q -m gpu > gpurun_out/r02_p2_pytest.log 2>&1
echo "pytest p2x2 rc=$?"; tail -3 gpurun_out/r02_p2_pytest.log
bash scripts/ab.sh "new p2x1 p2x2 p2x4 p2x8 new p2x2" "dam8m" 100
bash scripts/ab.sh "new p2x2 p2x4" "dam64m" 20
bash scripts/ab.sh "new p2x2 p2x4" "drop256k" 500
