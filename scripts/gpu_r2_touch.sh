#!/bin/bash
# Deterministic reorder without the id pass for cells whose population did not change (cell_touch):
# the whole GPU suite on the new build, then A/B against SPHB_TOUCH=0 (variant "notouch").
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_touch_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02_touch_pytest.log
bash scripts/ab.sh "notouch new notouch new" "dam8m" 100
bash scripts/ab.sh "notouch new" "dam64m" 20
bash scripts/ab.sh "notouch new" "drop256k" 500
