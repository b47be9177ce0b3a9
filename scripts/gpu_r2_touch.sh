#!/bin/bash
# Deterministic reorder without the id pass for cells whose population did not change (cell_touch), and the
# neighbour-list block leaving k_density as one bulk store: the whole GPU suite on the new build, then A/B
# against SPHB_TOUCH=0 (variant "notouch") and SPHB_LIST_BULK_STORE=0 (variant "nobulk").
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_touch_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02_touch_pytest.log
bash scripts/ab.sh "notouch nobulk new notouch nobulk new" "dam8m" 100
bash scripts/ab.sh "notouch nobulk new" "dam64m" 20
bash scripts/ab.sh "notouch nobulk new" "drop256k" 500
