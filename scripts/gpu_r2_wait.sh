#!/bin/bash
# k_force: one warp watches the mbarrier, the others wait at the CTA barrier (variant "pollall" = every warp polls);
# k_scatter_ids with four slots per thread.  GPU suite on the new build, then A/B.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_wait_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_wait_pytest.log
bash scripts/ab.sh "pollall new pollall new" "dam8m" 100
bash scripts/ab.sh "pollall new" "dam64m" 20
bash scripts/ab.sh "pollall new" "drop256k" 500
bash scripts/ab.sh "pollall new" "dam4m" 100
