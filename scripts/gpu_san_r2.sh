#!/bin/bash
# compute-sanitizer over the kernels changed in round 2 (small scene, cell marks forced on): the marks in
# k_advect_bin / k_bin_recv, the marked paths of k_scatter_ids / k_reorder (cell_start_prev), the bulk store of
# the neighbour-list block in k_density, the reference-arithmetic force pass, the slab re-cut kernels.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/san_r2.py <<'PY'
import numpy as np, pi_sph_fluid_b200 as pkg
prm = pkg.default_params(0.02)
f, b = pkg.scene_drop(prm), pkg.scene_boundary(prm)
with pkg.Simulation(prm) as sim:
    sim.upload(f, b); sim.init_boundary(); sim.compute_accel(30.0, -9.81); sim.step(40, 30.0, -9.81)
    out = sim.download()
    print("single ok", sim.reorder_marks(), float(out[0]["x"].mean()))
_, cols = pkg.grid_columns(prm)
with pkg.SlabGroup(prm, [0, 40, 60, cols], halo_capacity=4096) as grp:
    grp.upload(f, b); grp.init_boundary(); grp.compute_accel(30.0, -9.81); grp.step(20, 30.0, -9.81)
    grp.rebalance(); grp.step(10, 30.0, -9.81); grp.synchronize()
    print("group ok", grp.stats()["n_fluid"])
PY
for tool in memcheck racecheck synccheck; do
  SPHB_TOUCH_MIN_SLOTS=0 PYTHONPATH=. timeout 200 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_r2.py > $OUT/san_r2_${tool}.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok|Error|hazard" $OUT/san_r2_${tool}.log | head -8
done
