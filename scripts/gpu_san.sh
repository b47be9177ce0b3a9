#!/bin/bash
# compute-sanitizer session: memcheck and racecheck over the smoke scene, memcheck + synccheck over a
# 61k-particle drop (all six step kernels, staged tiles, list hand-over) and over in-process slabs.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/san_drop.py <<'PY'
import numpy as np, pi_sph_fluid_b200 as pkg
prm = pkg.default_params(0.005)
f, b = pkg.scene_drop(prm), pkg.scene_boundary(prm)
with pkg.Simulation(prm) as sim:
    sim.upload(f, b); sim.init_boundary(); sim.compute_accel(0.0, -9.81); sim.step(3, 0.0, -9.81)
    st = sim.step_stats(np.asarray([[0.0, -9.81]], np.float32))
    out = sim.download(); fr = sim.render()
    print("drop61k ok", len(f), st["max_speed"])
cuts = [0, 140, 170, pkg.grid_columns(prm)[1]]
with pkg.SlabGroup(prm, cuts) as g:
    g.upload(f, b); g.init_boundary(); g.compute_accel(0.0, -9.81); g.step(3, 0.0, -9.81); g.synchronize()
    print("slabs ok", g.stats()["n_fluid"])
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/san_${tool}_smoke.log 2>&1
  echo "$tool smoke rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" $OUT/san_${tool}_smoke.log | tail -3
done
for tool in memcheck synccheck; do
  PYTHONPATH=. timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_drop.py > $OUT/san_${tool}_drop.log 2>&1
  echo "$tool drop rc=$?"; grep -E "ERROR SUMMARY|ok" $OUT/san_${tool}_drop.log | tail -4
done
