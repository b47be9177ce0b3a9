#!/bin/bash
# compute-sanitizer (one tool) over a 61k-particle drop: all step kernels with staged tiles and the list
# hand-over, blocking and pipelined step statistics, render, in-process slabs.
# usage: scripts/gpu_san_drop.sh [tool]   (default synccheck)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
TOOL=${1:-synccheck}
cat > /tmp/san_drop.py <<'PY'
import ctypes, numpy as np, pi_sph_fluid_b200 as pkg
prm = pkg.default_params(0.005)
f, b = pkg.scene_drop(prm), pkg.scene_boundary(prm)
g = np.ascontiguousarray(np.tile([[0.0, -9.81]], (4, 1)), np.float32)
st = pkg.Stats(); ref = ctypes.byref(st)
with pkg.Simulation(prm) as sim:
    sim.upload(f, b); sim.init_boundary(); sim.compute_accel(0.0, -9.81); sim.step(2, 0.0, -9.81)
    a = sim.step_stats(g[:1])
    t1 = sim.step_stats_begin(g.ctypes.data, 1); t2 = sim.step_stats_begin(g.ctypes.data + 8, 1)
    sim.step_stats_end(t1, ref); sim.step_stats_end(t2, ref)
    out = sim.download(); fr = sim.render()
    print("drop61k ok", len(f), a["max_speed"], st.asdict()["steps"])
cuts = [0, 140, 170, pkg.grid_columns(prm)[1]]
with pkg.SlabGroup(prm, cuts) as grp:
    grp.upload(f, b); grp.init_boundary(); grp.compute_accel(0.0, -9.81); grp.step(2, 0.0, -9.81); grp.synchronize()
    print("slabs ok", grp.stats()["n_fluid"])
PY
PYTHONPATH=. timeout 130 compute-sanitizer --tool $TOOL --error-exitcode 9 python /tmp/san_drop.py > $OUT/san_drop61k_$TOOL.log 2>&1
echo "$TOOL rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok" $OUT/san_drop61k_$TOOL.log | tail -4
