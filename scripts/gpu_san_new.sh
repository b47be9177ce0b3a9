#!/bin/bash
# compute-sanitizer over the kernels added for the peer-store transport and the pipelined step
# statistics (small scene: k_halo_signal / k_bin_recv wait on a world-1 loopback, slot atomics,
# stats_fold_deliver in k_force's last CTA, in k_advect_bin's CTA 0 and in k_stats_deliver).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/san_new.py <<'PY'
import ctypes, numpy as np, pi_sph_fluid_b200 as pkg
prm = pkg.default_params(0.02)
f, b = pkg.scene_drop(prm), pkg.scene_boundary(prm)
g = np.ascontiguousarray(np.tile([[0.0, -9.81]], (8, 1)), np.float32)
st = pkg.Stats(); ref = ctypes.byref(st)
with pkg.Simulation(prm) as sim:
    sim.upload(f, b); sim.init_boundary(); sim.compute_accel(0.0, -9.81)
    a = sim.step_stats(g[:1])
    prev = None
    for i in range(4):
        t = sim.step_stats_begin(g.ctypes.data + 8 * i, 1)
        if prev is not None: sim.step_stats_end(prev, ref)
        prev = t
    sim.step_stats_end(prev, ref)
    print("stats ok", a["max_speed"], st.asdict()["steps"])
_, cols = pkg.grid_columns(prm)
s = pkg.Slab(prm, 0, 1, 0, cols)
s.connect_ipc([s.ipc_handle()])
s.upload(f, b, ids=np.arange(len(f), dtype=np.uint32)); s.init_boundary(); s.compute_accel(0.0, -9.81); s.step(3, 0.0, -9.81)
t = s.step_stats_begin(g.ctypes.data, 1); s.step_stats_end(t, ref)
print("slab ok", st.asdict()["n_fluid"]); s.disconnect_ipc(); s.close()
with pkg.SlabGroup(prm, [0, 40, 60, cols], halo_capacity=4096) as grp:
    grp.upload(f, b); grp.init_boundary(); grp.compute_accel(0.0, -9.81); grp.step(3, 0.0, -9.81); grp.synchronize()
    print("group ok", grp.stats()["n_fluid"])
PY
for tool in memcheck racecheck; do
  PYTHONPATH=. timeout 110 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_new.py > $OUT/san_new_${tool}.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok|Error|hazard" $OUT/san_new_${tool}.log | head -8
done
