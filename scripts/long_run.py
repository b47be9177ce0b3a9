#!/usr/bin/env python
"""Long run at full size on one GPU: the 4M-particle dam break (BASELINE configs[2]) for 20,000 leapfrog
steps, statistics every 2,000 steps -> gpurun_out/long_run_dam4m.json (argv[2] renames it).  Checks that mass is
exact, that nothing escapes or overflows and that no field goes non-finite while the column collapses; the
SHA-256 of the final state lets two runs be compared bit for bit (e.g. SPHB_TOUCH_MIN_SLOTS=1073741824 against
the default: the deterministic reorder with and without the cell marks)."""
import hashlib, json, os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import pi_sph_fluid_b200 as pkg

steps, every = (int(sys.argv[1]) if len(sys.argv) > 1 else 20000), 2000
R = 0.0005
prm = pkg.default_params(R)
fluid, boundary = pkg.scene_block(prm, 2 * R, 2.0, 2 * R, 0.5), pkg.scene_boundary(prm)
m_total = float(fluid["m"].astype("f8").sum())
rows = []
with pkg.Simulation(prm) as sim:
    sim.upload(fluid, boundary); sim.init_boundary(); sim.compute_accel(0.0, -9.81)
    t0 = time.perf_counter()
    for done in range(0, steps, every):
        sim.step(every, 0.0, -9.81)
        st = sim.stats()
        rows.append({"step": done + every, "t_sim": (done + every) * float(prm.dt), "mass_rel_err": abs(st["mass"] - m_total) / m_total,
                     "mom_x": st["mom_x"], "mom_y": st["mom_y"], "kinetic": st["kinetic"], "max_speed": st["max_speed"],
                     "max_rho": st["max_rho"], "min_rho": st["min_rho"], "n_escaped": st["n_escaped"],
                     "max_cell_count": st["max_cell_count"]})
        print(rows[-1], flush=True)
    wall = time.perf_counter() - t0
    f, du, dv = sim.download()
finite = bool(all(np.isfinite(f[k]).all() for k in ("x", "y", "u", "v", "rho", "p")) and np.isfinite(du).all() and np.isfinite(dv).all())
h = hashlib.sha256()
for k in ("x", "y", "u", "v", "rho", "p"):
    h.update(np.ascontiguousarray(f[k]).tobytes())
h.update(np.ascontiguousarray(du).tobytes()); h.update(np.ascontiguousarray(dv).tobytes())
out = {"workload": "dam_break_R0.0005_4M (BASELINE configs[2])", "n_fluid": int(len(fluid)), "steps": steps, "wall_s": round(wall, 2),
       "state_sha256": h.hexdigest(), "reorder_marks_env": os.environ.get("SPHB_TOUCH_MIN_SLOTS", "default (marks on: >= 2^20 slots)"),
       "updates_per_s_incl_stats": len(fluid) * steps / wall, "all_finite": finite,
       "x_front_max": float(f["x"].max()), "y_max": float(f["y"].max()), "samples": rows}
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / (sys.argv[2] if len(sys.argv) > 2 else "long_run_dam4m.json")).write_text(json.dumps(out, indent=1))
assert finite and all(r["mass_rel_err"] < 1e-12 and r["n_escaped"] == 0 for r in rows), "long run failed its checks"
print("long run ok:", steps, "steps,", round(wall, 1), "s, state sha256", h.hexdigest()[:16])
