"""Generate tests/golden/*.npz from the REFERENCE's own compiled code.

Run in the build container (needs /root/reference to (re)build oracle/_ref):

    python tests/golden/make_golden.py

Every array below is produced by oracle/_ref/libpisph_ref*_strict.so, i.e. the
reference translation unit pi_sph_fluid.c compiled with
`-O2 -fno-fast-math -ffp-contract=off` (IEEE semantics of the source) and
driven through oracle/ref_driver.c in the call order of the reference's main()
(:600-607, :610-641).  Scene arrays are the reference's own lattice (:484-540), built by
oracle_scene_* and cross-checked here against the counts the reference prints
(n_fluid = 269, n_boundary = 162 at R = 0.075).

Files
  drop_R0.075.npz   config 1 (default scene): t=0 and after 1 / 100 / 2000 steps
                    (2000 is post-impact: p up to ~1.8e5 Pa), neighbour lists in the
                    reference's visiting order at t=0 and step 2000 (plus acc_du/acc_dv: the reference's
                    calculate_accelerations re-run on the stored snapshot), and the 1-bpp
                    metaball frame at step 2000.
  drop_R0.02.npz    the same scene at R = 0.02 (N = 3848, sed-widened reference
                    variant): state after 5000 steps (post-impact) and one step later —
                    a non-lattice state for one-step parity of rho, p, a.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import Oracle, Reference, PARTICLE, build  # noqa: E402

OUT = Path(__file__).resolve().parent
G = (0.0, -9.81)   # :442-443 constant gravity without the MPU6050


def lists(ref, a, b, ctx_b):
    flat, off = [], [0]
    for i in range(len(a)):
        nb = ref.find_neighbors(a, b, i, ctx_b)
        flat.extend(nb.tolist())
        off.append(len(flat))
    return np.asarray(flat, np.int32), np.asarray(off, np.int32)


def run(R_tag, R, snaps, with_lists, with_frame):
    o = Oracle(R=R)
    ref = Reference(R=R_tag)
    fluid = o.scene_drop()
    boundary = o.scene_boundary()
    out = {"R": np.float32(R), "fluid_init": fluid.copy(), "boundary_init": boundary.copy()}
    cb = ref.init_boundary(boundary)
    cf = ref.ctx(len(fluid))
    du, dv = ref.compute_accel(fluid, boundary, cf, cb, *G)
    out["boundary"] = boundary.copy()
    step = 0
    for s in snaps:
        if s > step:
            ref.step(fluid, boundary, cf, cb, du, dv, s - step, *G, threads=4)
            step = s
        out[f"fluid_{s}"] = fluid.copy()
        out[f"du_{s}"] = du.copy()
        out[f"dv_{s}"] = dv.copy()
        if s in with_lists:
            # calculate_accelerations of the reference on the snapshot AS STORED (velocities after
            # the closing kick, :637-640) — du_/dv_ above were computed before that kick
            f2 = fluid.copy()
            out[f"acc_du_{s}"], out[f"acc_dv_{s}"] = ref.compute_accel(f2, boundary, cf, cb, *G)
            assert all(np.array_equal(f2[k].view("u4"), fluid[k].view("u4")) for k in ("rho", "p"))
            out[f"ff_list_{s}"], out[f"ff_off_{s}"] = lists(ref, fluid, fluid, cf)
            out[f"fb_list_{s}"], out[f"fb_off_{s}"] = lists(ref, fluid, boundary, cb)
        if s in with_frame:
            buf = np.zeros(1024, np.uint8)
            ref.draw_metaballs(buf, o.pixels(), fluid, cf)
            out[f"frame_{s}"] = buf
    return out


def main():
    build(ref=True)
    d = run(None, 0.075, [0, 1, 100, 2000], {0, 2000}, {0, 2000})
    assert len(d["fluid_init"]) == 269 and len(d["boundary_init"]) == 162
    np.savez_compressed(OUT / "drop_R0.075.npz", **d)
    d = run("0.02", 0.02, [0, 5000, 5001], {5000}, {5000})
    np.savez_compressed(OUT / "drop_R0.02.npz", **d)
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
