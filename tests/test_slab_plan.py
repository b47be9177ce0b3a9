"""CPU tests of the multi-GPU host logic: the slab planner of the C ABI, the slab
scene builders, and — over a world_size-2 `gloo` job — the halo/migration protocol itself, emulated
on the host with the oracle doing the physics (tests/slab_emul.py)."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import same_bits

ROOT = Path(__file__).resolve().parents[1]


def test_plan_cuts_properties(lib_built):
    pkg = lib_built
    rng = np.random.default_rng(7)
    for cols, world in [(77, 2), (77, 8), (635, 8), (8703, 8), (40, 1)]:
        hist = rng.integers(0, 50, cols).astype(np.uint64)
        hist[: cols // 5] = 0
        cuts = pkg.plan_cuts(hist, world, 4)
        assert cuts[0] == 0 and cuts[-1] == cols and len(cuts) == world + 1
        assert (np.diff(cuts) >= (4 if world > 1 else 1)).all()
        if cols >= 600:      # quantiles: every slab within a few columns' worth of the ideal share
            per = np.array([hist[cuts[r]:cuts[r + 1]].sum() for r in range(world)], float)
            assert np.abs(per - hist.sum() / world).max() <= 4 * hist.max()
    # all particles in one column: the minimum width still holds and the cuts stay ordered
    hist = np.zeros(100, np.uint64); hist[50] = 1000
    cuts = pkg.plan_cuts(hist, 8, 4)
    assert (np.diff(cuts) >= 4).all() and cuts[-1] == 100
    with pytest.raises(pkg.SphbError):
        pkg.plan_cuts(np.ones(10, np.uint64), 4, 4)
    # a cost per column (the cells a rank's scan walks): a dam break's dry half is shared out — the last rank
    # gets fewer particles for its many empty columns, and the total cost per rank is level
    hist = np.zeros(8703, np.uint64); hist[:4350] = 14700
    plain = pkg.plan_cuts(hist, 8, 4)
    cost = pkg.api.CELL_COST * 4352
    fair = pkg.plan_cuts(hist, 8, 4, column_cost=cost)
    assert plain[-2] < fair[-2] and fair[-1] == 8703       # the last cut moves right: fewer particles for the last rank
    per = np.array([hist[fair[r]:fair[r + 1]].sum() + cost * (fair[r + 1] - fair[r]) for r in range(8)])
    assert np.abs(per - per.mean()).max() <= 2 * (14700 + cost)
    assert hist[fair[7]:].sum() < hist[plain[7]:].sum()


def test_column_of_matches_the_reference_binning(lib_built, oracle_built, golden02):
    """sphb_column_of is :112 — the same float32 divide + truncation the oracle bins with."""
    pkg = lib_built
    o = oracle_built.Oracle(R=0.02)
    f = golden02["fluid_5000"]
    g = o.grid(len(f))
    prm = pkg.default_params(0.02)
    rows, cols = pkg.grid_columns(prm)
    assert (rows, cols) == (g.contents.n_cells, g.contents.m_cells)
    assert np.array_equal(pkg.columns_of(prm, f["x"]), o.cell_ids(g, f) % cols)
    hist = pkg.column_histogram(prm, f)
    assert hist.sum() == len(f) and np.array_equal(hist, np.bincount(o.cell_ids(g, f) % cols, minlength=cols))


def test_slab_scene_builder_tiles_the_block_scene(lib_built):
    pkg = lib_built
    R = 0.01
    prm = pkg.default_params(R)
    box = (2 * R, 2.0, 2 * R, 0.5)
    full = pkg.scene_block(prm, *box)
    hist = pkg.scene_block_column_hist(prm, *box)
    assert np.array_equal(hist, pkg.column_histogram(prm, full))
    cuts = pkg.plan_cuts(hist, 5)
    at = 0
    for r in range(5):
        part, base = pkg.scene_block_slab(prm, *box, int(cuts[r]), int(cuts[r + 1]))
        assert base == at and np.array_equal(part, full[base:base + len(part)])
        assert len(part) == hist[cuts[r]:cuts[r + 1]].sum()
        at += len(part)
    assert at == len(full)


def test_merge_stats(lib_built):
    import ctypes as C
    pkg = lib_built
    from pi_sph_fluid_b200.api import Stats
    per = (Stats * 3)()
    per[0].mass, per[0].n_fluid, per[0].max_rho, per[0].min_rho, per[0].max_speed = 1.0, 10, 1001.0, 990.0, 2.0
    per[1].n_fluid = 0                                   # an empty slab must not contribute min/max rho
    per[2].mass, per[2].n_fluid, per[2].max_rho, per[2].min_rho, per[2].max_speed, per[2].n_lost = 2.0, 5, 1003.0, 995.0, 1.0, 3
    out = Stats()
    assert pkg.lib().sphb_mg_merge_stats(per, 3, C.byref(out)) == 0
    assert (out.mass, out.n_fluid, out.max_rho, out.min_rho, out.max_speed, out.n_lost) == (3.0, 15, 1003.0, 990.0, 2.0, 3)


@pytest.mark.parametrize("world", [2, 3])
def test_halo_and_migration_protocol_over_gloo(lib_built, oracle_built, tmp_path, world):
    """world_size-2 and -3 gloo jobs (the middle rank of three has two neighbours): each rank runs the
    slab protocol on the host with the oracle as the physics; the union of the owned particles must
    equal the oracle's single-domain run bit for bit."""
    R, steps, g = 0.03, 120, (300.0, -9.81)      # huge sideways pull: particles cross the cuts
    port = 29650 + 7 * world
    procs = [subprocess.Popen([sys.executable, str(ROOT / "tests" / "slab_emul.py"), str(r), str(world), str(port),
                               str(R), str(steps), str(g[0]), str(g[1]), str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)[-4000:]

    o = oracle_built.Oracle(R=R, variant="chain")
    fluid, boundary = o.scene_drop(), o.scene_boundary()
    gb = o.init_boundary(boundary); gf = o.grid(len(fluid))
    du, dv = o.compute_accel(fluid, boundary, gf, gb, *g)
    o.step(fluid, boundary, gf, gb, du, dv, steps, *g)

    seen = np.zeros(len(fluid), int)
    migrated = 0
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        ids = d["ids"]
        seen[ids] += 1
        migrated += int(d["migrated"])
        for fld in ("x", "y", "u", "v", "rho", "p"):
            assert same_bits(d["fluid"][fld], fluid[fld][ids]), (r, fld)
        assert same_bits(d["du"], du[ids]) and same_bits(d["dv"], dv[ids])
    assert (seen == 1).all()
    assert migrated > 0
