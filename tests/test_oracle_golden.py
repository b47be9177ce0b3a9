"""The oracle against the golden vectors produced by the reference's own compiled code
(tests/golden/make_golden.py) and against the survey's known answers.  CPU only.

Bar: bit-for-bit — the strict oracle is a restatement of the same IEEE arithmetic."""
import numpy as np
import pytest

from conftest import G, same_bits

PARTICLE_FIELDS = ("x", "y", "u", "v", "m", "rho", "p")


def _state_equal(a, b):
    return all(same_bits(a[f], b[f]) for f in PARTICLE_FIELDS)


def test_scene_matches_reference_printout(oracle_built, golden075):
    # pi_sph_fluid.c:543-545 prints dt=0.000244 (4102 ticks/s), n_fluid=269, n_boundary=162
    o = oracle_built.Oracle()
    fluid, boundary = o.scene_drop(), o.scene_boundary()
    assert len(fluid) == 269 and len(boundary) == 162
    assert int(1 / o.dt) == 4102 and f"{o.dt:f}" == "0.000244"
    assert _state_equal(fluid, golden075["fluid_init"])
    assert _state_equal(boundary, golden075["boundary_init"])
    # SURVEY.md A.1 hex constants
    assert float(o.H).hex() == "0x1.8f5c2a0000000p-4"
    assert float(o.dt).hex() == "0x1.ff2e4a0000000p-13"
    assert float(o.prm.mass).hex() == "0x1.5ac9bc0000000p+2"


def test_appendix_b_known_answers(oracle_built):
    # SURVEY.md Appendix B (strict build of the reference source, config 1, t = 0)
    o = oracle_built.Oracle()
    fluid, boundary = o.scene_drop(), o.scene_boundary()
    gb = o.init_boundary(boundary)
    gf = o.grid(len(fluid))
    du, dv = o.compute_accel(fluid, boundary, gf, gb, *G)
    assert (gf.contents.n_cells, gf.contents.m_cells) == (11, 21)
    psi = boundary["m"]
    assert np.float32(psi.min()) == np.float32(9.7449789) and psi.argmin() == 0
    assert np.float32(psi.max()) == np.float32(22.7172413) and psi[50] == psi.max()
    assert np.float32(psi[2]) == np.float32(13.7315283)
    assert abs(psi.astype("f8").sum() - 3561.15394592) < 1e-6
    rho = fluid["rho"]
    assert np.float32(rho.min()) == np.float32(644.909729) and np.float32(rho.max()) == np.float32(973.388489)
    assert np.float32(rho[0]) == np.float32(644.910034) and np.float32(rho[134]) == np.float32(973.388367)
    assert abs(rho.astype("f8").sum() - 248186.33551) < 1e-4
    assert not fluid["p"].any()
    assert np.float32(du[0]) == np.float32(-7.32106829) and np.float32(dv[0]) == np.float32(-17.0121555)
    assert abs(dv.astype("f8").sum() - (-2638.89011192)) < 1e-6
    cells = o.cell_ids(gf, fluid)
    assert len(np.unique(cells)) == 53 and np.bincount(cells).max() == 9


def test_oracle_bit_exact_vs_golden_config1(oracle_built, golden075):
    g = golden075
    o = oracle_built.Oracle()
    fluid, boundary = g["fluid_init"].copy(), g["boundary_init"].copy()
    gb = o.init_boundary(boundary)
    assert _state_equal(boundary, g["boundary"])
    gf = o.grid(len(fluid))
    du, dv = o.compute_accel(fluid, boundary, gf, gb, *G)
    step = 0
    for snap in (0, 1, 100, 2000):
        o.step(fluid, boundary, gf, gb, du, dv, snap - step, *G)
        step = snap
        assert _state_equal(fluid, g[f"fluid_{snap}"]), snap
        assert same_bits(du, g[f"du_{snap}"]) and same_bits(dv, g[f"dv_{snap}"]), snap
    assert g["fluid_2000"]["p"].max() > 1e5        # the fixture really is post-impact
    assert o.ctr.neighbor_overflows == 0


@pytest.mark.parametrize("which", ["ff", "fb"])
def test_neighbour_lists_match_reference_order(oracle_built, golden075, which):
    g = golden075
    o = oracle_built.Oracle()
    for snap in (0, 2000):
        fluid, boundary = g[f"fluid_{snap}"].copy(), g["boundary"].copy()
        gf, gb = o.grid(len(fluid)), o.grid(len(boundary))
        o.grid_update(gf, fluid)
        o.grid_update(gb, boundary)
        flat, off = g[f"{which}_list_{snap}"], g[f"{which}_off_{snap}"]
        for i in range(len(fluid)):
            mine = o.neighbor_list(fluid, fluid if which == "ff" else boundary, i,
                                   gf if which == "ff" else gb, same=(which == "ff"))
            assert np.array_equal(mine, flat[off[i]:off[i + 1]]), (snap, i)


def test_oracle_bit_exact_vs_golden_R002_post_impact(oracle_built, golden02):
    g = golden02
    o = oracle_built.Oracle(R=0.02)
    boundary = g["boundary"].copy()
    fluid = g["fluid_5000"].copy()
    du, dv = g["du_5000"].copy(), g["dv_5000"].copy()
    gf, gb = o.grid(len(fluid)), o.grid(len(boundary))
    o.grid_update(gb, boundary)
    o.step(fluid, boundary, gf, gb, du, dv, 1, *G)
    assert _state_equal(fluid, g["fluid_5001"])
    assert same_bits(du, g["du_5001"]) and same_bits(dv, g["dv_5001"])
    assert len(fluid) == 3848 and fluid["p"].max() > 1e5


def test_metaball_frame_matches_reference(oracle_built, golden075):
    g = golden075
    o = oracle_built.Oracle()
    for snap in (0, 2000):
        fluid = g[f"fluid_{snap}"].copy()
        gf = o.grid(len(fluid))
        o.grid_update(gf, fluid)
        buf = np.zeros(1024, np.uint8)
        o.draw_metaballs(buf, o.pixels(), fluid, gf)
        assert np.array_equal(buf, g[f"frame_{snap}"])
        assert 0 < np.unpackbits(buf).sum() < 8192


def test_chain_flavour_differs_only_in_last_bits(oracle_built, golden02):
    """The 'chain' oracle replaces powf(x,3|4|7) by the multiply chains gcc -Ofast emits for
    the reference (SURVEY.md §8c allows either); it must stay within float rounding of the
    pinned flavour: rho <= 4 ulp, and identical neighbour structure."""
    g = golden02
    strict, chain = oracle_built.Oracle(R=0.02), oracle_built.Oracle(R=0.02, variant="chain")
    out = []
    for o in (strict, chain):
        fluid, boundary = g["fluid_5000"].copy(), g["boundary"].copy()
        gf, gb = o.grid(len(fluid)), o.grid(len(boundary))
        o.grid_update(gb, boundary)
        du, dv = o.compute_accel(fluid, boundary, gf, gb, *G)
        out.append(fluid)
    rel = np.abs(out[0]["rho"].astype("f8") - out[1]["rho"]) / out[0]["rho"]
    assert rel.max() < 3e-7


def test_gravity_mapping(oracle_built):
    o = oracle_built.Oracle()
    assert o.gravity_from_raw(16384, 0) == (np.float32(0.0), np.float32(-9.81))     # :439-440
    gx, gy = o.gravity_from_raw(0, 16384)
    assert gx == np.float32(9.81) and gy == np.float32(-0.0)


def test_oracle_thread_count_invariance(oracle_built):
    """Per-particle sums are sequential inside one thread, so results do not depend on the
    OpenMP team size — the oracle may use every host core."""
    res = []
    for threads in (1, 5):
        o = oracle_built.Oracle(R=0.02, threads=threads)
        fluid, boundary = o.scene_drop(), o.scene_boundary()
        gb = o.init_boundary(boundary)
        gf = o.grid(len(fluid))
        du, dv = o.compute_accel(fluid, boundary, gf, gb, *G)
        o.step(fluid, boundary, gf, gb, du, dv, 20, *G)
        res.append((fluid.copy(), du.copy(), dv.copy()))
    assert _state_equal(res[0][0], res[1][0]) and same_bits(res[0][1], res[1][1])
