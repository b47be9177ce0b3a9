"""The arithmetic the CUDA kernels execute (csrc/sph_math.cuh + sph_consts.h), compiled for
the HOST by tests/emu/emu.cpp and run in the kernels' visiting order, against the oracle.
CPU only; this is how formula or ordering mistakes are caught without a GPU.  The same
comparisons run on the real kernels in test_gpu_parity.py.

Tolerances (DESIGN.md "Parity"):
  neighbour sets / order, cell ids        exact
  rho, p vs the chain oracle              bit-for-bit (same IEEE ops, same order)
  rho vs the libm-powf oracle / golden    1e-6 relative   (north star: 1e-4)
  p   vs the libm-powf oracle / golden    max(1e-4*p, 160 Pa): 160 Pa = B*7*1e-6, the pressure
                                          change of a 1e-6 relative density change
  a   vs the chain oracle                 bit-for-bit in the default mode (force_pair_strict: the reference's
                                          own types and roundings); ||da|| <= 1e-4 * max(||a||, G) for
                                          fast_force = 1 and against the libm-powf golden (north star: 1e-4)
"""
import ctypes as C

import numpy as np
import pytest

from conftest import G, same_bits

TOL_A = 1e-4
TOL_RHO = 1e-6


def P(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def emu_accel(emu, prm, fluid, boundary, want_lists=False):
    f = fluid.copy()
    du, dv = np.zeros(len(f), np.float32), np.zeros(len(f), np.float32)
    cnt = np.zeros(len(f), np.int32) if want_lists else None
    lst = np.zeros((len(f), 64), np.int32) if want_lists else None
    emu.emu_compute_accel(C.byref(prm), P(f), len(f), P(boundary), len(boundary), C.c_float(G[0]),
                          C.c_float(G[1]), P(du), P(dv), P(cnt), P(lst), 64)
    return f, du, dv, cnt, lst


def accel_err(du, dv, rdu, rdv):
    d = np.hypot(du.astype("f8") - rdu, dv.astype("f8") - rdv)
    return d / np.maximum(np.hypot(rdu.astype("f8"), rdv), 9.81)


def oracle_accel(pyoracle, R, variant, fluid, boundary):
    o = pyoracle.Oracle(R=R, variant=variant)
    f, b = fluid.copy(), boundary.copy()
    gf, gb = o.grid(len(f)), o.grid(len(b))
    o.grid_update(gb, b)
    du, dv = o.compute_accel(f, b, gf, gb, *G)
    return f, du, dv


@pytest.mark.parametrize("name,R,snap", [("golden075", 0.075, 0), ("golden075", 0.075, 2000), ("golden02", 0.02, 5000)])
def test_one_pass_parity(request, oracle_built, lib_built, emu, name, R, snap):
    g = request.getfixturevalue(name)
    prm = lib_built.default_params(R)
    fluid, boundary = g[f"fluid_{snap}"], g["boundary"]
    ef, edu, edv, cnt, lst = emu_accel(emu, prm, fluid, boundary, want_lists=True)

    # neighbour sets and visiting order == the reference's (golden lists come from its find_neighbors)
    off, flat = g[f"ff_off_{snap}"], g[f"ff_list_{snap}"]
    for i in range(len(fluid)):
        assert np.array_equal(lst[i, :cnt[i]], flat[off[i]:off[i + 1]]), i

    # chain oracle: rho, p and (default mode) the accelerations bit-for-bit
    cf, cdu, cdv = oracle_accel(oracle_built, R, "chain", fluid, boundary)
    assert same_bits(ef["rho"], cf["rho"]) and same_bits(ef["p"], cf["p"])
    assert same_bits(edu, cdu) and same_bits(edv, cdv)
    # fast_force = 1: within 1e-4
    fprm = lib_built.default_params(R)
    fprm.fast_force = 1
    _, fdu, fdv, _, _ = emu_accel(emu, fprm, fluid, boundary)
    assert accel_err(fdu, fdv, cdu, cdv).max() < TOL_A
    assert not (same_bits(fdu, cdu) and same_bits(fdv, cdv))      # it really is the other arithmetic

    # reference-built golden (libm powf): rho 1e-6, p max(1e-4 p, 160 Pa)
    ref = g[f"fluid_{snap}"]     # rho/p in the fixture are the reference's
    assert (np.abs(ef["rho"].astype("f8") - ref["rho"]) / ref["rho"]).max() < TOL_RHO
    assert (np.abs(ef["p"].astype("f8") - ref["p"]) <= np.maximum(1e-4 * ref["p"], 160.0)).all()


def test_pseudomass_parity(oracle_built, lib_built, emu, golden075, golden02):
    for g, R in ((golden075, 0.075), (golden02, 0.02)):
        prm = lib_built.default_params(R)
        b = g["boundary_init"].copy()
        emu.emu_pseudomass(C.byref(prm), P(b), len(b))
        o = oracle_built.Oracle(R=R, variant="chain")
        ob = g["boundary_init"].copy()
        o.init_boundary(ob)
        assert same_bits(b["m"], ob["m"])                                    # chain: bit-for-bit
        assert (np.abs(b["m"] - g["boundary"]["m"]) / g["boundary"]["m"]).max() < TOL_RHO   # reference


def test_cell_ids_and_constants(oracle_built, lib_built, emu, golden02):
    g = golden02
    prm = lib_built.default_params(0.02)
    o = oracle_built.Oracle(R=0.02)
    f = g["fluid_5000"]
    out = np.zeros(len(f), np.int32)
    emu.emu_cell_ids(C.byref(prm), P(f), len(f), P(out))
    gf = o.grid(len(f))
    assert np.array_equal(out, o.cell_ids(gf, f))
    k = np.zeros(8, np.float32)
    emu.emu_consts(C.byref(prm), P(k))
    d2max, nf, support = k[0], k[1], k[7]
    # d2max is the largest float whose sqrt is still < 2H (:144)
    assert np.sqrt(np.float32(d2max)) < support <= np.sqrt(np.nextafter(np.float32(d2max), np.float32(np.inf)))
    assert nf == np.float32(o.lib.oracle_W(C.byref(o.prm), 0, 0, 0, 0))       # W(0) == nf, :274


def test_kick_drift_bit_exact_while_pressure_is_zero(emu, lib_built, oracle_built, golden075):
    """100 steps of config 1 from t = 0 (free fall, p == 0): every field bit-for-bit with the chain oracle
    (the whole step is the reference's arithmetic now); against the reference-built golden (libm powf,
    last-bit differences in powf(x, 3|4) feed the velocities) positions to 1e-6, velocities to 5e-6."""
    g = golden075
    prm = lib_built.default_params(0.075)
    f, du, dv = g["fluid_0"].copy(), g["du_0"].copy(), g["dv_0"].copy()
    b = g["boundary"]
    emu.emu_step(C.byref(prm), P(f), len(f), P(b), len(b), C.c_float(G[0]), C.c_float(G[1]), P(du), P(dv), 100)
    r = g["fluid_100"]
    assert np.abs(f["x"] - r["x"]).max() < 1e-6 and np.abs(f["y"] - r["y"]).max() < 1e-6
    assert max(np.abs(f["u"] - r["u"]).max(), np.abs(f["v"] - r["v"]).max()) < 5e-6
    assert same_bits(f["rho"], r["rho"]) or (np.abs(f["rho"] - r["rho"]) / r["rho"]).max() < TOL_RHO

    o = oracle_built.Oracle(R=0.075, variant="chain")
    of, ob = g["fluid_init"].copy(), g["boundary_init"].copy()
    gb = o.init_boundary(ob)
    gf = o.grid(len(of))
    odu, odv = o.compute_accel(of, ob, gf, gb, *G)
    ef, edu, edv = of.copy(), odu.copy(), odv.copy()
    o.step(of, ob, gf, gb, odu, odv, 100, *G)
    emu.emu_step(C.byref(prm), P(ef), len(ef), P(ob), len(ob), C.c_float(G[0]), C.c_float(G[1]), P(edu), P(edv), 100)
    for key in ("x", "y", "u", "v", "rho", "p"):
        assert same_bits(ef[key], of[key]), key
    assert same_bits(edu, odu) and same_bits(edv, odv)


def test_coincident_particles_give_nan_like_the_reference(oracle_built, lib_built, emu):
    """grad W is 0/0 for two distinct particles at the same point."""
    prm = lib_built.default_params(0.075)
    o = oracle_built.Oracle()
    f = o.scene_drop()[:40].copy()
    f[7]["x"], f[7]["y"] = f[3]["x"], f[3]["y"]
    b = o.scene_boundary(); o.init_boundary(b)
    ef, edu, edv, _, _ = emu_accel(emu, prm, f, b)
    _, odu, odv = oracle_accel(oracle_built, 0.075, "chain", f, b)
    assert np.array_equal(np.isnan(edu), np.isnan(odu)) and np.isnan(edu[[3, 7]]).all()
