"""The device instruction sequences of the force pass (csrc/sph_math.cuh force_pair_strict, kernels_pair.cu
force_pair_strict_packed) one pair at a time, against the HOST evaluation of the same header (tests/emu),
which test_device_math.py pins bit-for-bit to the chain oracle.  Bit-exact, NaNs matching NaNs.

What is exercised: r = sqrtf(d2) from rsqrt.approx + one correction; r/H, W/W_ref and (x/r)/H by Markstein's
correction with verified constants; x/r and y/r from one shared reciprocal; 0.1*pow4 and H*xu/(xx + 0.01 H^2)
through double; -0.01*C*mu/rho as one float division; packed (f32x2) evaluation of the two components."""
import ctypes as C

import numpy as np
import pytest

from conftest import same_bits_nan

pytestmark = pytest.mark.gpu


def make_pairs(rng, H, n, rest=False):
    p = np.zeros((n, 12), np.float32)
    xi = rng.uniform(0.05, 3.95, n); yi = rng.uniform(0.05, 1.95, n)
    r = 2 * H * np.sqrt(rng.uniform(1e-6, 1.0, n)); th = rng.uniform(0, 2 * np.pi, n)
    p[:, 0], p[:, 1] = xi, yi
    p[:, 2], p[:, 3] = xi + r * np.cos(th), yi + r * np.sin(th)
    if not rest:
        p[:, 4:8] = rng.normal(0, 1.0, (n, 4))
    p[:, 8] = rng.uniform(900, 1100, n); p[:, 10] = rng.uniform(900, 1100, n)
    p[:, 9] = rng.uniform(0, 2e5, n) / p[:, 8] ** 2; p[:, 11] = rng.uniform(0, 2e5, n) / p[:, 10] ** 2
    return p


@pytest.mark.parametrize("R", [0.075, 0.02, 0.002423, 0.0005, 0.00017677669])
def test_force_pair_sequences_bit_exact(lib_built, emu, R):
    pkg = lib_built
    prm = pkg.default_params(R)
    rng = np.random.default_rng(17)
    n = 1 << 20
    pairs = np.concatenate([make_pairs(rng, float(prm.H), n), make_pairs(rng, float(prm.H), n // 8, rest=True)])
    # edge cases: coincident particles (0/0), a pair on the x axis, on the y axis, at the support radius
    pairs[0, 2:4] = pairs[0, 0:2]
    pairs[1, 3] = pairs[1, 1]
    pairs[2, 2] = pairs[2, 0]
    pairs[3, 2], pairs[3, 3] = pairs[3, 0] + np.float32(2 * prm.H), pairs[3, 1]
    with pkg.Simulation(prm) as sim:
        for boundary in (0, 1):
            ref = np.zeros((len(pairs), 2), np.float32)
            emu.emu_force_pair(C.byref(prm), len(pairs), pairs.ctypes.data_as(C.c_void_p), boundary,
                               ref.ctypes.data_as(C.c_void_p))
            assert np.isnan(ref[0]).all() and np.isfinite(ref[1:]).all()
            got, shortcuts = sim.probe_force_pair(pairs, 1 + 4 * boundary)
            assert same_bits_nan(got, ref), ("general", boundary, int((got.view("u4") != ref.view("u4")).sum()))
            if shortcuts and not boundary:
                # two neighbours at once: rows 2t and 2t+1 share particle i
                two = pairs.copy()
                for cols in (slice(0, 2), slice(4, 6), slice(8, 10)):
                    two[1::2, cols] = two[0::2, cols]
                # keep j inside the support of the shared i
                two[1::2, 2:4] = two[1::2, 0:2] + (pairs[1::2, 2:4] - pairs[1::2, 0:2])
                ref2 = np.zeros((len(two), 2), np.float32)
                emu.emu_force_pair(C.byref(prm), len(two), two.ctypes.data_as(C.c_void_p), 0, ref2.ctypes.data_as(C.c_void_p))
                got, _ = sim.probe_force_pair(two, 3)
                bad = (got.view("u4") != ref2.view("u4")) & ~np.isnan(ref2)
                assert same_bits_nan(got, ref2), ("two at once", int(bad.sum()), two[bad.any(axis=1)][:3], got[bad.any(axis=1)][:3], ref2[bad.any(axis=1)][:3])
            if shortcuts:
                for variant in (0, 2):
                    got, _ = sim.probe_force_pair(pairs, variant + 4 * boundary)
                    bad = (got.view("u4") != ref.view("u4")) & ~np.isnan(ref)
                    assert same_bits_nan(got, ref), (variant, boundary, int(bad.sum()), pairs[bad.any(axis=1)][:3], got[bad.any(axis=1)][:3], ref[bad.any(axis=1)][:3])
