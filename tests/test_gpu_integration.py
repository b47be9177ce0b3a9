"""The reference's OWN main() (pi_sph_fluid.c:475-704) linked against libsphb200.so — the drop-in claim of
INTEGRATION.md 1, executed: oracle/Makefile `refmain` streams the reference source into gcc with its operator
functions cut out (the library's header declares them), REALTIME off (:10) and the loops bounded by
tests/ref_main/harness.c.  Scene construction (:484-540), the kick / drift loops (:615-624, :637-640), the
statistics and both pthreads are the reference's code; density, pressure, accelerations, grid updates, boundary
pseudo-mass and draw_metaballs run on the GPU through the compat tier.  After 300 iterations its state must equal
the chain oracle's bit for bit."""
import os
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import G, same_bits

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_reference_main_runs_on_the_library(oracle_built, lib_built, tmp_path):
    exe = ROOT / "oracle" / "_ref" / "pi_sph_fluid_main_b200"
    if Path("/root/reference/pi_sph_fluid.c").exists():
        exe = oracle_built.build_refmain()
    assert exe.exists(), "oracle/_ref/pi_sph_fluid_main_b200 is built by __graft_entry__.build() where /root/reference exists"
    steps = 300
    out = tmp_path / "state.bin"
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, SPHB_REFMAIN_STEPS=str(steps), SPHB_REFMAIN_OUT=str(out)))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    # the reference's own start-up lines (:543-545)
    assert "dt = 0.000244    (expected ticks/s) 4102" in r.stdout
    assert "n_fluid = 269" in r.stdout and "n_boundary = 162" in r.stdout
    assert f"refmain: {steps} iterations" in r.stdout
    raw = out.read_bytes()
    n = struct.unpack("i", raw[:4])[0]
    assert n == 269
    f = np.frombuffer(raw, dtype=oracle_built.PARTICLE, count=n, offset=4)
    du = np.frombuffer(raw, dtype=np.float32, count=n, offset=4 + 28 * n)
    dv = np.frombuffer(raw, dtype=np.float32, count=n, offset=4 + 32 * n)

    o = oracle_built.Oracle(R=0.075, variant="chain")
    of, ob = o.scene_drop(), o.scene_boundary()
    gb = o.init_boundary(ob)
    gf = o.grid(len(of))
    odu, odv = o.compute_accel(of, ob, gf, gb, *G)
    o.step(of, ob, gf, gb, odu, odv, steps, *G)
    for fld in ("x", "y", "u", "v", "m", "rho", "p"):
        assert same_bits(f[fld], of[fld]), fld
    assert same_bits(du, odu) and same_bits(dv, odv)
