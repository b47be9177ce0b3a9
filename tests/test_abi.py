"""The C ABI surface and the host-side logic of libsphb200.so.  CPU only: no compute calls —
without a GPU every compute entry point must fail loudly, never fall back."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, same_bits

FIELDS = ("x", "y", "u", "v", "m", "rho", "p")


def _declared_functions():
    names = []
    for h in ("sph_b200.h", "sph_b200_scene.h"):
        text = (ROOT / "include" / h).read_text()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
        text = re.sub(r"\btypedef\s+\w+\s*\(\s*\*[^;]*;", "", text)      # function-pointer typedefs declare no symbol
        names += re.findall(r"\b([a-z_][a-z0-9_]*)\s*\([^;{}]*\)\s*;", text)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(lib_built):
    L = lib_built.lib()
    declared = _declared_functions()
    assert len(declared) >= 40
    for must in ("sphb_step", "sphb_upload", "calculate_density", "calculate_accelerations", "draw_metaballs",
                 "update_neighbors_context", "alloc_neighbors_context", "calculate_boundary_pseudomass",
                 "calculate_particle_pressure", "sphb_scene_fill_drop"):
        assert must in declared
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, missing


def test_only_the_c_abi_is_exported(lib_built):
    out = subprocess.run(["nm", "-D", "--defined-only", str(lib_built.lib_path())], capture_output=True, text=True).stdout
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert syms and all(not s.startswith("_Z") for s in syms), [s for s in syms if s.startswith("_Z")][:5]


def test_sm100a_only_and_no_ptx_fallback(lib_built):
    out = subprocess.run(["cuobjdump", "--list-elf", "--list-ptx", str(lib_built.lib_path())], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out), out
    assert "ptx" not in out.lower().replace("--list-ptx", "")


def test_params_match_reference_defines(lib_built, oracle_built):
    for R in (0.075, 0.02, 0.002423):
        prm = lib_built.default_params(R)
        o = oracle_built.Oracle(R=R)
        for mine, ref in ((prm.H, o.prm.H), (prm.dt, o.prm.dt), (prm.vol, o.prm.vol),
                          (prm.cell_length, float(o.cell)), (prm.rho0, o.prm.rho0), (prm.c0, o.prm.c0)):
            assert np.float32(mine).tobytes() == np.float32(ref).tobytes()
    assert lib_built.api.Params.__name__ and C.sizeof(lib_built.Params) == 4 * 14 + 4 * 2 + 4 * 6


def test_scene_builders_match_reference_lattice(lib_built, oracle_built, golden075):
    prm = lib_built.default_params(0.075)
    fluid, boundary = lib_built.scene_drop(prm), lib_built.scene_boundary(prm)
    assert all(same_bits(fluid[f], golden075["fluid_init"][f]) for f in FIELDS)
    assert all(same_bits(boundary[f], golden075["boundary_init"][f]) for f in FIELDS)
    for R in (0.02, 0.005):
        prm = lib_built.default_params(R)
        o = oracle_built.Oracle(R=R)
        assert all(same_bits(lib_built.scene_drop(prm)[f], o.scene_drop()[f]) for f in FIELDS)
        assert all(same_bits(lib_built.scene_boundary(prm)[f], o.scene_boundary()[f]) for f in FIELDS)
        a, b = lib_built.scene_block(prm, R, 2.0, R, 0.5), o.scene_block(R, 2.0, R, 0.5)
        assert len(a) == len(b) > 0 and all(same_bits(a[f], b[f]) for f in FIELDS)
    # the cfg2 scene of BASELINE.json: 262,204 fluid + 4,954 boundary particles
    prm = lib_built.default_params(0.002423)
    assert lib_built.lib().sphb_scene_count_drop(C.byref(prm)) == 262204
    assert lib_built.lib().sphb_scene_count_boundary(C.byref(prm)) == 4954


def test_gravity_trace(lib_built, oracle_built):
    prm = lib_built.default_params()
    o = oracle_built.Oracle()
    for ax, ay in ((16384, 0), (0, 16384), (15396, -5604), (-120, 77)):
        assert lib_built.gravity_from_raw(prm, ax, ay) == o.gravity_from_raw(ax, ay)
    tr = lib_built.gravity_trace_tilt(prm, 20.0, 400, 50, 400)
    assert tr.shape == (400, 2) and tr[0, 0] == 0.0 and tr[0, 1] == np.float32(-9.81)
    assert (tr[:50] == tr[0]).all() and not (tr[50] == tr[0]).all()          # held per sample
    assert np.abs(np.hypot(tr[:, 0], tr[:, 1]) - 9.81).max() < 2e-3          # |g| stays G
    assert abs(np.degrees(np.arctan2(tr[:, 0], -tr[:, 1])).max() - 20.0) < 0.5


def test_no_cpu_fallback(lib_built):
    """Without a usable sm_100 GPU: create fails with SPHB_E_CUDA and a message; the C host
    driver exits non-zero.  (On the GPU box this test is a no-op.)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib_built.SphbError, match="no CPU fallback|no CUDA device|sm_"):
        lib_built.Simulation()
    r = subprocess.run([str(ROOT / "pi_sph_fluid_b200" / "host" / "sph_b200_main"), "--steps", "1"], capture_output=True, text=True)
    assert r.returncode != 0 and "failed" in r.stderr


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package or include/ may
    reference it."""
    for path in list((ROOT / "pi_sph_fluid_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if path.suffix in (".py", ".c", ".h", ".cu", ".cuh"):
            text = path.read_text()
            assert "pyoracle" not in text and "sph_oracle" not in text and "oracle/" not in text, path
