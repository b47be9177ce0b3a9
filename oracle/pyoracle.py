"""ctypes bindings for the CPU checkers under oracle/.  TEST INFRASTRUCTURE ONLY.

Two checkers live here:

* ``Oracle``    — our C restatement (oracle/sph_oracle.c) with runtime parameters.
* ``Reference`` — the reference's own translation unit compiled into ``oracle/_ref``
  (see oracle/Makefile) and driven through ``oracle/ref_driver.c``.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"

#: pi_sph_fluid.c:26-31 — 7 x f32, 28 bytes, no padding
PARTICLE = np.dtype(
    [("x", "f4"), ("y", "f4"), ("u", "f4"), ("v", "f4"), ("m", "f4"), ("rho", "f4"), ("p", "f4")]
)
assert PARTICLE.itemsize == 28


class OracleParams(C.Structure):
    _fields_ = [
        ("R", C.c_float), ("H", C.c_float), ("width", C.c_float), ("height", C.c_float),
        ("rho0", C.c_float), ("c0", C.c_float), ("g", C.c_float), ("dt", C.c_float),
        ("vol", C.c_float), ("mass", C.c_float), ("max_neighbors", C.c_int),
    ]


class OracleCounters(C.Structure):
    _fields_ = [("neighbor_overflows", C.c_longlong), ("max_neighbors_seen", C.c_int)]


class OracleGrid(C.Structure):
    _fields_ = [
        ("x_min", C.c_float), ("x_max", C.c_float), ("y_min", C.c_float), ("y_max", C.c_float),
        ("cell_length", C.c_float), ("n_cells", C.c_int), ("m_cells", C.c_int),
        ("n_particles", C.c_int),
        ("cells_head", C.POINTER(C.c_uint32)), ("cells_tail", C.POINTER(C.c_uint32)),
        ("particles_next", C.POINTER(C.c_uint32)), ("n_clamped", C.c_longlong),
    ]


def build(ref: bool = True) -> None:
    """Compile the checkers (``make -C oracle``).  The reference variants are only
    (re)built where /root/reference exists; elsewhere the prebuilt files are used."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", str(HERE), "-j8", *targets], check=True,
                   stdout=subprocess.DEVNULL)


def build_refmain() -> Path:
    """oracle/_ref/pi_sph_fluid_main_b200: the reference's own main() linked against libsphb200.so (oracle/Makefile
    `refmain`; needs the library built and, to (re)build, /root/reference).  Returns the binary's path."""
    subprocess.run(["make", "-C", str(HERE), "refmain"], check=True, stdout=subprocess.DEVNULL)
    return HERE / "_ref" / "pi_sph_fluid_main_b200"


def load_state(paths):
    """Reader of libsphb200's state files for the oracle side: one version-1 file (sphb_save_state) or the
    version-2 parts of a slab run (sphb_mg_save_state, any number of ranks).  -> dict(R, H, steps, fluid, du, dv,
    boundary) with the fluid in ORIGINAL particle order, ready for Oracle.step."""
    if isinstance(paths, (str, Path)):
        paths = [paths]
    parts, out = [], {}
    for p in paths:
        raw = Path(p).read_bytes()
        assert raw[:8] == b"SPHB200\0", f"{p}: not a libsphb200 state file"
        version, params_bytes, n, nb = np.frombuffer(raw, "<u4", 4, 8)
        steps = int(np.frombuffer(raw, "<u8", 1, 24)[0])
        prm = np.frombuffer(raw, "<f4", 14, 64)             # R, H, width, height, rho0, c0, g, dt, vol, x/y min/max, cell
        off = 64 + int(params_bytes)
        ids = np.arange(n, dtype=np.uint32)
        if version == 2:
            ids = np.frombuffer(raw, "<u4", n, off); off += 4 * n
        else:
            assert version == 1, f"{p}: unknown version {version}"
        fluid = np.frombuffer(raw, PARTICLE, n, off); off += 28 * n
        du = np.frombuffer(raw, "<f4", n, off); off += 4 * n
        dv = np.frombuffer(raw, "<f4", n, off); off += 4 * n
        if not out:
            out = dict(R=float(prm[0]), H=float(prm[1]), steps=steps, boundary=np.frombuffer(raw, PARTICLE, nb, off).copy())
        assert steps == out["steps"], "parts from different steps"
        parts.append((ids, fluid, du, dv))
    total = sum(len(i) for i, _, _, _ in parts)
    out["fluid"] = np.zeros(total, PARTICLE); out["du"] = np.zeros(total, np.float32); out["dv"] = np.zeros(total, np.float32)
    seen = np.zeros(total, bool)
    for ids, f, du, dv in parts:
        assert not seen[ids].any(), "a particle appears in two parts"
        out["fluid"][ids] = f; out["du"][ids] = du; out["dv"][ids] = dv; seen[ids] = True
    assert seen.all()
    return out


def cpu_level() -> str:
    """'v4' if the host CPU can run the AVX-512 builds, else 'v3'."""
    try:
        flags = Path("/proc/cpuinfo").read_text()
    except OSError:
        return "v3"
    need = ("avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl")
    return "v4" if all(f in flags for f in need) else "v3"


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _particles(a: np.ndarray) -> np.ndarray:
    assert a.dtype == PARTICLE and a.flags["C_CONTIGUOUS"], "need a contiguous PARTICLE array"
    return a


# --------------------------------------------------------------------------- Oracle


def _load_oracle(variant: str) -> C.CDLL:
    name = {"strict": "liboracle_strict.so", "chain": "liboracle_chain_strict.so",
            "fast": f"liboracle_fast_{cpu_level()}.so"}[variant]
    path = HERE / name
    if not path.exists():
        build(ref=False)
    lib = C.CDLL(str(path))
    lib.oracle_W.restype = C.c_float
    lib.oracle_W.argtypes = [C.c_void_p] + [C.c_float] * 4
    lib.oracle_euclid_dist.restype = C.c_float
    lib.oracle_euclid_dist.argtypes = [C.c_float] * 4
    lib.oracle_grid_alloc.restype = C.POINTER(OracleGrid)
    lib.oracle_grid_alloc.argtypes = [C.c_int] + [C.c_float] * 5
    lib.oracle_grid_fnv.restype = C.c_uint64
    lib.oracle_grid_fnv.argtypes = [C.c_void_p]
    lib.oracle_neighbor_list.restype = C.c_int
    lib.oracle_scene_count_block.argtypes = [C.c_void_p] + [C.c_float] * 4
    lib.oracle_scene_fill_block.argtypes = [C.c_void_p, C.c_void_p] + [C.c_float] * 4
    lib.oracle_compute_accel.argtypes = [
        C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
        C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_step.argtypes = [
        C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
        C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_accelerations.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_float, C.c_float, C.c_void_p]
    lib.oracle_make_params.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
    lib.oracle_grid_free.argtypes = [C.c_void_p]
    return lib


class Oracle:
    """Our restatement, one instance per scene geometry (R, width, height)."""

    def __init__(self, R: float = 0.075, width: float = 4.0, height: float = 2.0,
                 variant: str = "strict", max_neighbors: int = 48, threads: int | None = None):
        self.lib = _load_oracle(variant)
        self.variant = variant
        self.prm = OracleParams()
        self.lib.oracle_make_params(C.byref(self.prm), R, width, height)
        self.prm.max_neighbors = max_neighbors
        self.ctr = OracleCounters()
        if threads:
            self.lib.oracle_set_num_threads(threads)
        self.threads = self.lib.oracle_num_threads()
        self._grids = []

    # -- parameters as numpy float32 scalars
    @property
    def H(self): return np.float32(self.prm.H)
    @property
    def dt(self): return np.float32(self.prm.dt)
    @property
    def cell(self): return np.float32(2) * np.float32(self.prm.H)   # :596  2*H

    def __del__(self):
        for g in getattr(self, "_grids", []):
            self.lib.oracle_grid_free(g)

    # -- scenes
    def scene_drop(self) -> np.ndarray:
        n = self.lib.oracle_scene_count_drop(C.byref(self.prm))
        a = np.zeros(n, PARTICLE)
        self.lib.oracle_scene_fill_drop(C.byref(self.prm), _ptr(a))
        return a

    def scene_block(self, x0, x1, y0, y1) -> np.ndarray:
        n = self.lib.oracle_scene_count_block(C.byref(self.prm), x0, x1, y0, y1)
        a = np.zeros(n, PARTICLE)
        self.lib.oracle_scene_fill_block(C.byref(self.prm), _ptr(a), x0, x1, y0, y1)
        return a

    def scene_boundary(self) -> np.ndarray:
        n = self.lib.oracle_scene_count_boundary(C.byref(self.prm))
        a = np.zeros(n, PARTICLE)
        self.lib.oracle_scene_fill_boundary(C.byref(self.prm), _ptr(a))
        return a

    # -- grid
    def grid(self, n: int):
        g = self.lib.oracle_grid_alloc(n, 0.0, self.prm.width, 0.0, self.prm.height, float(self.cell))
        self._grids.append(g)
        return g

    def grid_update(self, g, particles): self.lib.oracle_grid_update(g, _ptr(_particles(particles)))

    def cell_ids(self, g, particles) -> np.ndarray:
        out = np.zeros(len(particles), np.int32)
        self.lib.oracle_cell_ids(g, _ptr(_particles(particles)), len(particles), _ptr(out))
        return out

    def neighbor_list(self, a, b, i, gb, same: bool, cap: int = 512) -> np.ndarray:
        out = np.zeros(cap, np.int32)
        n = self.lib.oracle_neighbor_list(C.byref(self.prm), _ptr(_particles(a)), _ptr(_particles(b)),
                                          int(same), int(i), gb, _ptr(out), cap)
        return out[:n].copy()

    def grid_fnv(self, g) -> int: return int(self.lib.oracle_grid_fnv(g))

    # -- operators
    def init_boundary(self, boundary):
        gb = self.grid(len(boundary))
        self.grid_update(gb, boundary)
        self.lib.oracle_boundary_pseudomass(C.byref(self.prm), _ptr(_particles(boundary)), gb, C.byref(self.ctr))
        return gb

    def density(self, fluid, boundary, gf, gb):
        self.lib.oracle_density(C.byref(self.prm), _ptr(_particles(fluid)), _ptr(_particles(boundary)), gf, gb, C.byref(self.ctr))

    def pressure(self, particles):
        self.lib.oracle_pressure(C.byref(self.prm), _ptr(_particles(particles)), len(particles))

    def accelerations(self, fluid, boundary, gf, gb, gx, gy):
        du = np.zeros(len(fluid), np.float32); dv = np.zeros(len(fluid), np.float32)
        self.lib.oracle_accelerations(C.byref(self.prm), _ptr(du), _ptr(dv), _ptr(_particles(fluid)),
                                      _ptr(_particles(boundary)), gf, gb, gx, gy, C.byref(self.ctr))
        return du, dv

    def compute_accel(self, fluid, boundary, gf, gb, gx, gy):
        du = np.zeros(len(fluid), np.float32); dv = np.zeros(len(fluid), np.float32)
        self.lib.oracle_compute_accel(C.byref(self.prm), _ptr(_particles(fluid)), len(fluid),
                                      _ptr(_particles(boundary)), gf, gb, gx, gy, _ptr(du), _ptr(dv), C.byref(self.ctr))
        return du, dv

    def step(self, fluid, boundary, gf, gb, du, dv, nsteps, gx=0.0, gy=-9.81, gxy=None):
        g = None
        if gxy is not None:
            gxy = np.ascontiguousarray(gxy, np.float32); assert gxy.shape == (nsteps, 2)
            g = _ptr(gxy)
        self.lib.oracle_step(C.byref(self.prm), _ptr(_particles(fluid)), len(fluid), _ptr(_particles(boundary)),
                             gf, gb, gx, gy, g, nsteps, _ptr(du), _ptr(dv), C.byref(self.ctr))

    def pixels(self) -> np.ndarray:
        a = np.zeros(64 * 128, PARTICLE)
        self.lib.oracle_pixel_pseudoparticles(C.byref(self.prm), _ptr(a))
        return a

    def draw_metaballs(self, buf, pixels, fluid, gf):
        assert buf.dtype == np.uint8 and buf.size == 1024
        self.lib.oracle_draw_metaballs(C.byref(self.prm), _ptr(buf), _ptr(_particles(pixels)),
                                       _ptr(_particles(fluid)), gf, C.byref(self.ctr))

    def gravity_from_raw(self, ax_raw: int, ay_raw: int):
        gx = C.c_float(); gy = C.c_float()
        self.lib.oracle_gravity_from_raw(C.byref(self.prm), int(ax_raw), int(ay_raw), C.byref(gx), C.byref(gy))
        return np.float32(gx.value), np.float32(gy.value)


# --------------------------------------------------------------------------- Reference


def reference_so(R: str | None = None, flavour: str = "strict") -> Path:
    """Path of a reference build.  R=None: the unmodified file (R=0.075, ushort links);
    R='0.002423' etc.: the sed-widened variants (oracle/Makefile REF_RS)."""
    fl = "strict" if flavour == "strict" else f"fast_{cpu_level()}"
    tag = "" if R is None else f"_R{R}"
    return REF_DIR / f"libpisph_ref{tag}_{fl}.so"


def reference_available(R: str | None = None, flavour: str = "strict") -> bool:
    return reference_so(R, flavour).exists() and (REF_DIR / "libref_driver.so").exists()


class Reference:
    """The reference's own compiled code, driven like its main() (ref_driver.c)."""

    def __init__(self, R: str | None = None, flavour: str = "strict"):
        so = reference_so(R, flavour)
        if not so.exists() and Path("/root/reference/pi_sph_fluid.c").exists():
            build(ref=True)
        if not so.exists():
            raise FileNotFoundError(f"{so} (build with `make -C oracle ref` where /root/reference exists)")
        self.so = so
        d = C.CDLL(str(REF_DIR / "libref_driver.so"))
        d.refdrv_open.restype = C.c_void_p
        d.refdrv_open.argtypes = [C.c_char_p]
        d.refdrv_alloc_ctx.restype = C.c_void_p
        d.refdrv_alloc_ctx.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 5
        d.refdrv_update_ctx.argtypes = [C.c_void_p] * 3
        d.refdrv_find_neighbors.restype = C.c_int
        d.refdrv_find_neighbors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        d.refdrv_init_boundary.argtypes = [C.c_void_p] * 3
        d.refdrv_compute_accel.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        d.refdrv_draw_metaballs.argtypes = [C.c_void_p] * 5
        d.refdrv_step.restype = C.c_double
        d.refdrv_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p]
        self.d = d
        self.h = d.refdrv_open(str(so).encode())
        if not self.h:
            raise OSError(f"cannot open {so}")
        self.max_threads = d.refdrv_max_threads()
        Rf = np.float32(0.075 if R is None else float(R))
        self.R = Rf
        self.H = np.float32(Rf * np.float32(1.3))                 # :12
        self.cell = np.float32(2) * self.H                         # :596
        self.dt = np.float32(np.float32(1.0) * self.H / np.float32(400.0))   # :19

    def ctx(self, n, width=4.0, height=2.0):
        return self.d.refdrv_alloc_ctx(self.h, n, 0.0, width, 0.0, height, float(self.cell))

    def update_ctx(self, ctx, particles): self.d.refdrv_update_ctx(self.h, ctx, _ptr(_particles(particles)))

    def find_neighbors(self, a, b, i, ctx_b) -> np.ndarray:
        out = np.zeros(48, np.int32)     # :21 — the reference's own cap; callers keep to sane states
        n = self.d.refdrv_find_neighbors(self.h, _ptr(out), _ptr(_particles(a)), _ptr(_particles(b)), int(i), ctx_b)
        return out[:n].copy()

    def init_boundary(self, boundary, width=4.0, height=2.0):
        cb = self.ctx(len(boundary), width, height)
        self.d.refdrv_init_boundary(self.h, _ptr(_particles(boundary)), cb)
        return cb

    def compute_accel(self, fluid, boundary, cf, cb, gx, gy):
        du = np.zeros(len(fluid), np.float32); dv = np.zeros(len(fluid), np.float32)
        self.d.refdrv_compute_accel(self.h, _ptr(_particles(fluid)), len(fluid), _ptr(_particles(boundary)),
                                    cf, cb, gx, gy, _ptr(du), _ptr(dv))
        return du, dv

    def step(self, fluid, boundary, cf, cb, du, dv, nsteps, gx=0.0, gy=-9.81, gxy=None, threads=4) -> float:
        g = None
        if gxy is not None:
            gxy = np.ascontiguousarray(gxy, np.float32); assert gxy.shape == (nsteps, 2)
            g = _ptr(gxy)
        return self.d.refdrv_step(self.h, _ptr(_particles(fluid)), len(fluid), _ptr(_particles(boundary)), cf, cb,
                                  float(self.dt), gx, gy, g, nsteps, threads, _ptr(du), _ptr(dv))

    def draw_metaballs(self, buf, pixels, fluid, cf):
        self.d.refdrv_draw_metaballs(self.h, _ptr(buf), _ptr(_particles(pixels)), _ptr(_particles(fluid)), cf)
