"""ctypes binding of include/sph_b200.h and include/sph_b200_scene.h.

``Simulation`` wraps the resident tier (what the reference's main loop would call);
``compat`` exposes the seven reference-named operators (pi_sph_fluid.c:82-411) with the
reference's argument meaning, operating in place on numpy arrays of ``PARTICLE`` records.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent

#: `struct particle`, pi_sph_fluid.c:26-31 — 7 x f32 = 28 bytes
PARTICLE = np.dtype(
    [("x", "f4"), ("y", "f4"), ("u", "f4"), ("v", "f4"), ("m", "f4"), ("rho", "f4"), ("p", "f4")]
)
assert PARTICLE.itemsize == 28

KERNEL_NAMES = ["advect_bin", "scan", "reorder", "density", "force", "other"]


class SphbError(RuntimeError):
    pass


class Params(C.Structure):
    """sphb_params (include/sph_b200.h)."""
    _fields_ = [
        ("R", C.c_float), ("H", C.c_float), ("width", C.c_float), ("height", C.c_float),
        ("rho0", C.c_float), ("c0", C.c_float), ("g", C.c_float), ("dt", C.c_float),
        ("vol", C.c_float),
        ("x_min", C.c_float), ("x_max", C.c_float), ("y_min", C.c_float), ("y_max", C.c_float),
        ("cell_length", C.c_float), ("deterministic", C.c_int), ("device", C.c_int),
        ("fast_force", C.c_int), ("reserved", C.c_int * 5),
    ]


class Stats(C.Structure):
    """sphb_stats (include/sph_b200.h)."""
    _fields_ = [
        ("mass", C.c_double), ("mom_x", C.c_double), ("mom_y", C.c_double), ("kinetic", C.c_double),
        ("max_speed", C.c_float), ("max_rho_err", C.c_float), ("last_rho_err_ref", C.c_float),
        ("min_rho", C.c_float), ("max_rho", C.c_float),
        ("n_escaped", C.c_uint), ("max_cell_count", C.c_uint), ("n_fluid", C.c_uint),
        ("n_boundary", C.c_uint), ("steps", C.c_ulonglong), ("n_lost", C.c_uint), ("n_overflow", C.c_uint),
    ]

    def asdict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


_LIB = None


def lib_path() -> Path:
    import os
    tag = os.environ.get("SPHB_LIB_VARIANT")        # kernel tuning experiments only (scripts/tune_pair.py)
    return PKG / (f"libsphb200_{tag}.so" if tag else "libsphb200.so")


def lib() -> C.CDLL:
    """Load libsphb200.so (built in-tree by ``python -m pi_sph_fluid_b200.build``)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not path.exists():
        raise SphbError(f"{path} is missing: build it with `python -m pi_sph_fluid_b200.build` "
                        "(there is no fallback implementation)")
    L = C.CDLL(str(path))
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    sig = {
        "sphb_default_params": (ci, [vp, cf, cf, cf]),
        "sphb_create": (ci, [vp, C.POINTER(vp)]),
        "sphb_destroy": (ci, [vp]),
        "sphb_upload": (ci, [vp, vp, ci, vp, ci]),
        "sphb_init_boundary": (ci, [vp]),
        "sphb_compute_accel": (ci, [vp, cf, cf]),
        "sphb_upload_accel": (ci, [vp, vp, vp]),
        "sphb_step": (ci, [vp, cf, cf, ci]),
        "sphb_step_trace": (ci, [vp, vp, ci]),
        "sphb_download": (ci, [vp, vp, vp, vp]),
        "sphb_download_boundary": (ci, [vp, vp]),
        "sphb_render": (ci, [vp, vp]),
        "sphb_render_splat": (ci, [vp, vp]),
        "sphb_render_counts": (ci, [vp, vp]),
        "sphb_splat_frame": (ci, [vp, vp, vp]),
        "sphb_get_stats": (ci, [vp, vp]),
        "sphb_step_stats": (ci, [vp, vp, ci, vp]),
        "sphb_step_stats_begin": (ci, [vp, vp, ci, vp]),
        "sphb_step_stats_end": (ci, [vp, C.c_ulonglong, vp]),
        "sphb_synchronize": (ci, [vp]),
        "sphb_save_state": (ci, [vp, C.c_char_p]),
        "sphb_load_state": (ci, [C.c_char_p, ci, C.POINTER(vp)]),
        "sphb_gravity_from_raw": (ci, [vp, ci, ci, vp, vp]),
        "sphb_cell_ids": (ci, [vp, vp]),
        "sphb_grid_shape": (ci, [vp, vp, vp]),
        "sphb_neighbor_lists": (ci, [vp, ci, ci, vp, vp]),
        "sphb_profile": (ci, [vp, ci]),
        "sphb_profile_read": (ci, [vp, vp, vp, ci]),
        "sphb_pair_stats": (ci, [vp, vp, vp]),
        "sphb_probe_force_pair": (ci, [vp, ci, vp, ci, vp, vp]),
        "sphb_handover_lists": (ci, [vp, ci, vp, vp, vp]),
        "sphb_flush_l2": (ci, [vp]),
        "sphb_reorder_marks": (ci, [vp, vp]),
        "sphb_stream": (vp, [vp]),
        "sphb_launch_count": (C.c_ulonglong, [vp]),
        "sphb_last_error": (C.c_char_p, []),
        "sphb_build_info": (C.c_char_p, []),
        "sphb_compat_set_params": (ci, [vp]),
        "sphb_compat_free_context": (None, [vp]),
        "sphb_compat_shutdown": (None, []),
        "alloc_neighbors_context": (vp, [ci, cf, cf, cf, cf, cf]),
        "update_neighbors_context": (None, [vp, vp]),
        "calculate_boundary_pseudomass": (None, [vp, vp]),
        "calculate_density": (None, [vp, vp, vp, vp]),
        "calculate_particle_pressure": (None, [vp, ci]),
        "calculate_accelerations": (None, [vp, vp, vp, vp, vp, vp, cf, cf]),
        "draw_metaballs": (None, [vp, vp, vp, vp]),
        "sphb_scene_count_drop": (ci, [vp]),
        "sphb_scene_fill_drop": (ci, [vp, vp]),
        "sphb_scene_count_block": (ci, [vp, cf, cf, cf, cf]),
        "sphb_scene_fill_block": (ci, [vp, cf, cf, cf, cf, vp]),
        "sphb_scene_count_boundary": (ci, [vp]),
        "sphb_scene_fill_boundary": (ci, [vp, vp]),
        "sphb_gravity_trace_tilt": (ci, [vp, cf, ci, ci, ci, vp]),
        "sphb_spacing_for_count": (cf, [C.c_double, C.c_double]),
        "sphb_scene_fill_block_slab": (ci, [vp, cf, cf, cf, cf, ci, ci, vp, vp]),
        "sphb_scene_block_column_hist": (ci, [vp, cf, cf, cf, cf, vp]),
        "sphb_grid_columns": (ci, [vp, vp, vp]),
        "sphb_column_of": (ci, [vp, cf]),
        "sphb_column_histogram": (ci, [vp, vp, ci, vp]),
        "sphb_mg_plan_cuts": (ci, [vp, ci, ci, ci, vp]),
        "sphb_mg_plan_cuts_cost": (ci, [vp, ci, ci, ci, C.c_double, vp]),
        "sphb_mg_configure": (ci, [vp, ci, ci, ci, ci, ci, ci]),
        "sphb_mg_unique_id": (ci, [vp]),
        "sphb_mg_connect_nccl": (ci, [vp, vp]),
        "sphb_mg_connect_local": (ci, [vp, ci]),
        "sphb_mg_ipc_handle": (ci, [vp, vp]),
        "sphb_mg_connect_ipc": (ci, [vp, vp, vp]),
        "sphb_mg_disconnect_ipc": (ci, [vp]),
        "sphb_mg_upload": (ci, [vp, vp, vp, C.c_uint, ci, vp, ci]),
        "sphb_mg_download": (ci, [vp, ci, vp, vp, vp, vp, vp]),
        "sphb_mg_upload_accel": (ci, [vp, vp, vp]),
        "sphb_mg_group_compute_accel": (ci, [vp, ci, cf, cf]),
        "sphb_mg_group_step": (ci, [vp, ci, cf, cf, vp, ci]),
        "sphb_mg_group_synchronize": (ci, [vp, ci]),
        "sphb_mg_merge_stats": (ci, [vp, ci, vp]),
        "sphb_mg_allreduce_stats": (ci, [vp, vp]),
        "sphb_mg_info": (ci, [vp, vp]),
        "sphb_mg_save_state": (ci, [vp, C.c_char_p]),
        "sphb_mg_load_state": (ci, [vp, vp, ci]),
        "sphb_mg_rebalance": (ci, [vp, ci, C.c_double, C.c_double, vp]),
        "sphb_mg_rebalance_host": (ci, [vp, ci, C.c_double, C.c_double, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


def _check(rc: int, what: str) -> int:
    if rc < 0:
        raise SphbError(f"{what} failed ({rc}): {lib().sphb_last_error().decode()}")
    return rc


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _particles(a: np.ndarray) -> np.ndarray:
    if a.dtype != PARTICLE or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("expected a C-contiguous array of PARTICLE records")
    return a


def default_params(R: float = 0.075, width: float = 4.0, height: float = 2.0,
                   deterministic: bool = True, device: int = 0, fast_force: bool = False) -> Params:
    prm = Params()
    _check(lib().sphb_default_params(C.byref(prm), R, width, height), "sphb_default_params")
    prm.deterministic = int(deterministic)
    prm.device = device
    prm.fast_force = int(fast_force)
    return prm


# ------------------------------------------------------------------------------- scenes

def scene_drop(prm: Params) -> np.ndarray:
    """The reference's initial drop, pi_sph_fluid.c:484-506."""
    n = _check(lib().sphb_scene_count_drop(C.byref(prm)), "sphb_scene_count_drop")
    a = np.zeros(n, PARTICLE)
    _check(lib().sphb_scene_fill_drop(C.byref(prm), _p(a)), "sphb_scene_fill_drop")
    return a


def scene_block(prm: Params, x0: float, x1: float, y0: float, y1: float) -> np.ndarray:
    n = _check(lib().sphb_scene_count_block(C.byref(prm), x0, x1, y0, y1), "sphb_scene_count_block")
    a = np.zeros(n, PARTICLE)
    _check(lib().sphb_scene_fill_block(C.byref(prm), x0, x1, y0, y1, _p(a)), "sphb_scene_fill_block")
    return a


def scene_boundary(prm: Params) -> np.ndarray:
    """The reference's four walls, pi_sph_fluid.c:513-540."""
    n = _check(lib().sphb_scene_count_boundary(C.byref(prm)), "sphb_scene_count_boundary")
    a = np.zeros(n, PARTICLE)
    _check(lib().sphb_scene_fill_boundary(C.byref(prm), _p(a)), "sphb_scene_fill_boundary")
    return a


def gravity_from_raw(prm: Params, ax_raw: int, ay_raw: int):
    gx, gy = C.c_float(), C.c_float()
    _check(lib().sphb_gravity_from_raw(C.byref(prm), ax_raw, ay_raw, C.byref(gx), C.byref(gy)), "sphb_gravity_from_raw")
    return np.float32(gx.value), np.float32(gy.value)


def gravity_trace_tilt(prm: Params, amplitude_deg: float, period_steps: int, hold_steps: int, nsteps: int) -> np.ndarray:
    out = np.zeros((nsteps, 2), np.float32)
    _check(lib().sphb_gravity_trace_tilt(C.byref(prm), amplitude_deg, period_steps, hold_steps, nsteps, _p(out)),
           "sphb_gravity_trace_tilt")
    return out


def spacing_for_count(area: float, n_target: float) -> float:
    return float(lib().sphb_spacing_for_count(area, n_target))


# ------------------------------------------------------------------------------- resident tier

class Simulation:
    """Resident-tier handle: the state of the reference's main() kept in HBM."""

    def __init__(self, prm: Params | None = None, **kw):
        self.prm = prm if prm is not None else default_params(**kw)
        self._h = C.c_void_p()
        _check(lib().sphb_create(C.byref(self.prm), C.byref(self._h)), "sphb_create")
        self.n_fluid = 0
        self.n_boundary = 0

    def close(self):
        if getattr(self, "_h", None):
            lib().sphb_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # :491-547 arrays -> HBM
    def upload(self, fluid: np.ndarray, boundary: np.ndarray | None = None):
        nb = 0 if boundary is None else len(boundary)
        _check(lib().sphb_upload(self._h, _p(_particles(fluid)), len(fluid),
                                 _p(_particles(boundary)) if nb else None, nb), "sphb_upload")
        self.n_fluid, self.n_boundary = len(fluid), nb

    def init_boundary(self):                       # :600-601
        _check(lib().sphb_init_boundary(self._h), "sphb_init_boundary")

    def compute_accel(self, gx: float = 0.0, gy: float = -9.81):   # :604-607
        _check(lib().sphb_compute_accel(self._h, gx, gy), "sphb_compute_accel")

    def upload_accel(self, du: np.ndarray, dv: np.ndarray):        # :492-493 restored from a checkpoint
        du = np.ascontiguousarray(du, np.float32); dv = np.ascontiguousarray(dv, np.float32)
        assert len(du) == len(dv) == self.n_fluid
        _check(lib().sphb_upload_accel(self._h, _p(du), _p(dv)), "sphb_upload_accel")

    def step(self, nsteps: int = 1, gx: float = 0.0, gy: float = -9.81):   # :612-641
        _check(lib().sphb_step(self._h, gx, gy, nsteps), "sphb_step")

    def step_trace(self, gravity_xy: np.ndarray):
        g = np.ascontiguousarray(gravity_xy, np.float32)
        assert g.ndim == 2 and g.shape[1] == 2
        _check(lib().sphb_step_trace(self._h, _p(g), len(g)), "sphb_step_trace")

    def step_stats(self, gravity_xy: np.ndarray) -> dict:      # :612-675
        """len(gravity_xy) steps; returns the statistics of the state after the last one."""
        g = np.ascontiguousarray(gravity_xy, np.float32)
        assert g.ndim == 2 and g.shape[1] == 2 and len(g) >= 1
        st = Stats()
        _check(lib().sphb_step_stats(self._h, _p(g), len(g), C.byref(st)), "sphb_step_stats")
        return st.asdict()

    def step_stats_into(self, g_addr: int, nsteps: int, st_ref) -> None:
        """step_stats without per-call allocations, for host loops that call it once per step:
        g_addr = address of nsteps float32 (gx, gy) pairs, st_ref = ctypes.byref(Stats()) made once."""
        rc = lib().sphb_step_stats(self._h, g_addr, nsteps, st_ref)
        if rc < 0:
            _check(rc, "sphb_step_stats")

    def step_stats_begin(self, g_addr: int, nsteps: int) -> int:
        """launches nsteps steps (g_addr as in step_stats_into) and returns the ticket of their statistics"""
        t = C.c_ulonglong(0)
        rc = lib().sphb_step_stats_begin(self._h, g_addr, nsteps, C.byref(t))
        if rc < 0:
            _check(rc, "sphb_step_stats_begin")
        return t.value

    def step_stats_end(self, ticket: int, st_ref) -> None:
        rc = lib().sphb_step_stats_end(self._h, ticket, st_ref)
        if rc < 0:
            _check(rc, "sphb_step_stats_end")

    def synchronize(self):
        _check(lib().sphb_synchronize(self._h), "sphb_synchronize")

    def save_state(self, path):
        _check(lib().sphb_save_state(self._h, str(path).encode()), "sphb_save_state")

    @classmethod
    def load_state(cls, path, device: int = -1) -> "Simulation":
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        _check(lib().sphb_load_state(str(path).encode(), device, C.byref(self._h)), "sphb_load_state")
        st = Stats()
        _check(lib().sphb_get_stats(self._h, C.byref(st)), "sphb_get_stats")
        self.n_fluid, self.n_boundary, self.prm = st.n_fluid, st.n_boundary, None
        return self

    def download(self, accel: bool = True):
        fluid = np.zeros(self.n_fluid, PARTICLE)
        du = np.zeros(self.n_fluid, np.float32) if accel else None
        dv = np.zeros(self.n_fluid, np.float32) if accel else None
        _check(lib().sphb_download(self._h, _p(fluid), _p(du), _p(dv)), "sphb_download")
        return (fluid, du, dv) if accel else fluid

    def download_into(self, fluid: np.ndarray, du: np.ndarray | None, dv: np.ndarray | None):
        _check(lib().sphb_download(self._h, _p(_particles(fluid)), _p(du), _p(dv)), "sphb_download")

    def download_boundary(self) -> np.ndarray:
        b = np.zeros(self.n_boundary, PARTICLE)
        _check(lib().sphb_download_boundary(self._h, _p(b)), "sphb_download_boundary")
        return b

    def render(self) -> np.ndarray:                # :649
        buf = np.zeros(1024, np.uint8)
        _check(lib().sphb_render(self._h, _p(buf)), "sphb_render")
        return buf

    def render_splat(self) -> np.ndarray:
        """sphb_render_splat: the frame for scenes whose pixels are much wider than the kernel support."""
        buf = np.zeros(1024, np.uint8)
        _check(lib().sphb_render_splat(self._h, _p(buf)), "sphb_render_splat")
        return buf

    def render_counts(self) -> np.ndarray:
        counts = np.zeros((64, 128), np.uint32)
        _check(lib().sphb_render_counts(self._h, _p(counts)), "sphb_render_counts")
        return counts

    def stats(self) -> dict:                       # :656-675
        st = Stats()
        _check(lib().sphb_get_stats(self._h, C.byref(st)), "sphb_get_stats")
        return st.asdict()

    def grid_shape(self):
        r, c = C.c_int(), C.c_int()
        _check(lib().sphb_grid_shape(self._h, C.byref(r), C.byref(c)), "sphb_grid_shape")
        return r.value, c.value

    def cell_ids(self) -> np.ndarray:
        out = np.zeros(self.n_fluid, np.int32)
        _check(lib().sphb_cell_ids(self._h, _p(out)), "sphb_cell_ids")
        return out

    def neighbor_lists(self, which: int = 0, cap: int = 64):
        n = self.n_boundary if which == 2 else self.n_fluid
        counts = np.zeros(n, np.int32)
        lists = np.zeros((n, cap), np.int32)
        over = _check(lib().sphb_neighbor_lists(self._h, which, cap, _p(counts), _p(lists)), "sphb_neighbor_lists")
        return counts, lists, over

    def handover_lists(self, cap: int = 64):
        """sphb_handover_lists -> (counts, lists, chunks whose plan travelled in the record)."""
        counts = np.zeros(self.n_fluid, np.int32)
        lists = np.zeros((self.n_fluid, cap), np.int32)
        whole = C.c_uint(0)
        _check(lib().sphb_handover_lists(self._h, cap, _p(counts), _p(lists), C.byref(whole)), "sphb_handover_lists")
        return counts, lists, int(whole.value)

    def probe_force_pair(self, pairs: np.ndarray, variant: int = 0):
        """sphb_probe_force_pair: (n, 12) float32 pairs -> ((n, 2) float32 pair terms, shortcuts verified?)."""
        pairs = np.ascontiguousarray(pairs, np.float32).reshape(-1, 12)
        out = np.zeros((len(pairs), 2), np.float32)
        ok = C.c_int(0)
        _check(lib().sphb_probe_force_pair(self._h, len(pairs), _p(pairs), variant, _p(out), C.byref(ok)), "sphb_probe_force_pair")
        return out, bool(ok.value)

    def pair_stats(self):
        c, a = C.c_double(), C.c_double()
        _check(lib().sphb_pair_stats(self._h, C.byref(c), C.byref(a)), "sphb_pair_stats")
        return c.value, a.value

    def profile(self, mode: int):
        _check(lib().sphb_profile(self._h, mode), "sphb_profile")

    def profile_read(self, reset: bool = True) -> dict:
        ms = (C.c_double * len(KERNEL_NAMES))()
        ln = (C.c_ulonglong * len(KERNEL_NAMES))()
        _check(lib().sphb_profile_read(self._h, ms, ln, int(reset)), "sphb_profile_read")
        return {k: {"ms": ms[i], "launches": int(ln[i])} for i, k in enumerate(KERNEL_NAMES)}

    def reorder_marks(self) -> int:
        """grid builds whose deterministic reorder used the cell marks (sphb_reorder_marks)"""
        n = C.c_ulonglong(0)
        _check(lib().sphb_reorder_marks(self._h, C.byref(n)), "sphb_reorder_marks")
        return int(n.value)

    def flush_l2(self):
        _check(lib().sphb_flush_l2(self._h), "sphb_flush_l2")

    @property
    def stream(self) -> int:
        return int(lib().sphb_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(lib().sphb_launch_count(self._h))


# ------------------------------------------------------------------------------- multi-GPU slabs

class MgInfo(C.Structure):
    """sphb_mg_info_t (include/sph_b200.h)."""
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("col_lo", C.c_int), ("col_hi", C.c_int),
                ("window_lo", C.c_int), ("window_hi", C.c_int), ("halo_capacity", C.c_int),
                ("particle_capacity", C.c_int), ("transport", C.c_int),
                ("message_bytes", C.c_ulonglong), ("bytes_sent", C.c_ulonglong), ("exchanges", C.c_ulonglong)]

    def asdict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


NCCL_ID_BYTES = 128
IPC_HANDLE_BYTES = 64


def grid_columns(prm: Params):
    r, c = C.c_int(), C.c_int()
    _check(lib().sphb_grid_columns(C.byref(prm), C.byref(r), C.byref(c)), "sphb_grid_columns")
    return r.value, c.value


def column_histogram(prm: Params, particles: np.ndarray) -> np.ndarray:
    _, cols = grid_columns(prm)
    hist = np.zeros(cols, np.uint64)
    _check(lib().sphb_column_histogram(C.byref(prm), _p(_particles(particles)), len(particles), _p(hist)),
           "sphb_column_histogram")
    return hist


def columns_of(prm: Params, x: np.ndarray) -> np.ndarray:
    """global cell column of each x (pi_sph_fluid.c:112), through the library's own host function"""
    f = lib().sphb_column_of
    return np.fromiter((f(C.byref(prm), float(v)) for v in np.asarray(x, np.float32)), np.int32, len(x))


CELL_COST = 0.034     # SPHB_CELL_COST: particle-equivalents per grid cell per step (the prefix scan)


def plan_cuts(hist: np.ndarray, world: int, min_width: int = 4, column_cost: float = 0.0) -> np.ndarray:
    """Cuts at the quantiles of particles-per-column (+ `column_cost` per column, e.g. CELL_COST * rows)."""
    hist = np.ascontiguousarray(hist, np.uint64)
    cuts = np.zeros(world + 1, np.int32)
    _check(lib().sphb_mg_plan_cuts_cost(_p(hist), len(hist), world, min_width, float(column_cost), _p(cuts)),
           "sphb_mg_plan_cuts_cost")
    return cuts


def scene_block_column_hist(prm: Params, x0, x1, y0, y1) -> np.ndarray:
    _, cols = grid_columns(prm)
    hist = np.zeros(cols, np.uint64)
    _check(lib().sphb_scene_block_column_hist(C.byref(prm), x0, x1, y0, y1, _p(hist)), "sphb_scene_block_column_hist")
    return hist


def scene_block_slab(prm: Params, x0, x1, y0, y1, col_lo: int, col_hi: int):
    """-> (particles of the block scene in cell columns [col_lo, col_hi), index of the first one)"""
    base = C.c_uint()
    n = _check(lib().sphb_scene_fill_block_slab(C.byref(prm), x0, x1, y0, y1, col_lo, col_hi, None, C.byref(base)),
               "sphb_scene_fill_block_slab")
    a = np.zeros(n, PARTICLE)
    _check(lib().sphb_scene_fill_block_slab(C.byref(prm), x0, x1, y0, y1, col_lo, col_hi, _p(a), C.byref(base)),
           "sphb_scene_fill_block_slab")
    return a, int(base.value)


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(NCCL_ID_BYTES)
    _check(lib().sphb_mg_unique_id(buf), "sphb_mg_unique_id")
    return buf.raw


class Slab(Simulation):
    """One rank of a multi-GPU run: a resident-tier handle restricted to a slab of cell columns."""

    def __init__(self, prm: Params, rank: int, world: int, col_lo: int, col_hi: int,
                 particle_capacity: int = 0, halo_capacity: int = 0):
        super().__init__(prm)
        _check(lib().sphb_mg_configure(self._h, rank, world, col_lo, col_hi, particle_capacity, halo_capacity),
               "sphb_mg_configure")
        self.rank, self.world = rank, world

    def connect_nccl(self, unique_id: bytes):
        _check(lib().sphb_mg_connect_nccl(self._h, C.c_char_p(unique_id)), "sphb_mg_connect_nccl")

    # peer-store transport across processes (sphb_mg_connect_ipc): every rank exports its receive block,
    # the ranks exchange the handles (any channel) and map their neighbours' blocks
    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        _check(lib().sphb_mg_ipc_handle(self._h, buf), "sphb_mg_ipc_handle")
        return buf.raw

    def connect_ipc(self, handles):
        """handles: every rank's ipc_handle(), indexed by rank."""
        left = handles[self.rank - 1] if self.rank > 0 else None
        right = handles[self.rank + 1] if self.rank < self.world - 1 else None
        _check(lib().sphb_mg_connect_ipc(self._h, left, right), "sphb_mg_connect_ipc")

    def disconnect_ipc(self):
        _check(lib().sphb_mg_disconnect_ipc(self._h), "sphb_mg_disconnect_ipc")

    def upload(self, fluid: np.ndarray, boundary: np.ndarray | None = None, ids: np.ndarray | None = None,
               id_base: int = 0):
        nb = 0 if boundary is None else len(boundary)
        if ids is not None:
            ids = np.ascontiguousarray(ids, np.uint32)
            assert len(ids) == len(fluid)
        _check(lib().sphb_mg_upload(self._h, _p(_particles(fluid)) if len(fluid) else None, _p(ids), id_base, len(fluid),
                                    _p(_particles(boundary)) if nb else None, nb), "sphb_mg_upload")
        self.n_fluid, self.n_boundary = len(fluid), nb

    def download(self, cap: int | None = None, accel: bool = True):
        """-> (ids, fluid, du, dv) of the particles this rank owns now, sorted by global id"""
        cap = int(cap if cap is not None else self.info()["particle_capacity"])
        fluid = np.zeros(cap, PARTICLE)
        ids = np.zeros(cap, np.uint32)
        du = np.zeros(cap, np.float32) if accel else None
        dv = np.zeros(cap, np.float32) if accel else None
        n = C.c_int()
        _check(lib().sphb_mg_download(self._h, cap, _p(fluid), _p(ids), _p(du), _p(dv), C.byref(n)), "sphb_mg_download")
        n = n.value
        order = np.argsort(ids[:n], kind="stable")
        return (ids[:n][order], fluid[:n][order], du[:n][order] if accel else None, dv[:n][order] if accel else None)

    def upload_accel(self, du: np.ndarray, dv: np.ndarray):
        du = np.ascontiguousarray(du, np.float32); dv = np.ascontiguousarray(dv, np.float32)
        assert len(du) == len(dv) == self.n_fluid
        _check(lib().sphb_mg_upload_accel(self._h, _p(du), _p(dv)), "sphb_mg_upload_accel")

    def download_into(self, fluid: np.ndarray, ids: np.ndarray, du: np.ndarray | None, dv: np.ndarray | None) -> int:
        """Owned particles into caller buffers (e.g. pinned), arrival order; returns how many."""
        n = C.c_int()
        _check(lib().sphb_mg_download(self._h, len(fluid), _p(_particles(fluid)), _p(ids), _p(du), _p(dv), C.byref(n)),
               "sphb_mg_download")
        return n.value

    def save_state(self, path):
        """sphb_mg_save_state: this rank's part of a slab run's state file."""
        _check(lib().sphb_mg_save_state(self._h, str(path).encode()), "sphb_mg_save_state")

    def load_parts(self, paths):
        """sphb_mg_load_state: the particles of ALL parts that fall into this (configured) rank's columns."""
        arr = (C.c_char_p * len(paths))(*[str(p).encode() for p in paths])
        _check(lib().sphb_mg_load_state(self._h, arr, len(paths)), "sphb_mg_load_state")
        self.n_boundary = 1      # the parts carry the boundary: sphb_init_boundary has work to do

    def rebalance(self, min_width: int = 4, column_cost: float = 0.0, min_imbalance: float = 0.0) -> bool:
        """sphb_mg_rebalance (collective, NCCL): re-cut the running slabs at the current quantiles; True if the cuts moved."""
        changed = C.c_int(0)
        _check(lib().sphb_mg_rebalance(self._h, min_width, float(column_cost), float(min_imbalance), C.byref(changed)),
               "sphb_mg_rebalance")
        return bool(changed.value)

    def rebalance_host(self, dist, min_width: int = 4, column_cost: float = 0.0, min_imbalance: float = 0.0) -> bool:
        """sphb_mg_rebalance_host with the bytes carried by a torch.distributed process group of CPU tensors
        (gloo): the same re-cut for host programs without NCCL between the ranks."""
        import torch
        rank, world = dist.get_rank(), dist.get_world_size()

        @C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int)
        def allreduce(_user, buf, count):
            a = np.ctypeslib.as_array(buf, shape=(count,))
            t = torch.from_numpy(a.view(np.int64))
            dist.all_reduce(t)
            return 0

        @C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), C.c_void_p,
                     C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong))
        def alltoallv(_user, send, sbytes, soff, recv, rbytes, roff):
            def view(base, off, nbytes):
                return torch.from_numpy(np.ctypeslib.as_array((C.c_ubyte * nbytes).from_address(base + off)))
            if sbytes[rank]:
                C.memmove(recv + roff[rank], send + soff[rank], sbytes[rank])
            reqs = []
            for r in range(world):
                if r != rank and rbytes[r]:
                    reqs.append(dist.irecv(view(recv, roff[r], rbytes[r]), src=r))
            for r in range(world):
                if r != rank and sbytes[r]:
                    reqs.append(dist.isend(view(send, soff[r], sbytes[r]), dst=r))
            for q in reqs:
                q.wait()
            return 0

        changed = C.c_int(0)
        _check(lib().sphb_mg_rebalance_host(self._h, min_width, float(column_cost), float(min_imbalance), allreduce, alltoallv,
                                            None, C.byref(changed)), "sphb_mg_rebalance_host")
        return bool(changed.value)

    def allreduce_stats(self) -> dict:
        st = Stats()
        _check(lib().sphb_get_stats(self._h, C.byref(st)), "sphb_get_stats")
        _check(lib().sphb_mg_allreduce_stats(self._h, C.byref(st)), "sphb_mg_allreduce_stats")
        return st.asdict()

    def info(self) -> dict:
        out = MgInfo()
        _check(lib().sphb_mg_info(self._h, C.byref(out)), "sphb_mg_info")
        return out.asdict()


class SlabGroup:
    """All slabs of a run inside this process (one host thread drives every GPU): the in-process
    transport, where the advect+bin kernel stores halo entries straight into the neighbour's
    receive buffer.  ``devices[r]`` is the CUDA device of slab r (they may coincide)."""

    def __init__(self, prm: Params, cuts, devices=None, particle_capacity: int = 0, halo_capacity: int = 0):
        world = len(cuts) - 1
        devices = list(devices) if devices is not None else [prm.device] * world
        self.cuts = [int(c) for c in cuts]
        self.slabs = []
        for r in range(world):
            p = Params.from_buffer_copy(prm)
            p.device = devices[r]
            self.slabs.append(Slab(p, r, world, self.cuts[r], self.cuts[r + 1], particle_capacity, halo_capacity))
        self._arr = (C.c_void_p * world)(*[s._h for s in self.slabs])
        _check(lib().sphb_mg_connect_local(self._arr, world), "sphb_mg_connect_local")
        self.prm = prm

    def close(self):
        for s in self.slabs:
            s.close()
        self.slabs = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def upload(self, fluid: np.ndarray, boundary: np.ndarray | None = None, accel=None):
        """Splits the whole scene by owned columns; ids are the indices into `fluid`.  `accel` =
        (du, dv) restores the accelerations too (restart / re-cut)."""
        col = columns_of(self.prm, fluid["x"]) if len(fluid) < 200000 else None
        if col is None:       # vectorised equivalent of sphb_column_of (float32 divide, truncate, clamp)
            _, cols = grid_columns(self.prm)
            d = (fluid["x"] - np.float32(self.prm.x_min)) / np.float32(self.prm.cell_length)
            col = np.clip(d.astype(np.int32), 0, cols - 1)
        self.n_fluid = len(fluid)
        self._boundary = boundary
        for r, s in enumerate(self.slabs):
            sel = np.nonzero((col >= self.cuts[r]) & (col < self.cuts[r + 1]))[0]
            s.upload(np.ascontiguousarray(fluid[sel]), boundary, ids=sel.astype(np.uint32))
            if accel is not None:
                s.upload_accel(accel[0][sel], accel[1][sel])

    def rebalance(self, min_width: int = 4):
        """Re-cut the slabs at the particle-count quantiles of the CURRENT state (the
        dam break drains the left slabs).  The state, accelerations included, moves to new slab
        contexts through the host, so the run continues bit-identically.  Returns the new cuts."""
        fluid, du, dv, _ = self.download()
        steps = [s.stats()["steps"] for s in self.slabs]
        cuts = plan_cuts(column_histogram(self.prm, fluid), len(self.slabs), min_width)
        devices = [s.prm.device for s in self.slabs]
        info = self.slabs[0].info()
        boundary = self._boundary
        self.close()
        self.__init__(self.prm, cuts, devices, 0, info["halo_capacity"])
        self.upload(fluid, boundary, accel=(du, dv))
        self.init_boundary()
        return self.cuts

    def save_state(self, prefix) -> list:
        """One part file per slab (sphb_mg_save_state); returns their paths."""
        paths = [f"{prefix}.part{r}" for r in range(len(self.slabs))]
        for s, p in zip(self.slabs, paths):
            s.save_state(p)
        return paths

    def load_parts(self, paths):
        """Every slab takes the particles of ALL parts that fall into its columns (any rank count wrote them)."""
        for s in self.slabs:
            s.load_parts(paths)
        self.n_fluid = sum(int(s.stats()["n_fluid"]) for s in self.slabs)

    def init_boundary(self):
        for s in self.slabs:
            s.init_boundary()

    def compute_accel(self, gx: float = 0.0, gy: float = -9.81):
        _check(lib().sphb_mg_group_compute_accel(self._arr, len(self.slabs), gx, gy), "sphb_mg_group_compute_accel")

    def step(self, nsteps: int = 1, gx: float = 0.0, gy: float = -9.81):
        _check(lib().sphb_mg_group_step(self._arr, len(self.slabs), gx, gy, None, nsteps), "sphb_mg_group_step")

    def step_trace(self, gravity_xy: np.ndarray):
        g = np.ascontiguousarray(gravity_xy, np.float32)
        _check(lib().sphb_mg_group_step(self._arr, len(self.slabs), 0.0, 0.0, _p(g), len(g)), "sphb_mg_group_step")

    def synchronize(self):
        _check(lib().sphb_mg_group_synchronize(self._arr, len(self.slabs)), "sphb_mg_group_synchronize")

    def download(self):
        """-> (fluid, du, dv) of the whole scene in original order, plus the owner rank of each particle"""
        fluid = np.zeros(self.n_fluid, PARTICLE)
        du = np.zeros(self.n_fluid, np.float32); dv = np.zeros(self.n_fluid, np.float32)
        owner = np.full(self.n_fluid, -1, np.int32)
        for r, s in enumerate(self.slabs):
            ids, f, a, b = s.download()
            assert (owner[ids] == -1).all(), "a particle is owned by two slabs"
            fluid[ids] = f; du[ids] = a; dv[ids] = b; owner[ids] = r
        return fluid, du, dv, owner

    def stats(self) -> dict:
        per = (Stats * len(self.slabs))()
        for r, s in enumerate(self.slabs):
            _check(lib().sphb_get_stats(s._h, C.byref(per[r])), "sphb_get_stats")
        out = Stats()
        _check(lib().sphb_mg_merge_stats(per, len(self.slabs), C.byref(out)), "sphb_mg_merge_stats")
        return out.asdict()


# ------------------------------------------------------------------------------- compat tier

class _Compat:
    """The reference's operator functions (pi_sph_fluid.c), same names, same argument order;
    arrays are numpy ``PARTICLE`` records modified in place exactly where the reference
    writes them."""

    def set_params(self, prm: Params):
        _check(lib().sphb_compat_set_params(C.byref(prm)), "sphb_compat_set_params")

    def alloc_neighbors_context(self, n_particles, x_min, x_max, y_min, y_max, cell_length):    # :82
        return C.c_void_p(lib().alloc_neighbors_context(n_particles, x_min, x_max, y_min, y_max, cell_length))

    def free_neighbors_context(self, ctx):
        lib().sphb_compat_free_context(ctx)

    def update_neighbors_context(self, ctx, particles):                                         # :104
        lib().update_neighbors_context(ctx, _p(_particles(particles)))

    def calculate_boundary_pseudomass(self, boundary, ctx_boundary):                            # :242
        lib().calculate_boundary_pseudomass(_p(_particles(boundary)), ctx_boundary)

    def calculate_density(self, fluid, boundary, ctx_fluid, ctx_boundary):                      # :263
        lib().calculate_density(_p(_particles(fluid)), _p(_particles(boundary)), ctx_fluid, ctx_boundary)

    def calculate_particle_pressure(self, particles, n_particles=None):                         # :294
        lib().calculate_particle_pressure(_p(_particles(particles)), len(particles) if n_particles is None else n_particles)

    def calculate_accelerations(self, du_dt, dv_dt, fluid, boundary, ctx_fluid, ctx_boundary, gravity_x, gravity_y):   # :303
        assert du_dt.dtype == np.float32 and dv_dt.dtype == np.float32
        lib().calculate_accelerations(_p(du_dt), _p(dv_dt), _p(_particles(fluid)), _p(_particles(boundary)),
                                      ctx_fluid, ctx_boundary, gravity_x, gravity_y)

    def draw_metaballs(self, draw_buffer, pixel_pseudoparticles, fluid, ctx_fluid):            # :380
        assert draw_buffer.dtype == np.uint8 and draw_buffer.size == 1024
        lib().draw_metaballs(_p(draw_buffer), _p(_particles(pixel_pseudoparticles)), _p(_particles(fluid)), ctx_fluid)


compat = _Compat()
