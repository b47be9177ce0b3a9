"""pi_sph_fluid_b200 — B200-native WCSPH hot path of colonelwatch/pi-sph-fluid.

The product is ``libsphb200.so`` (hand-written CUDA for sm_100a behind a C ABI, see
``include/sph_b200.h``) plus a plain-C host driver (``host/sph_main.c``).  This package is
only the thin ctypes binding the tests and ``bench.py`` use: every compute call goes
straight into the shared library, and there is no Python, PyTorch or CPU implementation
of any operator here.  If the library is missing or no sm_100 GPU is usable the calls
raise ``SphbError`` — nothing falls back.
"""
from .api import (  # noqa: F401
    PARTICLE,
    KERNEL_NAMES,
    Params,
    Simulation,
    Slab,
    SlabGroup,
    column_histogram,
    columns_of,
    grid_columns,
    nccl_unique_id,
    plan_cuts,
    scene_block_column_hist,
    scene_block_slab,
    SphbError,
    Stats,
    compat,
    default_params,
    gravity_from_raw,
    gravity_trace_tilt,
    lib,
    lib_path,
    scene_block,
    scene_boundary,
    scene_drop,
    spacing_for_count,
)

__all__ = [
    "PARTICLE", "KERNEL_NAMES", "Params", "Simulation", "Slab", "SlabGroup", "column_histogram", "columns_of",
    "grid_columns", "nccl_unique_id", "plan_cuts", "scene_block_column_hist", "scene_block_slab", "SphbError", "Stats", "compat",
    "default_params", "gravity_from_raw", "gravity_trace_tilt", "lib", "lib_path",
    "scene_block", "scene_boundary", "scene_drop", "spacing_for_count",
]
