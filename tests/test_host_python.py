"""Static checks of the host-side Python (bench.py, the ctypes binding, the torchrun check script):
they only run end to end on a GPU box, so a name that does not resolve must be caught here."""
import builtins
import symtable
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
FILES = ["bench.py", "__graft_entry__.py", "pi_sph_fluid_b200/api.py", "pi_sph_fluid_b200/build.py",
         "tests/mg_nccl_check.py", "tests/slab_emul.py", "scripts/long_run.py"]


@pytest.mark.parametrize("rel", FILES)
def test_no_undefined_globals(rel):
    path = ROOT / rel
    top = symtable.symtable(path.read_text(), str(path), "exec")
    module_names = {s.get_name() for s in top.get_symbols()}
    missing = []

    def walk(table):
        for s in table.get_symbols():
            name = s.get_name()
            if s.is_global() and s.is_referenced() and name not in module_names and not hasattr(builtins, name):
                missing.append(f"{table.get_name()}: {name}")
        for child in table.get_children():
            walk(child)
    walk(top)
    assert not missing, missing


def test_bench_slab_arm_releases_what_it_connects():
    """run_gpu_slabs maps its neighbours' receive blocks (peer-store transport) and must unmap them on
    every rank before any rank frees its own; the single-GPU arm has nothing to release."""
    src = (ROOT / "bench.py").read_text()
    single = src[src.index("def run_gpu("):src.index("def run_gpu_slabs(")]
    slabs = src[src.index("def run_gpu_slabs("):src.index("def main(")]
    assert "release(" not in single and "connect_ipc" not in single
    assert slabs.count("release(s") >= 2 and "disconnect_ipc" in slabs and "connect_ipc" in slabs


def test_bench_slab_sizing_covers_the_messages_and_the_library_minimum():
    """bench.py sizes a slab's message buffers from the column histogram around its cuts and its particle slots
    from the particles it uploads: every message must fit (two columns of the sender + slack), the slots must
    meet the library's minimum (n + two messages, sphb_mg_upload), and an end rank has one neighbour only."""
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    rng = np.random.default_rng(3)
    hist = np.zeros(400, np.int64)
    hist[10:250] = rng.integers(9000, 15000, 240)          # a block of fluid, the rest of the tank dry
    cuts = [0, 70, 130, 190, 400]
    world = len(cuts) - 1
    caps = [bench.slab_halo_capacity(hist, cuts, r, world) for r in range(world)]
    for r in range(world):
        lo, hi = cuts[r], cuts[r + 1]
        sends = ([int(hist[lo:lo + 2].sum())] if r > 0 else []) + ([int(hist[hi - 2:hi].sum())] if r < world - 1 else [])
        recvs = ([int(hist[lo - 2:lo].sum())] if r > 0 else []) + ([int(hist[hi:hi + 2].sum())] if r < world - 1 else [])
        assert all(2 * m <= caps[r] for m in sends + recvs)
    cap = max(caps)
    assert bench.slab_halo_capacity(hist, cuts, 0, 1) == 8192                 # no neighbour: the floor
    n = int(hist[cuts[1]:cuts[2]].sum())
    slots = bench.slab_particle_capacity(n, cap, bench.CAP_FACTOR)
    assert slots >= n + 2 * cap and slots <= 1.06 * n + 2 * cap + 1024
    assert bench.slab_particle_capacity(n, cap, 0.0) == 0                     # library default
