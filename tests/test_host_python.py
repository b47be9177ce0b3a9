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
