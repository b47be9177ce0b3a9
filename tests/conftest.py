"""Shared fixtures.  CPU tests (`-m "not gpu"`) cover the oracle against the reference-built
golden vectors, the host logic, the device arithmetic compiled for the host, and the ABI
surface.  GPU tests (`-m gpu`) are the parity tests proper and call through the C ABI."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
G = (0.0, -9.81)     # pi_sph_fluid.c:442-443


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by `pytest -m gpu` on the GPU box)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 B200s (run by `pytest -m multigpu` under `gpurun --gpus 2`)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _gpu_count() -> int:
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    ngpu = _gpu_count()
    if ngpu < 2:
        # tests that need two real devices are not part of a one-GPU (or CPU) session at all: they run under
        # `gpurun --gpus 2 -- pytest -m multigpu` and their log is committed under profiles/
        multi = [it for it in items if "multigpu" in it.keywords]
        if multi:
            items[:] = [it for it in items if "multigpu" not in it.keywords]
            config.hook.pytest_deselected(items=multi)
    if ngpu > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device here (GPU tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_built():
    from oracle import pyoracle
    pyoracle.build(ref=Path("/root/reference/pi_sph_fluid.c").exists())
    return pyoracle


@pytest.fixture(scope="session")
def lib_built():
    from pi_sph_fluid_b200 import build
    build.build_all()
    import pi_sph_fluid_b200 as pkg
    return pkg


@pytest.fixture(scope="session")
def emu(lib_built):
    """tests/emu/emu.cpp: the device arithmetic compiled for the host (test infra only)."""
    import ctypes as C
    src = ROOT / "tests" / "emu" / "emu.cpp"
    so = ROOT / "tests" / "emu" / "libemu.so"
    deps = [src, ROOT / "pi_sph_fluid_b200" / "csrc" / "sph_math.cuh", ROOT / "pi_sph_fluid_b200" / "csrc" / "sph_consts.h"]
    if not so.exists() or any(d.stat().st_mtime > so.stat().st_mtime for d in deps):
        subprocess.run(["/usr/bin/g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                        "-std=c++17", str(src), "-o", str(so)], check=True)
    return C.CDLL(str(so))


@pytest.fixture(scope="session")
def golden075():
    return np.load(GOLDEN / "drop_R0.075.npz")


@pytest.fixture(scope="session")
def golden02():
    return np.load(GOLDEN / "drop_R0.02.npz")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def same_bits(a, b) -> bool:
    return np.array_equal(bits(a), bits(b))


def same_bits_nan(a, b) -> bool:
    """Bit-for-bit, any NaN equal to any NaN (0/0 is 0xFFC00000 on x86 and 0x7FFFFFFF on the GPU)."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(bits(a)[~na], bits(b)[~nb])
