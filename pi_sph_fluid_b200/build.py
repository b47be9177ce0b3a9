"""Builds libsphb200.so (hand-written CUDA for sm_100a + the C ABI) and the C host driver.

In-tree build with plain nvcc/gcc so the artefacts travel with the repo snapshot:

    python -m pi_sph_fluid_b200.build            # library + host driver
    python -m pi_sph_fluid_b200.build --verbose  # also print ptxas resource usage
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libsphb200.so"
HOST_BIN = PKG / "host" / "sph_b200_main"

SOURCES = ["kernels_build.cu", "kernels_pair.cu", "kernels_aux.cu", "sphb_api.cu", "sphb_compat.cu", "sphb_mg.cu"]
HEADERS = ["sph_math.cuh", "sph_consts.h", "sphb_internal.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only; no other arch, no PTX fallback
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-fvisibility=hidden",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build_variant(tag: str, defines: dict) -> Path:
    """libsphb200_<tag>.so with -DSPHB_* overrides (kernel tuning experiments, scripts/tune_pair.py)."""
    out = PKG / f"libsphb200_{tag}.so"
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{k}={v}" for k, v in defines.items()], "-o", str(out),
           *[str(CSRC / s) for s in SOURCES], str(PKG / "host" / "scene.c"), "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed building variant {tag}")
    return out


def build_library(force: bool = False, verbose: bool = False) -> Path:
    deps = [CSRC / s for s in SOURCES + HEADERS] + [ROOT / "include" / "sph_b200.h", ROOT / "include" / "sph_b200_scene.h",
                                                    PKG / "host" / "scene.c", Path(__file__)]
    if not force and not _stale(LIB, deps):
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB), *[str(CSRC / s) for s in SOURCES], str(PKG / "host" / "scene.c")]
    cmd += ["-ldl"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building libsphb200.so")
    return LIB


def build_host(force: bool = False) -> Path:
    """The C host driver (pi_sph_fluid_b200/host/sph_main.c): plain C over the C ABI."""
    src = [PKG / "host" / "sph_main.c"]     # scene.c is part of libsphb200.so
    deps = src + [ROOT / "include" / "sph_b200.h", ROOT / "include" / "sph_b200_scene.h", LIB]
    if not force and not _stale(HOST_BIN, deps):
        return HOST_BIN
    cmd = ["/usr/bin/gcc", "-O2", "-Wall", "-std=gnu11", "-I", str(ROOT / "include"), *map(str, src),
           "-o", str(HOST_BIN), "-L", str(PKG), "-lsphb200", f"-Wl,-rpath,{PKG}", "-Wl,-rpath,$ORIGIN/..", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("gcc failed building the host driver")
    return HOST_BIN


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_library(force, verbose)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
    print(HOST_BIN)
