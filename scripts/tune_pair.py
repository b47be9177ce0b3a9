#!/usr/bin/env python
"""Kernel tuning sweep: builds libsphb200_<tag>.so variants (here, no GPU needed) or times them
(on the GPU box).   tune_pair.py build | run [workload ...]"""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
VARIANTS = {
    "base": {},
    "k0": {"SPHB_DENS_KIND": 0},
    "k0_l32": {"SPHB_DENS_KIND": 0, "SPHB_LIST_CAP": 32},
    "k0_l32_t512": {"SPHB_DENS_KIND": 0, "SPHB_LIST_CAP": 32, "SPHB_TILE_CAP": 512},
    "k0_l32_t512_mb": {"SPHB_DENS_KIND": 0, "SPHB_LIST_CAP": 32, "SPHB_TILE_CAP": 512, "SPHB_MINB_D": 12, "SPHB_MINB_F": 10},
    "l32_d32": {"SPHB_LIST_CAP": 32, "SPHB_DLIST_CAP": 32},
    "pt64": {"SPHB_PT": 64, "SPHB_TILE_CAP": 320},
    "pt64_k0_l32": {"SPHB_PT": 64, "SPHB_TILE_CAP": 320, "SPHB_DENS_KIND": 0, "SPHB_LIST_CAP": 32},
    "pt256": {"SPHB_PT": 256, "SPHB_TILE_CAP": 1024, "SPHB_LIST_CAP": 40, "SPHB_DLIST_CAP": 32},
    "pt256_k0_l32": {"SPHB_PT": 256, "SPHB_TILE_CAP": 1024, "SPHB_DENS_KIND": 0, "SPHB_LIST_CAP": 32},
}
if sys.argv[1] == "build":
    from pi_sph_fluid_b200 import build
    for tag, d in VARIANTS.items():
        build.build_variant(tag, d); print("built", tag)
else:
    wl = sys.argv[2:] or ["drop256k", "dam4m"]
    for tag in VARIANTS:
        for w in wl:
            env = dict(os.environ, SPHB_LIB_VARIANT=tag)
            r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--workload", w, "--steps", "100" if w != "drop256k" else "300",
                                "--no-cpu-baseline", "--no-e2e"], capture_output=True, text=True, env=env)
            try:
                j = json.loads(r.stdout.strip().splitlines()[-1])
                k = j["roofline"]["kernels"]
                print(f"{tag:18s} {w:9s} value={j['value']:.4e} ms/step={j['ms_per_step']:.4f} dens={k['density']['ms']:.4f} force={k['force']['ms']:.4f}", flush=True)
            except Exception as e:
                print(tag, w, "FAILED", r.stderr[-400:], flush=True)
