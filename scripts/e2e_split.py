#!/usr/bin/env python
"""Where the one-off part of the e2e leg goes at 64M particles (BASELINE configs[3], one GPU): host-timed
sphb_upload / sphb_init_boundary / sphb_compute_accel / sphb_download with pinned host buffers."""
import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import pi_sph_fluid_b200 as pkg

R = 0.00017677669529803097
prm = pkg.default_params(R)
fluid, boundary = pkg.scene_block(prm, 2 * R, 2.0, 2 * R, 1.0), pkg.scene_boundary(prm)
n = len(fluid)
fl = torch.empty(n * 7, dtype=torch.float32).pin_memory(); fl.numpy().view(pkg.PARTICLE)[:] = fluid
out = torch.empty(n * 7, dtype=torch.float32).pin_memory()
du = torch.empty(n, dtype=torch.float32).pin_memory(); dv = torch.empty(n, dtype=torch.float32).pin_memory()
fh, oh = fl.numpy().view(pkg.PARTICLE), out.numpy().view(pkg.PARTICLE)
with pkg.Simulation(prm) as sim:
    for rep in range(3):
        t = [time.perf_counter()]
        sim.upload(fh, boundary); sim.synchronize(); t.append(time.perf_counter())
        sim.init_boundary(); sim.synchronize(); t.append(time.perf_counter())
        sim.compute_accel(0.0, -9.81); sim.synchronize(); t.append(time.perf_counter())
        sim.step(5, 0.0, -9.81); sim.synchronize(); t.append(time.perf_counter())
        # the same 5 steps the way the e2e leg runs them: every step's statistics read on the host, one step behind
        import ctypes
        st = pkg.Stats(); ref = ctypes.byref(st)
        g = np.ascontiguousarray(np.tile([[0.0, -9.81]], (5, 1)), np.float32)
        t5 = time.perf_counter(); prev = None
        for i in range(5):
            tk = sim.step_stats_begin(g.ctypes.data + 8 * i, 1)
            if prev is not None: sim.step_stats_end(prev, ref)
            prev = tk
        sim.step_stats_end(prev, ref); sim.synchronize()
        t5 = (time.perf_counter() - t5) * 1e3
        t.append(time.perf_counter())
        sim.download_into(oh, du.numpy(), dv.numpy()); t.append(time.perf_counter())
        d = np.diff(t) * 1e3
        d = np.delete(d, 4)
        print("5 steps with per-step statistics: %.1f ms" % t5)
        print("n=%d  upload %.1f ms (%.1f GB/s)  init_boundary %.1f  compute_accel %.1f  5 steps %.1f  download %.1f ms (%.1f GB/s)"
              % (n, d[0], n * 28 / d[0] / 1e6, d[1], d[2], d[3], d[4], n * 36 / d[4] / 1e6), flush=True)
