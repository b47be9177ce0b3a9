#!/usr/bin/env python
"""Where does the end-to-end step time go?  (GPU box)"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import pi_sph_fluid_b200 as pkg
R = float(sys.argv[1]) if len(sys.argv) > 1 else 0.002423
prm = pkg.default_params(R)
fluid, boundary = pkg.scene_drop(prm), pkg.scene_boundary(prm)
sim = pkg.Simulation(prm)
sim.upload(fluid, boundary); sim.init_boundary(); sim.compute_accel(0.0, -9.81); sim.step(50, 0.0, -9.81); sim.synchronize()
K = 500
def timeit(name, f):
    sim.synchronize(); t0 = time.perf_counter()
    for _ in range(K): f()
    sim.synchronize(); dt = (time.perf_counter() - t0) / K * 1e6
    print(f"{name:40s} {dt:8.1f} us/iter")
g = np.asarray([[0.0, -9.81]], np.float32)
timeit("step(1) async (queue K)", lambda: sim.step(1, 0.0, -9.81))
timeit("step(1) + synchronize", lambda: (sim.step(1, 0.0, -9.81), sim.synchronize()))
timeit("step_trace(1) + synchronize", lambda: (sim.step_trace(g), sim.synchronize()))
timeit("step(1) + stats", lambda: (sim.step(1, 0.0, -9.81), sim.stats()))
timeit("stats only", lambda: sim.stats())
timeit("synchronize only", lambda: sim.synchronize())
timeit("step(10) + stats", lambda: (sim.step(10, 0.0, -9.81), sim.stats()))
timeit("step_stats(1)", lambda: sim.step_stats(g))
g10 = np.tile(g, (10, 1))
timeit("step_stats(10)", lambda: sim.step_stats(g10))
import ctypes
st = pkg.Stats(); st_ref = ctypes.byref(st); addr = g.ctypes.data
timeit("step_stats_into(1)", lambda: sim.step_stats_into(addr, 1, st_ref))
prev = [None]
def piped():
    t = sim.step_stats_begin(addr, 1)
    if prev[0] is not None:
        sim.step_stats_end(prev[0], st_ref)
    prev[0] = t
timeit("begin(1) + end(previous)", piped)
sim.step_stats_end(prev[0], st_ref)
def once(name, f, reps=5):
    ts = []
    for _ in range(reps):
        sim.synchronize(); t0 = time.perf_counter(); f(); sim.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{name:40s} {sorted(ts)[len(ts) // 2]:8.3f} ms (median of {reps})")
import torch
n = len(fluid)
fl = torch.empty(n * 7, dtype=torch.float32).pin_memory().numpy().view(pkg.PARTICLE); fl[:] = fluid
bd = torch.empty(len(boundary) * 7, dtype=torch.float32).pin_memory().numpy().view(pkg.PARTICLE); bd[:] = boundary
out = torch.empty(n * 7, dtype=torch.float32).pin_memory().numpy().view(pkg.PARTICLE)
du = torch.empty(n, dtype=torch.float32).pin_memory().numpy(); dv = torch.empty(n, dtype=torch.float32).pin_memory().numpy()
once("upload (pinned)", lambda: sim.upload(fl, bd))
once("init_boundary", lambda: sim.init_boundary())
once("compute_accel", lambda: sim.compute_accel(0.0, -9.81))
once("download_into (pinned)", lambda: sim.download_into(out, du, dv))
