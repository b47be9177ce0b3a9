#!/usr/bin/env python
"""bench.py — particle-updates/s of the WCSPH step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

One "step" = one leapfrog step (kick, drift, cell build, density+pressure, accelerations,
kick: pi_sph_fluid.c:612-641) over every fluid particle of the workload.

Workload at EVERY N (strong scaling, so the 1 -> 8 GPU curve is one configuration): BASELINE.json
configs[3], the 64M-particle 2-D dam break — on one GPU at N = 1 (it fits: ~170 B x 64M of 180 GB),
as x-slabs with halo exchange over NVLink at N = 2, 4, 8.  The N = 1 line also carries, as named
`secondary` entries measured in the same run, configs[1] (the reference's drop scene at 262,204
particles, with the reference's own compiled code as its CPU arm) and configs[2] (4M dam break), and
the cost of the reference-arithmetic force pass against fast_force.

Keys of the JSON line (see the round prompt for the contract):
  value         whole-job fluid-particle-updates/s, state resident in HBM, CUDA events on the
                library's stream around each step, L2 flushed between steps (untimed)
  e2e           same metric through the C ABI with HOST buffers: sphb_upload (pinned host ->
                HBM) + K x (sphb_step with that step's gravity + sphb_get_stats read-back) +
                sphb_download, all inside the timed region
  roofline      dominant kernel (k_force): algorithmic bytes / launch over its mean CUDA-event
                duration vs MEASURED_PEAKS.json hbm_gbs; `fp32` gives the binding roof
  cpu_baseline  the reference's own compiled code (oracle/_ref, OpenMP) on this host, bounded sample
  --impl reference   the reference arm: times oracle/_ref only (rank 0), same workload/metric
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "fluid-particle-updates/sec"
UNIT = "particle-updates/s"
G = (0.0, -9.81)
CAP_FACTOR = 1.05        # slabs: particle slots per rank / particles uploaded to it

# algorithmic HBM bytes / particle / launch (DESIGN.md "Kernels")
ALGO_BYTES = {"advect_bin": 44.0, "reorder": 40.0, "density": 16.0, "force": 40.0}
# algorithmic flops / particle: 6*C + 13*P (density), 6*C + 44*P (force); C, P measured by the run


def workload_spec(name: str):
    """-> dict(name, R, scene, args)"""
    if name == "drop256k":
        return dict(name="drop_R0.002423_262204_fluid_4954_boundary (BASELINE configs[1])", R=0.002423, scene="drop", ref_tag="0.002423")
    if name == "drop61k":
        return dict(name="drop_R0.005_61519_fluid", R=0.005, scene="drop", ref_tag="0.005")
    if name == "drop4k":
        return dict(name="drop_R0.02_3848_fluid", R=0.02, scene="drop", ref_tag="0.02")
    if name == "drop269":
        return dict(name="drop_R0.075_269_fluid (BASELINE configs[0])", R=0.075, scene="drop", ref_tag=None)
    if name == "dam4m":
        return dict(name="dam_break_R0.0005_4M (BASELINE configs[2])", R=0.0005, scene="dam", block=(2.0, 0.5), ref_tag=None)
    if name == "slosh16m":
        # BASELINE configs[4]: tank filled to y = 1 m, R = 5e-4 -> 16M particles, gravity from a synthetic
        # MPU6050 trace (+-20 degrees, one period per 2000 steps, samples held 50 steps)
        R = 5e-4
        return dict(name="slosh_tank_filled_to_1m_R5e-04_16M_tilt20deg (BASELINE configs[4])", R=R, scene="tank",
                    box=(2 * R, 4.0 - 2 * R, 2 * R, 1.0), tilt=(20.0, 2000, 50), ref_tag=None)
    if name.startswith("dam") and name.endswith("m"):
        # dam break, block x in [2R,2) x y in [2R,1): 2 m^2 -> R = sqrt(2 / N).  dam64m = BASELINE configs[3]
        n = float(name[3:-1]) * 1e6
        R = float(np.float32(np.sqrt(2.0 / n)))
        tag = " (BASELINE configs[3])" if name == "dam64m" else ""
        return dict(name=f"dam_break_2x1m_block_R{R:.4e}_{name[3:-1]}M{tag}", R=R, scene="dam", block=(2.0, 1.0), ref_tag=None)
    raise SystemExit(f"unknown workload {name}")


def read_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


def traffic_file():
    """The newest committed ncu traffic summary (profiles/rNN*_traffic.json, scripts/ncu_traffic.py)."""
    files = sorted((ROOT / "profiles").glob("r*_traffic.json"))
    return files[-1] if files else None


def read_traffic(workload: str) -> dict:
    """Per-kernel DRAM bytes per launch from the committed `ncu --set full` capture of this workload;
    {} when there is none."""
    p = traffic_file()
    try:
        return json.loads(p.read_text()).get(workload, {}) if p else {}
    except (OSError, ValueError):
        return {}


def make_roofline(workload, n, kern, force_ms, dens_ms, cand, acc, hbm_peak, sm_max_mhz, peak_src, sm_count, suffix=""):
    """The roofline object of the JSON line, for the DOMINANT kernel of the step (the pair kernel with
    the longer average launch, CUDA events on the library's stream).  `achieved` = algorithmic bytes of
    one launch (SURVEY.md 8d per-particle figure x particles) / that duration.  Both pair kernels are
    FP32-issue-bound, so the flop-side figures and ncu's issue-slot utilisation are listed next to it."""
    traffic = read_traffic(workload)
    ms = {"force": force_ms, "density": dens_ms}
    dom = "density" if dens_ms >= force_ms else "force"
    achieved = ALGO_BYTES[dom] * n / (ms[dom] * 1e-3) / 1e9
    fp32_peak = sm_count * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    flops = {"force": (6 * cand + 44 * acc) * n, "density": (6 * cand + 13 * acc) * n}
    tf = {k: flops[k] / (ms[k] * 1e-3) / 1e12 for k in ms}
    for name, d in kern.items():
        t = traffic.get(name)
        if t:
            d["ncu_dram_bytes_per_launch"] = t["traffic_bytes"]
            d["ncu_issue_active_pct"] = t["issue_active_pct"]
    t = traffic.get(dom)
    return {
        "kernel": f"k_{dom}{suffix}", "bound": "hbm", "achieved": round(achieved, 2), "peak": hbm_peak, "unit": "GB/s",
        "frac": round(achieved / hbm_peak, 5), "traffic": t["traffic_bytes"] if t else None,
        "traffic_source": f"profiles/{traffic_file().name} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)" if t else None,
        "peak_source": peak_src,
        "algorithmic_bytes_per_particle": ALGO_BYTES[dom], "algorithmic_bytes_per_launch": ALGO_BYTES[dom] * n,
        "note": "k_density and k_force are FP32-issue-bound, not HBM-bound (SURVEY.md §8d): see fp32 and the "
                "ncu issue-slot utilisation per kernel; DRAM traffic above the algorithmic bytes is the "
                "neighbour lists k_density hands to k_force",
        "fp32": {"force_TFLOPs": round(tf["force"], 3), "density_TFLOPs": round(tf["density"], 3),
                 "peak_TFLOPs": round(fp32_peak, 1), "peak_def": f"{sm_count} SM x 128 lanes x 2 x {sm_max_mhz:.0f} MHz",
                 "force_frac": round(tf["force"] / fp32_peak, 4), "density_frac": round(tf["density"] / fp32_peak, 4),
                 "pairs_per_particle": {"candidates": round(cand, 2), "accepted": round(acc, 2)}},
        "kernels": kern,
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def wait_first(self, timeout: float = 5.0):
        """nvidia-smi takes a while to start: block until its first sample (or the timeout)."""
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)
        self.rows.clear()                # that sample predates the load

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_scene(pkg, spec, deterministic=True, device=0, fast_force=False):
    prm = pkg.default_params(spec["R"], deterministic=deterministic, device=device, fast_force=fast_force)
    if spec["scene"] == "drop":
        fluid = pkg.scene_drop(prm)
    elif "box" in spec:
        fluid = pkg.scene_block(prm, *spec["box"])
    else:
        x1, y1 = spec["block"]
        fluid = pkg.scene_block(prm, 2 * spec["R"], x1, 2 * spec["R"], y1)     # 2R off the walls (DESIGN.md "Scenes")
    boundary = pkg.scene_boundary(prm)
    return prm, fluid, boundary


# ------------------------------------------------------------------------------------------ CPU arm

def host_cores() -> int:
    """The host threads the CPU arm uses: every core this process may run on.  Asked for explicitly
    (num_threads / omp_set_num_threads), because torchrun exports OMP_NUM_THREADS=1 to its workers."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def run_reference(spec, steps: int, warmup: int, budget_s: float = 90.0):
    """Times the reference's own compiled operators (oracle/_ref, shipped -Ofast flags, OpenMP
    over all host cores) on the workload.  Returns dict or None if unavailable."""
    from oracle import pyoracle
    tag = spec["ref_tag"]
    R = spec["R"]
    if spec["scene"] != "drop":
        return run_reference_block(spec, steps, warmup, budget_s)
    kind = "reference"
    try:
        if not pyoracle.reference_available(tag, "fast"):
            raise FileNotFoundError
        ref = pyoracle.Reference(R=tag, flavour="fast")
    except (FileNotFoundError, OSError):
        ref = None
    cores = host_cores()
    o = pyoracle.Oracle(R=R, variant="fast", threads=cores)
    fluid, boundary = o.scene_drop(), o.scene_boundary()
    if ref is not None:
        cb = ref.init_boundary(boundary); cf = ref.ctx(len(fluid))
        du, dv = ref.compute_accel(fluid, boundary, cf, cb, *G)
        run = lambda n: ref.step(fluid, boundary, cf, cb, du, dv, n, *G, threads=cores)
    else:
        kind = "port"
        gb = o.init_boundary(boundary); gf = o.grid(len(fluid))
        du, dv = o.compute_accel(fluid, boundary, gf, gb, *G)
        cores = o.threads

        def run(n):
            t0 = time.perf_counter(); o.step(fluid, boundary, gf, gb, du, dv, n, *G); return time.perf_counter() - t0
    t_probe = run(max(1, min(warmup, 3)))
    per_step = t_probe / max(1, min(warmup, 3))
    # at least ~2.5 s of timed steps whatever K is (a 0.2 s sample is noise), at most the budget
    n_run = int(max(1, min(max(steps, 2.5 / max(per_step, 1e-9)), budget_s / max(per_step, 1e-9))))
    t = run(n_run)
    return {"value": len(fluid) * n_run / t, "unit": UNIT, "cores": int(cores), "kind": kind,
            "sample": f"{n_run} steps (asked: {steps}; at least 2.5 s are timed) of the full {len(fluid)}-particle scene, {t:.2f} s",
            "ms_per_step": 1e3 * t / n_run, "n_fluid": int(len(fluid)), "steps_run": n_run}


def run_reference_block(spec, steps: int, warmup: int, budget_s: float, sample_particles: float = 1.0e6,
                        min_timed_s: float = 3.0):
    """Dam-break scenes are not in the reference (its main() hard-codes the drop and R is a macro),
    so the CPU arm for them is the oracle port (same operators, the reference's shipped -Ofast flags,
    OpenMP on all host cores) on a BOUNDED SAMPLE: the leftmost part of the same block at the same
    spacing, about `sample_particles` particles."""
    from oracle import pyoracle
    R = spec["R"]
    o = pyoracle.Oracle(R=R, variant="fast", threads=host_cores())
    x1, y1 = spec["block"] if "block" in spec else (spec["box"][1], spec["box"][3])
    ny = max(1.0, (y1 - 2 * R) / R)
    # the timed region should last a few seconds whatever K is: a sub-block of ~1M particles takes ~30 ms per
    # step on 16-32 host threads, so grow the sample when K is small (bounded by 8M particles and the block)
    sample_particles = min(8.0e6, max(sample_particles, min_timed_s * 3.0e7 / max(1, steps)))
    width = min(x1 - 2 * R, max(8 * 2.6 * R, sample_particles / ny * R))
    fluid = o.scene_block(2 * R, 2 * R + width, 2 * R, y1)
    boundary = o.scene_boundary()
    gb = o.init_boundary(boundary); gf = o.grid(len(fluid))
    du, dv = o.compute_accel(fluid, boundary, gf, gb, *G)

    def run(n):
        t0 = time.perf_counter(); o.step(fluid, boundary, gf, gb, du, dv, n, *G); return time.perf_counter() - t0
    nprobe = max(1, min(warmup, 3))
    per_step = run(nprobe) / nprobe
    n_run = int(max(1, min(steps, budget_s / max(per_step, 1e-9))))
    t = run(n_run)
    return {"value": len(fluid) * n_run / t, "unit": UNIT, "cores": int(o.threads), "kind": "port",
            "sample": f"{n_run} steps of a {len(fluid)}-particle sub-block (x in [2R, 2R+{width:.4f}) of the same block, same R), {t:.2f} s",
            "ms_per_step": 1e3 * t / n_run, "n_fluid": int(len(fluid)), "steps_run": n_run}


# ------------------------------------------------------------------------------------------ GPU arm

def run_gpu(args, spec, rank, world, K=None, W=None, want_e2e=True, fast_force=None, workload_key=None):
    """One GPU, any scene.  K / W / fast_force default to the command line's; the secondary entries of the
    N = 1 line call this again with their own."""
    import torch
    import pi_sph_fluid_b200 as pkg

    dev = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(dev)
    fast_force = args.fast_force if fast_force is None else fast_force
    prm, fluid, boundary = build_scene(pkg, spec, deterministic=not args.nondeterministic, device=dev,
                                       fast_force=fast_force)
    n = len(fluid)
    K = args.steps if K is None else K
    W = max(args.warmup if W is None else W, 3)
    workload_key = workload_key or args.workload

    sim = pkg.Simulation(prm)
    stream = torch.cuda.ExternalStream(sim.stream, device=dev)
    sim.upload(fluid, boundary)
    sim.init_boundary()
    sim.compute_accel(*G)
    # clocks / throttle reasons are sampled from here to the end of the per-kernel pass: the K timed
    # steps alone can be shorter than one nvidia-smi sample, so the sampler also sees the warm-up,
    # which is stretched to at least 0.4 s of the same steps (same kernels, same load)
    clocks = ClockSampler(dev)
    clocks.start()
    sim.step(W, *G)
    sim.synchronize()
    clocks.wait_first()
    t_w = time.perf_counter()
    while time.perf_counter() - t_w < 0.4:
        sim.step(50 if n < 4_000_000 else 5, *G)
        sim.synchronize()
    cand, acc = sim.pair_stats()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- value: K steps, per-step events on the library's stream, L2 flushed between steps
    launches0 = sim.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    t_wall0 = time.perf_counter()
    for a, b in ev:
        sim.flush_l2()
        a.record(stream)
        sim.step(1, *G)
        b.record(stream)
    sim.synchronize()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    gpu_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = sim.launch_count - launches0
    if world > 1:
        tmax = torch.tensor([gpu_ms], device=f"cuda:{dev}")
        torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
        gpu_ms = float(tmax.item())
    total_particles = n * world
    value = total_particles * K / (gpu_ms * 1e-3)

    # ---- warm-L2 variant (steady state of a resident simulation), for reference in config
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a.record(stream); sim.step(K, *G); b.record(stream)
    sim.synchronize()
    warm_ms = a.elapsed_time(b)

    # ---- per-kernel CUDA-event times (separate pass: events between kernels perturb the step)
    sim.profile(1)
    sim.profile_read(reset=True)
    for _ in range(K):
        sim.flush_l2()
        sim.step(1, *G)
    prof = sim.profile_read(reset=True)
    sim.profile(0)
    clk = clocks.stop()
    hbm_peak, sm_max_mhz, peak_src = read_peaks()
    kern = {}
    for name, d in prof.items():
        if d["launches"] == 0:
            continue
        ms = d["ms"] / d["launches"]
        kern[name] = {"ms": round(ms, 5), "launches_per_step": d["launches"] / K}
        if name in ALGO_BYTES:
            kern[name]["algo_GBps"] = round(ALGO_BYTES[name] * n / (ms * 1e-3) / 1e9, 2)
    force_ms = prof["force"]["ms"] / max(1, prof["force"]["launches"])
    dens_ms = prof["density"]["ms"] / max(1, prof["density"]["launches"])
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    roofline = make_roofline(workload_key, n, kern, force_ms, dens_ms, cand, acc, hbm_peak, sm_max_mhz, peak_src, sm_count)

    if args.no_e2e or not want_e2e:
        sim.close()
        return {"metric": METRIC, "value": value, "ms_per_step": gpu_ms / K, "roofline": roofline, "tuning_run": True,
                "steps": K, "n_fluid": n, "warm_l2_value": total_particles * K / (warm_ms * 1e-3), "clocks": clk,
                "gpu_launches": int(launches), "fast_force": bool(fast_force)}
    # ---- e2e: host buffers -> C ABI -> host buffers, copies inside the timed region
    fl_pin = torch.empty(n * 7, dtype=torch.float32).pin_memory()
    bd_pin = torch.empty(len(boundary) * 7, dtype=torch.float32).pin_memory()
    fl_host = fl_pin.numpy().view(pkg.PARTICLE); bd_host = bd_pin.numpy().view(pkg.PARTICLE)
    fl_host[:] = fluid; bd_host[:] = boundary
    out_pin = torch.empty(n * 7, dtype=torch.float32).pin_memory()
    du_pin = torch.empty(n, dtype=torch.float32).pin_memory(); dv_pin = torch.empty(n, dtype=torch.float32).pin_memory()
    out_host = out_pin.numpy().view(pkg.PARTICLE)
    trace = np.tile(np.asarray([G], np.float32), (K, 1))
    sim2 = pkg.Simulation(prm)
    sim2.upload(fl_host, bd_host); sim2.init_boundary(); sim2.compute_accel(*G); sim2.step(W, *G); sim2.synchronize()
    # three repetitions of the whole region, the median is reported (host-side jitter of a shared box
    # shows up here, not in the device-timed value); all three are listed
    runs, last = [], None
    for _rep in range(3):
        barrier()
        t0 = time.perf_counter()
        sim2.upload(fl_host, bd_host)                      # H2D: 28 B x (n_fluid + n_boundary)
        sim2.init_boundary()
        sim2.compute_accel(*G)
        st = pkg.Stats()
        st_ref, g_addr = ctypes.byref(st), trace.ctypes.data
        prev = None
        for s in range(K):
            # that step's gravity sample in (8 B); the step's statistics (:656-675) out: 136 B written by
            # the force pass into mapped pinned host memory.  The host launches step s, then reads the
            # statistics of step s - 1 (sphb_step_stats_begin / _end): every step's result is read, and the
            # GPU does not idle while the host picks it up.
            t = sim2.step_stats_begin(g_addr + 8 * s, 1)
            if prev is not None:
                sim2.step_stats_end(prev, st_ref)
            prev = t
        sim2.step_stats_end(prev, st_ref)
        last = st.asdict()
        sim2.download_into(out_host, du_pin.numpy(), dv_pin.numpy())     # D2H: 36 B x n_fluid
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            tmax = torch.tensor([e2e_s], device=f"cuda:{dev}")
            torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
            e2e_s = float(tmax.item())
        runs.append(e2e_s)
    e2e_s = sorted(runs)[1]
    h2d = (28 * (n + len(boundary))) / K + 8
    d2h = (36 * n) / K + 136
    e2e = {"value": total_particles * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": round(h2d, 1),
           "d2h_bytes_per_step": round(d2h, 1), "ms_per_step": round(1e3 * e2e_s / K, 5),
           "ms_per_step_runs": [round(1e3 * r / K, 5) for r in runs],
           "path": "sphb_upload + sphb_init_boundary + sphb_compute_accel + K x (sphb_step_stats_begin(1 step), sphb_step_stats_end(previous step)) + sphb_download, pinned host buffers; every step's statistics are read on the host, one step behind the launches; median of 3 repetitions",
           "last_step_stats": {"max_speed": last["max_speed"], "max_rho_err": last["max_rho_err"]}}
    sim2.close()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": gpu_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 (f64 at the reference's double sites: pi_sph_fluid.c:325, :332, :334, :616)",
        "data": ("synthetic (the reference's own lattice drop scene, pi_sph_fluid.c:484-540)" if spec["scene"] == "drop" else
                 "synthetic (dam-break block on the reference's lattice idiom; not a reference scene)"),
        "config": {"workload": spec["name"], "n_fluid": n, "n_boundary": int(len(boundary)), "R": spec["R"],
                   "deterministic_order": not args.nondeterministic,
                   "force_arithmetic": ("fast_force = 1: single precision, approximate rsqrt/rcp (outside the 1e-4 parity bar at >= 4M particles)"
                                        if fast_force else
                                        "the reference's own (default): du_dt, dv_dt and whole runs bit-identical with the oracle"),
                   "l2": "flushed between timed steps (256 MiB write on the same stream, untimed)"
                         + ("; the state (~170 B/particle) is also far larger than the 126 MB L2" if n > 2_000_000 else ""),
                   "warm_l2_value": total_particles * K / (warm_ms * 1e-3),
                   "timing": "CUDA events on the library stream around each step; max over ranks",
                   "wall_s_timed_region": round(t_wall, 4),
                   "scaling_note": scaling_note(),
                   "build": pkg.lib().sphb_build_info().decode()},
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
    }
    sim.close()
    return line


def scaling_note():
    """Every N runs the SAME configuration (BASELINE configs[3], the 64M-particle dam break): strong scaling,
    so value(N) / (N x value(1)) is the parallel efficiency.  `--workload` overrides it (tuning, other configs)."""
    return {"family": "strong scaling on BASELINE configs[3]: one 64M-particle dam break at every N",
            "n1": "the whole 64M-particle scene on one B200", "n_gt_1": "x-slabs of it, 64M / N particles per GPU"}


def slab_halo_capacity(hist, cuts, rank: int, world: int) -> int:
    """Entries one halo + migration message of this rank must hold.  A message carries the sender's two columns
    next to a cut (ghosts for the receiver) plus the few particles that crossed it: 2 x the fuller side of the
    fuller of this rank's cuts (the ranks then agree on the maximum)."""
    def side(c):
        return max(int(hist[max(c - 2, 0):c].sum()), int(hist[c:c + 2].sum()))
    est = max(side(int(cuts[rank])) if rank > 0 else 0, side(int(cuts[rank + 1])) if rank < world - 1 else 0)
    return max(8192, 2 * est)


def slab_particle_capacity(n: int, halo_cap: int, factor: float) -> int:
    """Particle slots of a rank that uploads n particles: factor x n + two messages' worth (the least the library
    takes is n + 2 x halo_cap); 0 = the library's own default (1.25 x + four messages)."""
    return int(factor * n) + 2 * halo_cap + 1024 if factor > 0 else 0


def run_gpu_slabs(args, spec, rank, world):
    """N > 1: the dam-break block cut into x-slabs at the particle-count quantiles, one process per GPU,
    halo + migration as peer stores over NVLink (or NCCL send/recv).  Strong scaling: the same scene at every N."""
    import torch
    import torch.distributed as dist
    import pi_sph_fluid_b200 as pkg

    dev = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(dev)
    R = spec["R"]
    prm = pkg.default_params(R, deterministic=not args.nondeterministic, device=dev, fast_force=args.fast_force)
    if "box" in spec:
        box = spec["box"]
    else:
        x1, y1 = spec["block"]
        box = (2 * R, x1, 2 * R, y1)
    hist = pkg.scene_block_column_hist(prm, *box)
    rows, _cols = pkg.grid_columns(prm)
    # cuts at the quantiles of (particles + what a column's cells cost the scan): the dry half of the tank is
    # shared out instead of being the last rank's burden (its k_scan ran 0.108 ms against 0.019 ms elsewhere)
    cuts = pkg.plan_cuts(hist, world, column_cost=pkg.api.CELL_COST * rows)
    n_total = int(hist.sum())
    part, base = pkg.scene_block_slab(prm, *box, int(cuts[rank]), int(cuts[rank + 1]))
    boundary = pkg.scene_boundary(prm)
    n = len(part)
    halo_cap = slab_halo_capacity(hist, cuts, rank, world)
    cap_t = torch.tensor([halo_cap], device=f"cuda:{dev}")
    dist.all_reduce(cap_t, op=dist.ReduceOp.MAX)
    halo_cap = int(cap_t.item())
    K, W = args.steps, max(args.warmup, 3)

    ident = [pkg.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)

    def make():
        # particle slots per rank: 1.05 x its particles + two messages' worth (the least the library takes; its
        # own default is 1.25 x + four messages, for long runs that do not re-cut).  Kernels are launched for
        # the slot capacity, so unused slots cost empty CTAs (~1.5 us per 1184 of them): at N = 8 the round-1
        # sizing (1.10 x + four messages of twice the size) left 1.8M empty slots per rank, ~3 % of the step.
        # An overflow would show in merged_stats.n_overflow.  SPHB_BENCH_CAP_FACTOR overrides (0: library default).
        capf = float(os.environ.get("SPHB_BENCH_CAP_FACTOR", str(CAP_FACTOR)))
        pcap = slab_particle_capacity(n, halo_cap, capf)
        s_ = pkg.Slab(prm, rank, world, int(cuts[rank]), int(cuts[rank + 1]), particle_capacity=pcap, halo_capacity=halo_cap)
        return s_
    # transport of the per-step halo + migration message: "ipc" = the advect+bin kernel stores the entries
    # straight into the neighbour's receive buffer over NVLink (CUDA IPC mapping, device-side completion
    # signal, no NCCL call in the step); "nccl" = ncclSend/ncclRecv.  If any rank cannot map its
    # neighbours' buffers every rank stays on NCCL, and the line says so.
    want_ipc = os.environ.get("SPHB_BENCH_TRANSPORT", "ipc") != "nccl"
    transport_note = []

    def connect(s_, ident_):
        s_.connect_nccl(ident_)
        if not want_ipc:
            return "nccl"
        handles = [None] * world
        dist.all_gather_object(handles, s_.ipc_handle())
        ok, why = 1, ""
        try:
            s_.connect_ipc(handles)
        except pkg.SphbError as e:
            ok, why = 0, str(e)
        flag = torch.tensor([ok], device=f"cuda:{dev}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            return "ipc"
        if ok:
            s_.disconnect_ipc()
        if why:
            transport_note.append(why)
        dist.barrier()
        return "nccl"

    def release(s_):
        if s_.info()["transport"] == 3:
            s_.synchronize()
            s_.disconnect_ipc()       # every rank unmaps its neighbours' blocks before anybody frees its own
            dist.barrier()
        s_.close()
    sim = make()
    transport = connect(sim, ident[0])
    stream = torch.cuda.ExternalStream(sim.stream, device=dev)
    fl_pin = torch.empty(max(n, 1) * 7, dtype=torch.float32).pin_memory()
    fl_host = fl_pin.numpy().view(pkg.PARTICLE)[:n]
    fl_host[:] = part
    sim.upload(fl_host, boundary, id_base=base)
    sim.init_boundary()
    # gravity: constant, or (slosh) one sample per step from the synthetic accelerometer trace; the trace
    # position carries on across warm-up, timed steps and the per-kernel pass
    tilt = spec.get("tilt")
    g_trace = pkg.gravity_trace_tilt(prm, tilt[0], tilt[1], tilt[2], W + 100 + 2 * K + 64) if tilt else None
    g_pos = [0]

    def advance(s_, nsteps):
        if g_trace is None:
            s_.step(nsteps, *G)
        else:
            s_.step_trace(g_trace[g_pos[0]:g_pos[0] + nsteps])
            g_pos[0] += nsteps
    g0 = tuple(float(v) for v in g_trace[0]) if tilt else G
    sim.compute_accel(*g0)
    clocks = ClockSampler(dev)           # from the warm-up on: the timed region alone may be shorter than a sample
    clocks.start()
    advance(sim, W)
    sim.synchronize()
    clocks.wait_first()
    advance(sim, 100)
    sim.synchronize()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=f"cuda:{dev}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: K steps between barriers, CUDA events on the library's stream, max over ranks
    launches0 = sim.launch_count
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    a.record(stream)
    advance(sim, K)
    b.record(stream)
    sim.synchronize()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    gpu_ms = max_over_ranks(a.elapsed_time(b))
    launches = sim.launch_count - launches0
    clk = clocks.stop()
    value = n_total * K / (gpu_ms * 1e-3)

    # ---- per-kernel CUDA-event times on this rank (separate pass)
    cand, acc = sim.pair_stats()
    sim.profile(1); sim.profile_read(reset=True)
    advance(sim, min(K, 20))
    prof = sim.profile_read(reset=True)
    sim.profile(0)
    st = sim.allreduce_stats()
    if st["n_lost"] or st["n_overflow"] or st["n_fluid"] != n_total:
        raise SystemExit(f"slab run lost particles or overflowed its slots (n_fluid {st['n_fluid']} of {n_total}, n_lost {st['n_lost']}, "
                         f"n_overflow {st['n_overflow']}): the line would not describe the configured workload")
    info = sim.info()
    n_local = sim.stats()["n_fluid"]
    hbm_peak, sm_max_mhz, peak_src = read_peaks()
    kern = {}
    for name, d in prof.items():
        if d["launches"]:
            ms = d["ms"] / d["launches"]
            kern[name] = {"ms": round(ms, 5)}
            if name in ALGO_BYTES:
                kern[name]["algo_GBps"] = round(ALGO_BYTES[name] * n_local / (ms * 1e-3) / 1e9, 2)
    # every rank's per-kernel times (ms per launch, rank order): "other" is the binning of the received
    # entries, which on the peer-store transport includes the wait for the neighbours' signals
    all_kern = [None] * world
    dist.all_gather_object(all_kern, {name: d["ms"] for name, d in kern.items()})
    per_rank_ms = {name: [round(k_.get(name, 0.0), 4) for k_ in all_kern] for name in kern}
    force_ms = prof["force"]["ms"] / max(1, prof["force"]["launches"])
    dens_ms = prof["density"]["ms"] / max(1, prof["density"]["launches"])
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    # per-GPU load is the 8M-particle dam-break slab: the committed ncu capture of dam8m is the matching one
    roofline = make_roofline(f"dam{max(1, round(n_total / world / 1e6))}m", n_local, kern, force_ms, dens_ms, cand, acc, hbm_peak, sm_max_mhz, peak_src,
                             sm_count, suffix=" (rank 0)")
    roofline["kernels_per_rank_ms"] = per_rank_ms

    # ---- e2e: host buffers -> C ABI -> host buffers on every rank
    pcap = info["particle_capacity"]
    out_host = torch.empty(pcap * 7, dtype=torch.float32).pin_memory().numpy().view(pkg.PARTICLE)
    ids_host = torch.empty(pcap, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    du_host = torch.empty(pcap, dtype=torch.float32).pin_memory().numpy()
    dv_host = torch.empty(pcap, dtype=torch.float32).pin_memory().numpy()
    sim2 = make()
    ident2 = [pkg.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident2, src=0)
    connect(sim2, ident2[0])
    sim2.upload(fl_host, boundary, id_base=base); sim2.init_boundary(); sim2.compute_accel(*g0); sim2.step(W, *g0); sim2.synchronize()
    trace = np.ascontiguousarray(g_trace[:K] if tilt else np.tile(np.asarray([G], np.float32), (K, 1)), np.float32)
    runs, last, n_out = [], None, 0
    for _rep in range(3):
        barrier()
        t0 = time.perf_counter()
        sim2.upload(fl_host, boundary, id_base=base)
        sim2.init_boundary()
        sim2.compute_accel(*g0)
        st_ = pkg.Stats()
        st_ref, g_addr = ctypes.byref(st_), trace.ctypes.data
        prev = None
        for s_ in range(K):
            t = sim2.step_stats_begin(g_addr + 8 * s_, 1)      # launch step s_, then read step s_ - 1's statistics
            if prev is not None:
                sim2.step_stats_end(prev, st_ref)
            prev = t
        sim2.step_stats_end(prev, st_ref)
        last = st_.asdict()
        n_out = sim2.download_into(out_host, ids_host, du_host, dv_host)
        torch.cuda.synchronize()
        runs.append(max_over_ranks(time.perf_counter() - t0))
    e2e_s = sorted(runs)[1]
    assert n_out > 0 or n == 0
    n_max = int(max_over_ranks(n))
    e2e = {"value": n_total * K / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": round(28 * (n_max + len(boundary)) / K + 8, 1),
           "d2h_bytes_per_step": round(40 * n_max / K + 136, 1), "ms_per_step": round(1e3 * e2e_s / K, 5),
           "ms_per_step_runs": [round(1e3 * r / K, 5) for r in runs],
           "path": "per rank: sphb_mg_upload + sphb_init_boundary + sphb_compute_accel + K x (sphb_step_stats_begin(1 step), sphb_step_stats_end(previous step)) + sphb_mg_download, pinned host buffers; every step's statistics are read on the host, one step behind the launches; bytes are the busiest rank's; median of 3 repetitions",
           "last_step_stats": {"max_speed": last["max_speed"], "max_rho_err": last["max_rho_err"]}}
    release(sim2)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": gpu_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 (f64 at the reference's double sites: pi_sph_fluid.c:325, :332, :334, :616)",
        "data": ("synthetic (filled tank on the reference's lattice idiom, gravity from a synthetic MPU6050 tilt trace mapped as pi_sph_fluid.c:439-440; not a reference scene)"
                                 if tilt else "synthetic (dam-break block on the reference's lattice idiom; not a reference scene)"),
        "config": {"workload": spec["name"], "n_fluid": n_total, "n_boundary": int(len(boundary)), "R": R,
                   "particles_per_gpu": [int(hist[int(cuts[r]):int(cuts[r + 1])].sum()) for r in range(world)],
                   "parallelism": f"x-slabs of cell columns, cuts at particle-count quantiles {[int(c) for c in cuts]}, "
                                  "2 ghost columns, one halo+migration message per neighbour per step "
                                  + ("stored by the advect+bin kernel into the neighbour's receive buffer over NVLink (CUDA IPC peer memory), completed by a device-side signal"
                                     if transport == "ipc" else "over NCCL send/recv"),
                   "transport": transport, **({"transport_fallback": transport_note[0][:200]} if transport_note else {}),
                   "halo_message_bytes": info["message_bytes"], "particle_slots_per_gpu": info["particle_capacity"],
                   "deterministic_order": not args.nondeterministic,
                   "force_arithmetic": ("fast_force = 1" if args.fast_force else
                                        "the reference's own (default): bit-identical with the oracle and with the 1-GPU run"),
                   "gravity": (f"tilt trace: +-{tilt[0]} deg, period {tilt[1]} steps, sample held {tilt[2]} steps, one (gx, gy) per step" if tilt else "constant (0, -9.81)"),
                   "l2": f"not flushed: per-GPU state (~100 B x {n_total / world / 1e6:.0f}M particles) is far larger than the 126 MB L2",
                   "timing": "CUDA events on the library stream around the K steps, barrier + synchronize both sides; max over ranks",
                   "wall_s_timed_region": round(t_wall, 4),
                   "merged_stats": {k2: st[k2] for k2 in ("n_fluid", "n_lost", "n_overflow", "n_escaped", "max_speed", "max_rho_err")},
                   "scaling_note": scaling_note(),
                   "build": pkg.lib().sphb_build_info().decode()},
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
    }
    release(sim)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--nondeterministic", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="tuning runs only: skip the end-to-end leg")
    ap.add_argument("--no-secondary", action="store_true", help="N = 1: skip the secondary entries (configs[1], configs[2], fast_force)")
    ap.add_argument("--fast-force", action="store_true",
                    help="sphb_params.fast_force = 1: the single-precision force arithmetic (outside the 1e-4 parity bar "
                         "at >= 4M particles) instead of the reference's own (default, bit-identical accelerations)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.steps is None:
        args.steps = 100 if world == 1 else 200
    if args.warmup is None:
        args.warmup = 5 if world == 1 else 10
    # every N: BASELINE configs[3], the 64M-particle dam break (strong scaling)
    args.workload = args.workload or "dam64m"
    spec = workload_spec(args.workload)

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = run_reference(spec, args.steps, max(args.warmup, 3))
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "no reference build for this workload"}))
            return 0
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": ("synthetic (the reference's own lattice drop scene)" if spec["scene"] == "drop" else
                         "synthetic (dam-break block on the reference's lattice idiom; bounded sub-block sample)"),
                "config": {"workload": spec["name"], "n_fluid": r["n_fluid"], "R": spec["R"],
                           "threads": r["cores"], "flags": "-Ofast -march=x86-64-v3|v4 -fopenmp (Makefile:2,4; -march pinned for portability)",
                           "same_config_note": ("the reference's main() hard-codes the drop scene and R (pi_sph_fluid.c:11, :484-506), so a dam break "
                                                "runs through the oracle port of its operators (same shipped flags) on a sub-block of the same block at "
                                                "the same spacing; particle-updates/s is a per-particle rate, so the sample size does not enter the ratio"
                                                if r["kind"] == "port" else "the reference's own compiled translation unit on the whole scene")},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
    if world == 1 and "tilt" in spec:
        raise SystemExit("bench.py: the sloshing workload is a slab (multi-GPU) bench line: launch it under torchrun with --gpus 2|4|8")
    line = run_gpu_slabs(args, spec, rank, world) if world > 1 else run_gpu(args, spec, rank, world)
    if rank == 0 and world == 1 and not args.no_e2e and not args.no_secondary:
        line["secondary"] = secondary_entries(args, spec, line)
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            r = run_reference(spec, 2000, 3, budget_s=15.0)
            line["cpu_baseline"] = ({k: r[k] for k in ("value", "unit", "cores", "kind", "sample")} if r else None)
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


def secondary_entries(args, spec, main_line):
    """Named secondary measurements of the N = 1 line, taken in the same process right after the main workload:
    BASELINE configs[1] and configs[2] (unless one of them IS the main workload), and the main workload's force
    pass in fast_force arithmetic — the cost of the reference-arithmetic force pass the headline is measured with."""
    out = {}
    for key, K2 in (("drop256k", 300), ("dam4m", 100)):
        if key == args.workload:
            continue
        sp = workload_spec(key)
        l2 = run_gpu(args, sp, 0, 1, K=K2, W=10, want_e2e=(key == "drop256k"), workload_key=key)
        ent = {"workload": sp["name"], "value": l2["value"], "unit": UNIT, "ms_per_step": l2["ms_per_step"], "steps": K2,
               "n_fluid": l2.get("n_fluid", l2.get("config", {}).get("n_fluid")),
               "kernels_ms": {k_: v["ms"] for k_, v in l2["roofline"]["kernels"].items()}}
        if "e2e" in l2:
            ent["e2e"] = l2["e2e"]
        if key == "drop256k" and not args.no_cpu_baseline:
            r = run_reference(sp, 2000, 3, budget_s=10.0)
            if r:
                ent["cpu_baseline"] = {k_: r[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}
        out[key] = ent
    if not args.fast_force:
        K3 = max(3, min(args.steps, 30))
        lf = run_gpu(args, spec, 0, 1, K=K3, W=3, want_e2e=False, fast_force=True)
        km, kf = main_line["roofline"]["kernels"], lf["roofline"]["kernels"]
        out["fast_force_on_main_workload"] = {
            "workload": spec["name"], "value": lf["value"], "ms_per_step": lf["ms_per_step"], "steps": K3,
            "force_ms": {"reference_arithmetic": km["force"]["ms"], "fast_force": kf["force"]["ms"]},
            "note": "fast_force = 1 is NOT the headline: its accelerations are outside 1e-4*max(|a|, G) at >= 4M particles "
                    "(tests/test_gpu_parity.py); the headline value is measured with the reference's arithmetic"}
    return out


if __name__ == "__main__":
    sys.exit(main())
