"""Parity of the slab path AT FULL SIZE (BASELINE configs[3] and [4]).  Run as

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/mg_full_check.py [workload] [steps]

workload: dam64m (default, configs[3]) | slosh16m (configs[4], tilt trace) | any bench.py workload name.
Every rank steps its slab (peer-store transport), then the ranks' owned particles are reduced to one 64-bit
hash per field (sum over particles of a mix of global id and the field's bits — independent of which rank owns
what); rank 0 then runs the WHOLE scene on its one GPU and hashes the same fields.  Equal hashes for x, y, u, v,
m, rho, p, du_dt, dv_dt = the N-rank run is bit-identical to the single GPU at this size.  Finally the oracle
(CPU restatement of the reference) recomputes rho, p and the accelerations of the particles in a strip of cell
columns straddling a cut from the final positions and velocities: bit-for-bit.
Prints "mg_full_check ok"."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import pi_sph_fluid_b200 as pkg  # noqa: E402
from bench import workload_spec  # noqa: E402

FIELDS = ("x", "y", "u", "v", "m", "rho", "p")
G = (0.0, -9.81)


def field_hashes(ids, f, du, dv):
    """{field: sum_i mix(id_i, bits_i) mod 2^64} + the particle count"""
    key = (ids.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
    out = {}
    with np.errstate(over="ignore"):
        for name, arr in [(k, f[k]) for k in FIELDS] + [("du_dt", du), ("dv_dt", dv)]:
            bits = np.ascontiguousarray(arr).view(np.uint32).astype(np.uint64)
            h = (key ^ (bits * np.uint64(0xC2B2AE3D27D4EB4F))) * np.uint64(0x165667B19E3779F9)
            h ^= h >> np.uint64(29)
            out[name] = int(h.sum(dtype=np.uint64))
    out["count"] = len(ids)
    return out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    name = sys.argv[1] if len(sys.argv) > 1 else "dam64m"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dev = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    spec = workload_spec(name)
    R = spec["R"]
    prm = pkg.default_params(R, device=dev)
    box = spec["box"] if "box" in spec else (2 * R, spec["block"][0], 2 * R, spec["block"][1])
    hist = pkg.scene_block_column_hist(prm, *box)
    rows, cols = pkg.grid_columns(prm)
    cuts = pkg.plan_cuts(hist, world, column_cost=pkg.api.CELL_COST * rows)
    boundary = pkg.scene_boundary(prm)
    part, base = pkg.scene_block_slab(prm, *box, int(cuts[rank]), int(cuts[rank + 1]))
    tilt = spec.get("tilt")
    trace = pkg.gravity_trace_tilt(prm, tilt[0], tilt[1], tilt[2], steps + 1) if tilt else np.tile(np.asarray([G], np.float32), (steps + 1, 1))
    g0, g_last = tuple(map(float, trace[0])), tuple(map(float, trace[steps]))
    halo = int(hist[max(int(cuts[rank]) - 2, 0):int(cuts[rank]) + 2].sum() + hist[int(cuts[rank + 1]) - 2:int(cuts[rank + 1]) + 2].sum())
    cap_t = torch.tensor([max(8192, 2 * halo)], device=f"cuda:{dev}")
    dist.all_reduce(cap_t, op=dist.ReduceOp.MAX)
    slab = pkg.Slab(prm, rank, world, int(cuts[rank]), int(cuts[rank + 1]), halo_capacity=int(cap_t.item()))
    ident = [pkg.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    slab.connect_nccl(ident[0])
    handles = [None] * world
    dist.all_gather_object(handles, slab.ipc_handle())
    slab.connect_ipc(handles)
    t0 = time.perf_counter()
    slab.upload(part, boundary, id_base=base)
    slab.init_boundary()
    slab.compute_accel(*g0)
    slab.step_trace(trace[:steps])
    slab.compute_accel(*g_last)          # rho, p, du_dt, dv_dt of the FINAL positions and velocities
    ids, f, du, dv = slab.download()
    st = slab.allreduce_stats()
    t_slabs = time.perf_counter() - t0
    h = field_hashes(ids, f, du, dv)
    names = list(h)
    t = torch.tensor([np.int64(np.uint64(h[k])) for k in names], dtype=torch.int64, device=f"cuda:{dev}")
    dist.all_reduce(t)                   # int64 addition wraps: sums mod 2^64
    merged = {k: int(np.uint64(np.int64(v))) for k, v in zip(names, t.tolist())}
    slab.synchronize()
    slab.disconnect_ipc()
    dist.barrier()
    slab.close()
    del part, f, du, dv, ids
    ok = True
    if rank == 0:
        full = pkg.scene_block(prm, *box)
        t0 = time.perf_counter()
        with pkg.Simulation(prm) as sim:
            sim.upload(full, boundary); sim.init_boundary(); sim.compute_accel(*g0)
            sim.step_trace(trace[:steps])
            sim.compute_accel(*g_last)
            rf, rdu, rdv = sim.download()
            rb = sim.download_boundary()
        t_one = time.perf_counter() - t0
        ref = field_hashes(np.arange(len(full), dtype=np.uint32), rf, rdu, rdv)
        bad = [k for k in names if merged[k] != ref[k]]
        ok = not bad and st["n_lost"] == 0 and st["n_overflow"] == 0 and st["n_fluid"] == len(full)
        print(f"{spec['name']}: {len(full)} particles, {steps} steps, {world} ranks ({t_slabs:.1f} s) vs one GPU ({t_one:.1f} s): "
              f"hashes of {len(names) - 1} fields + count {'EQUAL' if not bad else 'DIFFER in ' + str(bad)}; cuts {[int(c) for c in cuts]}")
        # ---- oracle spot check: a strip of 20 cell columns around the middle cut, all rows
        from oracle import pyoracle
        pyoracle.build(ref=False)
        c = int(cuts[max(1, world // 2)])
        col = np.clip(((rf["x"] - np.float32(prm.x_min)) / np.float32(prm.cell_length)).astype(np.int32), 0, cols - 1)
        sel = np.nonzero((col >= c - 10) & (col < c + 10))[0]
        inner = (col[sel] >= c - 6) & (col[sel] < c + 6)
        o = pyoracle.Oracle(R=R, variant="chain", max_neighbors=128)
        strip = rf[sel].copy()
        strip["rho"] = 0; strip["p"] = 0
        ob = rb.copy()
        gb = o.grid(len(ob)); o.grid_update(gb, ob)
        gf = o.grid(len(strip))
        t0 = time.perf_counter()
        odu, odv = o.compute_accel(strip, ob, gf, gb, *g_last)
        same = lambda a, b: bool(np.array_equal(np.ascontiguousarray(a).view("u4"), np.ascontiguousarray(b).view("u4")))
        o_ok = (same(strip["rho"][inner], rf["rho"][sel][inner]) and same(strip["p"][inner], rf["p"][sel][inner])
                and same(odu[inner], rdu[sel][inner]) and same(odv[inner], rdv[sel][inner]))
        print(f"oracle spot check: strip of columns [{c - 10}, {c + 10}) around cut {c}: {len(sel)} particles ({int(inner.sum())} compared, "
              f"{time.perf_counter() - t0:.1f} s): rho, p, du_dt, dv_dt {'bit-for-bit' if o_ok else 'DIFFER'}; "
              f"p max {float(rf['p'][sel].max()):.3g} Pa, |a| max {float(np.hypot(rdu[sel], rdv[sel]).max()):.3g}")
        ok = ok and o_ok and int(inner.sum()) > 1000
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{dev}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print("mg_full_check ok")


if __name__ == "__main__":
    main()
