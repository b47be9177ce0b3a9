"""The one-process-per-GPU transports of the slab path.  Run as

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/mg_nccl_check.py [nccl|ipc]

Every rank builds its slab of the scene, the ranks step together through sphb_step (halo +
migration over ncclSend/ncclRecv, or — "ipc" — stored by the advect+bin kernel straight into the
neighbour's receive buffer over NVLink and completed by a device-side signal), rank 0 gathers the
owned particles and compares them BIT FOR BIT with a single-GPU run of the same scene.
"ipc1dev": the same peer-store transport with every rank on CUDA device 0 — N processes sharing ONE GPU,
receive blocks mapped across processes with cudaIpc*, handles carried by gloo, no NCCL anywhere — so the
cross-process transport is testable on a one-GPU box.
A second argument `recut=K` re-cuts the running slabs every K steps (sphb_mg_rebalance over NCCL, or — "ipc1dev" —
sphb_mg_rebalance_host with the bytes carried by gloo): the run must stay bit-identical to the single GPU and the
spread of the ranks' particle counts must shrink.
Prints "mg_nccl_check ok" / "mg_ipc_check ok" / "mg_ipc1dev_check ok"."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import pi_sph_fluid_b200 as pkg  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    transport = sys.argv[1] if len(sys.argv) > 1 else "nccl"
    recut = int(sys.argv[2].split("=")[1]) if len(sys.argv) > 2 and sys.argv[2].startswith("recut=") else 0
    one_dev = transport == "ipc1dev"
    dev = 0 if one_dev else int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(dev)
    if one_dev:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    # strong sideways pull: particles cross the cuts (counted below).  Processes sharing one device are
    # time-sliced, and every step waits for the neighbour's signal: fewer steps there.
    R, steps, g = 0.01, (150 if one_dev else 600), ((2000.0, -9.81) if one_dev else (200.0, -9.81))
    prm = pkg.default_params(R, device=dev)
    box = (2 * R, 1.5, 2 * R, 0.6)
    cuts = pkg.plan_cuts(pkg.scene_block_column_hist(prm, *box), world)
    boundary = pkg.scene_boundary(prm)
    part, base = pkg.scene_block_slab(prm, *box, int(cuts[rank]), int(cuts[rank + 1]))

    slab = pkg.Slab(prm, rank, world, int(cuts[rank]), int(cuts[rank + 1]), halo_capacity=16384)
    if not one_dev:
        ident = [pkg.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        slab.connect_nccl(ident[0])
    if transport in ("ipc", "ipc1dev"):
        handles = [None] * world
        dist.all_gather_object(handles, slab.ipc_handle())
        slab.connect_ipc(handles)
        assert slab.info()["transport"] == 3
    slab.upload(part, boundary, id_base=base)
    slab.init_boundary()
    slab.compute_accel(*g)
    spread = []
    if recut:
        def counts():
            per = [None] * world
            dist.all_gather_object(per, slab.stats()["n_fluid"])
            return per
        done, n_recuts = 0, 0
        while done < steps:
            k = min(recut, steps - done)
            slab.step(k, *g)
            done += k
            if done < steps:
                before = counts()
                changed = slab.rebalance_host(dist) if one_dev else slab.rebalance()
                after = counts()
                n_recuts += int(changed)
                spread.append((max(before) - min(before), max(after) - min(after), changed))
                assert sum(after) == sum(before)
        if rank == 0:
            print(f"re-cuts: {n_recuts} of {len(spread)} calls moved the cuts; (max-min) owned before -> after: {spread}")
    else:
        slab.step(steps, *g)
    ids, f, du, dv = slab.download()
    if one_dev:
        per = [None] * world
        dist.all_gather_object(per, slab.stats())
        st = dict(per[0])
        for q in per[1:]:
            for key in ("n_fluid", "n_lost", "n_overflow", "kinetic"):
                st[key] += q[key]
            st["max_speed"] = max(st["max_speed"], q["max_speed"])
    else:
        st = slab.allreduce_stats()
    info = slab.info()
    gathered = [None] * world
    dist.all_gather_object(gathered, (ids, f, du, dv, base, len(part)))
    ok = True
    if rank == 0:
        full = pkg.scene_block(prm, *box)
        with pkg.Simulation(prm) as sim:
            sim.upload(full, boundary); sim.init_boundary(); sim.compute_accel(*g); sim.step(steps, *g)
            rf, rdu, rdv = sim.download()
            rst = sim.stats()
        out = np.zeros(len(full), pkg.PARTICLE); odu = np.zeros(len(full), np.float32); odv = np.zeros(len(full), np.float32)
        seen = np.zeros(len(full), np.int32)
        migrated = 0
        for i, ff, a, b, base_r, n_r in gathered:
            out[i] = ff; odu[i] = a; odv[i] = b; seen[i] += 1
            migrated += int(((i < base_r) | (i >= base_r + n_r)).sum())      # owned now, uploaded elsewhere
        ok = bool((seen == 1).all())
        if not ok:
            print(f"ownership: {int((seen == 0).sum())} particles owned by nobody, {int((seen > 1).sum())} by several ranks")
        for fld in out.dtype.names:
            same = out[fld].view("u4") == rf[fld].view("u4")
            if not same.all():
                print(f"field {fld}: {int((~same).sum())} of {len(same)} particles differ, first ids {np.nonzero(~same)[0][:8]}")
            ok &= bool(same.all())
        ok &= bool(np.array_equal(odu.view("u4"), rdu.view("u4")) and np.array_equal(odv.view("u4"), rdv.view("u4")))
        ok &= st["n_fluid"] == len(full) and st["n_lost"] == 0 and st["n_overflow"] == 0
        ok &= abs(st["kinetic"] - rst["kinetic"]) <= 1e-9 * abs(rst["kinetic"]) and st["max_speed"] == rst["max_speed"]
        moved = sum(int(((pkg.columns_of(prm, ff["x"]) < cuts[r]) | (pkg.columns_of(prm, ff["x"]) >= cuts[r + 1])).sum())
                    for r, (i, ff, a, b, _b, _n) in enumerate(gathered))
        # (with re-cuts the cuts may follow the flow so well that every particle stays with its first owner)
        ok &= migrated > 0 or world == 1 or recut > 0
        if recut:
            ok &= any(sp[2] for sp in spread)          # at least one call moved the cuts
        print(f"world {world} ({transport}): {len(full)} particles, {steps} steps, migrated {migrated}, owned-out-of-slab {moved}, "
              f"message {info['message_bytes']} B, sent {info['bytes_sent']} B, identical={ok}")
    flag = torch.tensor([1 if ok else 0]) if one_dev else torch.tensor([1 if ok else 0], device=f"cuda:{dev}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if transport in ("ipc", "ipc1dev"):
        slab.disconnect_ipc()          # unmap the neighbours' blocks on every rank before anybody frees its own
        dist.barrier()
    slab.close()
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print(f"mg_{transport}_check ok")


if __name__ == "__main__":
    main()
