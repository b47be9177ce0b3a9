"""Host-side emulation of the slab protocol (test infrastructure).

One rank of a world_size-N `gloo` job: it owns the particles of its cell columns, and every step
(1) advances them (kick + drift, the oracle's arithmetic), (2) sends to each neighbour every owned
particle that now lies within two columns of the cut on either side — the rule of
k_advect_bin<.., SLAB> in csrc/kernels_build.cu — (3) keeps what lies in its window, and
(4) evaluates density / pressure / accelerations with the ORACLE on owned + ghost particles.  If
the protocol is right, the owned results equal the oracle's single-domain results bit for bit.
No GPU is involved: this checks the decomposition logic the CUDA path implements."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def columns(x, cell, cols):
    c = ((x - np.float32(0.0)) / np.float32(cell)).astype(np.int32)      # :112, float32 divide, truncation
    return np.clip(c, 0, cols - 1)


def run_rank(rank, world, port, R, steps, g, out_dir):
    import torch.distributed as dist
    from oracle import pyoracle
    import pi_sph_fluid_b200 as pkg

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = pyoracle.Oracle(R=R, variant="chain", threads=2)
    prm = pkg.default_params(R)
    _, cols = pkg.grid_columns(prm)
    fluid = o.scene_drop()
    boundary = o.scene_boundary()
    gb = o.init_boundary(boundary)

    # every rank histograms a share of the scene; the all-reduced histogram gives identical cuts
    import torch
    share = fluid[rank::world]
    hist = torch.from_numpy(pkg.column_histogram(prm, np.ascontiguousarray(share)).astype(np.int64))
    dist.all_reduce(hist)
    cuts = pkg.plan_cuts(hist.numpy().astype(np.uint64), world)
    all_cuts = [None] * world
    dist.all_gather_object(all_cuts, cuts.tolist())
    assert all(c == all_cuts[0] for c in all_cuts)
    lo, hi = int(cuts[rank]), int(cuts[rank + 1])
    win_lo, win_hi = (lo - 2 if rank > 0 else 0), (hi + 2 if rank < world - 1 else cols)

    ids = np.arange(len(fluid), dtype=np.int64)
    col = columns(fluid["x"], o.cell, cols)
    mine = (col >= lo) & (col < hi)
    own_ids, own = ids[mine], fluid[mine].copy()
    own_du = np.zeros(len(own), np.float32); own_dv = np.zeros(len(own), np.float32)
    half_dt = 0.5 * np.float64(o.dt)
    migrated = 0

    for step in range(steps + 1):
        if step > 0:      # :615-624 — kick in double, drift in float
            own["u"] = (own["u"].astype(np.float64) + half_dt * own_du.astype(np.float64)).astype(np.float32)
            own["v"] = (own["v"].astype(np.float64) + half_dt * own_dv.astype(np.float64)).astype(np.float32)
            own["x"] = own["x"] + o.dt * own["u"]
            own["y"] = own["y"] + o.dt * own["v"]
        col = columns(own["x"], o.cell, cols)
        msgs = {}
        if rank > 0:
            sel = col < lo + 2
            msgs[rank - 1] = (own_ids[sel], own[sel])
        if rank < world - 1:
            sel = col >= hi - 2
            msgs[rank + 1] = (own_ids[sel], own[sel])
        got = []
        for peer in (rank - 1, rank + 1):       # ordered pairwise exchange: lower rank sends first
            if peer < 0 or peer >= world:
                continue
            box = [None]
            if rank < peer:
                dist.send_object_list([msgs[peer]], dst=peer)
                dist.recv_object_list(box, src=peer)
            else:
                dist.recv_object_list(box, src=peer)
                dist.send_object_list([msgs[peer]], dst=peer)
            got.append(box[0])
        keep = (col >= win_lo) & (col < win_hi)
        loc_ids = [own_ids[keep]] + [i for i, _ in got]
        loc = [own[keep]] + [p for _, p in got]
        loc_ids = np.concatenate(loc_ids); loc = np.concatenate(loc)
        lcol = columns(loc["x"], o.cell, cols)
        inwin = (lcol >= win_lo) & (lcol < win_hi)
        loc_ids, loc, lcol = loc_ids[inwin], loc[inwin], lcol[inwin]
        order = np.argsort(loc_ids, kind="stable")          # in-cell order = ascending global id
        loc_ids, loc, lcol = loc_ids[order], np.ascontiguousarray(loc[order]), lcol[order]
        assert len(np.unique(loc_ids)) == len(loc_ids)
        gf = o.grid(len(loc))
        du, dv = o.compute_accel(loc, boundary, gf, gb, *g)
        owned = (lcol >= lo) & (lcol < hi)
        migrated += int((~np.isin(loc_ids[owned], own_ids)).sum())
        own_ids, own = loc_ids[owned], loc[owned].copy()
        own_du, own_dv = du[owned], dv[owned]
        if step > 0:     # :637-640
            own["u"] = (own["u"].astype(np.float64) + half_dt * own_du.astype(np.float64)).astype(np.float32)
            own["v"] = (own["v"].astype(np.float64) + half_dt * own_dv.astype(np.float64)).astype(np.float32)
    np.savez(Path(out_dir) / f"rank{rank}.npz", ids=own_ids, fluid=own, du=own_du, dv=own_dv, migrated=migrated,
             cuts=cuts)
    dist.destroy_process_group()


if __name__ == "__main__":
    run_rank(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5]),
             (float(sys.argv[6]), float(sys.argv[7])), sys.argv[8])
