#!/usr/bin/env python
"""Row-major vs Morton (Z-order) cell ordering for the pair kernels' staging — the locality experiment the
reference lists as not implemented (README.md:175-176), SURVEY.md 8f-4.  CPU only (host scene builder + numpy).

The pair kernels give a CTA a CHUNK of 128 consecutive particles of the sorted arrays and stage the particles of
every cell in the 3x3 neighbourhoods of the chunk's cells.  What an ordering decides is (a) how many particles a
chunk stages per particle it owns (shared-memory / L2 read amplification) and (b) in how many contiguous runs of
the sorted arrays they lie (bulk copies to issue, and whether the reference's 3x3 walk — rows outer, columns
inner, :136-137 — is three contiguous runs per particle).  Both are counted here for the same scene under both
orderings; nothing is timed.   usage: python scripts/morton_experiment.py [R] > profiles/r02_morton_experiment.json"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import pi_sph_fluid_b200 as pkg  # noqa: E402

PT, TILE_CAP = 128, 624


def part1by1(v):
    v = v.astype(np.uint64) & np.uint64(0xFFFF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF)
    v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F)
    v = (v | (v << np.uint64(2))) & np.uint64(0x33333333)
    v = (v | (v << np.uint64(1))) & np.uint64(0x55555555)
    return v


def analyse(order_key, row, col, rows, cols, label):
    """order_key: per-particle sort key of the ordering (cell-granular).  Returns the statistics of its chunks."""
    perm = np.argsort(order_key, kind="stable")
    r_s, c_s = row[perm], col[perm]
    # rank of each CELL in the ordering and its population -> position of every cell's run in the sorted arrays
    cell = r_s.astype(np.int64) * cols + c_s
    uniq, first, count = np.unique(cell, return_index=True, return_counts=True)
    start = dict(zip(uniq.tolist(), first.tolist()))
    cnt = dict(zip(uniq.tolist(), count.tolist()))
    n = len(perm)
    staged, runs, over = [], [], 0
    for s0 in range(0, n, PT * 37):          # every 37th chunk: a few thousand samples
        cells = np.unique(cell[s0:s0 + PT])
        need = set()
        for cc in cells.tolist():
            rr, c0 = divmod(cc, cols)
            for dr in (-1, 0, 1):
                for dc in (-1, 0, 1):
                    r2, c2 = rr + dr, c0 + dc
                    if 0 <= r2 < rows and 0 <= c2 < cols and (r2 * cols + c2) in start:
                        need.add(r2 * cols + c2)
        segs = sorted((start[q], start[q] + cnt[q]) for q in need)
        total = sum(b - a for a, b in segs)
        nruns, end = 0, -1
        for a, b in segs:                    # cells adjacent in the sorted arrays merge into one run
            if a != end:
                nruns += 1
            end = b
        staged.append(total / min(PT, n - s0))
        runs.append(nruns)
        over += total > TILE_CAP
    return {"ordering": label, "chunks_sampled": len(staged),
            "staged_particles_per_owned_particle": round(float(np.mean(staged)), 3),
            "contiguous_runs_per_chunk_mean": round(float(np.mean(runs)), 2), "runs_per_chunk_max": int(max(runs)),
            "chunks_over_the_624_entry_tile_pct": round(100.0 * over / len(staged), 2)}


def main():
    R = float(sys.argv[1]) if len(sys.argv) > 1 else 0.001
    prm = pkg.default_params(R)
    fluid = pkg.scene_block(prm, 2 * R, 2.0, 2 * R, 0.5)          # the dam-break block of configs[2], at this R
    rows, cols = pkg.grid_columns(prm)
    cell = np.float32(prm.cell_length)
    row = np.clip((fluid["y"] / cell).astype(np.int32), 0, rows - 1)
    col = np.clip((fluid["x"] / cell).astype(np.int32), 0, cols - 1)
    out = {"scene": f"dam-break block 2 m x 0.5 m, R = {R}: {len(fluid)} particles, grid {rows} x {cols} cells, "
                    f"{len(fluid) / len(np.unique(row.astype(np.int64) * cols + col)):.2f} particles per occupied cell",
           "chunk": PT, "tile_capacity": TILE_CAP,
           "row_major": analyse(row.astype(np.int64) * cols + col, row, col, rows, cols, "row-major (shipped)"),
           "morton": analyse((part1by1(col) | (part1by1(row) << np.uint64(1))).astype(np.int64), row, col, rows, cols, "Morton / Z-order")}
    rm, mo = out["row_major"], out["morton"]
    out["reading"] = (
        f"Z-order makes a chunk a compact 2-D block, so it stages {mo['staged_particles_per_owned_particle']} particles per owned particle "
        f"against {rm['staged_particles_per_owned_particle']} for a row-major chunk (one strip of cells + the strips above and below) — "
        f"less shared-memory fill per CTA — but they lie in {mo['contiguous_runs_per_chunk_mean']} contiguous runs on average "
        f"(max {mo['runs_per_chunk_max']}) instead of {rm['contiguous_runs_per_chunk_mean']}: every run is one bulk copy with its own 16-byte alignment slack and "
        "its own cell_start window, and — decisive for parity — a particle's 3x3 walk is no longer three contiguous runs "
        "visited in the reference's order (rows outer, columns inner, :136-137), so the per-thread candidate loop would "
        "chase up to nine runs.  The staged tiles are served from L2 either way (DRAM traffic of k_density is the "
        "algorithmic 16 B/particle + the list hand-over, profiles/r02*_traffic.json), so the fill saved is not HBM "
        "traffic.  Row-major stays.")
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
