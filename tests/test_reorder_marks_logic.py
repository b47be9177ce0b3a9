"""The invariant behind the deterministic reorder's shortcut (csrc/kernels_build.cu: cell_touch, cell_start_prev),
restated in numpy and checked against a plain sort — no GPU, no library.

Claim: let the input be sorted by (cell, id).  Particles move; every particle whose cell changed (or that left the
window, or that arrives from outside) marks its old and its new cell.  Then a particle of an UNMARKED cell belongs
at  start_new[cell] + (slot - start_prev[cell])  in the order sorted by (new cell, id) — the reference's list order
inside a cell is ascending original index, pi_sph_fluid.c:110-123 — and only the marked cells need ranking by id.
"""
import numpy as np
import pytest


def build_with_marks(cell_prev, ids, cell_new, keep, arrivals_cell, arrivals_id, ncells):
    """One grid build the way the kernels do it.  cell_prev/ids: the previous sorted order; cell_new: each slot's
    cell after the drift; keep[s] False: the slot leaves (slab ghost dropped / left the window); arrivals: slots
    appended behind the resident ones (k_bin_recv)."""
    start_prev = np.searchsorted(cell_prev, np.arange(ncells + 1))
    touched = np.zeros(ncells, bool)
    moved = keep & (cell_new != cell_prev)
    touched[cell_prev[moved]] = True                 # k_advect_bin: both cells of a particle that changed cell
    touched[cell_new[moved]] = True
    touched[cell_prev[~keep]] = True                 # ... or left the window
    touched[arrivals_cell] = True                    # k_bin_recv
    key = np.concatenate([np.where(keep, cell_new, -1), arrivals_cell])
    all_ids = np.concatenate([ids, arrivals_id])
    live = key >= 0
    count = np.bincount(key[live], minlength=ncells)
    start = np.concatenate([[0], np.cumsum(count)])
    dst = np.full(len(key), -1)
    for s in np.nonzero(live)[0]:
        c = key[s]
        if not touched[c]:
            dst[s] = start[c] + (s - start_prev[c])                      # k_reorder, unmarked cell
        else:
            same = live & (key == c)
            dst[s] = start[c] + int((all_ids[same] < all_ids[s]).sum())   # ranking by id (k_scatter_ids + k_reorder)
    out_cell = np.full(int(live.sum()), -1)
    out_id = np.full(int(live.sum()), -1)
    assert len(np.unique(dst[live])) == live.sum()                        # a permutation of the kept slots
    out_cell[dst[live]] = key[live]
    out_id[dst[live]] = all_ids[live]
    return out_cell, out_id, touched


@pytest.mark.parametrize("seed", range(6))
def test_unmarked_cells_keep_their_previous_order(seed):
    rng = np.random.default_rng(seed)
    rows, cols = 12, 17
    ncells = rows * cols
    n = 1500
    cell = rng.integers(0, ncells, n)
    ids = rng.permutation(n)
    order = np.lexsort((ids, cell))
    cell, ids = cell[order], ids[order]
    next_id = n
    for step in range(8):
        # ~4 % of the particles hop to a neighbouring cell, a few leave, a few arrive
        hop = rng.random(len(cell)) < 0.04
        r, c = cell // cols, cell % cols
        r2 = np.clip(r + np.where(hop, rng.integers(-1, 2, len(cell)), 0), 0, rows - 1)
        c2 = np.clip(c + np.where(hop, rng.integers(-1, 2, len(cell)), 0), 0, cols - 1)
        cell_new = r2 * cols + c2
        keep = rng.random(len(cell)) > 0.01
        n_arr = int(rng.integers(0, 12))
        arr_cell = rng.integers(0, ncells, n_arr)
        # arrivals carry ids on either side of the residents' (their cells are ranked by id, so any id works)
        arr_id = np.where(rng.random(n_arr) < 0.5, np.arange(next_id, next_id + n_arr), -1 - np.arange(next_id, next_id + n_arr))
        next_id += n_arr
        out_cell, out_id, touched = build_with_marks(cell, ids, cell_new, keep, arr_cell, arr_id, ncells)
        want_cell = np.concatenate([cell_new[keep], arr_cell])
        want_id = np.concatenate([ids[keep], arr_id])
        o = np.lexsort((want_id, want_cell))
        assert np.array_equal(out_cell, want_cell[o]) and np.array_equal(out_id, want_id[o])
        assert 0 < touched.sum() < ncells                                 # both paths were exercised
        cell, ids = out_cell, out_id


def test_nothing_moves_nothing_is_marked():
    cell = np.repeat(np.arange(20), 5)
    ids = np.arange(100)
    out_cell, out_id, touched = build_with_marks(cell, ids, cell.copy(), np.ones(100, bool), np.zeros(0, int), np.zeros(0, int), 20)
    assert not touched.any() and np.array_equal(out_cell, cell) and np.array_equal(out_id, ids)
