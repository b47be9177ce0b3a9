"""The deterministic reorder's shortcut for cells nobody left or entered (csrc/kernels_build.cu: cell_touch,
cell_start_prev).  Production uses it for sets of >= 2^20 slots — the full-size tests
(4M dam break, 64M / 16M slab checks) run it there.  Here the small scenes of the parity suite run WITH the
marks (SPHB_TOUCH_MIN_SLOTS=0 in the environment while the sets are uploaded) and must give the same bits
as the oracle, as one GPU, and as the id pass for every cell: the order inside a cell is the reference's
list order (:110-123), so any slip shows up in the sums of rho and du_dt, dv_dt.
"""
import numpy as np
import pytest

import test_gpu_parity as P
import test_gpu_slabs as S
from conftest import G, same_bits

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def marks_for_every_set(monkeypatch):
    monkeypatch.setenv("SPHB_TOUCH_MIN_SLOTS", "0")


def test_marks_are_used_and_equal_the_id_pass(lib_built, monkeypatch):
    """600 steps of the drop with a sideways pull (particles change cells all the time): the run with the marks
    equals the run that ranks every cell by id, bit for bit, and the hook says which one ran."""
    pkg = lib_built
    prm = pkg.default_params(0.02)
    fluid, boundary = pkg.scene_drop(prm), pkg.scene_boundary(prm)
    g = (40.0, -9.81)
    out = {}
    for tag, slots in (("marks", "0"), ("ids", str(1 << 30))):
        monkeypatch.setenv("SPHB_TOUCH_MIN_SLOTS", slots)
        with pkg.Simulation(prm) as sim:
            sim.upload(fluid, boundary)
            sim.init_boundary()
            sim.compute_accel(*g)
            sim.step(600, *g)
            out[tag] = sim.download() + (sim.reorder_marks(),)
    assert out["marks"][3] == 600 and out["ids"][3] == 0       # the first build (unsorted input) never uses them
    for fld in P.FIELDS:
        assert same_bits(out["marks"][0][fld], out["ids"][0][fld]), fld
    assert same_bits(out["marks"][1], out["ids"][1]) and same_bits(out["marks"][2], out["ids"][2])
    assert not np.array_equal(out["marks"][0]["x"], fluid["x"])


def test_multi_step_against_oracle_with_marks(oracle_built, lib_built, golden075):
    P.test_multi_step_against_oracle(oracle_built, lib_built, golden075)


def test_handed_over_lists_with_marks(oracle_built, lib_built, golden075, golden02):
    P.test_handed_over_lists_are_the_reference_neighbour_lists(oracle_built, lib_built, golden075, golden02)


def test_dam_break_scene_with_marks(oracle_built, lib_built):
    P.test_dam_break_scene_steps(oracle_built, lib_built)


def test_crowded_and_sparse_cells_with_marks(oracle_built, lib_built):
    P.test_edge_crowded_cell_flushes_the_neighbour_list(oracle_built, lib_built)
    P.test_edge_sparse_scene_uses_unstaged_tiles(oracle_built, lib_built)


def test_escaped_particles_with_marks(lib_built, golden075):
    P.test_edge_escaped_particles_are_clamped_and_counted(lib_built, golden075)


@pytest.mark.parametrize("world", [3])
def test_slabs_with_marks(lib_built, world):
    S.test_slabs_equal_single_gpu_bit_for_bit_drop(lib_built, world)
    S.test_slabs_dam_break_with_empty_slabs_and_trace(lib_built)


def test_cross_process_slabs_and_recut_with_marks(lib_built):
    S.test_cross_process_peer_store_transport_on_one_device(lib_built, 3)
    S.test_cross_process_recut_on_one_device(lib_built)


def test_state_parts_with_marks(lib_built, oracle_built, tmp_path):
    S.test_state_parts_continue_on_another_rank_count_and_in_the_oracle(lib_built, oracle_built, tmp_path)
