"""Multi-GPU slab path on the GPU box.

The decomposition is exact by construction: in deterministic mode every owned particle sees the
same neighbours in the same order as in the single-GPU run, so the N-slab result must equal the
1-GPU result BIT FOR BIT (positions, velocities, rho, p, accelerations), through halo exchange and
migration.  The in-process transport runs any number of slabs on one device, so these tests need
one GPU; so does the cross-process peer-store (CUDA IPC) transport with several processes on device 0.
tests/test_multigpu.py (marker `multigpu`, needs >= 2 GPUs) covers NCCL and IPC across real devices.
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import same_bits

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
FIELDS = ("x", "y", "u", "v", "m", "rho", "p")


def single_gpu(pkg, prm, fluid, boundary, steps, g):
    with pkg.Simulation(prm) as sim:
        sim.upload(fluid, boundary)
        sim.init_boundary()
        sim.compute_accel(*g)
        sim.step(steps, *g)
        f, du, dv = sim.download()
        st = sim.stats()
    return f, du, dv, st


def assert_identical(f, du, dv, rf, rdu, rdv):
    for fld in FIELDS:
        assert same_bits(f[fld], rf[fld]), fld
    assert same_bits(du, rdu) and same_bits(dv, rdv)


@pytest.mark.parametrize("world", [2, 3, 5])
def test_slabs_equal_single_gpu_bit_for_bit_drop(lib_built, world):
    pkg = lib_built
    prm = pkg.default_params(0.02)
    fluid, boundary = pkg.scene_drop(prm), pkg.scene_boundary(prm)
    g = (80.0, -9.81)                    # strong sideways pull: ~1 cell of drift, particles cross the cuts
    steps = 600
    rf, rdu, rdv, rst = single_gpu(pkg, prm, fluid, boundary, steps, g)

    cuts = pkg.plan_cuts(pkg.column_histogram(prm, fluid), world)
    col0 = pkg.columns_of(prm, fluid["x"])
    with pkg.SlabGroup(prm, cuts, halo_capacity=4096) as grp:
        grp.upload(fluid, boundary)
        grp.init_boundary()
        grp.compute_accel(*g)
        grp.step(steps, *g)
        f, du, dv, owner = grp.download()
        st = grp.stats()
    assert (owner >= 0).all()
    assert_identical(f, du, dv, rf, rdu, rdv)
    # ownership follows the column of the final position, and some particles changed slab
    col1 = pkg.columns_of(prm, f["x"])
    assert np.array_equal(owner, np.searchsorted(cuts, col1, side="right") - 1)
    assert (np.searchsorted(cuts, col0, side="right") != np.searchsorted(cuts, col1, side="right")).sum() > 0
    assert st["n_lost"] == 0 and st["n_overflow"] == 0 and st["n_fluid"] == len(fluid)
    assert st["mass"] == pytest.approx(rst["mass"], rel=1e-12)
    assert st["kinetic"] == pytest.approx(rst["kinetic"], rel=1e-9)
    assert st["max_speed"] == rst["max_speed"] and st["max_rho"] == rst["max_rho"] and st["min_rho"] == rst["min_rho"]


def test_slabs_dam_break_with_empty_slabs_and_trace(lib_built):
    """Dam-break block in the left half: quantile cuts leave the last slab wide and empty at t=0;
    gravity comes from a tilt trace (one sample per step)."""
    pkg = lib_built
    R = 0.01
    prm = pkg.default_params(R)
    fluid = pkg.scene_block(prm, 2 * R, 1.0, 2 * R, 0.6)
    boundary = pkg.scene_boundary(prm)
    steps = 300
    trace = pkg.gravity_trace_tilt(prm, 25.0, 200, 10, steps)
    with pkg.Simulation(prm) as sim:
        sim.upload(fluid, boundary); sim.init_boundary(); sim.compute_accel(float(trace[0][0]), float(trace[0][1]))
        sim.step_trace(trace)
        rf, rdu, rdv = sim.download()
    _, cols = pkg.grid_columns(prm)
    # hand-made cuts: two slabs inside the block, one straddling its face, one empty
    cuts = np.array([0, 12, 30, 60, cols], np.int32)
    with pkg.SlabGroup(prm, cuts, halo_capacity=8192) as grp:
        grp.upload(fluid, boundary)
        assert grp.slabs[3].n_fluid == 0
        grp.init_boundary()
        grp.compute_accel(float(trace[0][0]), float(trace[0][1]))
        grp.step_trace(trace)
        f, du, dv, owner = grp.download()
        st = grp.stats()
        info = [s.info() for s in grp.slabs]
    assert_identical(f, du, dv, rf, rdu, rdv)
    assert st["n_lost"] == 0 and st["n_overflow"] == 0
    assert info[1]["window_lo"] == 10 and info[1]["window_hi"] == 32 and info[0]["window_lo"] == 0
    assert all(i["exchanges"] == steps + 1 for i in info)


def test_scene_slab_builder_feeds_slabs_directly(lib_built):
    """Each rank builds only its own part of the block scene (sphb_scene_fill_block_slab): the
    union is the full scene, ids are the full scene's indices."""
    pkg = lib_built
    R = 0.01
    prm = pkg.default_params(R)
    box = (2 * R, 2.0, 2 * R, 0.5)
    full = pkg.scene_block(prm, *box)
    boundary = pkg.scene_boundary(prm)
    cuts = pkg.plan_cuts(pkg.scene_block_column_hist(prm, *box), 4)
    g = (0.0, -9.81)
    rf, rdu, rdv, _ = single_gpu(pkg, prm, full, boundary, 50, g)
    with pkg.SlabGroup(prm, cuts, halo_capacity=8192) as grp:
        n = 0
        for r, s in enumerate(grp.slabs):
            part, base = pkg.scene_block_slab(prm, *box, int(cuts[r]), int(cuts[r + 1]))
            assert np.array_equal(part, full[base:base + len(part)])
            s.upload(part, boundary, id_base=base)
            n += len(part)
        assert n == len(full)
        grp.n_fluid = n
        grp.init_boundary()
        grp.compute_accel(*g)
        grp.step(50, *g)
        f, du, dv, _ = grp.download()
    assert_identical(f, du, dv, rf, rdu, rdv)


def test_rebalance_recuts_at_current_quantiles_and_continues_exactly(lib_built):
    """The fluid drifts sideways, the initial cuts go stale; re-cutting mid-run (state + du/dv move to
    new slab contexts) must not change a single bit of the trajectory."""
    pkg = lib_built
    prm = pkg.default_params(0.02)
    fluid, boundary = pkg.scene_drop(prm), pkg.scene_boundary(prm)
    g = (120.0, -9.81)
    rf, rdu, rdv, _ = single_gpu(pkg, prm, fluid, boundary, 900, g)
    _, cols = pkg.grid_columns(prm)
    cuts0 = np.array([0, 30, 36, cols], np.int32)            # deliberately lopsided
    with pkg.SlabGroup(prm, cuts0, halo_capacity=4096) as grp:
        grp.upload(fluid, boundary)
        grp.init_boundary()
        grp.compute_accel(*g)
        grp.step(500, *g)
        per0 = [s.stats()["n_fluid"] for s in grp.slabs]
        cuts1 = grp.rebalance()
        grp.step(400, *g)
        f, du, dv, owner = grp.download()
        per2 = [s.stats()["n_fluid"] for s in grp.slabs]
    assert_identical(f, du, dv, rf, rdu, rdv)
    assert list(cuts1) != list(cuts0)
    assert max(per0) - min(per0) > max(per2) - min(per2)        # better balanced after the re-cut


def test_slab_overflow_is_reported(lib_built):
    pkg = lib_built
    prm = pkg.default_params(0.02)
    fluid, boundary = pkg.scene_drop(prm), pkg.scene_boundary(prm)
    cuts = pkg.plan_cuts(pkg.column_histogram(prm, fluid), 2)
    with pkg.SlabGroup(prm, cuts, halo_capacity=16) as grp:      # far too small for a 2-column halo
        grp.upload(fluid, boundary)
        grp.init_boundary()
        grp.compute_accel(0.0, -9.81)
        st = grp.stats()
    assert st["n_overflow"] > 0


def test_slab_api_misuse(lib_built):
    pkg = lib_built
    prm = pkg.default_params(0.02)
    _, cols = pkg.grid_columns(prm)
    with pytest.raises(pkg.SphbError):
        pkg.Slab(prm, 0, 2, 0, 3)                 # narrower than 4 columns
    with pytest.raises(pkg.SphbError):
        pkg.Slab(prm, 0, 2, 4, cols)              # rank 0 must start at column 0
    s = pkg.Slab(prm, 0, 2, 0, 40)
    with pytest.raises(pkg.SphbError):
        pkg.Simulation.upload(s, pkg.scene_drop(prm))     # plain upload on a slab context
    fluid = pkg.scene_drop(prm)
    uneven = fluid[:10].copy()
    uneven["m"][3] *= 1.5
    with pytest.raises(pkg.SphbError):
        s.upload(uneven, pkg.scene_boundary(prm), ids=np.arange(10))     # slabs need the reference's uniform mass (:502)
    s.upload(fluid[:10], pkg.scene_boundary(prm), ids=np.arange(10))
    s.init_boundary()
    with pytest.raises(pkg.SphbError):
        s.compute_accel(0.0, -9.81)               # world 2, not connected
    s.close()


@pytest.mark.parametrize("world", [2, 3])
def test_cross_process_peer_store_transport_on_one_device(lib_built, world):
    """The peer-store transport ACROSS PROCESSES on a one-GPU box: `world` processes share device 0, each maps
    its neighbours' receive blocks with cudaIpc* (handles carried by gloo, no NCCL), k_advect_bin stores the
    halo + migration entries into the neighbour process's buffer and k_bin_recv waits for the device-side
    signal — bit-identical to the single-GPU run, with particles migrating between the processes."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29640 + world),
                        str(ROOT / "tests" / "mg_nccl_check.py"), "ipc1dev"],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "mg_ipc1dev_check ok" in r.stdout


def test_state_parts_continue_on_another_rank_count_and_in_the_oracle(lib_built, oracle_built, tmp_path):
    """sphb_mg_save_state / sphb_mg_load_state: a dam break saved by 3 slabs after 60 steps continues on 2 slabs
    with other cuts, on one slab over all columns, and — read by oracle/pyoracle.load_state — in the oracle: each
    of them ends bit-identical to the uninterrupted single-GPU run of 120 steps."""
    pkg = lib_built
    R, g = 0.01, (30.0, -9.81)
    prm = pkg.default_params(R)
    box = (2 * R, 1.5, 2 * R, 0.6)
    fluid, boundary = pkg.scene_block(prm, *box), pkg.scene_boundary(prm)
    rf, rdu, rdv, rst = single_gpu(pkg, prm, fluid, boundary, 120, g)
    hist = pkg.column_histogram(prm, fluid)
    with pkg.SlabGroup(prm, pkg.plan_cuts(hist, 3)) as grp:
        grp.upload(fluid, boundary); grp.init_boundary(); grp.compute_accel(*g); grp.step(60, *g)
        parts = grp.save_state(tmp_path / "dam")
        mid, _, _, _ = grp.download()
    _, cols = pkg.grid_columns(prm)
    for cuts in (pkg.plan_cuts(pkg.column_histogram(prm, mid), 2), np.asarray([0, cols], np.int32)):
        with pkg.SlabGroup(prm, cuts) as grp:
            grp.load_parts(parts)
            assert grp.n_fluid == len(fluid)
            grp.init_boundary()
            grp.step(60, *g)
            f, du, dv, _ = grp.download()
            st = grp.stats()
        assert_identical(f, du, dv, rf, rdu, rdv)
        assert st["steps"] == 120 and st["n_lost"] == 0 and st["n_overflow"] == 0
    # the oracle reads the same parts and continues
    sv = oracle_built.load_state(parts)
    assert sv["steps"] == 60 and len(sv["fluid"]) == len(fluid) and np.float32(sv["R"]) == np.float32(R)
    o = oracle_built.Oracle(R=R, variant="chain")
    of, ob, odu, odv = sv["fluid"].copy(), sv["boundary"].copy(), sv["du"].copy(), sv["dv"].copy()
    gb = o.init_boundary(ob)
    gf = o.grid(len(of))
    o.step(of, ob, gf, gb, odu, odv, 60, *g)
    assert_identical(of, odu, odv, rf, rdu, rdv)
    # a part of another scene is refused
    with pytest.raises(pkg.SphbError):
        with pkg.SlabGroup(pkg.default_params(0.02), np.asarray([0, 77], np.int32)) as grp:
            grp.load_parts(parts)


def test_cross_process_recut_on_one_device(lib_built):
    """sphb_mg_rebalance_host across three processes (device 0, peer-store halo transport, bytes of the re-cut
    carried by gloo): a dam break under a strong sideways pull is re-cut every 40 steps — bit-identical to the
    single-GPU run, and every re-cut that moves the cuts narrows the spread of the ranks' particle counts."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3",
                        "--master-addr", "127.0.0.1", "--master-port", "29647",
                        str(ROOT / "tests" / "mg_nccl_check.py"), "ipc1dev", "recut=40"],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "mg_ipc1dev_check ok" in r.stdout and "re-cuts:" in r.stdout


def test_ipc_transport_single_process_loopback(lib_built):
    """World 1 has no neighbour: connecting the peer-store transport needs no handle and the slab
    steps like a single-GPU run; a handle without a neighbour is refused."""
    pkg = lib_built
    prm = pkg.default_params(0.02)
    fluid, boundary = pkg.scene_drop(prm), pkg.scene_boundary(prm)
    _, cols = pkg.grid_columns(prm)
    rf, rdu, rdv, _ = single_gpu(pkg, prm, fluid, boundary, 50, (0.0, -9.81))
    s = pkg.Slab(prm, 0, 1, 0, cols)
    h = s.ipc_handle()
    assert len(h) == pkg.api.IPC_HANDLE_BYTES and any(h)
    with pytest.raises(pkg.SphbError):
        pkg.api._check(pkg.lib().sphb_mg_connect_ipc(s._h, h, None), "sphb_mg_connect_ipc")
    s.connect_ipc([h])
    assert s.info()["transport"] == 3
    s.upload(fluid, boundary, ids=np.arange(len(fluid), dtype=np.uint32))
    s.init_boundary()
    s.compute_accel(0.0, -9.81)
    s.step(50, 0.0, -9.81)
    ids, f, du, dv = s.download()
    s.disconnect_ipc()
    s.close()
    out = np.zeros(len(fluid), pkg.PARTICLE); odu = np.zeros(len(fluid), np.float32); odv = np.zeros(len(fluid), np.float32)
    out[ids] = f; odu[ids] = du; odv[ids] = dv
    assert_identical(out, odu, odv, rf, rdu, rdv)
