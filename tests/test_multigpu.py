"""Transports of the slab path across REAL devices (>= 2 GPUs).  Not part of `pytest -m gpu` (the driver's box
has one GPU): run as `pytest -m multigpu tests/test_multigpu.py` under `gpurun --gpus 2`; the log of that run
is committed under profiles/.  The cross-process peer-store transport itself is also covered on one device by
tests/test_gpu_slabs.py::test_cross_process_peer_store_transport_on_one_device."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.multigpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("world,transport", [(2, "nccl"), (2, "ipc")])
def test_process_per_gpu_transports_under_torchrun(lib_built, world, transport):
    """One process per GPU (needs 2 GPUs: `pytest -m multigpu` under `gpurun --gpus 2`; the log of that run is
    committed under profiles/): halo + migration over ncclSend/ncclRecv, and as peer stores into the
    neighbour's receive buffer (CUDA IPC) completed by a device-side signal — both bit-identical to
    the single-GPU run."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29631" if transport == "nccl" else "29633",
                        str(ROOT / "tests" / "mg_nccl_check.py"), transport],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"mg_{transport}_check ok" in r.stdout




@pytest.mark.parametrize("world", [2])
def test_recut_over_nccl_under_torchrun(lib_built, world):
    """sphb_mg_rebalance (ncclAllReduce of the column counts, ncclSend/ncclRecv of the particles) every 100 steps of
    a dam break that drains to the right: bit-identical to the single GPU, spread of the ranks' counts shrinks."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29635",
                        str(ROOT / "tests" / "mg_nccl_check.py"), "ipc", "recut=100"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "mg_ipc_check ok" in r.stdout and "re-cuts:" in r.stdout


@pytest.mark.parametrize("workload,steps", [("dam64m", 5), ("slosh16m", 20)])
def test_full_size_configs_bit_identical_to_one_gpu(lib_built, workload, steps):
    """BASELINE configs[3] (64M dam break) and configs[4] (16M sloshing tank, tilt trace) on every visible GPU
    (8 on the bench box) against ONE GPU at the same size: per-field 64-bit hashes of x, y, u, v, m, rho, p,
    du_dt, dv_dt equal, plus an oracle spot check of a strip straddling a cut (tests/mg_full_check.py; the
    8-GPU log is committed as profiles/r02_mg_full_check_8gpu.log)."""
    import torch
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29637",
                        str(ROOT / "tests" / "mg_full_check.py"), workload, str(steps)],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "mg_full_check ok" in r.stdout and "EQUAL" in r.stdout and "bit-for-bit" in r.stdout
