"""The multi-GPU query step THROUGH THE C-ABI (mlc_comm_init, mlc_sharded_query_batch[_device],
mlc_sharded_knn_device — NCCL inside the library), one process per GPU under torch.distributed.run, each
rank checked against the CPU oracle. world = 1 runs everywhere (a one-rank communicator exercises the
whole code path incl. the collectives); world = 2 needs two GPUs (gpurun --gpus 2)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [1, 2, 4])
def test_sharded_query_step_through_the_c_abi(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "sharded_nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"SHARDED_OK world={world}" in r.stdout
