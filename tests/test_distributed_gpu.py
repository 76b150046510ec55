"""Multi-GPU parity on hardware: R-rank sharded retrieval and loss (NCCL) against the single-GPU
result on the concatenated inputs -- index- and value-exact for the top-k, 1e-3 for the rest.

The checks themselves live in ``tests/dist_gpu_check.py`` (one process per GPU); these tests launch
it with ``torch.distributed.run`` for every world size the box offers and keep the log under
``gpurun_out/`` (copied to ``profiles/`` when committed).  On a single-GPU box they skip."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_paths_equal_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    out = r.stdout + r.stderr
    logdir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(logdir):
        with open(os.path.join(logdir, f"dist_check_{world}gpu.log"), "w") as f:
            f.write(out)
    assert r.returncode == 0 and "DIST_CHECK PASS" in out, out[-4000:]
