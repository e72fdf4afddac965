"""Multi-GPU product path: compute_opacity sharded over NCCL ranks (needs >= 2 GPUs; the
single-GPU test box skips it, the gloo world-2 tests of test_host.py cover the host logic)."""
import os
import subprocess
import sys

import pytest

import helpers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2])
def test_sharded_compute_opacity_nccl(world, tmp_path):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29700 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(helpers.ROOT, "tests", "sharded_table_worker.py"), str(tmp_path)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    for r in range(world):
        assert f"OK rank {r}" in res.stdout
