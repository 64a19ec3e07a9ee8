"""Multi-GPU parity on hardware (needs >= 2 visible GPUs; skipped otherwise): the sharded CPT fit over NCCL is
bit-identical to a single-GPU recount of the concatenated shards -- for the in-place first reduction and for the delta
reduction of a second call -- and row-sharded queries equal the unsharded run."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_fit_and_queries_over_nccl(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, {torch.cuda.device_count()} visible")
    script = os.path.join(ROOT, "tests", "_nccl_worker.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29540 + world), script, str(tmp_path)],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-4000:]
    for r in range(world):
        assert (tmp_path / f"ok_{r}").exists()
