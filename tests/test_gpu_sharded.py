"""GPU run of the partitioned-table / partitioned-CSR path.  With one visible GPU this exercises the
real kernels (gs_bucket_by_owner, gs_gather_rows, gs_sample_csr) with world = 1; with >= 2 GPUs it also
launches the 2-rank NCCL check (tests/multigpu_sharded_check.py) under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "tests", "multigpu_sharded_check.py")


def test_sharded_path_world1():
    env = dict(os.environ, WORLD_SIZE="1", RANK="0", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, SCRIPT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "SHARDED-CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_sharded_path_two_ranks_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", SCRIPT],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "SHARDED-CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
