"""Multi-GPU parity (needs >= 2 visible GPUs; skipped otherwise): spawns torchrun on tests/dist_gpu_check.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_partitioned_ppr_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist_gpu_check.py"), "16"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0 and "DIST CHECK PASS" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


def test_golden_graphs_through_the_row_partitioned_filters():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    for env in ({}, {"PGB_HSELL_BLOCK_COLS": "64", "PGB_HSELL_BLOCKS": "4", "PGB_HSELL_MIN_ENTRIES": "4"}):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
               "127.0.0.1", "--master-port", "29535", os.path.join(ROOT, "tests", "dist_golden_check.py")]
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, **env))
        assert res.returncode == 0 and "DIST GOLDEN PASS" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
