"""CPU, gloo, world_size 2: host-side logic of the N > 1 (row-block partitioned) path."""
import os
import subprocess
import sys

from conftest import ROOT


def test_partition_logic_two_ranks_gloo(built):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "_partition_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "PARTITION_OK" in out.stdout


def test_partition_rows_edge_cases(built):
    lb = built
    assert lb.partition_rows(10, 1, 0) == (0, 10, 10)
    assert lb.partition_rows(10, 4, 3) == (9, 10, 3)      # ragged last block
    assert lb.partition_rows(3, 8, 7) == (3, 3, 1)        # more ranks than rows: empty blocks at the end
    import pytest
    with pytest.raises(lb.LoradsError):
        lb.partition_rows(10, 2, 2)
