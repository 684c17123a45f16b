"""GPU, >= 2 devices: the row-block partitioned path (NCCL all-gather of the direction's rows + all-reduce of the
scalar packs) against the single-GPU path on the same problem.  Skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,graph,halo", [(2, "random", "0"), (2, "random", "1"), (2, "torus", None), (4, "random", None),
                                              (4, "torus", None), (8, "random", None)])
def test_partitioned_run_matches_single_gpu(built, world, graph, halo):
    """exchange modes: all-gather of every row (LORADS_HALO=0), halo exchange of the referenced rows only (=1), or the
    library's own choice (unset)"""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, LORADS_TEST_GRAPH=graph)
    env.pop("LORADS_HALO", None)
    if halo is not None:
        env["LORADS_HALO"] = halo
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(29740 + world), os.path.join(ROOT, "tests", "_multi_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0 and "MULTI_GPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
