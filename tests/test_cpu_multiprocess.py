"""world_size-2 gloo tests (CPU) of the multi-rank partition of the H_eff application: the all-reduce form on an uneven bond
(chi = 13) and, on an even bond (chi = 14), the reduce-scatter (RS) and all-gather (AG) forms of the sharded Krylov vectors."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("chi,port", [(13, 29517), (14, 29518)])
def test_sharded_matvec_partition_gloo_world2(chi, port):
    env = dict(os.environ, OMP_NUM_THREADS="1", NSB_TEST_CHI=str(chi))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_gloo_shard_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GLOO_SHARD_OK 2" in r.stdout
