"""N>1 host logic on CPU: world_size-2 gloo run of chain sharding, broadcast and gathers."""
import os
import subprocess
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from rfsurfhmc_b200 import distributed as D
dist.init_process_group("gloo")
rank, ws = D.world()
N = 7
ids = D.shard_chains(N)
dobs = np.arange(5.0) if rank == 0 else np.zeros(5)
dobs = D.bcast_array(dobs)
assert np.array_equal(dobs, np.arange(5.0))
local = np.stack([np.full(3, float(i)) for i in ids]) if len(ids) else np.zeros((0, 3))
allv = D.gather_chains(local, N)
assert allv.shape == (N, 3) and np.array_equal(allv[:, 0], np.arange(N)), allv
t = D.max_over_ranks(10.0 + rank)
assert t == 10.0 + ws - 1
if rank == 0:
    print("GLOO_OK", ids.tolist())
dist.destroy_process_group()
'''


def test_shard_chains_partition():
    from rfsurfhmc_b200.distributed import shard_chains
    for n, ws in ((16384, 8), (7, 2), (3, 4), (1, 1)):
        parts = [shard_chains(n, r, ws) for r in range(ws)]
        allids = np.concatenate(parts)
        assert np.array_equal(allids, np.arange(n))
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO_OK [0, 1, 2, 3]" in out.stdout
