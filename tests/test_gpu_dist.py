"""The same replicated/sharded logic on real GPUs: torchrun, NCCL, world size =
number of visible GPUs (skipped below 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
import drrt_b200
from drrt_b200 import dist as ddist, synth
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank()
rep = ddist.ReplicatedKDTree(drrt_b200.KDTree(device=local), src=0)
pos = synth.box_points(60000, 301)
rep.KDInsert(pos if rank == 0 else None)
q = synth.box_points(1001, 302)
off, ids, dst = rep.KDFindWithinRange(1.5, q)
rep.SetObstacles(synth.obstacles(64, 303) if rank == 0 else None)
s, e = synth.random_edges(5001, 304)
v, ln = rep.ExplicitEdgeCheck(s, e)
if rank == 0:
    one = drrt_b200.KDTree(device=0); one.KDInsert(pos); one.SetObstacles(**synth.obstacles(64, 303))
    o2, i2, d2 = one.KDFindWithinRange(1.5, q)
    v2, l2 = one.ExplicitEdgeCheck(s, e)
    assert np.array_equal(off, o2) and np.array_equal(ids, i2) and np.array_equal(dst, d2)
    assert np.array_equal(v, v2) and np.array_equal(ln, l2)
    print("DIST_OK", dist.get_world_size())
dist.destroy_process_group()
'''


def test_nccl_replicated_sharded(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          "--nproc-per-node", str(min(n, 8)), "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "DIST_OK" in out.stdout
