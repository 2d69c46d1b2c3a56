"""world_size-2 gloo test of the multi-GPU host logic (scene sharding + result gather): no GPU, no data-path
collective -- exactly what bench.py does with NCCL on the GPU box."""
import os
import socket
import subprocess
import sys

import numpy as np

from cloth_sim_self_collision_b200 import batch as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, os.environ["REPO_ROOT"])
import torch.distributed as dist
from cloth_sim_self_collision_b200 import batch as B
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
for total in (0, 1, 5, 8):
    mine = B.shard_scenes(total, world, rank)
    vals = [B.state_checksum(np.full(7, float(s))) for s in mine]   # stands for the per-scene result of the step loop
    allv = B.gather_scene_values(dist, vals, total)
    expect = [B.state_checksum(np.full(7, float(s))) for s in range(total)]
    assert np.array_equal(allv, np.array(expect)), (rank, total, allv, expect)
t = B.max_over_ranks(dist, [10.0 + rank, 5.0 - rank])
assert t == [10.0 + world - 1, 5.0], t
dist.barrier()
dist.destroy_process_group()
print("rank %d ok" % rank)
"""


def test_shard_scenes_partitions_everything():
    for total in (0, 1, 7, 64):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                sh = B.shard_scenes(total, world, r)
                seen += list(sh)
                assert len(sh) in (total // world, total // world + 1)
            assert seen == list(range(total))
    assert B.scene_gravity((0.0, 0.05, -10.0), 3) == (0.0, 0.05 + 0.001 * 3, -10.0)


def test_gather_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, REPO_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2",
               OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o
