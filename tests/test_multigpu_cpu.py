"""N > 1 host logic on CPU: two gloo ranks broadcast the coefficient block, shard a
stack and run the collective form of the optional sinogram exchange:
gather_rows with uneven shards, broadcast_bytes (the IPC-handle exchange of the fused
form; the peer mapping itself needs two GPUs: tests/two_rank_sinogram_check.py)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from discorpy_b200 import multigpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_without_gaps():
    for total in (0, 1, 7, 8, 64, 2048, 2049):
        for world in (1, 2, 3, 4, 8):
            spans = [multigpu.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a0, a1), (b0, b1) in zip(spans[:-1], spans[1:]):
                assert a1 == b0 and a1 >= a0
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        multigpu.shard_range(8, 2, 2)


def test_param_block_roundtrip():
    fact = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
    coef = [1.02, 0.01, -15.0, 0.005, 1.01, -8.0, 8e-6, -5e-6]
    vec = multigpu.pack_params(12.5, -3.25, fact, coef)
    assert vec.dtype == np.float64 and vec.nbytes <= 256
    back = multigpu.unpack_params(vec)
    assert back == dict(xcenter=12.5, ycenter=-3.25, list_fact=fact, list_coef=coef)
    assert multigpu.unpack_params(multigpu.pack_params(1, 2, [1.0]))["list_coef"] == []
    with pytest.raises(ValueError):
        multigpu.pack_params(0, 0, [0.0] * 17)


WORKER = r"""
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
from discorpy_b200 import multigpu
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
params = dict(xcenter=1283.4, ycenter=1275.9,
              list_fact=[1.0, -2e-5, 6e-8, -1e-10, 5e-14]) if rank == 0 else None
got = multigpu.broadcast_params(params, src=0)
lo, hi = multigpu.shard_range(11, rank, world)
# sinogram exchange, collective form: rank r owns rows [lo, hi) of an 11 x 7 array
full = np.arange(77, dtype=np.float32).reshape(11, 7) * 0.5
got_full = multigpu.gather_rows(full[lo:hi].copy(), 11)
assert got_full.dtype == np.float32 and np.array_equal(got_full, full), got_full
try:
    multigpu.gather_rows(full[:1], 11)
    raise SystemExit("gather_rows accepted a block of the wrong size")
except ValueError:
    pass
blob = multigpu.broadcast_bytes(bytes(range(72)) if rank == 1 else b"", 72, src=1)
assert blob == bytes(range(72))
# one file per rank: two ranks writing to the shared stdout pipe can interleave
with open(os.path.join(%(out)r, "rank%%d.json" %% rank), "w") as f:
    json.dump(dict(rank=rank, world=world, params=got, shard=[lo, hi]), f)
dist.barrier()
dist.destroy_process_group()
"""


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_gloo_ranks_broadcast_and_shard(tmp_path):
    import json
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, out=str(tmp_path)))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    for attempt in range(3):            # a rendezvous port can be taken between probe and bind
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", str(_free_port()), str(script)]
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
        if res.returncode == 0:
            break
    assert res.returncode == 0, res.stderr[-2000:]
    rows = [json.loads(f.read_text()) for f in sorted(tmp_path.glob("rank*.json"))]
    assert sorted(r["rank"] for r in rows) == [0, 1]
    want = dict(xcenter=1283.4, ycenter=1275.9,
                list_fact=[1.0, -2e-5, 6e-8, -1e-10, 5e-14], list_coef=[])
    for r in rows:
        assert r["params"] == want        # bit-identical on every rank
    shards = sorted(tuple(r["shard"]) for r in rows)
    assert shards == [(0, 6), (6, 11)]
