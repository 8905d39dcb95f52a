"""N > 1 host logic on CPU: two gloo ranks (tests/gloo_comm.py stands in for the NCCL
communicator of the C ABI) broadcast the coefficient block, shard a stack and run the
collective form of the optional sinogram exchange: gather_rows with uneven
shards, broadcast_bytes (the IPC-handle exchange of the fused form; the peer mapping itself
needs two GPUs: tests/two_rank_sinogram_check.py), and the unique-id rendezvous of NcclComm."""
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
sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np
from discorpy_b200 import multigpu
from gloo_comm import GlooComm
comm = multigpu.set_default_comm(GlooComm())
rank, world = comm.rank, comm.world
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
assert comm.allreduce_max([float(rank), 5.0 - rank]) == [float(world - 1), 5.0]
# one file per rank: two ranks writing to the shared stdout pipe can interleave
with open(os.path.join(%(out)r, "rank%%d.json" %% rank), "w") as f:
    json.dump(dict(rank=rank, world=world, params=got, shard=[lo, hi]), f)
comm.barrier()
comm.close()
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


def test_unique_id_rendezvous_over_tcp():
    """NcclComm.from_env hands the 128-byte NCCL unique id from rank 0 to the others over a TCP
    connection on MASTER_ADDR; here three 'ranks' as threads with a fake id maker."""
    import threading
    port = _free_port()
    uid = bytes(range(128))
    got = {}

    def run(rank):
        got[rank] = multigpu._exchange_unique_id(rank, 3, lambda: uid, "127.0.0.1", port, timeout=30)

    threads = [threading.Thread(target=run, args=(r,)) for r in (2, 1, 0)]   # clients first
    for t in threads:
        t.start()
    for t in threads:
        t.join(60)
    assert got == {0: uid, 1: uid, 2: uid}


def test_package_does_not_import_torch():
    import re
    pkg = os.path.join(ROOT, "discorpy_b200")
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if name.endswith(".py"):
                text = open(os.path.join(dirpath, name)).read()
                assert not re.search(r"^\s*(import|from)\s+torch", text, re.M), name
