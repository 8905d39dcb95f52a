"""Host-side ceiling of the end-to-end figure at N GPUs: every rank moves pinned 64 MiB blocks
host -> device and device -> host at once on two streams, no kernels, all ranks at the same time.
The sum over ranks is what this host can move between pinned memory and N devices; the e2e line
of bench.py cannot beat it.  Launch like bench.py:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29544 tools/pcie_probe_multi.py        (torchrun is only the launcher)
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import discorpy_b200 as dcb                                    # noqa: E402
from discorpy_b200 import _cabi, multigpu                      # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
comm = multigpu.NcclComm.from_env() if world > 1 else None
rank = comm.rank if comm else 0
if comm is None:
    dcb.set_device(0)
cores = dcb.bind_host_to_device(int(os.environ.get("LOCAL_RANK", "0")))
NB, REPS = 64 << 20, 12
h_in = dcb.pinned_empty((NB // 4,), np.float32)
h_out = dcb.pinned_empty((NB // 4,), np.float32)
h_in[:] = 1.0
d_a, d_b = dcb.DeviceArray((1, NB // 4)), dcb.DeviceArray((1, NB // 4))
s1, s2 = dcb.Stream(), dcb.Stream()


def h2d():
    _cabi.call("dcb_h2d", ctypes.c_void_p(d_a.ptr), ctypes.c_void_p(h_in.ctypes.data), NB,
               ctypes.c_void_p(s1.handle))


def d2h():
    _cabi.call("dcb_d2h", ctypes.c_void_p(h_out.ctypes.data), ctypes.c_void_p(d_b.ptr), NB,
               ctypes.c_void_p(s2.handle))


def timed(fns):
    for f in fns:
        f()
    dcb.synchronize()
    if comm:
        comm.barrier()
    t0 = time.perf_counter()
    for _ in range(REPS):
        for f in fns:
            f()
    dcb.synchronize()
    dt = time.perf_counter() - t0
    if comm:
        dt = comm.allreduce_max([dt])[0]
        comm.barrier()
    return world * REPS * NB / dt / 1e9           # GB/s summed over ranks, per direction


res = {"n_gpus": world, "bytes_per_copy": NB,
       "h2d_gbs_sum": timed([h2d]), "d2h_gbs_sum": timed([d2h]),
       "duplex_gbs_each_way_sum": timed([h2d, d2h]),
       "host_cores_bound_per_rank": len(cores) if cores else None}
# one 4096^2 float32 image is 64 MiB each way
res["e2e_ceiling_mpix_s"] = res["duplex_gbs_each_way_sum"] * 1e9 / 4 / 1e6
if rank == 0:
    print(json.dumps(res), flush=True)
if comm:
    comm.close()
