"""Sinogram exchange over N ranks, run under torchrun (used as a launcher only: the
collectives are NCCL through the C ABI, discorpy_b200.multigpu.NcclComm) with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29611 tests/two_rank_sinogram_check.py [--depth 257] [--size 2560] [--time]

Each rank holds `depth / N` device-resident slices of a synthetic stack (the stateless generator, so the
whole stack can be re-created on the host for the oracle).  For several row indices the (depth, W)
sinogram of `unwarp_slice_backward` is assembled on rank 0 twice: by the remap kernels storing straight
into rank 0's buffer (SinogramWindow, peer stores over NVLink) and by the plain NCCL all-gather
(gather_rows).  Both must be bit-identical to the oracle's `unwarp_slice_backward` of the full stack.
--time adds a device-side timing of the two forms (max over ranks).  Test infrastructure: uses oracle/.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import discorpy_b200 as dcb
from discorpy_b200 import multigpu
from discorpy_b200.device import DeviceArray, synthetic_host
from discorpy_b200.post import postprocessing as post
from oracle import oracle_np

ap = argparse.ArgumentParser()
ap.add_argument("--depth", type=int, default=37)      # odd on purpose: shards differ by one slice
ap.add_argument("--size", type=int, default=640)
ap.add_argument("--time", action="store_true")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--no-check", action="store_true", help="timing only (large stacks: no host copy for the oracle)")
args = ap.parse_args()

comm = multigpu.NcclComm.from_env()          # binds LOCAL_RANK's device, NCCL through the C ABI
rank, world = comm.rank, comm.world

D, H, W = args.depth, args.size, args.size
params = multigpu.broadcast_params(
    dict(xcenter=W / 2 + 3.4, ycenter=H / 2 - 4.1,
         list_fact=[1.0, -2e-5 * 2560 / W, 6e-8 * (2560 / W) ** 2, -1e-10 * (2560 / W) ** 3,
                    5e-14 * (2560 / W) ** 4]) if rank == 0 else None)
lo, hi = multigpu.shard_range(D, rank, world)
shard = DeviceArray((hi - lo, H, W)).fill_synthetic(seed=4, offset=lo * H * W)
window = multigpu.SinogramWindow(D, W, owner=0)
indices = [] if args.no_check else [0, H // 3, H // 2, H - 1]
report = dict(rank=rank, world=world, depth=D, size=W, shard=[lo, hi], cases=[])

full_host = None
if rank == 0 and not args.no_check:
    full_host = synthetic_host(D * H * W, seed=4).reshape(D, H, W)

for index in indices:
    # fused: every rank's kernel writes its rows into rank 0's buffer
    multigpu.unwarp_slice_backward_sharded(shard, params, index, window)
    window.fence()
    # collective: local rows, then an NCCL all-gather
    rows = DeviceArray((hi - lo, W))
    post._unwarp_slice_into(shard, params["xcenter"], params["ycenter"], params["list_fact"], index,
                            multigpu._Rows(rows.ptr, rows.pitch, rows.shape))
    gathered = multigpu.gather_rows(rows, D)
    if rank == 0:
        want = oracle_np.unwarp_slice_backward(full_host, params["xcenter"], params["ycenter"],
                                               params["list_fact"], index)
        fused = window.array.to_host()
        coll = gathered.to_host()
        report["cases"].append(dict(index=index,
                                    fused_mismatches=int(np.count_nonzero(fused != want)),
                                    collective_mismatches=int(np.count_nonzero(coll != want))))
    comm.barrier()


def timed(fn, reps):
    fn()
    comm.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    comm.barrier()
    return comm.allreduce_max([(time.perf_counter() - t0) / reps])[0] * 1e3


if args.time:
    index = H // 2
    rows = DeviceArray((hi - lo, W))

    def fused():
        multigpu.unwarp_slice_backward_sharded(shard, params, index, window)
        window.fence()

    def collective():
        post._unwarp_slice_into(shard, params["xcenter"], params["ycenter"], params["list_fact"],
                                index, multigpu._Rows(rows.ptr, rows.pitch, rows.shape))
        multigpu.gather_rows(rows, D)
        dcb.current_stream().sync()

    def local_only():
        post._unwarp_slice_into(shard, params["xcenter"], params["ycenter"], params["list_fact"],
                                index, multigpu._Rows(rows.ptr, rows.pitch, rows.shape))
        dcb.current_stream().sync()

    report["ms"] = dict(fused_peer_stores=timed(fused, args.reps),
                        kernel_plus_nccl_all_gather=timed(collective, args.reps),
                        kernel_only_local_rows=timed(local_only, args.reps),
                        sinogram_bytes=D * W * 4)

window.close()
ok = True
if rank == 0:
    ok = all(c["fused_mismatches"] == 0 and c["collective_mismatches"] == 0 for c in report["cases"])
    report["ok"] = ok
    print(json.dumps(report))
comm.close()
sys.exit(0 if ok else 1)
