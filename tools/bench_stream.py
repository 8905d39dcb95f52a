#!/usr/bin/env python
"""End-to-end throughput of the streaming Z-stack entry (host stack in, host stack out,
both PCIe directions inside the timed region): pinned source (direct DMA) and pageable
source (staged through pinned buffers).  Usage: bench_stream.py [--depth D] [--size N]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import discorpy_b200 as dcb                                    # noqa: E402
from discorpy_b200.post import streaming                       # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=96)
    ap.add_argument("--size", type=int, default=2560)
    ap.add_argument("--block", type=int, default=8)
    args = ap.parse_args()
    d, n = args.depth, args.size
    dcb.set_device(0)
    fact = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
    xc, yc = n / 2 + 3.4, n / 2 - 4.1
    rng = np.random.default_rng(0)
    one = rng.random((n, n), dtype=np.float32)
    for kind in ("pinned", "pageable"):
        if kind == "pinned":
            src = dcb.pinned_empty((d, n, n), np.float32)
            dst = dcb.pinned_empty((d, n, n), np.float32)
        else:
            src = np.empty((d, n, n), np.float32)
            dst = np.empty((d, n, n), np.float32)
        src[:] = one
        for spb in (args.block, 2 * args.block):
            streaming.unwarp_chunk_slices_backward_stream(src[:2 * spb], xc, yc, fact, out=dst[:2 * spb],
                                                          slices_per_block=spb)      # warm-up
            t0 = time.perf_counter()
            streaming.unwarp_chunk_slices_backward_stream(src, xc, yc, fact, out=dst, slices_per_block=spb)
            dt = time.perf_counter() - t0
            print(json.dumps(dict(source=kind, depth=d, size=n, slices_per_block=spb, seconds=dt,
                                  Mpix_s=d * n * n / 1e6 / dt, GBs_each_way=d * n * n * 4 / 1e9 / dt)), flush=True)
        del src, dst


if __name__ == "__main__":
    main()
