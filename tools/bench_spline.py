#!/usr/bin/env python
"""Device-timed cost of the spline path (dcb_spline_prefilter + dcb_spline_remap) on a
4096 x 4096 float32 image, 5-term radial model, per order / mode; with --cpu the reference's
own code path for the same call (NumPy float64 coordinates + scipy.ndimage.map_coordinates,
postprocessing.py:138-147) is timed beside it on one host core.
Usage: bench_spline.py [--size N] [--cpu]"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import discorpy_b200 as dcb                                    # noqa: E402
from discorpy_b200 import _cabi                                # noqa: E402

COEF_DOT_05 = [1.00227490554, -2.99523692178e-05, 8.99519088e-08,
               -1.57066461911e-10, 8.08880211618e-14]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    n = args.size
    dcb.set_device(0)
    fact = [COEF_DOT_05[i] / 3.0 ** i for i in range(5)]
    model = _cabi.make_radial(n / 2 + 2.37, n / 2 - 7.19, fact)
    src = dcb.DeviceArray((n, n)).fill_synthetic(seed=2)
    dst = dcb.DeviceArray((n, n))
    stream = dcb.current_stream()
    sh = ctypes.c_void_p(stream.handle)
    for order, mode in ((3, "reflect"), (3, "nearest"), (2, "reflect"), (5, "reflect"), (1, "reflect")):
        need = ctypes.c_size_t()
        mc = _cabi.MODES[mode]
        _cabi.call("dcb_spline_workspace_bytes", n, n, order, mc, ctypes.byref(need))
        work = dcb.device.device_pool.take(need.value)

        def pre():
            _cabi.call("dcb_spline_prefilter", ctypes.c_void_p(src.ptr), 0, n, n, src.pitch, order,
                       mc, ctypes.c_void_p(work.ptr), need.value, sh)

        def remap():
            _cabi.call("dcb_spline_remap", ctypes.c_void_p(work.ptr), n, n, order, mc,
                       ctypes.c_void_p(dst.ptr), 0, dst.pitch, _cabi.MAP_RADIAL, ctypes.byref(model),
                       None, None, None, 0, 0, None, 0, 0.0, 0.0, sh)
        res = {}
        for name, fn in (("prefilter", pre), ("remap", remap)):
            fn()
            e0, e1 = dcb.Event(), dcb.Event()
            e0.record(stream)
            for _ in range(args.reps):
                fn()
            e1.record(stream)
            e1.sync()
            res[name + "_ms"] = e0.elapsed_ms(e1) / args.reps
        row = dict(size=n, order=order, mode=mode, workspace_MB=need.value / 1e6, **res)
        row["total_ms"] = row["prefilter_ms"] + row["remap_ms"]
        row["Mpix_s"] = n * n / 1e6 / (row["total_ms"] * 1e-3)
        if args.cpu and order == 3 and mode == "reflect":
            from scipy.ndimage import map_coordinates
            mat = src.to_host()
            t0 = time.perf_counter()
            xu, yu = np.meshgrid(np.arange(n) - model.xc, np.arange(n) - model.yc)
            ru = np.sqrt(xu ** 2 + yu ** 2)
            fmat = np.sum(np.asarray([a * ru ** i for i, a in enumerate(fact)]), axis=0)
            xd = np.float32(np.clip(model.xc + fmat * xu, 0, n - 1))
            yd = np.float32(np.clip(model.yc + fmat * yu, 0, n - 1))
            map_coordinates(mat, (np.reshape(yd, (-1, 1)), np.reshape(xd, (-1, 1))), order=3, mode=mode)
            row["cpu_reference_path_s_1core"] = time.perf_counter() - t0
        print(json.dumps(row), flush=True)
        dcb.device.device_pool.give(work)


if __name__ == "__main__":
    main()
