#!/usr/bin/env python
"""Device-timed figures for the other rows of the path: perspective (a4), the
combined radial->perspective entry (a5, BASELINE config 3), one unwarped
sinogram (a2) -- inputs resident in HBM, CUDA events."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import discorpy_b200 as dcb                                    # noqa: E402
from discorpy_b200 import _cabi                                # noqa: E402
import discorpy_b200.post.postprocessing as post               # noqa: E402

PEAK = 6451.5


def timed(fn, reps=10):
    fn()
    s = dcb.current_stream()
    a, b = dcb.Event(), dcb.Event()
    a.record(s)
    for _ in range(reps):
        fn()
    b.record(s)
    b.sync()
    return a.elapsed_ms(b) / reps


def main():
    dcb.set_device(0)
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = PEAK
    fact = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
    coef = [1.02, 0.01, -15.0, 0.005, 1.01, -8.0, 8e-6, -5e-6]
    for n in (2048, 4096):
        imgs = [dcb.DeviceArray((n, n)).fill_synthetic(seed=3, offset=i * n * n) for i in range(8)]
        k = [0]

        def nxt():
            k[0] = (k[0] + 1) % len(imgs)
            return imgs[k[0]]
        for blend in ("exact", "lerp32"):
            post.config["blend"] = {"exact": dcb.BLEND_EXACT, "lerp32": dcb.BLEND_LERP32}[blend]
            ms = timed(lambda: post.correct_perspective_image(nxt(), coef))
            print(json.dumps({"row": "a4 perspective", "n": n, "blend": blend, "us": ms * 1e3,
                              "frac": 8.0 * n * n / (ms * 1e-3) / 1e9 / peak}))
            ms = timed(lambda: post.unwarp_image_backward_perspective(
                nxt(), n / 2 + 6.2, n / 2 - 4.4, fact if n == 2048 else [fact[i] / 2.0 ** i for i in range(5)], coef))
            print(json.dumps({"row": "a5 combined (two kernels)", "n": n, "blend": blend, "us": ms * 1e3,
                              "frac_vs_8B_per_px": 8.0 * n * n / (ms * 1e-3) / 1e9 / peak}))
        post.config["blend"] = dcb.BLEND_EXACT
        del imgs
        dcb.device.device_pool.clear()
    # a2: one sinogram of a 256-slice stack of 2560^2 (config 4 shard)
    D, H, W = 128, 2560, 2560
    stack = dcb.DeviceArray((D, H, W)).fill_synthetic(seed=4)
    for index in (0, 1275, 2559):
        ms = timed(lambda: post.unwarp_slice_backward(stack, 1283.4, 1275.9, fact, index), reps=20)
        print(json.dumps({"row": "a2 slice", "D": D, "W": W, "index": index, "us": ms * 1e3,
                          "Mpix_s_out": D * W / 1e6 / (ms * 1e-3)}))


if __name__ == "__main__":
    main()
