#!/usr/bin/env python
"""Colour-frame row: util.unwarp_color_image_backward on a
2160 x 2560 x 3 uint8 frame with BASELINE config 1's coefficients -- end to end
through the public API with host buffers, next to the reference's code path
(oracle restatement: NumPy coordinates + one map_coordinates per channel) on
one host core, which is how the reference runs it."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import discorpy_b200 as dcb                                    # noqa: E402
from discorpy_b200.util import utility as util                 # noqa: E402

XC, YC = 588.692801577 * 2, 462.092631791 * 2.3
FACT = [1.00227490554, -2.99523692178e-05 / 2, 8.99519088e-08 / 4,
        -1.57066461911e-10 / 8, 8.08880211618e-14 / 16]


def main():
    dcb.set_device(0)
    rng = np.random.default_rng(1)
    frame = rng.integers(0, 256, (2160, 2560, 3), dtype=np.uint8)
    pinned = dcb.pinned_copy(frame)
    res = {}
    for name, arr in (("pageable", frame), ("pinned", pinned)):
        for _ in range(2):
            out = util.unwarp_color_image_backward(arr, XC, YC, FACT)
        ts = []
        for _ in range(8):
            t0 = time.perf_counter()
            out = util.unwarp_color_image_backward(arr, XC, YC, FACT)
            ts.append(time.perf_counter() - t0)
        res["gpu_e2e_ms_" + name] = 1e3 * min(ts)
    mpix = frame.shape[0] * frame.shape[1] * frame.shape[2] / 1e6
    res["gpu_e2e_Msamples_s_pinned"] = mpix / (res["gpu_e2e_ms_pinned"] * 1e-3)
    from oracle import oracle_np as orc
    t0 = time.perf_counter()
    want = orc.unwarp_color_image_backward(frame, XC, YC, FACT)
    res["cpu_reference_path_ms_1core"] = 1e3 * (time.perf_counter() - t0)
    res["cpu_Msamples_s_1core"] = mpix / (res["cpu_reference_path_ms_1core"] * 1e-3)
    res["pixels_differing_from_oracle"] = int(np.count_nonzero(out != want))
    res["frame"] = "2160x2560x3 uint8, 5-term model"
    print(json.dumps(res))


if __name__ == "__main__":
    main()
