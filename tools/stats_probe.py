"""Fraction of tile rows the single-image kernel sends through the patch path, per blend / order,
on the BASELINE config-2 geometry (and others): tools/stats_probe.py [cfg ...]"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import discorpy_b200 as dcb
from discorpy_b200 import _cabi
import discorpy_b200.post.postprocessing as post

CFG = {
    "2": (4096, 4096, 2050.37, 2040.81, [1.00227490554, -2.99523692178e-05 / 3, 8.99519088e-08 / 9,
                                          -1.57066461911e-10 / 27, 8.08880211618e-14 / 81]),
    "3": (2048, 2048, 1030.2, 1019.6, [1.0, -2e-5, 6e-8, -1e-10, 5e-14]),
    "4": (2560, 2560, 1283.4, 1275.9, [1.0, -2e-5, 6e-8, -1e-10, 5e-14]),
    "1": (2160, 2560, 588.692801577, 462.092631791,
          [1.00227490554, -2.99523692178e-05, 8.99519088e-08, -1.57066461911e-10, 8.08880211618e-14]),
    "5": (8192, 8192, 4100.3, 4090.8, [1.0, -1e-5, 3e-8, -2e-11, 5e-15, -8e-19, 6e-23, -2e-27, 3e-32]),
}
dcb.set_device(0)
for name in (sys.argv[1:] or ["2"]):
    H, W, xc, yc, fact = CFG[name]
    rng = np.random.default_rng(1)
    mat = dcb.DeviceArray.from_host(rng.random((H, W), dtype=np.float32))
    for order in (1, 0):
        dcb.plan_cache_clear()
        dcb.image_stats(True, reset=True)
        out = post.unwarp_image_backward(mat, xc, yc, fact, order=order)
        dcb.synchronize()
        st = dcb.image_stats(True, reset=True)
        tot = max(1, st["rows"])
        print("cfg %s order %d: rows verified in full %.1f %%, in part %.1f %%, blend redo %d, odd tiles %d  %s"
              % (name, order, 100.0 * st["rows_patch"] / tot, 100.0 * st["rows_partial"] / tot,
                 st["rows_blend_redo"], st["tiles_odd"], dcb.last_plan()))
    dcb.image_stats(False)
