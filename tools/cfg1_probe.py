"""Kernel time and patch-path coverage of the single-image kernel on the BASELINE config-1 geometry
(2160 x 2560, data/coef_dot_05.txt: a calibration whose corrected image has clipped borders) and on
configs 3 / 4 as single images."""
import os, sys, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import discorpy_b200 as dcb
from discorpy_b200 import _cabi
import discorpy_b200.post.postprocessing as post
CFG = {
    "1": (2160, 2560, 1252.18562214, 1008.91135307,
          [1.00015076e+00, -1.6217848e-05, 2.29152003e-08, -1.68852813e-11, 1.73565939e-15]),
    "3": (2048, 2048, 1030.2, 1019.6, [1.0, -2e-5, 6e-8, -1e-10, 5e-14]),
    "4": (2560, 2560, 1283.4, 1275.9, [1.0, -2e-5, 6e-8, -1e-10, 5e-14]),
}
dcb.set_device(0)
s = dcb.current_stream()
for name, (H, W, xc, yc, fact) in CFG.items():
    srcs = [dcb.DeviceArray((H, W)).fill_synthetic(seed=i) for i in range(12)]
    dsts = [dcb.DeviceArray((H, W)) for _ in range(12)]
    dcb.plan_cache_clear()
    dcb.image_stats(True, reset=True)
    post.unwarp_image_backward(srcs[0], xc, yc, fact)
    dcb.synchronize()
    st = dcb.image_stats(False, reset=True)
    model = _cabi.make_radial(xc, yc, fact); opt = _cabi.make_options(1)
    def once():
        for a, b in zip(srcs, dsts):
            _cabi.call("dcb_unwarp_image_backward_f32", ctypes.c_void_p(a.ptr), ctypes.c_void_p(b.ptr), H, W, a.pitch, b.pitch,
                       ctypes.byref(model), ctypes.byref(opt), ctypes.c_void_p(s.handle))
    once(); s.sync()
    e0, e1 = dcb.Event(), dcb.Event()
    e0.record(s)
    for _ in range(5): once()
    e1.record(s); e1.sync()
    us = e0.elapsed_ms(e1) * 1e3 / 60
    tot = max(1, st["rows"])
    print("cfg %s %dx%d: %.1f us per image (%.2f us per Mpixel; config 2 runs at 2.59); rows on the patch path %.1f %%, partial %.1f %%, tile not eligible %.1f %%, none verified %.1f %%  %s"
          % (name, H, W, us, us / (H * W / 1e6), 100.0 * st["rows_patch"] / tot, 100.0 * st["rows_partial"] / tot,
             100.0 * st["rows_tile_not_eligible"] / tot, 100.0 * st["rows_no_segment_verified"] / tot, dcb.last_plan()), flush=True)
