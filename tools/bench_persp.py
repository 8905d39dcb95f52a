#!/usr/bin/env python
"""Device-timed single-image kernels through the C ABI (no Python between launches):
perspective (BASELINE config 3 coefficients), radial config 2, and radial config 1
geometry (19.9 % of the coordinates clipped) at 4096^2 / 2048^2."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import discorpy_b200 as dcb
from discorpy_b200 import _cabi
dcb.set_device(0)
COEF = [1.02, 0.01, -15.0, 0.005, 1.01, -8.0, 8e-6, -5e-6]
C5 = [1.00227490554, -2.99523692178e-05, 8.99519088e-08, -1.57066461911e-10, 8.08880211618e-14]
lib = _cabi.load()
stream = dcb.current_stream()
sh = ctypes.c_void_p(stream.handle)
for n in (4096, 2048):
    srcs = [dcb.DeviceArray((n, n)).fill_synthetic(seed=3, offset=i * n * n) for i in range(8)]
    dsts = [dcb.DeviceArray((n, n)) for _ in range(8)]
    persp = _cabi.make_persp(COEF)
    rad2 = _cabi.make_radial(n / 2 + 2.37, n / 2 - 7.19, [C5[i] / (3.0 * n / 4096) ** i for i in range(5)])
    rad1 = _cabi.make_radial(588.692801577 * n / 2560, 462.092631791 * n / 2560, [C5[i] / (n / 2560) ** i for i in range(5)])
    for blend in ("exact", "lerp32"):
        opt = _cabi.make_options(1, {"exact": dcb.BLEND_EXACT, "lerp32": dcb.BLEND_LERP32}[blend], dcb.PATH_AUTO)
        def run_p(i):
            _cabi.check(lib.dcb_correct_perspective_image_f32(ctypes.c_void_p(srcs[i].ptr), ctypes.c_void_p(dsts[i].ptr), n, n, srcs[i].pitch, dsts[i].pitch, ctypes.byref(persp), ctypes.byref(opt), sh))
        def run_r(model):
            def f(i):
                _cabi.check(lib.dcb_unwarp_image_backward_f32(ctypes.c_void_p(srcs[i].ptr), ctypes.c_void_p(dsts[i].ptr), n, n, srcs[i].pitch, dsts[i].pitch, ctypes.byref(model), ctypes.byref(opt), sh))
            return f
        for name, fn in (("perspective cfg3", run_p), ("radial cfg2", run_r(rad2)), ("radial cfg1 geometry (clipped)", run_r(rad1))):
            for i in range(8):
                fn(i)
            e0, e1 = dcb.Event(), dcb.Event()
            e0.record(stream)
            for rep in range(5):
                for i in range(8):
                    fn(i)
            e1.record(stream)
            e1.sync()
            us = e0.elapsed_ms(e1) * 1e3 / 40
            print(json.dumps(dict(kernel=name, n=n, blend=blend, us=round(us, 2), us_per_4096sq=round(us * (4096.0 / n) ** 2, 2), lib=os.path.basename(_cabi.LIB_PATH))), flush=True)
