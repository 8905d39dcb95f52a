"""What a new calibration costs: the plan-build call of the single-image kernel through the C ABI
(preallocated buffers), host time to enqueue and time until done.  DCB_TRACE_PLAN=1 prints where
the library spends it.  tools/plan_probe.py"""
import ctypes
import os
import sys
import time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import discorpy_b200 as dcb
from discorpy_b200 import _cabi

H = W = 4096
xc, yc = 2050.37, 2040.81
fact = [1.00227490554, -2.99523692178e-05 / 3, 8.99519088e-08 / 9, -1.57066461911e-10 / 27,
        8.08880211618e-14 / 81]
dcb.set_device(0)
src = dcb.DeviceArray((H, W)).fill_synthetic(seed=1)
dst = dcb.DeviceArray((H, W))
fn = _cabi.load().dcb_unwarp_image_backward_f32
opt = _cabi.make_options(1, dcb.BLEND_EXACT, dcb.PATH_AUTO)
sh = ctypes.c_void_p(dcb.current_stream().handle)


def call(model):
    _cabi.check(fn(ctypes.c_void_p(src.ptr), ctypes.c_void_p(dst.ptr), H, W, src.pitch, dst.pitch,
                   ctypes.byref(model), ctypes.byref(opt), sh))


call(_cabi.make_radial(xc, yc, fact))
dcb.synchronize()
for rep in range(4):
    if rep % 2 == 0:
        dcb.plan_cache_clear()
    dcb.synchronize()
    model = _cabi.make_radial(xc + 0.001 * (rep + 1), yc, fact)
    t0 = time.perf_counter()
    call(model)
    t1 = time.perf_counter()
    dcb.synchronize()
    t2 = time.perf_counter()
    call(model)
    dcb.synchronize()
    t3 = time.perf_counter()
    print("new model %d (cache %s): enqueue %.0f us, done after %.0f us; second call %.0f us"
          % (rep, "cleared" if rep % 2 == 0 else "kept", (t1 - t0) * 1e6, (t2 - t0) * 1e6, (t3 - t2) * 1e6))
