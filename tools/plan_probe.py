"""What a new calibration costs: the plan-build call of the single-image kernel, timed on the host
and (kernels only) with the plan cache off.  tools/plan_probe.py"""
import os
import sys
import time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import discorpy_b200 as dcb
import discorpy_b200.post.postprocessing as post

H = W = 4096
xc, yc = 2050.37, 2040.81
fact = [1.00227490554, -2.99523692178e-05 / 3, 8.99519088e-08 / 9, -1.57066461911e-10 / 27,
        8.08880211618e-14 / 81]
dcb.set_device(0)
src = dcb.DeviceArray((H, W)).fill_synthetic(seed=1)
post.unwarp_image_backward(src, xc, yc, fact)
dcb.synchronize()
for rep in range(3):
    dcb.plan_cache_clear()
    dcb.synchronize()
    t0 = time.perf_counter()
    out = post.unwarp_image_backward(src, xc + 0.001 * rep, yc, fact)
    t1 = time.perf_counter()
    dcb.synchronize()
    t2 = time.perf_counter()
    out2 = post.unwarp_image_backward(src, xc + 0.001 * rep, yc, fact)
    dcb.synchronize()
    t3 = time.perf_counter()
    print("cold call: enqueue %.0f us, done after %.0f us; warm call %.0f us" % ((t1 - t0) * 1e6, (t2 - t0) * 1e6, (t3 - t2) * 1e6))
s = dcb.current_stream()
e0, e1 = dcb.Event(), dcb.Event()
os.environ["DCB_PLAN_CACHE"] = "0"
