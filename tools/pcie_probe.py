"""H2D / D2H bandwidth of pinned and pageable 64 MiB buffers, and the phases of
one host-buffer unwarp call (diagnostics for the e2e number)."""
import ctypes, time, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import discorpy_b200 as dcb
from discorpy_b200 import _cabi
import discorpy_b200.post.postprocessing as post
dcb.set_device(0)
n = 64 << 20
d = dcb.device.DeviceBuffer(n)
pin = dcb.pinned_empty((n // 4,), np.float32); pin[:] = 1.0
pag = np.ones(n // 4, np.float32)
s = dcb.current_stream()
vp = ctypes.c_void_p
for name, h in (("pinned", pin), ("pageable", pag)):
    for kind in ("h2d", "d2h"):
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            if kind == "h2d":
                _cabi.call("dcb_h2d", vp(d.ptr), vp(h.ctypes.data), n, vp(s.handle))
            else:
                _cabi.call("dcb_d2h", vp(h.ctypes.data), vp(d.ptr), n, vp(s.handle))
            s.sync()
            ts.append(time.perf_counter() - t0)
        print("%-8s %s 64 MiB: best %.2f ms = %.1f GB/s" % (name, kind, min(ts) * 1e3, n / min(ts) / 1e9))
H = W = 4096
img = dcb.pinned_empty((H, W), np.float32); img[:] = np.random.default_rng(0).random((H, W), dtype=np.float32)
fact = [1.00227490554, -2.99523692178e-05 / 3, 8.99519088e-08 / 9, -1.57066461911e-10 / 27, 8.08880211618e-14 / 81]
for bands in (1, 2, 4, 8, 16):
    post.config["bands"] = bands
    ts = []
    for _ in range(6):
        t0 = time.perf_counter()
        out = post.unwarp_image_backward(img, 2050.37, 2040.81, fact)
        ts.append(time.perf_counter() - t0)
    print("host call, %2d bands: best %.2f ms, median %.2f ms -> %.0f Mpix/s" % (bands, min(ts) * 1e3, sorted(ts)[3] * 1e3, H * W / 1e6 / min(ts)))
