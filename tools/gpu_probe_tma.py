"""TMA geometry probe: which (box, start coordinate) combinations complete and
deliver the right bytes.  Diagnostics only."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import discorpy_b200 as dcb
from discorpy_b200 import _cabi

dcb.set_device(0)
H, W, D = 64, 200, 2
host = np.arange(D * H * W, dtype=np.float32).reshape(D, H, W)
src = dcb.DeviceArray.from_host(host)
out = dcb.DeviceArray((256, 256))
for (bw, bh) in ((64, 8), (48, 28), (132, 36), (200, 16), (204, 8), (4, 4), (256, 20)):
    for (x0, y0, z0) in ((0, 0, 0), (4, 0, 0), (8, 5, 0), (12, 0, 1), (100, 9, 1), (-4, -3, 0), (192, 60, 1), (196, 63, 1)):
        st = ctypes.c_int(-1)
        try:
            _cabi.call("dcb_selftest_tma", ctypes.c_void_p(src.ptr), D, H, W, src.pitch,
                       src.slice_stride, bw, bh, x0, y0, z0, ctypes.c_void_p(out.ptr),
                       ctypes.byref(st))
        except Exception as e:
            print("box %dx%d at (%d,%d,%d): EXC %s" % (bw, bh, x0, y0, z0, e)); sys.exit(0)
        got = out.to_host().ravel()[:bw * bh].reshape(bh, bw)
        want = np.zeros((bh, bw), np.float32)
        ys = np.arange(y0, y0 + bh); xs = np.arange(x0, x0 + bw)
        vy = (ys >= 0) & (ys < H); vx = (xs >= 0) & (xs < W)
        want[np.ix_(vy, vx)] = host[z0][np.ix_(ys[vy], xs[vx])]
        print("box %3dx%-3d at (%4d,%3d,%d): status %d, data %s" % (
            bw, bh, x0, y0, z0, st.value, "OK" if np.array_equal(got, want) else "MISMATCH"))
