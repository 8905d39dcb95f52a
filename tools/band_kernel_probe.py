"""How long does one 512-row band launch of the single-image kernel take on the device: launched
back to back, launched into an idle GPU (host sleeps between launches), and with both copy
engines busy?  (diagnostics for the host-buffer pipeline)"""
import sys, os, time, ctypes, threading, numpy as np
sys.path.insert(0, os.getcwd())
import discorpy_b200 as dcb
from discorpy_b200 import _cabi
dcb.set_device(0)
vp = ctypes.c_void_p
H = W = 4096
fact = [1.00227490554, -2.99523692178e-05 / 3, 8.99519088e-08 / 9, -1.57066461911e-10 / 27, 8.08880211618e-14 / 81]
model = _cabi.make_radial(2050.37, 2040.81, fact)
opt = _cabi.make_options(1)
src = dcb.DeviceArray((H, W)).fill_synthetic(seed=1)
dst = dcb.DeviceArray((H, W))
s = dcb.device.Stream()

def band(b, nb, stream):
    rows = H // nb
    _cabi.call("dcb_unwarp_stack_backward_f32", vp(src.ptr), vp(dst.ptr + b * rows * dst.pitch), 1, H, W, 0, H,
               src.pitch, src.pitch * H, dst.pitch, dst.pitch * rows, b * rows, rows, 1,
               ctypes.byref(model), ctypes.byref(opt), vp(stream.handle))

def measure(nb, gap_s, label):
    for b in range(nb):
        band(b, nb, s)
    s.sync()
    res = []
    for rep in range(5):
        evs = []
        for b in range(nb):
            e0, e1 = dcb.device.Event(), dcb.device.Event()
            e0.record(s); band(b, nb, s); e1.record(s)
            evs.append((e0, e1))
            if gap_s:
                s.sync(); time.sleep(gap_s)
        s.sync()
        res.append([e0.elapsed_ms(e1) * 1e3 for e0, e1 in evs])
    med = np.median(np.array(res), axis=0)
    print("%-44s %d bands: us per band launch %s  (sum %.1f)" % (label, nb, " ".join("%.1f" % v for v in med), med.sum()), flush=True)

for nb in (1, 8, 16):
    measure(nb, 0, "back to back")
    measure(nb, 200e-6, "idle GPU between launches (200 us sleeps)")
    measure(nb, 2e-3, "idle GPU between launches (2 ms sleeps)")

# both copy engines busy
n = 64 << 20
dA, dB = dcb.device.DeviceBuffer(n), dcb.device.DeviceBuffer(n)
hin = dcb.pinned_empty((n // 4,), np.float32); hin[:] = 1.0
hout = dcb.pinned_empty((n // 4,), np.float32)
up, dn = dcb.device.Stream(), dcb.device.Stream()
stop = False
def copier():
    while not stop:
        _cabi.call("dcb_h2d", vp(dA.ptr), vp(hin.ctypes.data), n, vp(up.handle))
        _cabi.call("dcb_d2h", vp(hout.ctypes.data), vp(dB.ptr), n, vp(dn.handle))
        up.sync(); dn.sync()
th = threading.Thread(target=copier); th.start()
time.sleep(0.05)
for nb in (1, 8):
    measure(nb, 0, "copies running, back to back")
    measure(nb, 200e-6, "copies running, 200 us sleeps")
stop = True; th.join()
