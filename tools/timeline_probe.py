"""Per-CTA timeline of the single-image kernel on the BASELINE config-2 geometry
(dcb_image_timeline): when the CTAs start, when their first tile is ready, when they finish.
tools/timeline_probe.py [order] [blend]"""
import ctypes
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import discorpy_b200 as dcb
from discorpy_b200 import _cabi
import discorpy_b200.post.postprocessing as post

order = int(sys.argv[1]) if len(sys.argv) > 1 else 1
blend = sys.argv[2] if len(sys.argv) > 2 else "exact"
H = W = 4096
xc, yc = 2050.37, 2040.81
fact = [1.00227490554, -2.99523692178e-05 / 3, 8.99519088e-08 / 9, -1.57066461911e-10 / 27,
        8.08880211618e-14 / 81]
dcb.set_device(0)
rng = np.random.default_rng(1)
srcs = [dcb.DeviceArray.from_host(rng.random((H, W), dtype=np.float32)) for _ in range(4)]
post.config["blend"] = {"exact": dcb.BLEND_EXACT, "lerp64": dcb.BLEND_LERP64,
                        "lerp32": dcb.BLEND_LERP32}[blend]
keep = []
for i in range(8):
    keep.append(post.unwarp_image_backward(srcs[i % 4], xc, yc, fact, order=order))
dcb.synchronize()
dcb.image_stats(True, reset=True)
keep = keep[-4:]
for i in range(6):
    keep.append(post.unwarp_image_backward(srcs[i % 4], xc, yc, fact, order=order))
dcb.synchronize()
grid = dcb.last_plan()["grid"]
buf = (ctypes.c_uint64 * (8 * grid))()
_cabi.call("dcb_image_timeline", buf, grid)
dcb.image_stats(False)
t = np.array(buf, dtype=np.uint64).reshape(grid, 8)
start, first, end = (t[:, i].astype(np.int64) for i in range(3))
sm = (t[:, 3] & np.uint64(0xffffffff)).astype(np.int64)
ntl = (t[:, 3] >> np.uint64(32)).astype(np.int64)
t0 = start.min()
q = lambda a: "min %7.2f  p10 %7.2f  med %7.2f  p90 %7.2f  max %7.2f" % tuple(
    np.percentile(a, [0, 10, 50, 90, 100]) / 1e3)
print("order %d blend %s grid %d  (us after the first CTA start)" % (order, blend, grid))
print("CTA start          ", q(start - t0))
print("first tile ready   ", q(first - t0))
print("first tile - start ", q(first - start))
print("CTA end            ", q(end - t0))
print("CTA busy (end-first)", q(end - first))
print("kernel span %.2f us; mean CTA idle before first tile %.2f us, after its end %.2f us"
      % ((end.max() - t0) / 1e3, np.mean(first - t0) / 1e3, np.mean(end.max() - end) / 1e3))
per_tile = (end - first) / np.maximum(ntl, 1)
print("tiles per CTA: %s; us per tile " % np.unique(ntl), q(per_tile))
wait, wmax = t[:, 4].astype(np.int64), t[:, 5].astype(np.int64)
print("warp 0 waits for tiles 1.. (total)", q(wait), " longest single wait", q(wmax))
os.makedirs("gpurun_out", exist_ok=True)
np.savetxt("gpurun_out/timeline_%d_%s.csv" % (order, blend),
           np.stack([np.arange(grid), sm, ntl, start - t0, first - t0, end - t0, wait, wmax], 1),
           fmt="%d", delimiter=",", header="cta,sm,tiles,start_ns,first_ns,end_ns,wait_ns,wmax_ns")
# per-SM spread
order_ = np.argsort(end)
print("slowest CTAs (end us, SM, tiles):", [(round((end[i] - t0) / 1e3, 2), int(sm[i]), int(ntl[i])) for i in order_[-6:]])
print("fastest CTAs (end us, SM, tiles):", [(round((end[i] - t0) / 1e3, 2), int(sm[i]), int(ntl[i])) for i in order_[:6]])

# per-warp event log of CTA 0: sampling warps (start, end of tile i), producers (buffer free, box landed, widened)
log = (ctypes.c_uint64 * (4 * 10 * 64))()
_cabi.call("dcb_image_timeline", log, -1)
lg = np.array(log, dtype=np.uint64).reshape(4, 10, 64).astype(np.int64)
for cta in (0, 1):
    L = lg[cta]
    print("CTA %d event log (us after the first CTA start)" % cta)
    for w in range(10):
        ev = L[w][L[w] > 0]
        if ev.size == 0:
            continue
        print("  warp %d:" % w, " ".join("%.2f" % ((e - t0) / 1e3) for e in ev[:45]))
