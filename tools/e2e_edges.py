"""Sweep of unequal row-band schedules of the host-buffer pipeline (DCB_BAND_EDGES: band heights in
64ths of the image) with the per-band device timeline (DCB_PIPE_TRACE=1) of a few of them; pinned
4096^2 float32 in, pinned out, post.unwarp_image_backward."""
import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
import discorpy_b200 as dcb
import discorpy_b200.post.postprocessing as post
dcb.set_device(0)
H = W = 4096
fact = [1.00227490554, -2.99523692178e-05 / 3, 8.99519088e-08 / 9, -1.57066461911e-10 / 27, 8.08880211618e-14 / 81]
xc, yc = 2050.37, 2040.81
host_in = []
for i in range(4):
    a = dcb.pinned_empty((H, W), np.float32); a[:] = 1.0 + i
    host_in.append(a)

SCHEDS = {
    "auto": "",
    "equal8": "8,8,8,8,8,8,8,8",
    "head": "2,2,4,8,8,8,8,8,8,8",
    "tail": "8,8,8,8,8,8,8,4,2,2",
    "both": "2,2,4,8,8,8,8,8,8,4,2,2",
    "both_fine": "1,1,2,4,8,8,8,8,8,8,4,2,1,1",
    "equal16": ",".join(["4"] * 16),
    "head6": "2,2,4,6,6,6,6,6,6,6,6,4,2,2",
    "ramp": "1,1,2,2,4,4,6,6,8,8,8,6,4,2,1,1",
    "big_mid": "2,2,4,8,16,16,8,4,2,2",
}

def run(name, n=40):
    os.environ["DCB_BAND_EDGES"] = SCHEDS[name]
    for _ in range(3):
        post.unwarp_image_backward(host_in[0], xc, yc, fact)
    ts = []
    for k in range(n):
        t1 = time.perf_counter()
        out = post.unwarp_image_backward(host_in[k % 4], xc, yc, fact)
        ts.append(time.perf_counter() - t1)
    ts.sort()
    print("%-10s %-40s best %.3f  median %.3f ms  (%.2f Gpixel/s)" % (name, SCHEDS[name] or "(library default)", ts[0] * 1e3, ts[n // 2] * 1e3, H * W / ts[n // 2] / 1e9), flush=True)

def duplex_probe():
    """Ceiling of a chunked pipeline: 64 MiB up and 64 MiB down at once, in chunks, on one or two
    streams per direction (no kernels, no dependencies between the directions)."""
    import ctypes
    from discorpy_b200 import _cabi
    vp = ctypes.c_void_p
    n = 64 << 20
    dA, dB = dcb.device.DeviceBuffer(n), dcb.device.DeviceBuffer(n)
    hin = dcb.pinned_empty((n // 4,), np.float32); hin[:] = 1.0
    hout = dcb.pinned_empty((n // 4,), np.float32)
    ups = [dcb.device.Stream(), dcb.device.Stream()]
    dns = [dcb.device.Stream(), dcb.device.Stream()]
    for chunk_mb in (64, 16, 8, 4, 2):
        for nstr in (1, 2):
            c = chunk_mb << 20
            best = 1e9
            for rep in range(6):
                dcb.synchronize()
                t0 = time.perf_counter()
                for k in range(n // c):
                    _cabi.call("dcb_h2d", vp(dA.ptr + k * c), vp(hin.ctypes.data + k * c), c, vp(ups[k % nstr].handle))
                    _cabi.call("dcb_d2h", vp(hout.ctypes.data + k * c), vp(dB.ptr + k * c), c, vp(dns[k % nstr].handle))
                for q in ups + dns:
                    q.sync()
                best = min(best, time.perf_counter() - t0)
            print("duplex probe: %2d MiB chunks, %d stream(s) per direction: %.3f ms = %.1f GB/s each way" % (chunk_mb, nstr, best * 1e3, n / best / 1e9), flush=True)

if os.environ.get("DCB_PIPE_PROBE") == "1":
    duplex_probe()
    for direct in ("0", "1"):
        os.environ["DCB_PIPE_DIRECT"] = direct
        print("== DCB_PIPE_DIRECT", direct, flush=True)
        for name in ("auto", "equal8", "head6", "equal16"):
            run(name)
    os.environ["DCB_PIPE_DIRECT"] = "0"
    os.environ["DCB_PIPE_NOKERNEL"] = "1"
    print("== copies only (DCB_PIPE_NOKERNEL)", flush=True)
    for name in ("auto", "equal8", "equal16"):
        run(name)
elif os.environ.get("DCB_PIPE_TRACE") == "1":
    for name in sys.argv[1:] or ["equal8", "both"]:
        os.environ["DCB_BAND_EDGES"] = SCHEDS[name]
        print("== trace", name, flush=True)
        for _ in range(4):
            post.unwarp_image_backward(host_in[0], xc, yc, fact)
else:
    for rep in range(2):
        for name in SCHEDS:
            run(name)
