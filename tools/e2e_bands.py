"""Sweep of the row-band count of the host-buffer pipeline (post.config["bands"]) and of the number of
host threads calling post.unwarp_image_backward at once; pinned 4096^2 float32 in, pinned out."""
import sys, os, time, threading, numpy as np
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

def loop(n, tid, hold):
    for k in range(n):
        hold[tid] = post.unwarp_image_backward(host_in[(k + tid) % 4], xc, yc, fact)

def run(bands, nthreads, n=40):
    post.config["bands"] = bands
    hold = [None] * nthreads
    for _ in range(3):
        loop(1, 0, hold)
    ths = [threading.Thread(target=loop, args=(n, t, hold)) for t in range(nthreads)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    dcb.synchronize()
    tot = time.perf_counter() - t0
    print("bands %2d threads %d: %.3f ms per image (%.2f Gpixel/s)" % (bands, nthreads, tot / (n * nthreads) * 1e3, n * nthreads * H * W / tot / 1e9), flush=True)

for rep in range(2):
    for b in (0, 4, 8, 12, 16, 24, 32):
        run(b, 1)
for b in (8, 16):
    for nt in (2, 3):
        run(b, nt)
