"""End-to-end time of post.unwarp_image_backward (pinned float32 in, pinned out) for several image
sizes, with the library's band schedule and with equal bands (post.config["bands"] = 8)."""
import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
import discorpy_b200 as dcb
import discorpy_b200.post.postprocessing as post
dcb.set_device(0)
fact = [1.00227490554, -2.99523692178e-05 / 3, 8.99519088e-08 / 9, -1.57066461911e-10 / 27, 8.08880211618e-14 / 81]
for (H, W) in ((2160, 2560), (2560, 2560), (4096, 4096), (6000, 8000), (8192, 8192)):
    f = [fact[i] * (4096.0 / W) ** i for i in range(5)]
    a = dcb.pinned_empty((H, W), np.float32); a[:] = 1.0
    for bands in (0, 8):
        post.config["bands"] = bands
        for _ in range(3):
            post.unwarp_image_backward(a, W / 2 + 2.4, H / 2 - 7.2, f)
        ts = []
        for k in range(20):
            t1 = time.perf_counter()
            post.unwarp_image_backward(a, W / 2 + 2.4, H / 2 - 7.2, f)
            ts.append(time.perf_counter() - t1)
        ts.sort()
        print("%5d x %5d  bands %s: median %.3f ms  (%.2f Gpixel/s, %.1f GB/s each way)" % (H, W, "auto   " if bands == 0 else "8 equal", ts[10] * 1e3, H * W / ts[10] / 1e9, H * W * 4 / ts[10] / 1e9), flush=True)
post.config["bands"] = 0
