"""End-to-end time (pinned float32 in, pinned out) of correct_perspective_image and of the combined
radial -> projective entry: the banded host pipeline against one band (upload, kernels, download
in series -- what these two calls did before)."""
import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
import discorpy_b200 as dcb
import discorpy_b200.post.postprocessing as post
dcb.set_device(0)
fact = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
coef = [1.02, 0.01, -15.0, 0.005, 1.01, -8.0, 8e-6, -5e-6]
for (H, W) in ((2048, 2048), (4096, 4096)):
    f = [fact[i] * (2048.0 / W) ** i for i in range(5)]
    a = dcb.pinned_empty((H, W), np.float32); a[:] = np.random.default_rng(1).random((H, W), dtype=np.float32)
    for name, fn in (("correct_perspective_image", lambda: post.correct_perspective_image(a, coef)),
                     ("unwarp_image_backward_perspective", lambda: post.unwarp_image_backward_perspective(a, W / 2 + 6.2, H / 2 - 4.4, f, coef))):
        for bands in (1, 0):
            post.config["bands"] = bands
            for _ in range(3):
                fn()
            ts = []
            for k in range(20):
                t1 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t1)
            ts.sort()
            print("%5d^2 %-34s %s: median %.3f ms (%.2f Gpixel/s)" % (H, name, "one band " if bands == 1 else "banded   ", ts[10] * 1e3, H * W / ts[10] / 1e9), flush=True)
post.config["bands"] = 0
