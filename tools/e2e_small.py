"""Equal-band count sweep of the host-buffer pipeline for images below 32 MiB."""
import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
import discorpy_b200 as dcb
import discorpy_b200.post.postprocessing as post
dcb.set_device(0)
fact = [1.00227490554, -2.99523692178e-05 / 3, 8.99519088e-08 / 9, -1.57066461911e-10 / 27, 8.08880211618e-14 / 81]
for (H, W) in ((512, 640), (1024, 1024), (1200, 1600), (2048, 2048), (2160, 2560), (2560, 2560), (2800, 2800)):
    f = [fact[i] * (4096.0 / W) ** i for i in range(5)]
    a = dcb.pinned_empty((H, W), np.float32); a[:] = 1.0
    out = []
    for bands in (0, 1, 2, 3, 4, 6, 8, 12, 16):
        post.config["bands"] = bands
        for _ in range(3):
            post.unwarp_image_backward(a, W / 2 + 2.4, H / 2 - 7.2, f)
        ts = []
        for k in range(30):
            t1 = time.perf_counter()
            post.unwarp_image_backward(a, W / 2 + 2.4, H / 2 - 7.2, f)
            ts.append(time.perf_counter() - t1)
        ts.sort()
        out.append("%s %.3f" % ("auto" if bands == 0 else str(bands), ts[15] * 1e3))
    print("%5d x %5d (%5.1f MiB) median ms by band count: %s" % (H, W, H * W * 4 / 2 ** 20, "  ".join(out)), flush=True)
post.config["bands"] = 0
