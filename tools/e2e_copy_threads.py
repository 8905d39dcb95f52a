"""Pageable (ordinary NumPy) 4096^2 float32 input: end-to-end time against the number of host
threads staging the row bands into pinned memory (DCB_COPY_THREADS, read when the pool starts)."""
import sys, os, time, subprocess
if len(sys.argv) > 1:
    import numpy as np
    sys.path.insert(0, os.getcwd())
    import discorpy_b200 as dcb
    import discorpy_b200.post.postprocessing as post
    dcb.set_device(0)
    H = W = 4096
    fact = [1.00227490554, -2.99523692178e-05 / 3, 8.99519088e-08 / 9, -1.57066461911e-10 / 27, 8.08880211618e-14 / 81]
    rng = np.random.default_rng(0)
    ins = [rng.random((H, W), dtype=np.float32) for _ in range(24)]
    for i in range(3):
        post.unwarp_image_backward(ins[i], 2050.37, 2040.81, fact)
    ts = []
    for k in range(3, 24):     # every array new to the library (no in-place page-locking)
        t0 = time.perf_counter()
        post.unwarp_image_backward(ins[k], 2050.37, 2040.81, fact)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print("DCB_COPY_THREADS=%-3s best %.2f ms  median %.2f ms  (%.2f Gpixel/s)" % (sys.argv[1], ts[0] * 1e3, ts[len(ts) // 2] * 1e3, H * W / ts[len(ts) // 2] / 1e9), flush=True)
else:
    for n in ("2", "4", "6", "8", "12", "16", "24", "32"):
        subprocess.run([sys.executable, __file__, n], env=dict(os.environ, DCB_COPY_THREADS=n))
