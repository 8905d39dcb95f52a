"""post.unwarp_image_backward end to end with pageable (ordinary NumPy) input against pinned input."""
import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
import discorpy_b200 as dcb
import discorpy_b200.post.postprocessing as post
dcb.set_device(0)
H = W = 4096
fact = [1.00227490554, -2.99523692178e-05 / 3, 8.99519088e-08 / 9, -1.57066461911e-10 / 27, 8.08880211618e-14 / 81]
xc, yc = 2050.37, 2040.81
rng = np.random.default_rng(0)
for kind in ("pinned", "pageable", "pageable uint16", "pageable uint8"):
    ins = []
    for i in range(4):
        if kind == "pinned":
            a = dcb.pinned_empty((H, W), np.float32); a[:] = rng.random((H, W), dtype=np.float32)
        elif kind == "pageable":
            a = rng.random((H, W), dtype=np.float32)
        elif kind == "pageable uint16":
            a = rng.integers(0, 65535, (H, W), dtype=np.uint16)
        else:
            a = rng.integers(0, 255, (H, W), dtype=np.uint8)
        ins.append(a)
    for _ in range(3):
        out = post.unwarp_image_backward(ins[0], xc, yc, fact)
    ts = []
    for k in range(20):
        t0 = time.perf_counter()
        out = post.unwarp_image_backward(ins[k % 4], xc, yc, fact)
        ts.append(time.perf_counter() - t0)
    print("%-16s best %.2f ms  median %.2f ms  -> %s" % (kind, min(ts) * 1e3, sorted(ts)[10] * 1e3, out.dtype), flush=True)

# a tomography chunk: 32 projections of 2560^2, every row, uint16 and float32
D, H2 = 32, 2560
f5 = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
for dt in (np.uint16, np.float32):
    stack = (rng.integers(0, 60000, (D, H2, H2)).astype(dt) if dt == np.uint16
             else rng.random((D, H2, H2), dtype=np.float32))
    for _ in range(2):
        out = post.unwarp_chunk_slices_backward(stack, 1283.4, 1275.9, f5, 0, H2 - 1)
    ts = []
    for k in range(5):
        t0 = time.perf_counter()
        out = post.unwarp_chunk_slices_backward(stack, 1283.4, 1275.9, f5, 0, H2 - 1)
        ts.append(time.perf_counter() - t0)
    print("chunk %-8s 32 x 2560^2 pageable: best %.1f ms median %.1f ms -> %s" % (np.dtype(dt).name, min(ts) * 1e3, sorted(ts)[2] * 1e3, out.dtype), flush=True)
