import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
import discorpy_b200 as dcb
import discorpy_b200.post.postprocessing as post
dcb.set_device(0)
H = W = 4096
fact = [1.00227490554, -2.99523692178e-05 / 3, 8.99519088e-08 / 9, -1.57066461911e-10 / 27, 8.08880211618e-14 / 81]
xc, yc = 2050.37, 2040.81
def run(nin, nimg_dev, label):
    devs = [dcb.DeviceArray((H, W)).fill_synthetic(seed=i) for i in range(nimg_dev)]
    host_in = []
    for i in range(nin):
        a = dcb.pinned_empty((H, W), np.float32); a[:] = 1.0 + i
        host_in.append(a)
    for _ in range(2):
        post.unwarp_image_backward(host_in[0], xc, yc, fact)
    ts = []
    t0 = time.perf_counter()
    for k in range(40):
        t1 = time.perf_counter()
        out = post.unwarp_image_backward(host_in[k % nin], xc, yc, fact)
        ts.append(time.perf_counter() - t1)
    dcb.synchronize()
    tot = time.perf_counter() - t0
    print("%-28s avg %.2f ms  best %.2f  median %.2f  worst %.2f" % (label, tot / 40 * 1e3, min(ts) * 1e3, sorted(ts)[20] * 1e3, max(ts) * 1e3))
run(1, 0, "1 input")
run(4, 0, "4 inputs")
run(4, 32, "4 inputs, 2 GiB on device")
