import sys, os, threading
sys.path.insert(0, os.getcwd())
import numpy as np
import discorpy_b200 as dcb
import discorpy_b200.post.postprocessing as post
from oracle import oracle_spline as osp
rng = np.random.default_rng(0)
mat = rng.random((300, 521), dtype=np.float32)
fact = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
res = {}
def work(dev):
    dcb.set_device(dev)
    a = post.unwarp_image_backward(mat, 250.3, 140.8, fact)
    b = post.unwarp_image_backward(mat, 250.3, 140.8, fact, order=3)
    res[dev] = (a, b, dcb.device_info(dev)["name"])
ts = [threading.Thread(target=work, args=(d,)) for d in range(dcb.device_count())]
[t.start() for t in ts]; [t.join() for t in ts]
w1 = osp.unwarp_image_backward(mat, 250.3, 140.8, fact, 1, "reflect")
w3 = osp.unwarp_image_backward(mat, 250.3, 140.8, fact, 3, "reflect")
for d, (a, b, name) in sorted(res.items()):
    print("device", d, name, "order1 ok", np.array_equal(a, w1), "order3 ok", np.array_equal(b, w3))
