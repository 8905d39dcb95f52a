import os, sys
sys.path.insert(0, ".")
import numpy as np
import discorpy_b200 as dcb
import discorpy_b200.post.postprocessing as post
dcb.set_device(0)
COEF = [1.02, 0.01, -15.0, 0.005, 1.01, -8.0, 8e-6, -5e-6]
for n in (2048, 4096):
    mat = dcb.DeviceArray((n, n)).fill_synthetic(seed=3)
    dcb.plan_cache_clear(); dcb.image_stats(True, reset=True)
    post.correct_perspective_image(mat, COEF); dcb.synchronize()
    st = dcb.image_stats(False, reset=True)
    print(n, st, dcb.last_plan())
