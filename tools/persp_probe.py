import sys, os, json
sys.path.insert(0, os.getcwd())
import numpy as np
import discorpy_b200 as dcb
import discorpy_b200.post.postprocessing as post
dcb.set_device(0)
n = 4096
coef = [1.02, 0.01, -15.0, 0.005, 1.01, -8.0, 8e-6, -5e-6]
img = dcb.DeviceArray((n, n)).fill_synthetic(seed=3)
for rep in range(3):
    out = post.correct_perspective_image(img, coef)
dcb.synchronize()
print(dcb.last_plan())
