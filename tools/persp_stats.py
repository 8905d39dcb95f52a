"""Share of tile rows the projective map's patch path verifies, and why the others were not:
tools/persp_stats.py"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import discorpy_b200 as dcb
import discorpy_b200.post.postprocessing as post
dcb.set_device(0)
COEFS = {"cfg3": [1.02, 0.01, -15.0, 0.005, 1.01, -8.0, 8e-6, -5e-6],
         "strong": [0.93, -0.04, 40.0, 0.02, 0.97, 25.0, -6e-5, 4e-5]}
for shape in ((2048, 2048), (4096, 4096), (1200, 1664)):
    for name, coef in COEFS.items():
        mat = dcb.DeviceArray(shape).fill(0.0)
        dcb.plan_cache_clear()
        dcb.image_stats(True, reset=True)
        post.correct_perspective_image(mat, coef)
        dcb.synchronize()
        st = dcb.image_stats(False, reset=True)
        print(shape, name, {k: v for k, v in st.items() if k.startswith("rows")}, dcb.last_plan())
