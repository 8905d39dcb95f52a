"""Run with DCB_PLAN_CACHE_MB=16 (tests/test_gpu_parity.py does, in a subprocess: the limit is read
once per process): more calibrations than the plan cache holds, so plans are evicted and their
slab memory is reused while launches keep coming; every result must equal the oracle's."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import discorpy_b200 as dcb
import discorpy_b200.post.postprocessing as post
from oracle import oracle_c

dcb.set_device(0)
rng = np.random.default_rng(82)
mat = rng.random((1536, 2048), dtype=np.float32)
dev = dcb.DeviceArray.from_host(mat)
fact = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
built0 = dcb.plan_cache_clear()
outs = []
models = [(1020.3 + 3.7 * i, 760.9 - 2.1 * i) for i in range(24)]
for rnd in range(2):
    for xc, yc in models:
        outs.append(((xc, yc), post.unwarp_image_backward(dev, xc, yc, fact)))
bad = 0
for (xc, yc), o in outs[::5] + outs[-3:]:
    want = oracle_c.unwarp_image_backward(mat, xc, yc, fact, 1)
    bad += int(np.count_nonzero(o.to_host() != want))
built = dcb.plan_cache_clear() - built0
print("plans built %d for %d launches of %d models, %d differing pixels" % (built, len(outs), len(models), bad))
assert bad == 0
assert built > len(models), "the cache limit did not force any eviction"
print("ok")
