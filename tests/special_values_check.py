"""GPU vs oracle on images holding NaN, +-Inf, +-0.0, denormals and near-overflow values (orders 0 / 1):
values must agree (NaN where NaN); reports how many results differ only in their bit pattern
(sign of zero, NaN payload).  Test infrastructure: uses oracle/."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import discorpy_b200.post.postprocessing as post          # noqa: E402
from oracle import oracle_spline as osp                   # noqa: E402

rng = np.random.default_rng(0)
bad = bits = total = 0
for it in range(40):
    h, w = int(rng.integers(5, 300)), int(rng.integers(5, 400))
    mat = (rng.random((h, w)) * 100 - 20).astype(np.float32)
    k = rng.integers(0, h * w, max(12, h * w // 50))
    vals = [np.nan, np.inf, -np.inf, 0.0, -0.0, 1e-42, 3e38, -3e38]
    mat.ravel()[k] = rng.choice(vals, k.size).astype(np.float32)
    fact = [float(rng.uniform(0.8, 1.2)), float(rng.normal() * 0.2 / max(h, w)), float(rng.normal() * 0.2 / max(h, w) ** 2)]
    xc, yc = float(rng.uniform(0, w)), float(rng.uniform(0, h))
    for order in (0, 1):
        with np.errstate(all="ignore"):
            want = osp.unwarp_image_backward(mat, xc, yc, fact, order, "reflect")
        got = post.unwarp_image_backward(mat, xc, yc, fact, order=order)
        total += want.size
        if not np.array_equal(got, want, equal_nan=True):
            bad += 1
            print("VALUE mismatch case", it, "order", order, int(np.count_nonzero(~((got == want) | (np.isnan(got) & np.isnan(want))))))
        d = got.view(np.uint32) != want.view(np.uint32)
        bits += int(d.sum())
print("special values: %d cases with value mismatches; %d of %d results differ in bit pattern only" % (bad, bits, total))
sys.exit(1 if bad else 0)
