#!/usr/bin/env python
"""Golden vectors for unwarp_chunk_slices_backward when rows of the chunk sample OUTSIDE the row
window the reference crops to (postprocessing.py:289-301 takes it from the first and last row only;
SciPy then reflects the coordinates into the cropped slice).  Runs the real reference from
/root/reference (this container only) and stores inputs + outputs in tests/golden/chunk_window.npz.
Found by tests/fuzz_parity.py (seed 20261017, case 1885) and a seeded search for excursions below
the window and further than one window height away."""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import discorpy.post.postprocessing as rpost          # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_np as orc                   # noqa: E402


def excursion(h, w, xc, yc, fact, a, b):
    y0, y1 = orc.chunk_row_window(h, w, xc, yc, fact, a, b)
    yd, _ = orc.radial_coords(h, w, xc, yc, fact, row0=a, nrows=b - a + 1)
    return y0, y1, float((y0 - yd).max()), float((yd - (y1 - 1)).max())


cases = []
rng = np.random.default_rng(20261017)
# 1: the fuzz case (int8, 3 x 135 x 88, rows 1..75, rows up to 2.5 above the window)
mat = None
for it in range(1886):
    h, w = int(rng.integers(1, 260)), int(rng.integers(1, 330))
    dt = rng.choice(["float32", "float32", "uint8", "uint16", "int16", "int8", "float64"])
    if dt in ("float32", "float64"):
        mat = (rng.random((h, w)) * 400 - 100).astype(dt)
    else:
        info = np.iinfo(dt)
        mat = rng.integers(info.min, info.max, (h, w), dtype=dt, endpoint=True)
    nt = int(rng.integers(1, 8))
    scale = max(h, w)
    fact = [float(rng.uniform(0.6, 1.4))] + [float(rng.normal() * 0.3 / scale ** i) for i in range(1, nt)]
    xc = float(rng.uniform(-0.5, 1.5) * w)
    yc = float(rng.uniform(-0.5, 1.5) * h)
    kind = rng.choice(["radial", "radial", "persp", "chunk"])
    rng.choice([0, 1, 1, 1, 2, 3, 3, 4, 5])
    rng.choice(8)
    if kind == "persp":
        for _ in range(8):
            rng.normal()
    elif kind == "chunk":
        if dt == "float64" or h < 2:
            continue
        fact = [float(rng.uniform(0.9, 1.1))] + [float(rng.normal() * 0.03 / scale ** i) for i in range(1, nt)]
        d = int(rng.integers(1, 5))
        stack = np.stack([np.roll(mat, k, axis=1) for k in range(d)])
        a = int(rng.integers(0, h))
        b = int(rng.integers(a, h))
cases.append((stack, xc, yc, fact, a, b))
# 2, 3: seeded search: excursion below the window / further than a window height
rng = np.random.default_rng(7)
want_below, want_far = True, True
while want_below or want_far:
    h, w = int(rng.integers(20, 120)), int(rng.integers(8, 90))
    nt = int(rng.integers(3, 7))
    scale = max(h, w)
    fact = [float(rng.uniform(0.8, 1.2))] + [float(rng.normal() * 0.4 / scale ** i) for i in range(1, nt)]
    xc, yc = float(rng.uniform(-0.5, 1.5) * w), float(rng.uniform(-0.5, 1.5) * h)
    a = int(rng.integers(0, h)); b = int(rng.integers(a, h))
    y0, y1, above, below = excursion(h, w, xc, yc, fact, a, b)
    n = y1 - y0
    take = False
    if want_below and below > 1.5 and n >= 3:
        want_below, take = False, True
    elif want_far and max(above, below) > 2.2 * n and n >= 2:
        want_far, take = False, True
    if take:
        dt = np.float32 if len(cases) == 1 else np.uint16
        stack = (rng.random((2, h, w)) * 1000).astype(dt)
        cases.append((stack, xc, yc, fact, a, b))

out = {}
for i, (stack, xc, yc, fact, a, b) in enumerate(cases):
    h, w = stack.shape[1:]
    ref = rpost.unwarp_chunk_slices_backward(stack, xc, yc, fact, a, b)
    mine = orc.unwarp_chunk_slices_backward(stack, xc, yc, fact, a, b)
    print("case", i, stack.dtype, stack.shape, "rows", a, b, "window/excursions", excursion(h, w, xc, yc, fact, a, b),
          "oracle == reference:", np.array_equal(ref, mine))
    out["stack%d" % i] = stack
    out["par%d" % i] = np.array([xc, yc, a, b] + list(fact), dtype=np.float64)
    out["ref%d" % i] = ref
out["n"] = np.array(len(cases))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "chunk_window.npz"), **out)
print("wrote tests/golden/chunk_window.npz")
