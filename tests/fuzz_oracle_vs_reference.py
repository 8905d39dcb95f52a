#!/usr/bin/env python
"""Randomised differential test of the ORACLE against the real reference (CPU only; needs
/root/reference, i.e. it runs in the build container, not on the GPU box): the same case generator
as tests/fuzz_parity.py, the reference's own functions on one side, oracle/ on the other.
Usage: python tests/fuzz_oracle_vs_reference.py [N] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import discorpy.post.postprocessing as rpost                   # noqa: E402
import discorpy.util.utility as rutil                          # noqa: E402
from oracle import oracle_np as orc                            # noqa: E402
from oracle import oracle_spline as osp                        # noqa: E402

MODES = osp.MODES


def run(n, seed):
    rng = np.random.default_rng(seed)
    bad = done = 0
    for it in range(n):
        h, w = int(rng.integers(1, 200)), int(rng.integers(1, 240))
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
        kind = rng.choice(["radial", "persp", "chunk", "chunk_wild", "slice", "both", "color"])
        order = int(rng.choice([0, 1, 1, 1, 2, 3, 5]))
        mode = str(rng.choice(MODES))
        coef = [float(c) for c in (1 + rng.normal() * 0.05, rng.normal() * 0.05, rng.normal() * 5,
                                   rng.normal() * 0.05, 1 + rng.normal() * 0.05, rng.normal() * 5,
                                   rng.normal() * 1e-4, rng.normal() * 1e-4)]
        d = int(rng.integers(1, 4))
        stack = np.stack([np.roll(mat, k, axis=1) for k in range(d)])
        a = int(rng.integers(0, h))
        b = int(rng.integers(a, h))
        try:
            if kind == "radial":
                ref = rpost.unwarp_image_backward(mat, xc, yc, fact, order=order, mode=mode)
                got = osp.unwarp_image_backward(mat, xc, yc, fact, order, mode)
            elif kind == "persp":
                ref = rpost.correct_perspective_image(mat, coef, order=order, mode=mode)
                got = osp.correct_perspective_image(mat, coef, order, mode)
            elif kind in ("chunk", "chunk_wild"):
                if kind == "chunk":   # mild model: rows stay inside the reference's window
                    fact = [float(rng.uniform(0.9, 1.1))] + [float(rng.normal() * 0.03 / scale ** i) for i in range(1, nt)]
                y0, y1 = orc.chunk_row_window(h, w, xc, yc, fact, a, b)
                if y1 <= y0:      # empty window: SciPy reads past an empty slice, undefined
                    continue
                ref = rpost.unwarp_chunk_slices_backward(stack, xc, yc, fact, a, b)
                got = orc.unwarp_chunk_slices_backward(stack, xc, yc, fact, a, b)
            elif kind == "slice":
                ref = rpost.unwarp_slice_backward(stack, xc, yc, fact, a)
                got = orc.unwarp_slice_backward(stack, xc, yc, fact, a)
            elif kind == "both":
                if order > 1:
                    continue
                ref = rpost.correct_perspective_image(rpost.unwarp_image_backward(mat, xc, yc, fact, order=order), coef, order=order)
                got = orc.unwarp_image_backward_perspective(mat, xc, yc, fact, coef, order=order)
            else:
                if order > 1:
                    continue
                chan = int(rng.integers(1, 4))
                frame = np.stack([np.roll(mat, 3 * k, axis=0) for k in range(chan)], axis=2)
                pad = [0, int(rng.integers(0, 9)), tuple(int(v) for v in rng.integers(0, 7, 4))][int(rng.integers(0, 3))]
                ref = rutil.unwarp_color_image_backward(frame, xc, yc, fact, order=order, pad=pad)
                got = orc.unwarp_color_image_backward(frame, xc, yc, fact, order=order, pad=pad)
            done += 1
            same = got.dtype == ref.dtype and got.shape == ref.shape and np.array_equal(got, ref, equal_nan=True)
        except Exception as exc:
            same = False
            print("EXC", kind, type(exc).__name__, str(exc)[:160])
        if not same:
            bad += 1
            print("MISMATCH case %d: %s %s %dx%d order %d mode %s nt %d rows %d..%d" % (it, kind, dt, h, w, order, mode, nt, a, b), flush=True)
    print("oracle vs reference: %d cases compared, %d not bit-identical" % (done, bad))
    return bad


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 300, int(sys.argv[2]) if len(sys.argv) > 2 else 1) else 0)
