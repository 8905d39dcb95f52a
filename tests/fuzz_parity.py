#!/usr/bin/env python
"""Randomised differential test of the CUDA path against the oracle (run on a GPU box; lives under
tests/ because only tests may use the oracle):
random shapes, centres (also far outside the image), polynomial lengths and strengths,
orders 0 / 1 / 2..5, modes, dtypes, perspective coefficients, row chunks and single slices of
stacks, the combined radial + perspective entry, colour frames with padding.
Prints every case whose outputs are not bit-identical.  Usage: python tests/fuzz_parity.py [N] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import discorpy_b200 as dcb                                    # noqa: E402
import discorpy_b200.post.postprocessing as post               # noqa: E402
import discorpy_b200.util.utility as util                      # noqa: E402
from oracle import oracle_np as orc                            # noqa: E402
from oracle import oracle_spline as osp                        # noqa: E402

MODES = osp.MODES


def run(n, seed):
    """Returns the number of cases (of n) whose outputs are not bit-identical."""
    rng = np.random.default_rng(seed)
    dcb.device.ensure_init()
    bad = 0
    for it in range(n):
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
        kind = rng.choice(["radial", "radial", "persp", "chunk", "slice", "both", "color"])
        order = int(rng.choice([0, 1, 1, 1, 2, 3, 3, 4, 5]))
        mode = str(rng.choice(MODES))
        # row bands of the host-buffer pipelines (float32 host images: radial, projective, both)
        post.config["bands"] = int(rng.choice([0, 0, 1, 2, 3, 5, 9, 32]))
        try:
            if kind == "radial":
                got = post.unwarp_image_backward(mat, xc, yc, fact, order=order, mode=mode)
                want = osp.unwarp_image_backward(mat, xc, yc, fact, order, mode)
            elif kind == "persp":
                coef = [1 + rng.normal() * 0.05, rng.normal() * 0.05, rng.normal() * 5,
                        rng.normal() * 0.05, 1 + rng.normal() * 0.05, rng.normal() * 5,
                        rng.normal() * 1e-4, rng.normal() * 1e-4]
                coef = [float(c) for c in coef]
                got = post.correct_perspective_image(mat, coef, order=order, mode=mode)
                want = osp.correct_perspective_image(mat, coef, order, mode)
            elif kind == "slice":
                if dt == "float64":
                    continue
                d = int(rng.integers(1, 5))
                stack = np.stack([np.roll(mat, k, axis=1) for k in range(d)])
                index = int(rng.integers(0, h))
                got = post.unwarp_slice_backward(stack, xc, yc, fact, index)
                want = orc.unwarp_slice_backward(stack, xc, yc, fact, index)
            elif kind == "color":
                if dt == "float64" or order > 1:
                    continue
                chan = int(rng.integers(1, 5))
                frame = np.stack([np.roll(mat, 3 * k, axis=0) for k in range(chan)], axis=2)
                pad = [0, int(rng.integers(0, 9)), tuple(int(v) for v in rng.integers(0, 7, 4))][int(rng.integers(0, 3))]
                got = util.unwarp_color_image_backward(frame, xc, yc, fact, order=order, pad=pad)
                want = orc.unwarp_color_image_backward(frame, xc, yc, fact, order=order, pad=pad)
            elif kind == "both":
                if dt == "float64" or order > 1:
                    continue
                coef = [float(c) for c in (1 + rng.normal() * 0.05, rng.normal() * 0.05, rng.normal() * 5,
                                           rng.normal() * 0.05, 1 + rng.normal() * 0.05, rng.normal() * 5,
                                           rng.normal() * 1e-4, rng.normal() * 1e-4)]
                got = post.unwarp_image_backward_perspective(mat, xc, yc, fact, coef, order=order)
                want = orc.unwarp_image_backward_perspective(mat, xc, yc, fact, coef, order=order)
            else:
                if dt == "float64" or h < 2:
                    continue
                # the reference's row window (postprocessing.py:289-301) only makes sense for maps
                # that keep the row order: use a mild model for the chunk cases
                fact = [float(rng.uniform(0.9, 1.1))] + [float(rng.normal() * 0.03 / scale ** i) for i in range(1, nt)]
                d = int(rng.integers(1, 5))
                stack = np.stack([np.roll(mat, k, axis=1) for k in range(d)])
                a = int(rng.integers(0, h))
                b = int(rng.integers(a, h))
                y0, y1 = orc.chunk_row_window(h, w, xc, yc, fact, a, b)
                if y1 <= y0:      # empty window: undefined in the reference, both sides raise
                    continue
                got = post.unwarp_chunk_slices_backward(stack, xc, yc, fact, a, b)
                want = orc.unwarp_chunk_slices_backward(stack, xc, yc, fact, a, b)
            same = got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want, equal_nan=True)
        except Exception as exc:                      # report and go on
            same = False
            got = want = None
            print("EXC", type(exc).__name__, str(exc)[:200])
        if not same:
            bad += 1
            nd = -1 if got is None or got.shape != want.shape else int(np.count_nonzero(got != want))
            print("MISMATCH case %d: %s %s %dx%d order %d mode %s nt %d xc %.3f yc %.3f: %d samples differ"
                  % (it, kind, dt, h, w, order, mode, nt, xc, yc, nd), flush=True)
    post.config["bands"] = 0
    print("fuzz: %d cases, %d not bit-identical" % (n, bad))
    return bad


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 300, int(sys.argv[2]) if len(sys.argv) > 2 else 1)
