"""
make_golden_spline.py -- golden vectors for spline orders 2..5 and float64
images, generated through the REAL reference (``/root/reference`` + the installed
SciPy).  Build container only:

    python oracle/make_golden_spline.py

For every case the unmodified reference function is called, the restatement in
``oracle/oracle_spline.py`` is asserted BIT-IDENTICAL to it, and the reference
output is stored in ``tests/golden/spline_reference_outputs.npz`` (parameters
in ``spline_cases.json``; inputs are regenerated from seeds by
``make_golden.make_input``).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.make_golden import make_input, GOLDEN          # noqa: E402
from oracle import oracle_spline as osp                    # noqa: E402

MODES = osp.MODES
PERSP = [1.02, 0.01, -3.0, 0.005, 1.01, -2.0, 8e-5, -5e-5]


def cases():
    f2 = [1.0, 3.0e-3]
    f5 = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
    out, n = [], 0

    def add(**kw):
        nonlocal n
        kw["id"] = "spl%03d" % n
        kw["seed"] = 700 + n
        n += 1
        out.append(kw)
    # every order x every mode on a small noisy float32 image (edges matter here)
    for order in (2, 3, 4, 5):
        for mode in MODES:
            add(fn="image", shape=[41, 57], kind="noise", dtype="float32", xc=27.3,
                yc=19.6, fact=f2, order=order, mode=mode)
    # default mode, other shapes / models / dtypes
    for order in (2, 3, 4, 5):
        add(fn="image", shape=[96, 131], kind="signed", dtype="float32", xc=131 / 2 + 0.37,
            yc=96 / 2 - 1.81, fact=f5, order=order, mode="reflect")
        add(fn="image", shape=[64, 64], kind="smooth", dtype="float32", xc=-7.25, yc=67.5,
            fact=[1.02, -1.5e-3, 4e-5, -2e-7], order=order, mode="reflect")
    for dtype in ("float64", "uint8", "int8", "uint16", "int16"):
        for order, mode in ((3, "reflect"), (5, "mirror"), (2, "nearest"), (3, "grid-constant")):
            add(fn="image", shape=[48, 61], kind="signed" if dtype.startswith("int") else "noise",
                dtype=dtype, xc=30.3, yc=22.9, fact=f2, order=order, mode=mode)
    # float64 images at order 0 / 1 (float64 in, float64 out)
    for order in (0, 1):
        add(fn="image", shape=[48, 61], kind="signed", dtype="float64", xc=30.3, yc=22.9,
            fact=f5, order=order, mode="reflect")
    # degenerate shapes
    for shape in ([1, 17], [23, 1], [2, 2], [3, 9]):
        add(fn="image", shape=shape, kind="noise", dtype="float32", xc=shape[1] / 3.0,
            yc=shape[0] / 1.7, fact=[0.8], order=3, mode="reflect")
        add(fn="image", shape=shape, kind="noise", dtype="float32", xc=shape[1] / 3.0,
            yc=shape[0] / 1.7, fact=[0.8], order=4, mode="grid-wrap")
    # perspective, with and without a precomputed map
    for order, mode in ((3, "reflect"), (2, "mirror"), (5, "grid-wrap"), (4, "nearest")):
        add(fn="persp", shape=[50, 70], kind="noise255", dtype="float32", coef=PERSP,
            order=order, mode=mode, use_map=False)
    add(fn="persp", shape=[50, 70], kind="noise255", dtype="uint8", coef=PERSP, order=3,
        mode="reflect", use_map=True)
    add(fn="persp", shape=[50, 70], kind="noise", dtype="float64", coef=PERSP, order=1,
        mode="reflect", use_map=False)
    # colour frames through util.unwarp_color_image_backward
    add(fn="color", shape=[40, 52, 3], kind="noise255", dtype="uint8", xc=25.1, yc=21.7,
        fact=f2, order=3, mode="reflect")
    add(fn="color", shape=[40, 52, 3], kind="noise", dtype="float32", xc=25.1, yc=21.7,
        fact=f2, order=3, mode="mirror")
    return out


def main():
    sys.path.insert(0, "/root/reference")
    import discorpy.post.postprocessing as ref_post
    import discorpy.util.utility as ref_util
    cs = cases()
    outs = {}
    for c in cs:
        mat = make_input(c["kind"], tuple(c["shape"]), c["seed"], c["dtype"])
        if c["fn"] == "image":
            ref = ref_post.unwarp_image_backward(mat, c["xc"], c["yc"], c["fact"],
                                                 order=c["order"], mode=c["mode"])
            got = osp.unwarp_image_backward(mat, c["xc"], c["yc"], c["fact"], c["order"],
                                            c["mode"])
        elif c["fn"] == "persp":
            mi = ref_post._generate_perspective_map(mat, c["coef"]) if c["use_map"] else None
            ref = ref_post.correct_perspective_image(mat, c["coef"], order=c["order"],
                                                     mode=c["mode"], map_index=mi)
            got = osp.correct_perspective_image(mat, c["coef"], c["order"], c["mode"], mi)
        else:
            ref = ref_util.unwarp_color_image_backward(mat, c["xc"], c["yc"], c["fact"],
                                                       order=c["order"], mode=c["mode"])
            got = osp.unwarp_color_image_backward(mat, c["xc"], c["yc"], c["fact"],
                                                  c["order"], c["mode"])
        ref = np.ascontiguousarray(ref)
        assert ref.dtype == got.dtype == mat.dtype, (c["id"], ref.dtype, got.dtype)
        same = np.array_equal(ref, got, equal_nan=True)
        assert same, "%s: oracle differs from the reference in %d samples" % (
            c["id"], int(np.count_nonzero(ref != got)))
        outs[c["id"]] = ref
    np.savez_compressed(os.path.join(GOLDEN, "spline_reference_outputs.npz"), **outs)
    with open(os.path.join(GOLDEN, "spline_cases.json"), "w") as f:
        json.dump(cs, f, indent=0)
    print("%d spline cases, oracle == reference bit for bit" % len(cs))


if __name__ == "__main__":
    main()
