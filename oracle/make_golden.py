"""
make_golden.py -- generate ``tests/golden/*.npz`` from the REAL reference.

Run only in the build container (needs ``/root/reference``; the GPU box does
not have it):

    python oracle/make_golden.py

For every case it (1) calls the unmodified reference function imported from
``/root/reference/discorpy/post/postprocessing.py``, (2) calls the NumPy
restatement in ``oracle/oracle_np.py`` on the same input and asserts the two
agree BIT FOR BIT, and (3) stores the case parameters, the input seed and the
reference output.  ``tests/test_oracle.py`` re-checks the restatement against
these stored outputs wherever the repo travels; ``tests/test_gpu_parity.py``
checks the CUDA path against them.

Inputs are regenerated from the seed by ``make_input`` below (also imported by
the tests), so the fixtures only hold outputs.
"""
import os
import sys
import json
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")

COEF_DOT_05 = dict(  # data/coef_dot_05.txt (values only; parsed, not copied)
    xc=588.692801577, yc=462.092631791,
    fact=[1.00227490554, -2.99523692178e-05, 8.99519088e-08,
          -1.57066461911e-10, 8.08880211618e-14])


def make_input(kind, shape, seed, dtype="float32"):
    """Deterministic synthetic inputs shared by generator and tests."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        a = rng.random(shape, dtype=np.float32)
    elif kind == "noise255":
        a = np.floor(rng.random(shape, dtype=np.float32) * 256.0)
    elif kind == "smooth":
        yy, xx = np.meshgrid(np.arange(shape[-2]), np.arange(shape[-1]),
                             indexing="ij")
        a = np.sin(0.05 * xx) * np.cos(0.033 * yy) + 1.5
        if len(shape) == 3:
            a = a[None] * (1.0 + 0.1 * np.arange(shape[0])[:, None, None])
        a = np.float32(a)
    elif kind == "signed":
        a = rng.standard_normal(shape, dtype=np.float32) * 1000.0
    else:
        raise ValueError(kind)
    dtype = np.dtype(dtype)
    if dtype.kind in "iu":
        info = np.iinfo(dtype)
        a = np.clip(a * (info.max / max(1.0, float(np.abs(a).max()))),
                    info.min, info.max)
    return np.ascontiguousarray(a.astype(dtype))


def image_cases():
    f2 = [1.0, 3.0e-3]
    f5 = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
    strong = [1.02, -1.5e-3, 4e-5, -2e-7]
    cases = []
    n = 0
    for shape, kind in (((64, 64), "noise"), ((37, 53), "noise255"),
                        ((128, 200), "smooth"), ((1, 17), "noise"),
                        ((23, 1), "noise"), ((96, 131), "signed")):
        for (xc, yc, fact) in ((shape[1] // 2, shape[0] // 2, f2),
                               (shape[1] / 2 + 0.37, shape[0] / 2 - 1.81, f5),
                               (-7.25, shape[0] + 3.5, strong),
                               (shape[1] / 3.0, shape[0] / 1.7, [0.8])):
            for order in (0, 1):
                cases.append(dict(id="img%03d" % n, shape=list(shape),
                                  kind=kind, seed=100 + n, xc=xc, yc=yc,
                                  fact=fact, order=order, dtype="float32"))
                n += 1
    # the coefficient file of BASELINE config 1 on a crop-sized image
    for order in (0, 1):
        cases.append(dict(id="img%03d" % n, shape=[200, 320], kind="noise255",
                          seed=100 + n, xc=COEF_DOT_05["xc"] / 4,
                          yc=COEF_DOT_05["yc"] / 4, fact=COEF_DOT_05["fact"],
                          order=order, dtype="float32"))
        n += 1
    # other dtypes (oracle-only until the CUDA path grows them)
    for dtype in ("float64", "uint8", "uint16", "int16"):
        for order in (0, 1):
            cases.append(dict(id="img%03d" % n, shape=[48, 61],
                              kind="signed" if dtype == "int16" else "noise",
                              seed=100 + n, xc=30.3, yc=22.9, fact=f2,
                              order=order, dtype=dtype))
            n += 1
    return cases


def stack_cases():
    f2 = [1.0, 3.0e-3]
    f5 = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
    f3n = [1.0, -4.0e-3, 1e-5]
    cases = []
    n = 0
    for shape, kind in (((10, 64, 64), "smooth"), ((3, 45, 77), "noise"),
                        ((5, 96, 130), "noise255")):
        d, h, w = shape
        for (xc, yc, fact) in ((w // 2, h // 2, f2),
                               (w / 2 + 0.3, h / 2 - 0.7, f5),
                               (w / 2 - 3.1, h / 2 + 2.2, f3n)):
            for index in (0, h // 2, h - 1, h // 3):
                cases.append(dict(id="slc%03d" % n, fn="slice",
                                  shape=list(shape), kind=kind, seed=500 + n,
                                  xc=xc, yc=yc, fact=fact, index=int(index)))
                n += 1
            for (start, stop) in ((0, h - 1), (h // 2 - 5, h // 2 + 5),
                                  (h // 4, h // 4)):
                cases.append(dict(id="chk%03d" % n, fn="chunk",
                                  shape=list(shape), kind=kind, seed=500 + n,
                                  xc=xc, yc=yc, fact=fact, start=int(start),
                                  stop=int(stop)))
                n += 1
    return cases


def persp_cases():
    cases = []
    n = 0
    coefs = ([1.02, 0.01, -1.5, 0.005, 1.01, -0.8, 8e-5, -5e-5],
             [0.9, -0.05, 3.0, 0.04, 0.95, 2.0, -3e-4, 2e-4],
             [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0],
             # 35 degree rotation about the centre, x1.5 (demo_07-like)
             [0.5461, 0.3824, -10.0, -0.3824, 0.5461, 25.0, 0.0, 0.0])
    fact = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
    for shape, kind in (((64, 64), "noise"), ((57, 91), "smooth"),
                        ((120, 80), "noise255")):
        for c in coefs:
            for order in (0, 1):
                cases.append(dict(id="per%03d" % n, fn="persp",
                                  shape=list(shape), kind=kind, seed=900 + n,
                                  coef=c, order=order))
                n += 1
            cases.append(dict(id="cmb%03d" % n, fn="combined",
                              shape=list(shape), kind=kind, seed=900 + n,
                              coef=c, order=1, xc=shape[1] / 2 + 0.4,
                              yc=shape[0] / 2 - 0.9, fact=fact))
            n += 1
    return cases


def dtype_and_color_cases():
    """Rows SURVEY.md 8(f) ranks 1 and 2: colour frames (H, W, C) through
    ``util.unwarp_color_image_backward`` and integer images through the other
    functions of the path (output dtype = input dtype)."""
    f2 = [1.0, 3.0e-3]
    f5 = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
    coef = [1.02, 0.01, -1.5, 0.005, 1.01, -0.8, 8e-5, -5e-5]
    cases = []
    n = 0
    for shape, dtype, kind in (((40, 56, 3), "uint8", "noise"),
                               ((33, 47, 4), "float32", "noise"),
                               ((64, 50, 3), "uint16", "noise"),
                               ((29, 31), "uint8", "noise"),
                               ((36, 44, 2), "int16", "signed")):
        for pad, pad_mode in ((0, "constant"), (3, "edge"),
                              ((1, 2, 3, 4), "reflect")):
            for order in (0, 1):
                cases.append(dict(id="col%03d" % n, fn="color",
                                  shape=list(shape), kind=kind, seed=1300 + n,
                                  dtype=dtype, xc=shape[1] / 2 + 0.6,
                                  yc=shape[0] / 2 - 1.3, fact=f5 if n % 2 else f2,
                                  order=order,
                                  pad=list(pad) if isinstance(pad, tuple) else pad,
                                  pad_mode=pad_mode))
                n += 1
    for dtype, kind in (("uint8", "noise"), ("uint16", "noise"),
                        ("int16", "signed"), ("int8", "signed")):
        cases.append(dict(id="dty%03d" % n, fn="chunk", shape=[3, 52, 71],
                          kind=kind, seed=1300 + n, dtype=dtype, xc=36.2,
                          yc=25.4, fact=f2, start=5, stop=40))
        n += 1
        for order in (0, 1):
            cases.append(dict(id="dty%03d" % n, fn="persp", shape=[57, 63],
                              kind=kind, seed=1300 + n, dtype=dtype, coef=coef,
                              order=order))
            n += 1
        cases.append(dict(id="dty%03d" % n, fn="combined", shape=[57, 63],
                          kind=kind, seed=1300 + n, dtype=dtype, coef=coef,
                          order=1, xc=31.9, yc=27.6, fact=f5))
        n += 1
        cases.append(dict(id="dty%03d" % n, fn="slice", shape=[3, 52, 71],
                          kind=kind, seed=1300 + n, dtype=dtype, xc=36.2,
                          yc=25.4, fact=f2, index=17))
        n += 1
    return cases


def main():
    sys.path.insert(0, "/root/reference")
    sys.path.insert(0, ROOT)
    import discorpy.post.postprocessing as ref   # the real reference
    import discorpy.util.utility as ref_util
    from oracle import oracle_np as orc
    os.makedirs(GOLDEN, exist_ok=True)

    def check(a, b, cid):
        assert a.dtype == b.dtype and a.shape == b.shape, (cid, a.dtype,
                                                            b.dtype)
        if not np.array_equal(a, b, equal_nan=True):
            bad = np.count_nonzero(a != b)
            raise AssertionError("%s: oracle != reference on %d px" % (cid,
                                                                       bad))

    out = {}
    meta = []
    for c in image_cases():
        mat = make_input(c["kind"], tuple(c["shape"]), c["seed"], c["dtype"])
        r = ref.unwarp_image_backward(mat, c["xc"], c["yc"], c["fact"],
                                      order=c["order"])
        o = orc.unwarp_image_backward(mat, c["xc"], c["yc"], c["fact"],
                                      order=c["order"])
        check(r, o, c["id"])
        out[c["id"]] = r
        c["fn"] = "image"
        meta.append(c)
    for c in stack_cases():
        mat = make_input(c["kind"], tuple(c["shape"]), c["seed"])
        if c["fn"] == "slice":
            r = ref.unwarp_slice_backward(mat, c["xc"], c["yc"], c["fact"],
                                          c["index"])
            o = orc.unwarp_slice_backward(mat, c["xc"], c["yc"], c["fact"],
                                          c["index"])
        else:
            r = ref.unwarp_chunk_slices_backward(mat, c["xc"], c["yc"],
                                                 c["fact"], c["start"],
                                                 c["stop"])
            o = orc.unwarp_chunk_slices_backward(mat, c["xc"], c["yc"],
                                                 c["fact"], c["start"],
                                                 c["stop"])
        check(r, o, c["id"])
        out[c["id"]] = r
        meta.append(c)
    for c in persp_cases():
        mat = make_input(c["kind"], tuple(c["shape"]), c["seed"])
        if c["fn"] == "persp":
            r = ref.correct_perspective_image(mat, c["coef"], order=c["order"])
            o = orc.correct_perspective_image(mat, c["coef"], order=c["order"])
            # the map_index= route must give the same thing
            mi = ref._generate_perspective_map(mat, c["coef"])
            r2 = ref.correct_perspective_image(mat, c["coef"],
                                               order=c["order"], map_index=mi)
            check(r, r2, c["id"] + "/map_index")
        else:
            t = ref.unwarp_image_backward(mat, c["xc"], c["yc"], c["fact"])
            r = ref.correct_perspective_image(t, c["coef"])
            o = orc.unwarp_image_backward_perspective(
                mat, c["xc"], c["yc"], c["fact"], c["coef"])
        check(r, o, c["id"])
        out[c["id"]] = r
        meta.append(c)
    for c in dtype_and_color_cases():
        mat = make_input(c["kind"], tuple(c["shape"]), c["seed"], c["dtype"])
        fn = c["fn"]
        if fn == "color":
            pad = tuple(c["pad"]) if isinstance(c["pad"], list) else c["pad"]
            r = ref_util.unwarp_color_image_backward(
                mat, c["xc"], c["yc"], c["fact"], order=c["order"], pad=pad,
                pad_mode=c["pad_mode"])
            o = orc.unwarp_color_image_backward(
                mat, c["xc"], c["yc"], c["fact"], order=c["order"], pad=pad,
                pad_mode=c["pad_mode"])
        elif fn == "chunk":
            r = ref.unwarp_chunk_slices_backward(mat, c["xc"], c["yc"], c["fact"],
                                                 c["start"], c["stop"])
            o = orc.unwarp_chunk_slices_backward(mat, c["xc"], c["yc"], c["fact"],
                                                 c["start"], c["stop"])
        elif fn == "slice":
            r = ref.unwarp_slice_backward(mat, c["xc"], c["yc"], c["fact"],
                                          c["index"])
            o = orc.unwarp_slice_backward(mat, c["xc"], c["yc"], c["fact"],
                                          c["index"])
        elif fn == "persp":
            r = ref.correct_perspective_image(mat, c["coef"], order=c["order"])
            o = orc.correct_perspective_image(mat, c["coef"], order=c["order"])
        else:
            t = ref.unwarp_image_backward(mat, c["xc"], c["yc"], c["fact"])
            r = ref.correct_perspective_image(t, c["coef"])
            o = orc.unwarp_image_backward_perspective(
                mat, c["xc"], c["yc"], c["fact"], c["coef"])
        check(np.ascontiguousarray(r), np.ascontiguousarray(o), c["id"])
        out[c["id"]] = np.ascontiguousarray(r)
        meta.append(c)
    np.savez_compressed(os.path.join(GOLDEN, "reference_outputs.npz"), **out)
    with open(os.path.join(GOLDEN, "cases.json"), "w") as f:
        json.dump(dict(
            generated_by="oracle/make_golden.py",
            reference="DiamondLightSource/discorpy 1.7.0 @ /root/reference",
            numpy=np.__version__,
            scipy=__import__("scipy").__version__,
            cases=meta), f, indent=1)
    npx = sum(int(v.size) for v in out.values())
    print("wrote %d cases, %d output pixels; oracle == reference bit-for-bit"
          % (len(meta), npx))


if __name__ == "__main__":
    main()
