"""The oracle against the reference's outputs (CPU only).

tests/golden/reference_outputs.npz was produced by oracle/make_golden.py from
the unmodified reference in the build container; here the NumPy and the C
restatement are re-checked against it wherever the repo travels.
"""
import os

import numpy as np
import pytest

from conftest import load_cases, golden_output
from oracle import oracle_np as orc
from oracle import oracle_c
from oracle.make_golden import make_input

CASES = load_cases()


def _run_np(c, mat):
    fn = c["fn"]
    if fn == "image":
        return orc.unwarp_image_backward(mat, c["xc"], c["yc"], c["fact"],
                                         order=c["order"])
    if fn == "slice":
        return orc.unwarp_slice_backward(mat, c["xc"], c["yc"], c["fact"],
                                         c["index"])
    if fn == "chunk":
        return orc.unwarp_chunk_slices_backward(mat, c["xc"], c["yc"],
                                                c["fact"], c["start"],
                                                c["stop"])
    if fn == "persp":
        return orc.correct_perspective_image(mat, c["coef"], order=c["order"])
    if fn == "combined":
        return orc.unwarp_image_backward_perspective(
            mat, c["xc"], c["yc"], c["fact"], c["coef"])
    if fn == "color":
        pad = tuple(c["pad"]) if isinstance(c["pad"], list) else c["pad"]
        return np.ascontiguousarray(orc.unwarp_color_image_backward(
            mat, c["xc"], c["yc"], c["fact"], order=c["order"], pad=pad,
            pad_mode=c["pad_mode"]))
    raise AssertionError(fn)


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_numpy_oracle_matches_reference_bit_for_bit(case):
    mat = make_input(case["kind"], tuple(case["shape"]), case["seed"],
                     case.get("dtype", "float32"))
    got = _run_np(case, mat)
    want = golden_output(case["id"])
    assert got.dtype == want.dtype and got.shape == want.shape
    assert np.array_equal(got, want, equal_nan=True)


F32 = [c for c in CASES if c.get("dtype", "float32") == "float32"
       and c["fn"] != "color"]


@pytest.mark.parametrize("case", F32, ids=[c["id"] for c in F32])
def test_c_oracle_matches_reference(case):
    mat = make_input(case["kind"], tuple(case["shape"]), case["seed"])
    fn = case["fn"]
    want = golden_output(case["id"])
    if fn == "image":
        got = oracle_c.unwarp_image_backward(mat, case["xc"], case["yc"],
                                             case["fact"], case["order"])
    elif fn == "slice":
        h, w = mat.shape[1:]
        yd, _ = orc.radial_coords_row(h, w, case["xc"], case["yc"],
                                      case["fact"], case["index"])
        ylo = int(np.floor(yd.min()))
        yhi = int(np.ceil(yd.max()))
        got = oracle_c.unwarp_stack_backward(
            mat, case["xc"], case["yc"], case["fact"], case["index"], 1,
            coord_round=False, ylo=ylo, yhi=yhi)[:, 0, :]
    elif fn == "chunk":
        h, w = mat.shape[1:]
        ylo, yend = orc.chunk_row_window(h, w, case["xc"], case["yc"],
                                         case["fact"], case["start"],
                                         case["stop"])
        got = oracle_c.unwarp_stack_backward(
            mat, case["xc"], case["yc"], case["fact"], case["start"],
            case["stop"] - case["start"] + 1, coord_round=True, ylo=ylo,
            yhi=yend - 1)
    elif fn == "persp":
        got = oracle_c.correct_perspective_image(mat, case["coef"],
                                                 case["order"])
    else:
        tmp = oracle_c.unwarp_image_backward(mat, case["xc"], case["yc"],
                                             case["fact"], 1)
        got = oracle_c.correct_perspective_image(tmp, case["coef"], 1)
    assert got.shape == want.shape
    # libm pow vs NumPy power may flip a float32 coordinate very rarely
    # (oracle_c.c header); on these fixtures it does not.
    assert np.array_equal(got, want)


def test_sampler_matches_scipy_map_coordinates():
    """The vectorised sampler restates SciPy's C loop: compare directly."""
    from scipy.ndimage import map_coordinates
    rng = np.random.default_rng(7)
    mat = rng.standard_normal((83, 117)).astype(np.float32) * 100
    yd = np.float32(rng.random(20000) * 82)
    xd = np.float32(rng.random(20000) * 116)
    yd[:50] = np.float32(np.arange(50) % 83)        # exact integers
    xd[:50] = 116.0                                  # last column
    yd[50:100] = 82.0                                # last row
    xd[100:150] = np.float32(np.arange(50)) + np.float32(0.5)   # halves
    for order in (0, 1):
        want = map_coordinates(mat, (yd, xd), order=order, mode="reflect")
        got = orc.sample(mat, yd, xd, order)
        assert np.array_equal(got, want)
    for mode in ("constant", "nearest", "mirror", "wrap", "grid-wrap"):
        want = map_coordinates(mat, (yd, xd), order=1, mode=mode)
        assert np.array_equal(orc.sample(mat, yd, xd, 1), want), mode


@pytest.mark.skipif(not os.path.isdir("/root/reference/discorpy"),
                    reason="the reference tree only exists in the build container")
def test_numpy_oracle_vs_live_reference_large():
    import importlib
    import sys
    sys.path.insert(0, "/root/reference")
    try:
        ref = importlib.import_module("discorpy.post.postprocessing")
    finally:
        sys.path.remove("/root/reference")
    rng = np.random.default_rng(11)
    mat = rng.random((700, 900), dtype=np.float32)
    fact = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
    for order in (0, 1):
        a = ref.unwarp_image_backward(mat, 451.3, 340.9, fact, order=order)
        b = orc.unwarp_image_backward(mat, 451.3, 340.9, fact, order=order)
        assert np.array_equal(a, b)
    coef = [1.02, 0.01, -15.0, 0.005, 1.01, -8.0, 8e-6, -5e-6]
    assert np.array_equal(ref.correct_perspective_image(mat, coef),
                          orc.correct_perspective_image(mat, coef))


def test_chunk_rows_outside_the_reference_window():
    """unwarp_chunk_slices_backward crops every slice to a row window taken from the first and last
    row of the chunk (postprocessing.py:289-301); rows in between can sample outside it and SciPy
    reflects them into the crop.  Golden outputs of the real reference
    (oracle/make_golden_chunk_window.py): the oracle restates the reflection bit for bit, and the
    host-side test of the product flags exactly these cases."""
    import os
    import discorpy_b200.post.postprocessing as post
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "chunk_window.npz"))
    for i in range(int(z["n"])):
        stack, par, ref = z["stack%d" % i], z["par%d" % i], z["ref%d" % i]
        xc, yc, a, b, fact = float(par[0]), float(par[1]), int(par[2]), int(par[3]), [float(v) for v in par[4:]]
        got = orc.unwarp_chunk_slices_backward(stack, xc, yc, fact, a, b)
        assert got.dtype == ref.dtype and np.array_equal(got, ref)
        h, w = stack.shape[1:]
        y0, y1 = orc.chunk_row_window(h, w, xc, yc, fact, a, b)
        assert post._rows_leave_window(h, w, xc, yc, fact, a, b, y0, y1)
    cc = np.linspace(-200, 200, 9001).astype(np.float32)
    for n in (1, 2, 5, 30):
        assert np.array_equal(post._fold_coordinate(cc, n, "reflect"), orc.reflect_coordinate(cc, n))
    # the BASELINE config-4 model keeps every row inside its window
    f4 = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
    assert not post._rows_leave_window(2560, 2560, 1283.4, 1275.9, f4, 100, 300,
                                       *orc.chunk_row_window(2560, 2560, 1283.4, 1275.9, f4, 100, 300))


def test_chunk_with_an_empty_row_window_is_refused():
    """A model that maps the last chunk row above the first gives the reference an empty slice (SciPy
    then reads past it: undefined values); the oracle and the product's host check both refuse."""
    import discorpy_b200.post.postprocessing as post
    fact = [0.7121505980847279, -0.0034137040107838647, -2.6659489737081743e-05,
            -2.2097870183732417e-08, -1.5716877223813336e-10, -7.430580089833948e-14]
    stack = np.zeros((2, 168, 6), dtype=np.float32)
    y0, y1 = orc.chunk_row_window(168, 6, 5.782255233984425, -9.275425402650226, fact, 10, 95)
    assert y1 <= y0
    with pytest.raises(ValueError):
        orc.unwarp_chunk_slices_backward(stack, 5.782255233984425, -9.275425402650226, fact, 10, 95)
    with pytest.raises(ValueError):     # raised before any GPU work
        post.unwarp_chunk_slices_backward(stack, 5.782255233984425, -9.275425402650226, fact, 10, 95)


@pytest.mark.skipif(not os.path.isdir("/root/reference/discorpy"),
                    reason="the reference tree only exists in the build container")
def test_oracle_against_the_real_reference_randomised():
    """Where /root/reference is present (the build container, not the GPU box): 200 random cases of
    tests/fuzz_oracle_vs_reference.py -- the reference's own functions against the oracle, every
    function, order, mode and dtype -- must be bit-identical."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    res = subprocess.run([sys.executable, os.path.join(here, "fuzz_oracle_vs_reference.py"), "200", "3"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    assert "0 not bit-identical" in res.stdout


def test_forward_oracle_equals_reference_golden():
    """oracle_np.unwarp_image_forward against the real reference's outputs
    (oracle/make_golden_forward.py, postprocessing.py:151-185)."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "forward.npz"))
    k = 0
    while "in%d" % k in z:
        par = z["par%d" % k]
        got = orc.unwarp_image_forward(z["in%d" % k], par[0], par[1], list(par[2:]))
        assert np.array_equal(got, z["out%d" % k]), k
        k += 1
    assert k == 4


def test_cfg1_real_image_oracle_equals_reference_golden():
    """BASELINE config 1 on the reference's own files (tests/golden/cfg1, written by
    oracle/make_golden_cfg1.py with the real reference): the oracle reproduces the full result
    (SHA-256) and the stored subsample."""
    import hashlib
    import json
    import os
    from PIL import Image
    d = os.path.join(os.path.dirname(__file__), "golden", "cfg1")
    meta = json.load(open(os.path.join(d, "meta.json")))
    mat = np.array(Image.open(os.path.join(d, "dot_pattern_01.jpg")), dtype=np.float32)
    if hashlib.sha256(mat.tobytes()).hexdigest() != meta["input_sha256"]:
        pytest.skip("this PIL / libjpeg decodes the JPEG differently from the build container")
    sub = np.load(os.path.join(d, "reference_subsample.npz"))
    got = orc.unwarp_image_backward(mat, meta["xcenter"], meta["ycenter"], meta["list_fact"], order=1)
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == meta["output_sha256_order1"]
    assert np.array_equal(got[::7, ::7], sub["order1"])
