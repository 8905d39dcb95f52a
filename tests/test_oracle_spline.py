"""CPU: the spline oracle (oracle/oracle_spline.py) against the reference's golden
outputs (tests/golden/spline_*, generated through the real discorpy functions by
oracle/make_golden_spline.py) and, where SciPy is importable, live against
scipy.ndimage -- bit for bit in both cases."""
import json
import os
from decimal import Decimal, getcontext

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import oracle_spline as osp
from oracle.make_golden import make_input

with open(os.path.join(GOLDEN_DIR, "spline_cases.json")) as f:
    CASES = json.load(f)
_OUT = None


def golden(case_id):
    global _OUT
    if _OUT is None:
        _OUT = np.load(os.path.join(GOLDEN_DIR, "spline_reference_outputs.npz"))
    return _OUT[case_id]


def run_oracle(c, mat):
    if c["fn"] == "image":
        return osp.unwarp_image_backward(mat, c["xc"], c["yc"], c["fact"], c["order"], c["mode"])
    if c["fn"] == "persp":
        mi = None
        if c["use_map"]:
            from oracle import oracle_np
            mi = oracle_np.persp_coords(mat.shape[0], mat.shape[1], c["coef"])
        return osp.correct_perspective_image(mat, c["coef"], c["order"], c["mode"], mi)
    return osp.unwarp_color_image_backward(mat, c["xc"], c["yc"], c["fact"], c["order"], c["mode"])


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_oracle_matches_reference_golden(case):
    mat = make_input(case["kind"], tuple(case["shape"]), case["seed"], case["dtype"])
    got = np.ascontiguousarray(run_oracle(case, mat))
    want = golden(case["id"])
    assert got.dtype == want.dtype and got.shape == want.shape
    assert np.array_equal(got, want, equal_nan=True)


def test_poles_are_the_correctly_rounded_radicals():
    """sqrt(8)-3, sqrt(3)-2, ... evaluated with 60 digits and rounded once."""
    getcontext().prec = 60
    D = Decimal
    exact = {
        2: [D(8).sqrt() - 3],
        3: [D(3).sqrt() - 2],
        4: [(D(664) - D(438976).sqrt()).sqrt() + D(304).sqrt() - 19,
            (D(664) + D(438976).sqrt()).sqrt() - D(304).sqrt() - 19],
        5: [(D("67.5") - D("4436.25").sqrt()).sqrt() + D("26.25").sqrt() - D("6.5"),
            (D("67.5") + D("4436.25").sqrt()).sqrt() - D("26.25").sqrt() - D("6.5")],
    }
    for order, vals in exact.items():
        for z, v in zip(osp.poles(order), vals):
            assert z == float(v), (order, z.hex(), float(v).hex())


def test_oracle_matches_live_scipy():
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(11)
    for shape in ((19, 31), (2, 5), (40, 3)):
        for dt in (np.float32, np.float64, np.uint8, np.int16):
            if np.dtype(dt).kind == "f":
                mat = (rng.random(shape) * 300 - 50).astype(dt)
            else:
                info = np.iinfo(dt)
                mat = rng.integers(info.min, info.max, shape, dtype=dt, endpoint=True)
            h, w = shape
            n = 500
            yd = np.float32(rng.random(n) * (h - 1))
            xd = np.float32(rng.random(n) * (w - 1))
            yd[:20], xd[20:40], yd[40:60], xd[60:80] = 0, 0, h - 1, w - 1
            for order in (2, 3, 4, 5):
                for mode in osp.MODES:
                    ref = ndi.spline_filter(mat, order, output=np.float64, mode=mode)
                    assert np.array_equal(ref, osp.spline_filter(mat, order, mode))
                    ref = ndi.map_coordinates(mat, (yd, xd), order=order, mode=mode)
                    assert np.array_equal(ref, osp.sample_spline(mat, yd, xd, order, mode))
