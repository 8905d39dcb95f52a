"""Host-side behaviour that needs no GPU: the C ABI loads and exports what the
header declares, argument validation mirrors the reference, the point-list
helpers agree with the reference's arithmetic."""
import ctypes
import os
import re

import numpy as np
import pytest

import discorpy_b200
from discorpy_b200 import _cabi
import discorpy_b200.post.postprocessing as post

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "discorpy_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dcb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.load()
    names = _declared_symbols()
    assert len(names) >= 35
    for name in names:
        assert hasattr(lib, name), name
    # and the ctypes table covers the header (plus dcb_last_error)
    assert set(names) == set(_cabi.SIGNATURES) | {"dcb_last_error"}


def test_version_and_struct_layout():
    lib = _cabi.load()
    assert lib.dcb_version() == 1            # 0 * 1000 + 1
    assert ctypes.sizeof(_cabi.Radial) == 8 + 8 + 4 + 4 + 16 * 8
    assert ctypes.sizeof(_cabi.Persp) == 64
    assert ctypes.sizeof(_cabi.Options) == 16
    assert lib.dcb_last_error() is not None


def test_argument_errors_come_before_any_device_work():
    # reference :211-212 / :281-288 / :486-487 and scipy's messages
    with pytest.raises(ValueError, match="Input must be a 3D data"):
        post.unwarp_slice_backward(np.zeros((4, 4), np.float32), 1, 1, [1.0], 2)
    with pytest.raises(ValueError, match="Input must be a 3D data"):
        post.unwarp_chunk_slices_backward(np.zeros((4, 4), np.float32), 1, 1,
                                          [1.0], 0, 1)
    st = np.zeros((2, 8, 8), np.float32)
    with pytest.raises(ValueError, match="out of the range"):
        post.unwarp_chunk_slices_backward(st, 4, 4, [1.0], 0, 8)
    with pytest.raises(ValueError, match="out of the range"):     # the -1 wart
        post.unwarp_chunk_slices_backward(st, 4, 4, [1.0], 0, -1)
    with pytest.raises(ValueError, match="Eight coefficients"):
        post.correct_perspective_image(np.zeros((4, 4), np.float32), [1.0] * 7)
    with pytest.raises(ValueError, match="Eight coefficients"):
        post.correct_perspective_line([np.zeros((3, 2))], [1.0] * 9)
    with pytest.raises(ValueError):           # unpack of a 3-D shape, like :137
        post.unwarp_image_backward(np.zeros((2, 4, 4), np.float32), 1, 1, [1.0])
    img = np.zeros((4, 4), np.float32)
    with pytest.raises(RuntimeError, match="boundary mode not supported"):
        post.unwarp_image_backward(img, 1, 1, [1.0], mode="bogus")
    with pytest.raises(RuntimeError, match="spline order not supported"):
        post.unwarp_image_backward(img, 1, 1, [1.0], order=6)
    with pytest.raises(NotImplementedError, match="float16"):
        post.unwarp_image_backward(img.astype(np.float16), 1, 1, [1.0])
    with pytest.raises(NotImplementedError, match="int32"):
        post.unwarp_image_backward(img.astype(np.int32), 1, 1, [1.0], order=3)


def test_no_silent_cpu_fallback_without_gpu():
    if discorpy_b200.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(discorpy_b200.DcbError):
        post.unwarp_image_backward(np.zeros((4, 4), np.float32), 1, 1, [1.0])
    with pytest.raises(discorpy_b200.DcbError):
        post.correct_perspective_image(np.zeros((4, 4), np.float32),
                                       [1, 0, 0, 0, 1, 0, 0, 0])


def _lines():
    x0, y0 = 33.5, 35.5
    fact = [1.0, -2.0e-3]
    lines = [np.asarray([[64.0 - y, x] for x in np.arange(1, 64, 2.0)])
             for y in np.arange(1, 64, 2.0)]
    dlines = []
    for line in lines:
        xu = line[:, 1] - x0
        yu = line[:, 0] - y0
        ru = np.sqrt(xu ** 2 + yu ** 2)
        f = fact[0] + fact[1] * ru
        dlines.append(np.stack([y0 + yu * f, x0 + xu * f], axis=1))
    return x0, y0, fact, lines, dlines


def test_line_helpers_behave_like_reference_tests():
    # the properties tests/test_postprocessing.py:61-75, 125-160 assert
    x0, y0, fact, lines, dlines = _lines()
    fwd = post.unwarp_line_forward(dlines, x0, y0, [1.0, 2.0e-3])
    err = max(np.max(np.abs(a - b)) for a, b in zip(fwd, lines))
    assert err <= 1.0
    bwd = post.unwarp_line_backward(dlines[:4], x0, y0, fact)
    err = max(np.max(np.abs(a - b)) for a, b in zip(bwd, lines[:4]))
    assert err <= 1.0
    res_h = post.calc_residual_hor(lines, x0, y0)
    assert res_h.shape == (32 * 32, 2) and np.max(res_h[:, 1]) < 1e-9
    assert np.all(np.diff(res_h[:, 0]) >= 0)
    vlines = [np.asarray([[y, x] for y in np.arange(1, 64, 2.0)])
              for x in np.arange(1, 64, 2.0)]
    res_v = post.calc_residual_ver(vlines, x0, y0)
    assert np.max(res_v[:, 1]) < 1e-9
    assert post.check_distortion(res_h) is False
    big = post.calc_residual_hor(dlines, x0, y0)
    big[:, 1] *= 50
    assert post.check_distortion(big) is True


@pytest.mark.skipif(not os.path.isdir("/root/reference/discorpy"),
                    reason="the reference tree only exists in the build container")
def test_line_helpers_equal_reference():
    import importlib
    import sys
    sys.path.insert(0, "/root/reference")
    try:
        ref = importlib.import_module("discorpy.post.postprocessing")
    finally:
        sys.path.remove("/root/reference")
    x0, y0, fact, lines, dlines = _lines()
    for a, b in zip(post.unwarp_line_forward(dlines, x0, y0, [1.0, 2e-3]),
                    ref.unwarp_line_forward(dlines, x0, y0, [1.0, 2e-3])):
        assert np.array_equal(a, b)
    for a, b in zip(post.unwarp_line_backward(dlines[:3], x0, y0, fact),
                    ref.unwarp_line_backward(dlines[:3], x0, y0, fact)):
        assert np.array_equal(a, b)
    assert np.array_equal(post.calc_residual_hor(dlines, x0, y0),
                          ref.calc_residual_hor(dlines, x0, y0))
    assert np.array_equal(post.calc_residual_ver(dlines, x0, y0),
                          ref.calc_residual_ver(dlines, x0, y0))
    coef = [1.02, 0.01, -1.5, 0.005, 1.01, -0.8, 8e-5, -5e-5]
    for a, b in zip(post.correct_perspective_line(lines, coef),
                    ref.correct_perspective_line(lines, coef)):
        assert np.array_equal(a, b)
    img = np.random.default_rng(0).random((40, 50)).astype(np.float32)
    assert np.array_equal(post.unwarp_image_forward(img, 25, 20, [1.0, -6e-3]),
                          ref.unwarp_image_forward(img, 25, 20, [1.0, -6e-3]))
    ya, xa = post._generate_perspective_map(img, coef)
    yb, xb = ref._generate_perspective_map(img, coef)
    assert np.array_equal(ya, yb) and np.array_equal(xa, xb)


def test_synthetic_generator_restatement_is_deterministic():
    from discorpy_b200.device import synthetic_host
    a = synthetic_host(1000, seed=4, offset=10)
    b = synthetic_host(1010, seed=4, offset=0)[10:]
    assert np.array_equal(a, b)
    assert a.dtype == np.float32 and 0.0 <= a.min() and a.max() < 1.0
    assert abs(float(a.mean()) - 0.5) < 0.05


def test_install_as_discorpy_shadows_only_post(monkeypatch):
    if not os.path.isdir("/root/reference/discorpy"):
        pytest.skip("needs an importable reference discorpy")
    import sys
    monkeypatch.syspath_prepend("/root/reference")
    for k in [k for k in sys.modules if k == "discorpy" or k.startswith("discorpy.")]:
        monkeypatch.delitem(sys.modules, k)
    discorpy_b200.install_as_discorpy()
    import discorpy.post.postprocessing as p2
    import discorpy.proc.processing as proc
    assert p2 is post
    assert proc.post is post
    assert "reference" in proc.__file__


def test_fold_coordinate_is_scipys_boundary_mapping():
    """post._fold_coordinate (+ clamped taps) against the installed SciPy, orders 0 and 1, the
    modes the explicit-coordinate path implements for coordinates outside the image."""
    from scipy.ndimage import map_coordinates
    import discorpy_b200.post.postprocessing as post
    rng = np.random.default_rng(0)
    for trial in range(40):
        h, w = int(rng.integers(2, 40)), int(rng.integers(2, 40))
        img = rng.random((h, w)).astype(np.float32)
        yd = rng.uniform(-3.5 * h, 4.5 * h, 400)
        xd = rng.uniform(-3.5 * w, 4.5 * w, 400)
        yd[:40], xd[:40] = np.round(yd[:40]), np.round(xd[:40])
        if trial % 2:
            yd, xd = yd.astype(np.float32), xd.astype(np.float32)
        for order in (0, 1):
            for mode in ("reflect", "grid-mirror", "mirror", "wrap", "nearest", "constant"):
                want = map_coordinates(img, (yd, xd), order=order, mode=mode)
                fy, fx, zero = post._explicit_coordinates(yd, xd, h, w, order, mode)
                got = map_coordinates(img, (np.clip(fy, 0, h - 1), np.clip(fx, 0, w - 1)),
                                      order=order, mode="nearest")
                if zero is not None:
                    got = np.where(zero, np.float32(0), got)
                assert np.array_equal(got, want), (trial, order, mode)
    with pytest.raises(NotImplementedError):
        post._explicit_coordinates(np.array([-2.0]), np.array([1.0]), 8, 8, 1, "grid-wrap")


def test_host_copy_2d_matches_numpy():
    """dcb_host_copy_2d (the copy pool with non-temporal stores that stages pageable data, no GPU
    needed): any alignment, pitch and size, both ways of splitting the work over the threads."""
    import ctypes
    rng = np.random.default_rng(5)
    cases = [(1, 1), (3, 63), (1, 70001), (5, 1 << 20), (40, 4097), (64, 300000), (2, (3 << 20) + 5),
             (700, 16384), (9, 65536)]
    for rows, width in cases:
        for so, do in ((0, 0), (1, 3), (13, 64), (64, 7)):
            sp, dp = width + int(rng.integers(0, 200)), width + int(rng.integers(0, 200))
            src = rng.integers(0, 256, rows * sp + so + 64, dtype=np.uint8)
            dst = np.full(rows * dp + do + 64, 0xA5, dtype=np.uint8)
            want = dst.copy()
            for r in range(rows):
                want[do + r * dp: do + r * dp + width] = src[so + r * sp: so + r * sp + width]
            _cabi.call("dcb_host_copy_2d", ctypes.c_void_p(dst.ctypes.data + do), dp,
                       ctypes.c_void_p(src.ctypes.data + so), sp, width, rows)
            assert np.array_equal(dst, want), (rows, width, so, do)


def test_host_band_schedule():
    """dcb_host_band_edges: the row bands of the host-buffer pipeline for any image shape --
    strictly increasing from 0 to H, at most 32 bands, 2 / 2 / 4 MiB bands at both ends of large
    images, the requested number of equal bands otherwise (no GPU needed)."""
    import ctypes
    import os

    def edges(h, w, nb=0):
        buf = (ctypes.c_int * 33)()
        cnt = ctypes.c_int(0)
        _cabi.call("dcb_host_band_edges", h, w, nb, buf, ctypes.byref(cnt))
        return list(buf[:cnt.value + 1])

    shapes = [(1, 1), (7, 3), (512, 640), (2048, 2048), (2160, 2560), (4096, 4096), (8192, 8192),
              (3000, 5000), (16, 600000), (40, 2000000), (100000, 100), (8388607, 1), (33, 8388607)]
    for (h, w) in shapes:
        for nb in (0, 1, 2, 5, 8, 31, 32, 100):
            e = edges(h, w, nb)
            assert e[0] == 0 and e[-1] == h and 1 <= len(e) - 1 <= 32, (h, w, nb, e)
            assert all(b > a for a, b in zip(e, e[1:])), (h, w, nb, e)
            if nb > 0:
                assert len(e) - 1 <= min(nb, h)
    assert [b - a for a, b in zip(edges(4096, 4096), edges(4096, 4096)[1:])] == \
        [128, 128, 256] + [384] * 8 + [256, 128, 128]
    assert len(edges(2160, 2560)) - 1 == 6 and len(edges(1024, 1024)) - 1 == 1
    os.environ["DCB_BAND_EDGES"] = "8,8,16,32"
    try:
        assert edges(4096, 4096) == [0, 512, 1024, 2048, 4096]
        assert edges(4096, 4096, 4) == [0, 1024, 2048, 3072, 4096]      # explicit counts ignore it
        os.environ["DCB_BAND_EDGES"] = "1,1"
        assert edges(640, 640) == [0, 10, 20, 640]
    finally:
        del os.environ["DCB_BAND_EDGES"]
