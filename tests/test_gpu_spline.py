"""GPU parity of the spline path (orders 2..5, float64 images; csrc/spline.cuh
through the Python drop-in) against the reference's golden outputs and the
oracle.  The kernels restate SciPy operation by operation, so the bar is the
same as for order 1: bit-identical on the golden fixtures; on the larger seeded
images at most one float32-rounded COORDINATE per 4 Mpixel may differ (Horner +
FMA here, term-by-term NumPy power in the reference), everything else within
1e-5 * max(1, max|mat|)."""
import numpy as np
import pytest

from conftest import tolerance
from test_oracle_spline import CASES, golden
from oracle import oracle_spline as osp
from oracle import oracle_np
from oracle.make_golden import make_input

import discorpy_b200 as dcb
import discorpy_b200.post.postprocessing as post
from discorpy_b200.util import utility as util

pytestmark = pytest.mark.gpu


def run_gpu(c, mat):
    if c["fn"] == "image":
        return post.unwarp_image_backward(mat, c["xc"], c["yc"], c["fact"], order=c["order"],
                                          mode=c["mode"])
    if c["fn"] == "persp":
        mi = post._generate_perspective_map(mat, c["coef"]) if c["use_map"] else None
        return post.correct_perspective_image(mat, c["coef"], order=c["order"], mode=c["mode"],
                                              map_index=mi)
    return util.unwarp_color_image_backward(mat, c["xc"], c["yc"], c["fact"], order=c["order"],
                                            mode=c["mode"])


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_spline_golden_vectors_bit_exact(case):
    mat = make_input(case["kind"], tuple(case["shape"]), case["seed"], case["dtype"])
    got = run_gpu(case, mat)
    want = golden(case["id"])
    assert got.dtype == want.dtype and got.shape == want.shape
    assert isinstance(got, np.ndarray) and got.flags.c_contiguous
    nbad = int(np.count_nonzero(got != want))
    assert nbad == 0, "%d of %d samples differ, max %g" % (
        nbad, got.size, float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64)))))


@pytest.mark.parametrize("order,mode", [(3, "reflect"), (2, "mirror"), (5, "nearest"),
                                         (4, "grid-wrap")])
def test_spline_seeded_image_against_oracle(order, mode):
    shape, xc, yc = (700, 900), 455.3, 341.8
    fact = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
    rng = np.random.default_rng(order)
    mat = np.floor(rng.random(shape, dtype=np.float32) * 256.0)
    want = osp.unwarp_image_backward(mat, xc, yc, fact, order, mode)
    got = post.unwarp_image_backward(mat, xc, yc, fact, order=order, mode=mode)
    diff = np.abs(got.astype(np.float64) - want.astype(np.float64))
    n_ne = int(np.count_nonzero(got != want))
    assert n_ne <= 1, "%d samples not identical (max %g)" % (n_ne, float(diff.max()))


@pytest.mark.parametrize("staged", ["auto", "0", "1"])
def test_prefilter_coefficients_bit_exact(staged, monkeypatch):
    """dcb_spline_prefilter alone against the oracle's spline_filter (== SciPy's), through
    the one-thread-per-line kernel, the shared-memory staged kernel, and the library's own
    choice between them."""
    import ctypes
    from discorpy_b200 import _cabi, device as dev
    if staged != "auto":
        monkeypatch.setenv("DCB_SPLINE_STAGED", staged)
    rng = np.random.default_rng(5)
    for shape, dt in (((130, 77), np.float32), ((64, 200), np.float64), ((1, 33), np.float32),
                      ((45, 1), np.float32), ((300, 521), np.float32), ((2, 3), np.float32),
                      ((65, 129), np.float64)):
        mat = (rng.random(shape) * 200 - 30).astype(dt)
        h, w = shape
        for order in (2, 3, 4, 5):
            for mode in osp.MODES:
                kind, npad = osp.filter_kind(mode)
                src = mat
                if npad:
                    src = np.pad(mat, npad, mode="edge" if mode == "nearest" else "constant")
                want = osp.spline_filter(src, order, mode)
                need = ctypes.c_size_t()
                _cabi.call("dcb_spline_workspace_bytes", h, w, order, _cabi.MODES[mode],
                           ctypes.byref(need))
                stream = dev.current_stream()
                sh = ctypes.c_void_p(stream.handle)
                with dev.borrowed(need.value) as work, dev.borrowed(mat.nbytes) as dsrc:
                    _cabi.call("dcb_h2d", ctypes.c_void_p(dsrc.ptr),
                               ctypes.c_void_p(mat.ctypes.data), mat.nbytes, sh)
                    stream.sync()
                    _cabi.call("dcb_spline_prefilter", ctypes.c_void_p(dsrc.ptr),
                               int(dt is np.float64), h, w, w * mat.itemsize, order,
                               _cabi.MODES[mode], ctypes.c_void_p(work.ptr), need.value, sh)
                    got = np.empty(want.shape, np.float64)
                    _cabi.call("dcb_d2h", ctypes.c_void_p(got.ctypes.data),
                               ctypes.c_void_p(work.ptr), got.nbytes, sh)
                    stream.sync()
                nbad = int(np.count_nonzero(got != want))
                assert nbad == 0, (shape, np.dtype(dt).name, order, mode, nbad,
                                   float(np.max(np.abs(got - want))))


def test_float64_image_orders_0_1_and_mapping():
    rng = np.random.default_rng(9)
    mat = rng.standard_normal((120, 150)) * 50.0
    fact = [1.0, 3.0e-3]
    for order in (0, 1):
        got = post.unwarp_image_backward(mat, 70.2, 61.9, fact, order=order)
        want = oracle_np.unwarp_image_backward(mat, 70.2, 61.9, fact, order)
        assert got.dtype == np.float64 and np.array_equal(got, want)
    coef = [1.02, 0.01, -3.0, 0.005, 1.01, -2.0, 8e-5, -5e-5]
    got = post.correct_perspective_image(mat, coef, order=3, mode="mirror")
    want = osp.correct_perspective_image(mat, coef, 3, "mirror")
    assert got.dtype == np.float64 and np.array_equal(got, want)


def test_spline_device_array_roundtrip():
    rng = np.random.default_rng(3)
    mat = rng.random((90, 140), dtype=np.float32)
    fact = [1.0, 3.0e-3]
    dsrc = dcb.DeviceArray.from_host(mat)
    out = post.unwarp_image_backward(dsrc, 66.6, 41.2, fact, order=3)
    assert isinstance(out, dcb.DeviceArray)
    want = osp.unwarp_image_backward(mat, 66.6, 41.2, fact, 3, "reflect")
    assert np.array_equal(out.to_host(), want)
