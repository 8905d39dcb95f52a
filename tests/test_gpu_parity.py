"""GPU parity: the CUDA path (through the Python drop-in -> ctypes -> C ABI)
against the reference's golden outputs and against the oracle.

Bars (SURVEY.md 8d / BASELINE north_star):
  * order 0 (nearest): bit-exact;
  * order 1: |gpu - ref| <= 1e-5 * max(1, max|mat|).  With the default
    DCB_BLEND_EXACT the result is expected to be bit-identical as well; the
    tests assert equality on the golden fixtures and report/limit the number
    of non-identical pixels on the big seeded images, where an fp64
    last-bit difference in the coordinate polynomial (Horner + FMA here,
    term-by-term NumPy `power` in the reference) can flip a float32-rounded
    coordinate with probability ~1e-8 per pixel (hard part 6).
"""
import numpy as np
import pytest

from conftest import load_cases, golden_output, tolerance
from oracle import oracle_np as orc
from oracle import oracle_c
from oracle.make_golden import make_input

import discorpy_b200 as dcb
import discorpy_b200.post.postprocessing as post

pytestmark = pytest.mark.gpu

GPU_DTYPES = ("float32", "uint8", "int8", "uint16", "int16")
CASES = [c for c in load_cases() if c.get("dtype", "float32") in GPU_DTYPES]
FACT5 = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
COEF_DOT_05 = [1.00227490554, -2.99523692178e-05, 8.99519088e-08,
               -1.57066461911e-10, 8.08880211618e-14]


def _run_gpu(c, mat):
    fn = c["fn"]
    if fn == "image":
        return post.unwarp_image_backward(mat, c["xc"], c["yc"], c["fact"],
                                          order=c["order"])
    if fn == "slice":
        return post.unwarp_slice_backward(mat, c["xc"], c["yc"], c["fact"],
                                          c["index"])
    if fn == "chunk":
        return post.unwarp_chunk_slices_backward(mat, c["xc"], c["yc"],
                                                 c["fact"], c["start"],
                                                 c["stop"])
    if fn == "persp":
        return post.correct_perspective_image(mat, c["coef"], order=c["order"])
    if fn == "combined":
        return post.unwarp_image_backward_perspective(
            mat, c["xc"], c["yc"], c["fact"], c["coef"])
    if fn == "color":
        from discorpy_b200.util import utility as util
        pad = tuple(c["pad"]) if isinstance(c["pad"], list) else c["pad"]
        return np.ascontiguousarray(util.unwarp_color_image_backward(
            mat, c["xc"], c["yc"], c["fact"], order=c["order"], pad=pad,
            pad_mode=c["pad_mode"]))
    raise AssertionError(fn)


@pytest.fixture(autouse=True)
def _default_config():
    post.config["blend"] = dcb.BLEND_EXACT
    post.config["path"] = dcb.PATH_AUTO
    yield
    post.config["blend"] = dcb.BLEND_EXACT
    post.config["path"] = dcb.PATH_AUTO


@pytest.mark.parametrize("path", [dcb.PATH_AUTO, dcb.PATH_DIRECT],
                         ids=["auto", "direct"])
@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_golden_vectors_bit_exact(case, path):
    post.config["path"] = path
    mat = make_input(case["kind"], tuple(case["shape"]), case["seed"],
                     case.get("dtype", "float32"))
    got = _run_gpu(case, mat)
    want = golden_output(case["id"])
    assert got.dtype == want.dtype and got.shape == want.shape
    assert isinstance(got, np.ndarray) and got.flags.c_contiguous
    nbad = int(np.count_nonzero(got != want))
    assert nbad == 0, "%d of %d pixels differ, max %g" % (
        nbad, got.size, float(np.max(np.abs(got - want))))


def _compare(got, want, mat, order, flips_allowed):
    """Returns (n_not_identical, n_above_tol)."""
    diff = np.abs(got.astype(np.float64) - want.astype(np.float64))
    n_ne = int(np.count_nonzero(got != want))
    n_tol = int(np.count_nonzero(diff > tolerance(mat)))
    if order == 0:
        assert n_ne <= flips_allowed, "nearest: %d pixels differ" % n_ne
    else:
        assert n_tol <= flips_allowed, (
            "%d pixels above tol (max %g), %d not identical"
            % (n_tol, float(diff.max()), n_ne))
    return n_ne, n_tol


@pytest.mark.parametrize("shape,xc,yc,fact", [
    ((1024, 1024), 510.3, 500.7, FACT5),
    ((1536, 2050), 1030.2, 760.4, FACT5),              # W % 4 != 0 -> pitched
    ((801, 1283), 588.692801577, 462.092631791, COEF_DOT_05),
    ((2160, 2560), 588.692801577, 462.092631791, COEF_DOT_05),   # BASELINE cfg 1 geometry
    ((1000, 1000), 500, 500, [1.0, 3.0e-3]),           # integer centre
    ((2048, 2048), 1030.2, 1019.6,
     [1.0, -1e-5, 3e-8, -2e-11, 5e-15, -8e-19, 6e-23, -2e-27, 3e-32]),
])
@pytest.mark.parametrize("order", [0, 1])
def test_seeded_images_against_oracle(shape, xc, yc, fact, order):
    rng = np.random.default_rng(shape[0] + shape[1] + order)
    mat = np.floor(rng.random(shape, dtype=np.float32) * 256.0)   # 0..255 like a camera frame
    want = oracle_c.unwarp_image_backward(mat, xc, yc, fact, order)
    outs = {}
    for name, path in (("auto", dcb.PATH_AUTO), ("direct", dcb.PATH_DIRECT)):
        post.config["path"] = path
        outs[name] = post.unwarp_image_backward(mat, xc, yc, fact, order=order)
    # the two memory paths implement the same arithmetic: always identical
    assert np.array_equal(outs["auto"], outs["direct"])
    # <= 1 flipped coordinate per ~4 Mpix tolerated (expected ~0.03)
    _compare(outs["auto"], want, mat, order, flips_allowed=1 + mat.size // (4 << 20))


def test_numpy_oracle_full_compare_one_image():
    """Same check against the NumPy oracle (the one pinned bit-for-bit to the
    reference), one mid-size image."""
    rng = np.random.default_rng(5)
    mat = rng.random((1200, 1600), dtype=np.float32)
    for order in (0, 1):
        want = orc.unwarp_image_backward(mat, 801.7, 590.2, FACT5, order)
        got = post.unwarp_image_backward(mat, 801.7, 590.2, FACT5, order=order)
        _compare(got, want, mat, order, flips_allowed=1)


@pytest.mark.parametrize("blend,rel", [(dcb.BLEND_LERP64, 1e-5),
                                       (dcb.BLEND_LERP32, 1e-5)])
def test_fast_blends_within_tolerance(blend, rel):
    rng = np.random.default_rng(9)
    mat = rng.random((1024, 1280), dtype=np.float32)          # [0,1) data
    want = orc.unwarp_image_backward(mat, 640.4, 511.6, FACT5, 1)
    post.config["blend"] = blend
    got = post.unwarp_image_backward(mat, 640.4, 511.6, FACT5)
    diff = np.abs(got.astype(np.float64) - want)
    assert int(np.count_nonzero(diff > rel)) <= 1
    if blend == dcb.BLEND_LERP32:
        assert float(np.median(diff)) < 1e-7


def test_integer_image_fast_blends_round_to_integers():
    """Integer images with the float32 / float64 lerp blends: the result is rounded half away
    from zero like SciPy's (never a truncated float), and differs from the oracle's exact
    blend by at most one count where the blend lands within 1e-3 of a tie."""
    rng = np.random.default_rng(19)
    frame = rng.integers(0, 65536, (900, 1100), dtype=np.uint16)
    want = orc.unwarp_image_backward(frame, 551.3, 447.9, FACT5, 1)
    exactf = orc.unwarp_image_backward(frame.astype(np.float32), 551.3, 447.9, FACT5, 1).astype(np.float64)
    for blend in (dcb.BLEND_LERP32, dcb.BLEND_LERP64):
        post.config["blend"] = blend
        got = post.unwarp_image_backward(frame, 551.3, 447.9, FACT5)
        assert got.dtype == np.uint16
        d = got.astype(np.int64) - want.astype(np.int64)
        assert int(np.max(np.abs(d))) <= 1
        # a differing pixel is one whose unrounded blend sits next to a tie
        frac = np.abs(exactf - np.floor(exactf) - 0.5)
        tol = 0.05 if blend == dcb.BLEND_LERP32 else 1e-6     # float32 blend of 16-bit counts
        assert np.all(frac[d != 0] <= tol)
        # and the rounding really happened: rounded (not truncated) float32 blends agree
        assert abs(float(np.mean(got.astype(np.float64) - exactf))) < 5e-3


def test_baseline_config2_full_size():
    """BASELINE config 2 at full size (4096 x 4096, 5 terms),
    full compare with the C oracle plus size-independent properties."""
    xc, yc = 2050.37, 2040.81
    fact = [COEF_DOT_05[i] / 3.0 ** i for i in range(5)]
    rng = np.random.default_rng(2)
    mat = rng.random((4096, 4096), dtype=np.float32)
    for order in (0, 1):
        want = oracle_c.unwarp_image_backward(mat, xc, yc, fact, order)
        got = post.unwarp_image_backward(mat, xc, yc, fact, order=order)
        n_ne, n_tol = _compare(got, want, mat, order, flips_allowed=4)
        print("cfg2 order %d: %d px not identical, %d above tol" % (order, n_ne, n_tol))
    # identity model: output == input exactly
    ident = post.unwarp_image_backward(mat, xc, yc, [1.0])
    assert np.array_equal(ident, mat)
    # a constant image stays constant under any model
    const = np.full((4096, 4096), 3.25, dtype=np.float32)
    assert np.all(post.unwarp_image_backward(const, xc, yc, fact) == 3.25)
    # idempotence of the data path: same input twice -> same bytes
    again = post.unwarp_image_backward(mat, xc, yc, fact)
    assert np.array_equal(again, got)


def test_baseline_config3_combined_two_pass():
    rng = np.random.default_rng(3)
    mat = rng.random((2048, 2048), dtype=np.float32)
    fact = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
    coef = [1.02, 0.01, -15.0, 0.005, 1.01, -8.0, 8e-6, -5e-6]
    tmp = oracle_c.unwarp_image_backward(mat, 1030.2, 1019.6, fact, 1)
    want = oracle_c.correct_perspective_image(tmp, coef, 1)
    got = post.unwarp_image_backward_perspective(mat, 1030.2, 1019.6, fact, coef)
    _compare(got, want, mat, 1, flips_allowed=2)
    # and the two public calls chained give the same bytes as the fused entry
    chained = post.correct_perspective_image(
        post.unwarp_image_backward(mat, 1030.2, 1019.6, fact), coef)
    assert np.array_equal(chained, got)


def test_perspective_map_index_route_and_mapping():
    rng = np.random.default_rng(13)
    mat = rng.random((300, 421), dtype=np.float32)
    coef = [0.9, -0.05, 3.0, 0.04, 0.95, 2.0, -3e-4, 2e-4]
    direct = post.correct_perspective_image(mat, coef)
    mi = post._generate_perspective_map(mat, coef)
    via_index = post.correct_perspective_image(mat, coef, map_index=mi)
    assert np.array_equal(direct, via_index)
    assert np.array_equal(direct, orc.correct_perspective_image(mat, coef))
    yd, xd = mi
    got = post._mapping(mat, xd.reshape(300, 421), yd.reshape(300, 421))
    assert np.array_equal(got, direct)
    # float64 coordinates (the slice path's kind) through the generic entry
    y64 = rng.random(5000) * 299
    x64 = rng.random(5000) * 420
    got = post._mapping(mat, x64, y64)
    assert np.array_equal(got, orc.sample(mat, y64, x64, 1))
    # coordinates outside the image: SciPy's 'reflect' (what _mapping hands them to, :251)
    from scipy.ndimage import map_coordinates
    got = post._mapping(mat, x64 + 1000.0, y64)
    assert np.array_equal(got, map_coordinates(mat, (y64, x64 + 1000.0), order=1, mode="reflect"))


def test_stack_paths_against_oracle():
    """BASELINE config 4 geometry on a host-sized subset: D' = 8 slices."""
    rng = np.random.default_rng(4)
    stack = rng.random((8, 640, 2560), dtype=np.float32)
    xc, yc = 1283.4, 318.9
    for index in (0, 317, 639):
        want = orc.unwarp_slice_backward(stack, xc, yc, FACT5, index)
        got = post.unwarp_slice_backward(stack, xc, yc, FACT5, index)
        assert got.dtype == np.float32 and got.shape == (8, 2560)
        diff = np.abs(got.astype(np.float64) - want)
        assert int(np.count_nonzero(diff > 1e-5)) == 0
        assert int(np.count_nonzero(got != want)) <= 2      # fp64 last-bit coordinate effects
    want = orc.unwarp_chunk_slices_backward(stack, xc, yc, FACT5, 100, 227)
    got = post.unwarp_chunk_slices_backward(stack, xc, yc, FACT5, 100, 227)
    assert got.shape == (8, 128, 2560)
    assert int(np.count_nonzero(got != want)) <= 1
    # chunk over all rows == unwarp_image_backward per slice
    full = post.unwarp_chunk_slices_backward(stack, xc, yc, FACT5, 0, 639)
    for z in (0, 5):
        assert np.array_equal(full[z], post.unwarp_image_backward(stack[z], xc, yc, FACT5))
    # uint16 stacks: SciPy rounds every slice to uint16 before the reference stores it
    # into the float32 sinogram (:227-228); the kernel rounds in fp64 the same way
    st16 = (stack[:2] * 60000).astype(np.uint16)
    got = post.unwarp_slice_backward(st16, xc, yc, FACT5, 300)
    want = orc.unwarp_slice_backward(st16, xc, yc, FACT5, 300)
    assert got.dtype == np.float32 and int(np.count_nonzero(got != want)) <= 1


def test_baseline_config4_full_frame_stack():
    """BASELINE config 4 at its frame size: 2560 x 2560 slices with the config's own
    centre and model (>= 4 slices, >= 3 indices through
    unwarp_slice_backward, float64 coordinates) plus the chunk path over all rows."""
    rng = np.random.default_rng(44)
    stack = rng.random((4, 2560, 2560), dtype=np.float32)
    xc, yc = 1283.4, 1275.9
    for index in (0, 1275, 2559, 77):
        want = orc.unwarp_slice_backward(stack, xc, yc, FACT5, index)
        got = post.unwarp_slice_backward(stack, xc, yc, FACT5, index)
        assert got.shape == (4, 2560)
        diff = np.abs(got.astype(np.float64) - want)
        assert int(np.count_nonzero(diff > 1e-5)) == 0
        assert int(np.count_nonzero(got != want)) <= 2
    full = post.unwarp_chunk_slices_backward(stack, xc, yc, FACT5, 0, 2559)
    for z in range(4):
        want = oracle_c.unwarp_image_backward(stack[z], xc, yc, FACT5, 1)
        _compare(full[z], want, stack[z], 1, flips_allowed=2)
    # linearity of the data path in the image (same geometry): T(2a) == 2 T(a) exactly
    twice = post.unwarp_chunk_slices_backward(stack[:1] * 2.0, xc, yc, FACT5, 0, 2559)
    assert np.array_equal(twice[0], full[0] * 2.0)


def test_baseline_config5_fisheye_8192():
    """BASELINE config 5 geometry at full size: one 8192 x 8192 image, 9-term
    fisheye-strength model (parity on one image per GPU), through
    the single-image kernel and through the Z-stack kernel (batch of 2)."""
    fact = [1.0, -1e-5, 3e-8, -2e-11, 5e-15, -8e-19, 6e-23, -2e-27, 3e-32]
    xc, yc = 4100.3, 4090.8
    rng = np.random.default_rng(5)
    mat = rng.random((8192, 8192), dtype=np.float32)
    for order in (0, 1):
        want = oracle_c.unwarp_image_backward(mat, xc, yc, fact, order)
        got = post.unwarp_image_backward(mat, xc, yc, fact, order=order)
        n_ne, n_tol = _compare(got, want, mat, order, flips_allowed=16)
        print("cfg5 order %d: %d px not identical, %d above tol" % (order, n_ne, n_tol))
    batch = np.stack([mat[:4096], mat[4096:]])          # two (4096, 8192) frames
    out = post.unwarp_chunk_slices_backward(batch, xc, 2045.4, fact, 0, 4095)
    for z in range(2):
        assert np.array_equal(out[z], post.unwarp_image_backward(batch[z], xc, 2045.4, fact))


def test_device_resident_api_and_sharding_gives_same_bytes():
    rng = np.random.default_rng(21)
    stack = rng.random((6, 256, 512), dtype=np.float32)
    dstack = dcb.DeviceArray.from_host(stack)
    out = post.unwarp_chunk_slices_backward(dstack, 255.5, 127.2, FACT5, 0, 255)
    assert isinstance(out, dcb.DeviceArray)
    whole = out.to_host()
    from discorpy_b200.multigpu import shard_range
    for world in (2, 4):
        parts = []
        for rank in range(world):
            lo, hi = shard_range(6, rank, world)
            if hi > lo:
                parts.append(post.unwarp_chunk_slices_backward(
                    stack[lo:hi], 255.5, 127.2, FACT5, 0, 255))
        assert np.array_equal(np.concatenate(parts), whole)
    img = dcb.DeviceArray.from_host(stack[0])
    res = post.unwarp_image_backward(img, 255.5, 127.2, FACT5)
    assert np.array_equal(res.to_host(), whole[0])


def test_banded_host_pipeline_gives_the_same_bytes():
    """dcb_unwarp_image_backward_host_f32: upload / compute / download in row
    bands must not change a single byte, whatever the band count, also when a
    band's source rows lie far from its output rows (strong distortion)."""
    rng = np.random.default_rng(77)
    mat = rng.random((1500, 1100), dtype=np.float32)
    pinned = dcb.pinned_copy(mat)
    for xc, yc, fact in ((551.3, 760.2, FACT5),
                         (300.0, 1400.5, [0.7, 6e-4, -2e-7]),      # strong, off-centre
                         (551.3, 760.2, [1.9, -1.5e-3, 6e-7])):    # folds rows back
        want = orc.unwarp_image_backward(mat, xc, yc, fact, 1)
        outs = []
        for bands in (1, 2, 5, 13):
            post.config["bands"] = bands
            try:
                outs.append(post.unwarp_image_backward(pinned if bands != 5 else mat, xc, yc, fact))
            finally:
                post.config["bands"] = 0
        for o in outs[1:]:
            assert np.array_equal(o, outs[0])
        _compare(outs[0], want, mat, 1, flips_allowed=1)
    # the default band count on a 32 MiB image (4 bands)
    big = rng.random((2048, 4096), dtype=np.float32)
    got = post.unwarp_image_backward(big, 2050.4, 1000.6, FACT5)
    dev = post.unwarp_image_backward(dcb.DeviceArray.from_host(big), 2050.4, 1000.6, FACT5)
    assert np.array_equal(got, dev.to_host())


def test_edge_shapes_and_values():
    for shape in ((1, 1), (1, 300), (300, 1), (2, 2), (33, 129), (7, 4097)):
        rng = np.random.default_rng(shape[0] * 31 + shape[1])
        mat = rng.standard_normal(shape).astype(np.float32)
        xc, yc = shape[1] / 2.0 + 0.25, shape[0] / 2.0 - 0.5
        for order in (0, 1):
            want = orc.unwarp_image_backward(mat, xc, yc, [1.01, -1e-4], order)
            got = post.unwarp_image_backward(mat, xc, yc, [1.01, -1e-4], order=order)
            assert np.array_equal(got, want), (shape, order)
    # a NaN pixel contaminates every output whose 2x2 footprint touches it
    mat = np.ones((64, 64), np.float32)
    mat[20, 30] = np.nan
    got = post.unwarp_image_backward(mat, 32, 32, [1.0])
    want = orc.unwarp_image_backward(mat, 32, 32, [1.0])
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert int(np.isnan(got).sum()) == 4
    # non-contiguous input view (a colour channel, demo_07.py:25)
    rgb = np.random.default_rng(3).random((50, 60, 3)).astype(np.float32)
    got = post.unwarp_image_backward(rgb[:, :, 1], 30.2, 24.9, [1.0, 2e-3])
    want = orc.unwarp_image_backward(rgb[:, :, 1], 30.2, 24.9, [1.0, 2e-3])
    assert np.array_equal(got, want)
    # float64 images: float64 in, float64 out (the spline path); float16 is still refused
    f64 = post.unwarp_image_backward(rgb[:, :, 1].astype(np.float64), 30.2, 24.9, [1.0, 2e-3])
    assert f64.dtype == np.float64
    assert np.array_equal(f64, orc.unwarp_image_backward(rgb[:, :, 1].astype(np.float64), 30.2,
                                                         24.9, [1.0, 2e-3]))
    with pytest.raises(NotImplementedError, match="dtype"):
        post.unwarp_image_backward(np.zeros((8, 8), np.float16), 4, 4, [1.0])


def test_integer_frames_full_compare():
    """uint8 camera frame (BASELINE config 1 geometry) and a uint16 colour frame:
    output dtype = input dtype, SciPy's integer rounding, bit-exact."""
    rng = np.random.default_rng(31)
    frame = rng.integers(0, 256, (2160, 2560), dtype=np.uint8)
    for order in (0, 1):
        want = orc.unwarp_image_backward(frame, 588.692801577, 462.092631791,
                                         COEF_DOT_05, order)
        got = post.unwarp_image_backward(frame, 588.692801577, 462.092631791,
                                         COEF_DOT_05, order=order)
        assert got.dtype == np.uint8
        assert int(np.count_nonzero(got != want)) <= 1
    from discorpy_b200.util import utility as util
    rgb = rng.integers(0, 65536, (600, 800, 3), dtype=np.uint16)
    want = orc.unwarp_color_image_backward(rgb, 402.3, 297.8, FACT5, pad=16, pad_mode="edge")
    before = dcb.launch_count()
    got = util.unwarp_color_image_backward(rgb, 402.3, 297.8, FACT5, pad=16, pad_mode="edge")
    assert dcb.launch_count() == before + 3          # unpack, ONE remap for all channels, pack
    assert got.dtype == np.uint16 and got.shape == want.shape == (632, 832, 3)
    assert int(np.count_nonzero(got != want)) <= 1


def test_tma_path_is_actually_taken_and_launches_counted():
    rng = np.random.default_rng(1)
    mat = rng.random((1024, 1024), dtype=np.float32)
    before = dcb.launch_count()
    post.config["path"] = dcb.PATH_TMA
    post.unwarp_image_backward(mat, 512.2, 511.1, FACT5)
    plan = dcb.last_plan()
    assert plan["path"] == dcb.PATH_TMA and plan["box_w"] >= 128 and plan["smem_bytes"] > 0
    assert dcb.launch_count() == before + 1     # (the one-off plan build is counted apart)
    post.config["path"] = dcb.PATH_DIRECT
    post.unwarp_image_backward(mat, 512.2, 511.1, FACT5)
    assert dcb.last_plan()["path"] == dcb.PATH_DIRECT


def test_forward_unwarp_on_device_matches_numpy_scatter():
    """unwarp_image_forward for a DeviceArray == the host (reference-style NumPy) result,
    including vacant pixels and collisions (last writer in C order wins)."""
    rng = np.random.default_rng(21)
    for shape, xc, yc, fact in (((300, 421), 200.3, 140.8, [1.0, 4.0e-4]),      # expands: holes
                                ((256, 256), 128.0, 128.0, [0.9, -5.0e-4]),       # shrinks: collisions
                                ((97, 130), 70.2, 33.3, [1.0, -2e-3, 1e-5])):
        mat = rng.random(shape, dtype=np.float32) + 1.0
        want = post.unwarp_image_forward(mat, xc, yc, fact)                # host NumPy path
        got = post.unwarp_image_forward(dcb.DeviceArray.from_host(mat), xc, yc, fact)
        assert isinstance(got, dcb.DeviceArray)
        assert np.array_equal(got.to_host(), want)
        assert np.count_nonzero(want == 0) > 0 or fact[0] < 1.0


def test_custom_sqrt_is_correctly_rounded():
    import ctypes
    from discorpy_b200 import _cabi
    bad = ctypes.c_uint64(123)
    _cabi.call("dcb_selftest_sqrt", 1 << 26, 12345, ctypes.byref(bad))
    assert bad.value == 0, "%d of 2^26 inputs not correctly rounded" % bad.value


def test_fast_sqrt_is_within_one_ulp():
    """the 5-operation sqrt of the single-image kernel: never more than one ulp off"""
    import ctypes
    from discorpy_b200 import _cabi
    differ, bad = ctypes.c_uint64(0), ctypes.c_uint64(123)
    _cabi.call("dcb_selftest_sqrt_fast", 1 << 26, 12345, ctypes.byref(differ), ctypes.byref(bad))
    assert bad.value == 0, "%d of 2^26 inputs more than one ulp off" % bad.value
    assert differ.value < (1 << 26) // 4


def test_synthetic_fill_matches_host_restatement():
    from discorpy_b200.device import synthetic_host
    arr = dcb.DeviceArray((64, 256)).fill_synthetic(seed=4, offset=1000)
    got = arr.to_host().ravel()
    assert np.array_equal(got, synthetic_host(64 * 256, 4, 1000))


def test_patch_path_equals_exact_path_and_plans_are_cached():
    """The single-image kernel's plan (csrc/remap_image.cuh: verified row patches) is built once
    per (model, geometry) and reused; the patch path must give the very bytes of the exact path
    (DCB_IMG_FAST=0 switches it off) for every blend and order, on a geometry where most rows
    are verified and on one where few are."""
    import os
    rng = np.random.default_rng(77)
    for shape, xc, yc, fact in (((1000, 1408), 700.3, 512.9, FACT5),
                                ((700, 900), 120.7, 655.1, [1.0, 3e-4, -2e-7])):
        mat = dcb.DeviceArray.from_host(rng.random(shape, dtype=np.float32) * 100.0)
        for order, blend in ((1, dcb.BLEND_EXACT), (1, dcb.BLEND_LERP64), (1, dcb.BLEND_LERP32),
                             (0, dcb.BLEND_EXACT)):
            post.config["blend"] = blend
            try:
                dcb.plan_cache_clear()
                built0 = dcb.plan_cache_clear()
                dcb.image_stats(True, reset=True)
                cold = post.unwarp_image_backward(mat, xc, yc, fact, order=order).to_host()
                st = dcb.image_stats(False, reset=True)
                before = dcb.launch_count()
                warm = post.unwarp_image_backward(mat, xc, yc, fact, order=order).to_host()
                assert dcb.launch_count() == before + 1
                assert dcb.plan_cache_clear() == built0 + 1          # one build, reused
                os.environ["DCB_IMG_FAST"] = "0"
                try:
                    exact = post.unwarp_image_backward(mat, xc, yc, fact, order=order).to_host()
                finally:
                    del os.environ["DCB_IMG_FAST"]
            finally:
                post.config["blend"] = dcb.BLEND_EXACT
            assert np.array_equal(cold, warm)
            if blend != dcb.BLEND_LERP32:          # lerp32: same coordinates, float32 blend differs
                assert np.array_equal(cold, exact), (shape, order, blend)
            else:
                assert np.max(np.abs(cold - exact)) <= 1e-5 * 100.0
            assert st["rows"] == shape[0] * ((shape[1] + 127) // 128)      # tile rows of 128 pixels
    # the first geometry is mostly verified rows
    dcb.plan_cache_clear()
    dcb.image_stats(True, reset=True)
    post.unwarp_image_backward(dcb.DeviceArray.from_host(rng.random((1000, 1408), dtype=np.float32)),
                               700.3, 512.9, FACT5)
    st = dcb.image_stats(False, reset=True)
    assert st["rows_patch"] > 0.3 * st["rows"], st          # (86 % on the 2048^2 config-3 geometry)


def test_patch_path_with_odd_values_in_the_image():
    """Negative values, Inf, NaN, zeros and tiny magnitudes switch single tiles to SciPy's own
    summation inside the patch path; bytes must equal the oracle's (equal NaN positions)."""
    rng = np.random.default_rng(78)
    mat = rng.random((640, 1152), dtype=np.float32)
    mat[100:140, 200:260] *= -1.0
    mat[300, 500] = np.inf
    mat[301, 777] = np.nan
    mat[400:420, :] = 0.0
    mat[500:520, 100:400] *= 1e-38
    want = orc.unwarp_image_backward(mat, 580.2, 318.6, FACT5)
    got = post.unwarp_image_backward(mat, 580.2, 318.6, FACT5)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.array_equal(got[ok], want[ok])


def test_forward_unwarp_on_device_equals_the_oracle_and_the_reference_golden():
    """dcb_unwarp_image_forward_f32 (csrc/forward.cuh) against the real reference's outputs
    (tests/golden/forward.npz) and the oracle -- not against the product's own host code."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "forward.npz"))
    k = 0
    while "in%d" % k in z:
        par = z["par%d" % k]
        mat, xc, yc, fact = z["in%d" % k], par[0], par[1], list(par[2:])
        got = post.unwarp_image_forward(dcb.DeviceArray.from_host(mat), xc, yc, fact).to_host()
        assert np.array_equal(got, z["out%d" % k]), k
        assert np.array_equal(got, orc.unwarp_image_forward(mat, xc, yc, fact)), k
        k += 1
    assert k == 4


def test_baseline_config1_on_the_reference_image():
    """BASELINE config 1 as the reference runs it (examples/unwarp.py:189): its own
    data/dot_pattern_01.jpg with data/coef_dot_05.txt.  The full 2160 x 2560 result must have
    the SHA-256 of the real reference's (tests/golden/cfg1/meta.json), orders 1 and 0."""
    import hashlib
    import json
    import os
    from PIL import Image
    from discorpy_b200.losa import loadersaver as losa
    d = os.path.join(os.path.dirname(__file__), "golden", "cfg1")
    meta = json.load(open(os.path.join(d, "meta.json")))
    mat = np.array(Image.open(os.path.join(d, "dot_pattern_01.jpg")), dtype=np.float32)
    if hashlib.sha256(mat.tobytes()).hexdigest() != meta["input_sha256"]:
        pytest.skip("this PIL / libjpeg decodes the JPEG differently from the build container")
    xc, yc, fact = losa.load_metadata_txt(os.path.join(d, "coef_dot_05.txt"))
    assert (xc, yc, list(fact)) == (meta["xcenter"], meta["ycenter"], meta["list_fact"])
    sub = np.load(os.path.join(d, "reference_subsample.npz"))
    for order in (1, 0):
        got = post.unwarp_image_backward(mat, xc, yc, fact, order=order)
        assert got.dtype == np.float32 and got.shape == (2160, 2560)
        assert np.array_equal(got[::7, ::7], sub["order%d" % order])
        sha = hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest()
        assert sha == meta["output_sha256_order%d" % order], order
    # the same frame as uint8 (what the JPEG holds): SciPy's integer rounding on the device
    got8 = post.unwarp_image_backward(mat.astype(np.uint8), xc, yc, fact)
    want8 = orc.unwarp_image_backward(mat.astype(np.uint8), xc, yc, fact)
    assert got8.dtype == np.uint8 and np.array_equal(got8, want8)


def test_explicit_coordinates_outside_the_image_follow_scipys_modes():
    """_mapping / map_index= with coordinates outside the image (reference :250-251, :489-491
    hand them to SciPy): 'reflect', 'grid-mirror', 'mirror', 'wrap', 'nearest', 'constant' on the
    device, bit-identical to scipy.ndimage.map_coordinates; the two grid-interpolating modes raise."""
    from scipy.ndimage import map_coordinates
    rng = np.random.default_rng(5)
    mat = rng.random((37, 53), dtype=np.float32)
    yd = rng.uniform(-90, 130, (37, 53))
    xd = rng.uniform(-120, 170, (37, 53))
    yd[::5, ::3] = np.round(yd[::5, ::3])
    got = post._mapping(mat, xd, yd)
    assert np.array_equal(got, map_coordinates(mat, (yd.ravel(), xd.ravel()), order=1,
                                               mode="reflect").reshape(yd.shape))
    coef = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0]
    for mode in ("reflect", "grid-mirror", "mirror", "wrap", "nearest", "constant"):
        for order in (0, 1):
            for cast in (np.float64, np.float32):
                idx = (yd.astype(cast).reshape(-1, 1), xd.astype(cast).reshape(-1, 1))
                want = map_coordinates(mat, idx, order=order, mode=mode).reshape(mat.shape)
                got = post.correct_perspective_image(mat, coef, order=order, mode=mode, map_index=idx)
                assert np.array_equal(got, want), (mode, order, cast)
    dev = post.correct_perspective_image(dcb.DeviceArray.from_host(mat), coef, map_index=idx)
    assert isinstance(dev, dcb.DeviceArray)
    for mode in ("grid-wrap", "grid-constant"):
        with pytest.raises(NotImplementedError):
            post.correct_perspective_image(mat, coef, mode=mode, map_index=idx)


def test_unwarp_slice_backward_takes_any_numeric_index():
    """The reference accepts fractional indices and indices outside the image (:214-220 clips the
    row coordinate); so does the drop-in, bit-identical to the oracle."""
    rng = np.random.default_rng(6)
    stack = rng.random((5, 96, 200), dtype=np.float32)
    for index in (40.5, -3, 97, 120.25, 95.999):
        got = post.unwarp_slice_backward(stack, 101.3, 47.1, FACT5, index)
        want = orc.unwarp_slice_backward(stack, 101.3, 47.1, FACT5, index)
        assert got.dtype == np.float32 and got.shape == (5, 200)
        assert np.max(np.abs(got - want)) <= 1e-5, index
        dev = post.unwarp_slice_backward(dcb.DeviceArray.from_host(stack), 101.3, 47.1, FACT5, index)
        assert np.array_equal(dev.to_host(), got)


def test_dependent_launches_back_to_back():
    """Programmatic dependent launch (csrc/remap_image.cuh: griddepcontrol): a launch whose source
    is the output of the launch right before it in the stream must see all of it.  Three passes
    chained on the device without a host synchronisation, twice (the second round reuses the
    cached plans and reads them ahead of the grid dependency), each result against the oracle's
    chain bit for bit; a model change in the middle builds a plan between two dependent launches."""
    rng = np.random.default_rng(79)
    mat = rng.random((1100, 1500), dtype=np.float32)
    models = [(760.3, 540.9, FACT5), (700.1, 600.2, [1.0, 1e-5, -3e-9]), (760.3, 540.9, FACT5)]
    want, stages = mat, []
    for xc, yc, fact in models:
        want = orc.unwarp_image_backward(want, xc, yc, fact)
        stages.append(want)
    for order_round in range(2):
        cur = dcb.DeviceArray.from_host(mat)
        outs = []
        for xc, yc, fact in models:
            cur = post.unwarp_image_backward(cur, xc, yc, fact)
            outs.append(cur)
        for k, (o, w) in enumerate(zip(outs, stages)):
            assert np.array_equal(o.to_host(), w), (order_round, k)


def test_launches_from_two_host_threads_share_no_scheduler_state():
    """Every host thread has its own stream; the single-image and Z-stack kernels take their tile
    counters from a ring of self-resetting slots (api.cu: image_sched_slot).  Two threads
    launching at the same time must both get the oracle's bytes."""
    import threading
    rng = np.random.default_rng(80)
    mats = [rng.random((900, 1300), dtype=np.float32) for _ in range(2)]
    stack = rng.random((6, 200, 700), dtype=np.float32)
    want_img = [orc.unwarp_image_backward(m, 640.2, 455.5, FACT5) for m in mats]
    want_chunk = orc.unwarp_chunk_slices_backward(stack, 351.2, 99.5, FACT5, 0, 199)
    errors = []

    def worker(i):
        try:
            dcb.set_device(0)
            dev = dcb.DeviceArray.from_host(mats[i])
            for _ in range(12):
                got = post.unwarp_image_backward(dev, 640.2, 455.5, FACT5).to_host()
                if not np.array_equal(got, want_img[i]):
                    errors.append(("image", i))
                got = post.unwarp_chunk_slices_backward(stack, 351.2, 99.5, FACT5, 0, 199)
                if not np.array_equal(got, want_chunk):
                    errors.append(("chunk", i))
        except Exception as exc:      # pragma: no cover
            errors.append(repr(exc))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors[:4]


def test_perspective_and_combined_host_pipelines():
    """dcb_correct_perspective_image_host_f32 / dcb_unwarp_image_backward_perspective_host_f32
    (host float32 in, host float32 out, banded): the bytes of the device-resident entries and of
    the oracle, for every band count, pinned and pageable sources, a keystone whose bands reach far
    across the image, a map that leaves the image, and a denominator that changes sign inside it
    (every band then waits for the whole upload)."""
    rng = np.random.default_rng(91)
    fact = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
    cases = [((1200, 1664), [1.02, 0.01, -15.0, 0.005, 1.01, -8.0, 8e-6, -5e-6]),
             ((777, 1031), [0.93, -0.04, 40.0, 0.02, 0.97, 25.0, -6e-5, 4e-5]),
             ((640, 512), [0.6, 0.5, -100.0, -0.5, 0.6, 300.0, 2e-4, -3e-4]),      # rotation + keystone
             ((400, 600), [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, -4.03e-3, 1.3e-4])]      # denominator crosses 0 (never exactly)
    for shape, coef in cases:
        mat = rng.random(shape, dtype=np.float32) * 9.0
        pinned = dcb.pinned_empty(shape, np.float32)
        pinned[:] = mat
        want = orc.correct_perspective_image(mat, coef)
        dev = post.correct_perspective_image(dcb.DeviceArray.from_host(mat), coef).to_host()
        assert np.array_equal(dev, want), (shape, coef)
        xc, yc = shape[1] / 2 + 3.3, shape[0] / 2 - 2.2
        want2 = orc.correct_perspective_image(orc.unwarp_image_backward(mat, xc, yc, fact), coef)
        for bands in (0, 1, 2, 3, 7, 16, 32):
            post.config["bands"] = bands
            try:
                src = mat if bands in (2, 7) else pinned
                got = post.correct_perspective_image(src, coef)
                got2 = post.unwarp_image_backward_perspective(src, xc, yc, fact, coef)
                got0 = post.correct_perspective_image(src, coef, order=0)
            finally:
                post.config["bands"] = 0
            assert np.array_equal(got, want), (shape, coef, bands)
            assert np.array_equal(got2, want2), (shape, coef, bands)
            assert np.array_equal(got0, orc.correct_perspective_image(mat, coef, order=0)), (shape, bands)
    # a 48 MiB image: the unequal band schedule (2 / 2 / 4 MiB at both ends), both stages
    big = rng.random((3072, 4096), dtype=np.float32)
    coef = [1.02, 0.01, -15.0, 0.005, 1.01, -8.0, 8e-6, -5e-6]
    got = post.unwarp_image_backward_perspective(big, 2050.4, 1500.6, FACT5, coef)
    dev = post.unwarp_image_backward_perspective(dcb.DeviceArray.from_host(big), 2050.4, 1500.6,
                                                 FACT5, coef)
    assert np.array_equal(got, dev.to_host())
    got = post.correct_perspective_image(big, coef)
    assert np.array_equal(got, post.correct_perspective_image(dcb.DeviceArray.from_host(big), coef).to_host())


def test_projective_patch_path_equals_exact_path():
    """correct_perspective_image on the patch path (verified row interpolants of 1 / denominator,
    csrc/remap_image.cuh) gives the bytes of the exact division chain (DCB_IMG_FAST=0) and of
    the oracle, for a mild and a strong keystone, orders 0 and 1."""
    import os
    rng = np.random.default_rng(81)
    mat = rng.random((1200, 1664), dtype=np.float32) * 50.0
    dev = dcb.DeviceArray.from_host(mat)
    for coef in ([1.02, 0.01, -15.0, 0.005, 1.01, -8.0, 8e-6, -5e-6],
                 [0.93, -0.04, 40.0, 0.02, 0.97, 25.0, -6e-5, 4e-5]):
        for order in (1, 0):
            dcb.plan_cache_clear()
            dcb.image_stats(True, reset=True)
            fast = post.correct_perspective_image(dev, coef, order=order).to_host()
            st = dcb.image_stats(False, reset=True)
            os.environ["DCB_IMG_FAST"] = "0"
            try:
                dcb.plan_cache_clear()
                exact = post.correct_perspective_image(dev, coef, order=order).to_host()
            finally:
                del os.environ["DCB_IMG_FAST"]
                dcb.plan_cache_clear()
            assert np.array_equal(fast, exact), (coef, order)
            assert np.array_equal(fast, orc.correct_perspective_image(mat, coef, order=order))
            if abs(coef[6]) < 1e-5:     # (the strong keystone magnifies beyond the 144-wide staged box:
                assert st["rows_patch"] > 0.5 * st["rows"], st      # most of its tiles stay exact)


def test_plan_cache_eviction_and_slab_reuse():
    """More calibrations than DCB_PLAN_CACHE_MB holds: plans are evicted, their slab memory reused
    (api.cu: plan_alloc / plan_release), results stay the oracle's.  A subprocess, because the
    cache limit is read once per process."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, DCB_PLAN_CACHE_MB="16")
    script = os.path.join(os.path.dirname(__file__), "plan_cache_eviction_check.py")
    res = subprocess.run([sys.executable, script], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
