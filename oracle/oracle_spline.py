"""
oracle_spline -- CPU restatement (NumPy) of what ``scipy.ndimage.map_coordinates``
does for spline orders 2..5: the float64 B-spline prefilter and the
(order+1)^2-tap interpolation, for coordinates that already lie inside the image.

*** TEST INFRASTRUCTURE, NOT PRODUCT *** (same rules as oracle_np.py).

The reference reaches this code through ``order=`` / ``mode=`` of
``discorpy/post/postprocessing.py:111-148`` (``:147``), ``:462-492`` (``:491``)
and ``discorpy/util/utility.py:278-342`` (``:333, :338``); ``demo_07.py:60``
uses order 3.  The arithmetic lives in SciPy (``scipy/ndimage/_interpolation.py
:375-476`` is the readable wrapper: pre-padding ``:212-227``, prefilter
``:467-469``; the C bodies ``spline_filter1d`` / ``geometric_transform`` are
binary only).  Parity status: PINNED against the installed SciPy 1.18.1 --
``tests/test_oracle_spline.py`` compares every function below bit-for-bit with
``scipy.ndimage.spline_filter`` / ``map_coordinates`` for all orders and all
eight boundary modes, and ``tests/golden/`` holds reference outputs generated
through the real ``discorpy`` functions.

Operation order matters for bit-exactness and is spelled out; nothing is fused.
"""
import math

import numpy as np

MODES = ("reflect", "grid-mirror", "constant", "grid-constant", "nearest",
         "mirror", "grid-wrap", "wrap")
NPAD = 12           # scipy/ndimage/_interpolation.py:214


# Poles of the B-spline prefilters (roots of the order-n B-spline's z-transform
# inside the unit circle, Unser et al. 1993): sqrt(8)-3, sqrt(3)-2, ... SciPy
# holds them as literals that are the correctly rounded doubles of the exact
# values -- evaluating ``math.sqrt(3.0) - 2.0`` in double loses 1-2 bits to
# cancellation and does NOT reproduce SciPy bit-for-bit.  The hex literals below
# are those correctly rounded doubles (``tests/test_oracle_spline.py`` re-derives
# them with 60-digit decimal arithmetic).
_POLES = {
    2: (float.fromhex("-0x1.5f619980c4337p-3"),),
    3: (float.fromhex("-0x1.126145e9ecd56p-2"),),
    4: (float.fromhex("-0x1.72036f2fc0817p-2"), float.fromhex("-0x1.c1c13efa52247p-7")),
    5: (float.fromhex("-0x1.b8e8be69086f0p-2"), float.fromhex("-0x1.610b778d2f346p-5")),
}


def poles(order):
    try:
        return list(_POLES[order])
    except KeyError:
        raise RuntimeError("spline order not supported")


def filter_kind(mode):
    """Which boundary condition the prefilter applies for a map_coordinates mode,
    and whether the input is pre-padded (``_interpolation.py:212-227``)."""
    if mode in ("reflect", "grid-mirror"):
        return "reflect", 0
    if mode == "nearest":
        return "reflect", NPAD
    if mode == "grid-wrap":
        return "wrap", 0
    if mode == "grid-constant":
        return "mirror", NPAD
    if mode in ("mirror", "constant", "wrap"):
        return "mirror", 0
    raise RuntimeError("boundary mode not supported")


def tap_kind(mode):
    """How a tap index outside [0, len) is folded back when the coordinate
    itself is in range (constant / wrap have no exact spline boundary: mirror)."""
    if mode in ("reflect", "grid-mirror"):
        return "reflect"
    if mode == "grid-wrap":
        return "wrap"
    return "mirror"        # mirror, constant, wrap; nearest / grid-constant never reach it (padded)


def filter_lines(c, order, kind):
    """In-place prefilter of the lines c[:, j] (axis 0 is the filtered axis),
    float64.  Gain first, then per pole: causal initialisation, forward
    recursion, anticausal initialisation, backward recursion."""
    n = c.shape[0]
    if n < 2:
        return c
    zs = poles(order)
    gain = 1.0
    for z in zs:
        gain *= (1.0 - z) * (1.0 - 1.0 / z)
    c *= gain
    for z in zs:
        if kind == "mirror":
            z_n_1 = math.pow(z, n - 1)
            c[0] = z_n_1 * c[n - 1] + c[0]
            z_i = z
            for i in range(1, n - 1):
                c[0] += z_i * (c[i] + z_n_1 * c[n - 1 - i])
                z_i *= z
            c[0] /= 1 - z_n_1 * z_n_1
        elif kind == "reflect":
            z_n = math.pow(z, n)
            c0 = c[0].copy()
            acc = c0 + z_n * c[n - 1]
            z_i = z
            for i in range(1, n):
                # the running sum lives in a register: c[n-1-i] at i = n-1 is the ORIGINAL c[0]
                acc = acc + z_i * (c[i] + z_n * c[n - 1 - i])
                z_i *= z
            c[0] = (acc * z) / (1 - z_n * z_n) + c0
        else:  # wrap
            z_i = z
            for i in range(1, n):
                c[0] += z_i * c[n - i]
                z_i *= z
            c[0] /= 1 - z_i
        for i in range(1, n):
            c[i] += z * c[i - 1]
        if kind == "mirror":
            c[n - 1] = (z * c[n - 2] + c[n - 1]) * z / (z * z - 1)
        elif kind == "reflect":
            c[n - 1] *= z / (z - 1)
        else:
            z_i = z
            for i in range(0, n - 1):
                c[n - 1] += z_i * c[i]
                z_i *= z
            c[n - 1] *= z / (z_i - 1)
        for i in range(n - 2, -1, -1):
            c[i] = z * (c[i + 1] - c[i])
    return c


def spline_filter(mat, order, mode):
    """``scipy.ndimage.spline_filter(mat, order, output=float64, mode=mode)`` of a
    2-D array: axis 0 first, then axis 1 (``_interpolation.py:185-188``)."""
    kind, _ = filter_kind(mode)
    c = np.array(mat, dtype=np.float64)
    filter_lines(c, order, kind)              # lines along axis 0
    ct = np.ascontiguousarray(c.T)
    filter_lines(ct, order, kind)             # lines along axis 1
    return np.ascontiguousarray(ct.T)


def weights(x, order):
    """(start offsets, weights[order+1]) of SciPy's B-spline basis at coordinate
    x (float64 array).  The last weight is 1 minus the others, in index order."""
    x = np.asarray(x, dtype=np.float64)
    if order & 1:
        fl = np.floor(x)
    else:
        fl = np.floor(x + 0.5)
    start = fl.astype(np.intp) - order // 2
    x = x - fl
    y = x
    z = 1.0 - x
    w = [None] * (order + 1)
    if order == 2:
        w[1] = 0.75 - x * x
        y = 0.5 - x
        w[0] = 0.5 * y * y
    elif order == 3:
        w[1] = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0
        w[2] = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0
        w[0] = z * z * z / 6.0
    elif order == 4:
        t = x * x
        w[2] = t * (t * 0.25 - 0.625) + 115.0 / 192.0
        y = 1.0 + x
        w[1] = y * (y * (y * (5.0 - y) / 6.0 - 1.25) + 5.0 / 24.0) + 55.0 / 96.0
        w[3] = z * (z * (z * (5.0 - z) / 6.0 - 1.25) + 5.0 / 24.0) + 55.0 / 96.0
        y = 0.5 - x
        t = y * y
        w[0] = t * t / 24.0
    elif order == 5:
        t = y * y
        w[2] = t * (t * (0.25 - y / 12.0) - 0.5) + 0.55
        t = z * z
        w[3] = t * (t * (0.25 - z / 12.0) - 0.5) + 0.55
        y = y + 1.0
        w[1] = y * (y * (y * (y * (y / 24.0 - 0.375) + 1.25) - 1.75) + 0.625) + 0.425
        z = z + 1.0
        w[4] = z * (z * (z * (z * (z / 24.0 - 0.375) + 1.25) - 1.75) + 0.625) + 0.425
        z = z - 1.0
        t = z * z
        w[0] = z * t * t / 120.0
    else:
        raise RuntimeError("spline order not supported")
    last = np.ones_like(x)
    for i in range(order):
        last = last - w[i]
    w[order] = last
    return start, w


def fold(idx, n, kind):
    """Tap index -> array index for taps that leave [0, n) (in-range coordinate)."""
    idx = np.asarray(idx)
    if n <= 1:
        return np.zeros_like(idx)
    if kind == "mirror":
        s2 = 2 * n - 2
        m = np.mod(idx, s2)
        return np.where(m >= n, s2 - m, m)
    if kind == "reflect":
        s2 = 2 * n
        m = np.mod(idx, s2)
        return np.where(m >= n, s2 - 1 - m, m)
    return np.mod(idx, n)      # wrap


def sample_spline(mat, yd, xd, order, mode="reflect", out_dtype=None):
    """``map_coordinates(mat, (yd, xd), order=order, mode=mode)`` for 2-D ``mat``,
    order 2..5 and coordinates inside ``[0, H-1] x [0, W-1]``: prefilter (of the
    padded image for nearest / grid-constant, coordinates shifted by the pad),
    then for every point  sum_i sum_j (c[i][j] * wy[i]) * wx[j]  accumulated
    row-major in float64, one cast at the end (``_cast_like_scipy``)."""
    from .oracle_np import _cast_like_scipy
    mat = np.asarray(mat)
    out_dtype = mat.dtype if out_dtype is None else np.dtype(out_dtype)
    kind, npad = filter_kind(mode)
    src = mat
    if npad:
        if mode == "nearest":
            src = np.pad(mat, npad, mode="edge")
        else:
            src = np.pad(mat, npad, mode="constant", constant_values=0.0)
    coef = spline_filter(src, order, mode)
    h, w = coef.shape
    shape = np.shape(yd)
    y = np.asarray(yd, dtype=np.float64).ravel() + npad
    x = np.asarray(xd, dtype=np.float64).ravel() + npad
    sy, wy = weights(y, order)
    sx, wx = weights(x, order)
    tk = tap_kind(mode)
    val = np.zeros(y.shape, dtype=np.float64)
    for i in range(order + 1):
        yi = fold(sy + i, h, tk)
        for j in range(order + 1):
            xj = fold(sx + j, w, tk)
            val = val + (coef[yi, xj] * wy[i]) * wx[j]
    return _cast_like_scipy(val, out_dtype).reshape(shape)


# --------------------------------------------------------------------------
# the public functions of the path at any order (0/1 delegate to oracle_np)
# --------------------------------------------------------------------------
def sample_any(mat, yd, xd, order, mode):
    from . import oracle_np
    if order <= 1:
        return oracle_np.sample(mat, yd, xd, order)
    return sample_spline(mat, yd, xd, order, mode)


def unwarp_image_backward(mat, xcenter, ycenter, list_fact, order=1, mode="reflect"):
    """``postprocessing.py:111-148`` at any spline order."""
    from . import oracle_np
    (height, width) = np.shape(mat)
    yd, xd = oracle_np.radial_coords(height, width, xcenter, ycenter, list_fact)
    return sample_any(np.asarray(mat), yd, xd, order, mode)


def correct_perspective_image(mat, list_coef, order=1, mode="reflect", map_index=None):
    """``postprocessing.py:462-492`` at any spline order."""
    from . import oracle_np
    if len(list_coef) != 8:
        raise ValueError("!!! Eight coefficients are required !!!")
    (height, width) = np.shape(mat)
    if map_index is None:
        yd, xd = oracle_np.persp_coords(height, width, list_coef)
    else:
        yd, xd = map_index
    out = sample_any(np.asarray(mat), np.reshape(yd, -1), np.reshape(xd, -1), order, mode)
    return out.reshape((height, width))


def unwarp_color_image_backward(mat, xcenter, ycenter, list_fact, order=1, mode="reflect"):
    """``discorpy/util/utility.py:278-342`` with ``pad=False`` at any spline order."""
    mat = np.asarray(mat)
    if mat.ndim == 2:
        return unwarp_image_backward(mat, xcenter, ycenter, list_fact, order, mode)
    planes = [unwarp_image_backward(mat[:, :, i], xcenter, ycenter, list_fact, order, mode)
              for i in range(mat.shape[-1])]
    return np.moveaxis(np.asarray(planes), 0, 2)
