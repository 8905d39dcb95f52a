"""
Drop-in replacement of ``discorpy.post.postprocessing`` whose image / stack
functions run on an NVIDIA B200 through ``libdiscorpy_b200.so``.

Same public names, argument meaning and error behaviour as the reference
module (``/root/reference/discorpy/post/postprocessing.py``, cited per
function).  The four hot functions --

- :func:`unwarp_image_backward`            (reference ``:111-148``)
- :func:`unwarp_slice_backward`            (``:188-229``)
- :func:`unwarp_chunk_slices_backward`     (``:255-313``)
- :func:`correct_perspective_image`        (``:462-492``)

plus ``_mapping`` (``:232-252``) and the additive
:func:`unwarp_image_backward_perspective` -- call hand-written sm_100a kernels;
there is NO CPU fallback for them (a missing library or GPU raises).  The
point-list helpers (``unwarp_line_*``, residuals, ``correct_perspective_line``)
work on a few hundred dots and stay host-side NumPy, re-implemented here so
that ``discorpy.proc`` keeps working when this module is installed in its
place.

Every hot function also accepts a :class:`discorpy_b200.DeviceArray` instead of
a NumPy array; it then returns a ``DeviceArray`` without synchronising, which
is how pipelines keep data resident in HBM.
"""
import ctypes

import numpy as np
from scipy import optimize

from .. import _cabi
from .. import device as _dev
from ..device import DeviceArray
from . import _spline

_MODES = ("reflect", "grid-mirror", "constant", "grid-constant", "nearest",
          "mirror", "grid-wrap", "wrap")

#: blend / path used by the hot functions; see include/discorpy_b200.h
config = {"blend": _cabi.BLEND_EXACT, "path": _cabi.PATH_AUTO, "bands": 0,
          # host float32 stacks whose row window is at least this large go through the
          # three-stream block pipeline of post/streaming.py (upload, kernel and download overlap)
          "stream_bytes": 256 << 20}


# ---------------------------------------------------------------------------
# argument policy shared by the hot functions
# ---------------------------------------------------------------------------
def _check_order_mode(order, mode):
    """Mirror what ``scipy.ndimage.map_coordinates`` raises for bad arguments
    (``scipy/ndimage/_interpolation.py:446-447``, ``_ni_support.py:59``)."""
    if mode not in _MODES:
        raise RuntimeError("boundary mode not supported")
    if order is None or order < 0 or order > 5:
        raise RuntimeError("spline order not supported")
    return int(order)


def _wants_spline(mat, order):
    """Orders 2..5 need SciPy's float64 B-spline prefilter, and float64 images
    need float64 input / output: both take the general spline path
    (``_spline.py`` / ``csrc/spline.cuh``) instead of the float32 order-0/1
    kernels."""
    return order > 1 or (not isinstance(mat, DeviceArray)
                         and np.asarray(mat).dtype == np.float64)


#: integer dtypes that embed exactly in float32; images of these types travel
#: as float32 and come back rounded the way SciPy rounds integer outputs
_INT_IMAGE_DTYPES = (np.dtype(np.uint8), np.dtype(np.int8),
                     np.dtype(np.uint16), np.dtype(np.int16))


def _as_f32_image(mat):
    """Returns ``(float32 C-contiguous array, flags, output dtype)``.

    float32 images go to the GPU as they are.  uint8 / int8 / uint16 / int16
    images are widened exactly on the host and flagged so that the kernels
    round order-1 results half away from zero while still in float64 -- the
    reference's 'same dtype out' rule (``scipy/ndimage/_ni_support.py:83``).
    Anything else (float64, float16, 32/64-bit integers) is refused loudly
    instead of silently going to the CPU."""
    if mat.dtype == np.float32:
        return np.ascontiguousarray(mat), 0, None
    if mat.dtype in _INT_IMAGE_DTYPES:
        return (np.ascontiguousarray(mat, dtype=np.float32),
                _cabi.FLAG_ROUND_INT, mat.dtype)
    raise NotImplementedError(
        "dtype %s is not implemented on the CUDA path yet (float32, uint8, "
        "int8, uint16 and int16 are); there is no CPU fallback" % mat.dtype)


def _narrow(out, dtype):
    """float32 result -> the input's integer dtype (values are integers)."""
    return out if dtype is None else out.astype(dtype)


def _opts(order=1, flags=0):
    return _cabi.make_options(order, config["blend"], config["path"], flags)


def _vp(ptr):
    return ctypes.c_void_p(ptr)


def _radial_factor_1d(ru, list_fact):
    """Sum_i a_i ru**i in the reference's operation order (host side, 1-D)."""
    acc = None
    for i, a in enumerate(list_fact):
        term = a * ru ** i
        acc = term if acc is None else acc + term
    return np.zeros_like(ru) if acc is None else acc


def _row_yd(height, width, xcenter, ycenter, list_fact, index):
    """Clipped float64 source row coordinate of output row ``index``
    (reference ``:214-220`` / ``:289-299``); 1-D, W values, host side."""
    xu = np.arange(0, width) - xcenter
    yu = index - ycenter
    ru = np.sqrt(xu ** 2 + yu ** 2)
    flist = _radial_factor_1d(ru, list_fact)
    return np.clip(ycenter + flist * yu, 0, height - 1)


# ---------------------------------------------------------------------------
# hot functions
# ---------------------------------------------------------------------------
def unwarp_image_backward(mat, xcenter, ycenter, list_fact, order=1,
                          mode="reflect"):
    """
    Unwarp an image using a backward model (reference ``:111-148``).

    Parameters
    ----------
    mat : array_like or DeviceArray
        2D array: float32, float64, uint8, int8, uint16 or int16 (the result
        has the same dtype, rounded like SciPy rounds it).
    xcenter, ycenter : float
        Center of distortion.
    list_fact : list of float
        Polynomial coefficients of the backward model.
    order : int, optional
        Spline order 0..5.  0 (nearest) and 1 (bilinear) take the tuned
        float32 kernels; 2..5 run SciPy's float64 B-spline prefilter and
        (order+1)^2-tap interpolation on the GPU (``csrc/spline.cuh``).
    mode : str, optional
        One of SciPy's eight boundary modes.  The coordinates are clipped to
        the image before sampling, so for order 0/1 all modes give the same
        result; for order >= 2 the mode selects the prefilter's boundary
        condition (and the 12-pixel pre-padding of 'nearest' /
        'grid-constant'), as in SciPy.

    Returns
    -------
    array_like
        2D array, same shape and dtype; distortion-corrected image.
    """
    on_device = isinstance(mat, DeviceArray)
    if not on_device:
        mat = np.asarray(mat)
    (height, width) = mat.shape          # ValueError for non-2D, like :137
    order = _check_order_mode(order, mode)
    model = _cabi.make_radial(xcenter, ycenter, list_fact)
    if _wants_spline(mat, order):
        return _spline.remap(mat, order, mode, _cabi.MAP_RADIAL, radial=model)[0]
    if not on_device and mat.dtype in _INT_IMAGE_DTYPES:
        # integer images cross PCIe in their own dtype (2-4 x fewer bytes); widening and SciPy's
        # round-half-away narrowing happen on the device, not in NumPy on one host core
        # (4096^2 uint16: 29 ms -> see profiles/r1/e2e_dtypes.txt)
        return _unwarp_frame_hwc(mat[:, :, None], xcenter, ycenter, list_fact,
                                 order)[:, :, 0]
    if not on_device:
        # host in, host out: banded upload / compute / download pipeline
        src, flags, out_dtype = _as_f32_image(mat)
        opt = _opts(order, flags)
        _dev.ensure_init()
        out = _dev.pinned_empty((height, width), np.float32)
        if src is mat:
            _dev.maybe_register(src)      # a frame buffer seen before: page-lock it in place
        _cabi.call("dcb_unwarp_image_backward_host_f32", _vp(src.ctypes.data),
                   _vp(out.ctypes.data), height, width, width * 4, width * 4,
                   ctypes.byref(model), ctypes.byref(opt), config["bands"])
        return _narrow(out, out_dtype)
    opt = _opts(order)
    stream = _dev.current_stream()
    dst = DeviceArray((height, width))
    _cabi.call("dcb_unwarp_image_backward_f32", _vp(mat.ptr), _vp(dst.ptr),
               height, width, mat.pitch, dst.pitch, ctypes.byref(model),
               ctypes.byref(opt), _vp(stream.handle))
    return dst


def unwarp_image_forward(mat, xcenter, ycenter, list_fact):
    """
    Unwarp an image using a forward model (reference ``:151-185``).  Scatter
    with vacant pixels, "only for assessment": NumPy arrays are handled host-side
    exactly like the reference does; a float32 :class:`DeviceArray` is scattered on
    the GPU (same positions, same last-writer-wins rule) and stays on the device.
    """
    if isinstance(mat, DeviceArray):
        # device-resident image: deterministic two-pass scatter on the GPU (forward.cuh)
        (height, width) = mat.shape
        model = _cabi.make_radial(xcenter, ycenter, list_fact)
        stream = _dev.current_stream()
        dst = DeviceArray((height, width))
        work = _dev.device_pool.take(max(height * width * 4, 16))
        _cabi.call("dcb_unwarp_image_forward_f32", _vp(mat.ptr), _vp(dst.ptr), height,
                   width, mat.pitch, dst.pitch, ctypes.byref(model), _vp(work.ptr),
                   _vp(stream.handle))
        import weakref
        weakref.finalize(dst, _dev.device_pool.give, work)   # until the stream is done with it
        return dst
    mat = np.asarray(mat)
    (height, width) = mat.shape
    xd = np.arange(width) - xcenter
    yd = np.arange(height) - ycenter
    xd_mat, yd_mat = np.meshgrid(xd, yd)
    rd_mat = np.sqrt(xd_mat ** 2 + yd_mat ** 2)
    fact_mat = _radial_factor_1d(rd_mat, list_fact)
    xu_mat = np.intp(np.round(np.clip(xcenter + fact_mat * xd_mat, 0,
                                      width - 1)))
    yu_mat = np.intp(np.round(np.clip(ycenter + fact_mat * yd_mat, 0,
                                      height - 1)))
    out = np.zeros_like(mat)
    out[yu_mat, xu_mat] = mat
    return out


def _stack_call(src, dst, depth, height, width, src_row0, src_rows, src_pitch,
                src_slice, row0, nrows, coord_round, model, stream, order=1,
                flags=0):
    opt = _opts(order, flags)
    _cabi.call("dcb_unwarp_stack_backward_f32", _vp(src), _vp(dst.ptr), depth,
               height, width, src_row0, src_rows, src_pitch, src_slice,
               dst.pitch, dst.slice_stride, row0, nrows, coord_round,
               ctypes.byref(model), ctypes.byref(opt), _vp(stream.handle))


def unwarp_slice_backward(mat3D, xcenter, ycenter, list_fact, index):
    """
    Generate an unwarped slice [:, index, :] of a 3D dataset, i.e. one
    unwarped sinogram of a 3D tomographic data (reference ``:188-229``).

    The coordinates stay float64 (never rounded to float32) and the output is
    always float32, as in the reference.  Only the row window the reference
    crops to (``:221-223``) is sent to the GPU.

    Returns
    -------
    array_like
        2D float32 array (depth, width).
    """
    dst = _unwarp_slice_into(mat3D, xcenter, ycenter, list_fact, index, None)
    if isinstance(mat3D, DeviceArray):
        return dst
    stream = _dev.current_stream()
    return dst.to_host(stream=stream).reshape(dst.shape)


def _unwarp_slice_into(mat3D, xcenter, ycenter, list_fact, index, dst):
    """Body of :func:`unwarp_slice_backward`: the (depth, width) sinogram is
    written to ``dst`` -- anything with ``ptr`` / ``pitch`` (bytes between the
    rows of the sinogram) in device-addressable memory, e.g. this rank's rows
    of a peer window (``multigpu.SinogramWindow``) -- or to a new DeviceArray
    when ``dst`` is None.  Asynchronous on the current stream."""
    on_device = isinstance(mat3D, DeviceArray)
    if len(mat3D.shape) < 3:
        raise ValueError("Input must be a 3D data")
    (depth, height, width) = mat3D.shape
    yd = _row_yd(height, width, xcenter, ycenter, list_fact, index)
    yd_min = int(np.int16(np.floor(np.amin(yd))))
    yd_max = int(np.int16(np.ceil(np.amax(yd)))) + 1
    if (int(index) != index or not 0 <= int(index) < height
            or (not on_device and np.dtype(mat3D.dtype) == np.float64)):
        # (also float64 stacks: SciPy samples them in float64, the result is rounded once into
        # the float32 sinogram -- the float64 sampler of the spline path does the same)
        # the reference takes any number (:214-220: the row coordinate is just clipped); the
        # tiled kernel works on output rows of the image, so such a row goes through the
        # explicit-coordinate kernel, one launch per slice
        return _unwarp_slice_any_index(mat3D, xcenter, ycenter, list_fact, index, yd,
                                       yd_min, yd_max, dst)
    index = int(index)
    model = _cabi.make_radial(xcenter, ycenter, list_fact)
    stream = _dev.current_stream()
    if dst is None:
        dst = DeviceArray((depth, width))
    # the kernel sees the sinogram as a stack of depth x (1, width) slices
    out = _SliceRows(dst.ptr, dst.pitch)
    if on_device:
        src_ptr = mat3D.ptr + yd_min * mat3D.pitch
        _stack_call(src_ptr, out, depth, height, width, yd_min,
                    yd_max - yd_min, mat3D.pitch, mat3D.slice_stride, index, 1,
                    0, model, stream)
        return dst
    # integer stacks: SciPy rounds each slice to the stack's dtype before the
    # reference stores it into the float32 sinogram (:227-228)
    win, flags, _ = _upload_native(mat3D[:, yd_min:yd_max, :], stream)
    _stack_call(win.ptr, out, depth, height, width, yd_min, yd_max - yd_min,
                win.pitch, win.slice_stride, index, 1, 0, model, stream,
                flags=flags)
    dst._keep = win     # the upload stays alive until the caller is done with dst
    return dst


def _unwarp_slice_any_index(mat3D, xcenter, ycenter, list_fact, index, yd, yd_min, yd_max, dst):
    """``unwarp_slice_backward`` for a row index that is fractional or lies outside the image
    (reference ``:214-228`` verbatim in its arithmetic: float64 coordinates, window crop, order 1,
    'reflect'); rare, one explicit-coordinate launch per slice."""
    (depth, height, width) = mat3D.shape
    xu = np.arange(0, width) - xcenter
    yu = index - ycenter
    flist = _radial_factor_1d(np.sqrt(xu ** 2 + yu ** 2), list_fact)
    xd = np.clip(xcenter + flist * xu, 0, width - 1)
    ydw = yd - yd_min
    host = mat3D.to_host() if isinstance(mat3D, DeviceArray) else mat3D
    sino = np.zeros((depth, width), dtype=np.float32)
    for i in range(depth):
        sino[i] = _map_coordinates(np.asarray(host[i, yd_min:yd_max, :]), ydw, xd, 1, "reflect")
    if dst is None:
        dst = DeviceArray((depth, width))
    stream = _dev.current_stream()
    _cabi.call("dcb_h2d_2d", _vp(dst.ptr), dst.pitch, _vp(sino.ctypes.data), width * 4,
               width * 4, depth, _vp(stream.handle))
    stream.sync()           # `sino` is an ordinary array: do not outlive it
    return dst


class _SliceRows:
    """Destination of the one-row-per-slice launch: rows ``pitch`` bytes apart."""

    def __init__(self, ptr, pitch):
        self.ptr, self.pitch, self.slice_stride = ptr, pitch, pitch


def _mapping(mat, xmat, ymat):
    """
    Apply a geometric transformation to a 2D array (reference ``:232-252``):
    bilinear sampling of ``mat`` at the given coordinates; coordinates outside the
    image are reflected like SciPy's default boundary does (``_fold_coordinate``).
    """
    xmat = np.asarray(xmat)
    ymat = np.asarray(ymat)
    out = _map_coordinates(np.asarray(mat), ymat.ravel(), xmat.ravel(), 1,
                           "reflect")
    return out.reshape(xmat.shape)


def _fold_coordinate(cc, n, mode):
    """SciPy's ``map_coordinate`` (``ni_interpolation.c``) for a float64 coordinate array and an
    axis of length ``n``: where the boundary mode sends a coordinate that lies outside
    ``[0, n-1]``.  Restated for 'reflect' / 'grid-mirror', 'mirror' and 'wrap' and pinned to the
    installed SciPy bit for bit (tests/test_host_api.py); after it, sampling with clamped taps
    (what the explicit-coordinate kernel does) gives SciPy's value for orders 0 and 1."""
    cc = np.array(cc, dtype=np.float64, copy=True)
    if n <= 1:
        cc[:] = 0.0
        return cc
    neg = cc < 0
    pos = cc > n - 1
    if mode in ("reflect", "grid-mirror"):
        sz2 = 2.0 * n
        v = cc[neg]
        far = v < -sz2
        v[far] = sz2 * np.trunc(-v[far] / sz2) + v[far]
        cc[neg] = np.where(v < -n, v + sz2, -v - 1.0)
        v = cc[pos]
        v = v - sz2 * np.trunc(v / sz2)
        cc[pos] = np.where(v >= n, sz2 - v - 1.0, v)
    elif mode == "mirror":
        sz2 = 2.0 * n - 2.0
        v = cc[neg]
        v = sz2 * np.trunc(-v / sz2) + v
        v = np.where(v <= 1 - n, v + sz2, -v)
        cc[neg] = np.where(v > n - 1, sz2 - v, v)      # (n-1, n): the mirrored twin inside
        v = cc[pos]
        v = v - sz2 * np.trunc(v / sz2)
        cc[pos] = np.where(v > n - 1, sz2 - v, v)
    elif mode == "wrap":
        sz = n - 1.0
        v = cc[neg]
        cc[neg] = v + sz * (np.trunc(-v / sz) + 1)
        v = cc[pos]
        cc[pos] = v - sz * np.trunc(v / sz)
    else:
        raise NotImplementedError(mode)
    return cc


def _explicit_coordinates(yd, xd, height, width, order, mode):
    """Explicit coordinates as SciPy would read them for order 0 / 1: inside the image nothing
    changes (and float32 arrays stay float32); outside, the boundary mode decides.  Returns
    ``(yd, xd, zero_mask)`` -- ``zero_mask`` marks the samples 'constant' answers with cval = 0."""
    outside = (yd < 0) | (yd > height - 1) | (xd < 0) | (xd > width - 1)
    outside |= np.isnan(yd) | np.isnan(xd)
    if mode == "nearest" or not outside.any():
        return yd, xd, None
    if mode == "constant":
        return yd, xd, outside
    if mode in ("grid-constant", "grid-wrap"):
        raise NotImplementedError(
            "%d coordinates lie outside the image and mode %r interpolates across the border "
            "(SciPy pads / wraps the grid there); the CUDA path implements 'reflect', "
            "'grid-mirror', 'mirror', 'wrap', 'nearest' and 'constant' for such coordinates"
            % (int(outside.sum()), mode))
    return (_fold_coordinate(yd, height, mode), _fold_coordinate(xd, width, mode), None)


def _map_coordinates(mat, yd, xd, order, mode):
    (height, width) = mat.shape
    if _wants_spline(mat, order):
        out, n_oob = _spline.remap(mat, order, mode, _cabi.MAP_COORDS, yd=yd, xd=xd)
        if n_oob and mode != "nearest":
            raise NotImplementedError(
                "%d coordinates lie outside the image; the CUDA path clamps them, "
                "which equals SciPy only for mode='nearest' (got %r)" % (n_oob, mode))
        return out
    src_np, flags, out_dtype = _as_f32_image(mat)
    yd, xd = np.asarray(yd).ravel(), np.asarray(xd).ravel()
    if yd.size != xd.size:
        raise RuntimeError("invalid shape for coordinate array")
    yd, xd, zero_mask = _explicit_coordinates(yd, xd, height, width, order, mode)
    kind = np.result_type(yd.dtype, xd.dtype)
    ctype = np.float32 if kind == np.float32 else np.float64
    yd = np.ascontiguousarray(yd, dtype=ctype).ravel()
    xd = np.ascontiguousarray(xd, dtype=ctype).ravel()
    n = yd.size
    stream = _dev.current_stream()
    src = DeviceArray.from_host(src_np, stream)
    itemsize = np.dtype(ctype).itemsize
    sh = _vp(stream.handle)
    with _dev.borrowed(max(n, 1) * itemsize) as dy, \
            _dev.borrowed(max(n, 1) * itemsize) as dx, \
            _dev.borrowed(max(n, 1) * 4) as dout, _dev.borrowed(16) as dflag:
        _cabi.call("dcb_h2d", _vp(dy.ptr), _vp(yd.ctypes.data), n * itemsize, sh)
        _cabi.call("dcb_h2d", _vp(dx.ptr), _vp(xd.ctypes.data), n * itemsize, sh)
        _cabi.call("dcb_memset", _vp(dflag.ptr), 0, 16, sh)
        opt = _opts(order, flags)
        _cabi.call("dcb_map_coordinates_f32", _vp(src.ptr), _vp(dout.ptr),
                   height, width, src.pitch, _vp(dy.ptr), _vp(dx.ptr),
                   int(ctype is np.float64), n, _vp(dflag.ptr),
                   ctypes.byref(opt), sh)
        out = _dev.pinned_empty((n,), np.float32)
        flag = np.zeros(4, dtype=np.uint32)
        _cabi.call("dcb_d2h", _vp(out.ctypes.data), _vp(dout.ptr), n * 4, sh)
        _cabi.call("dcb_d2h", _vp(flag.ctypes.data), _vp(dflag.ptr), 16, sh)
        stream.sync()
    # (coordinates still outside here are 'nearest', 'constant' -- zeroed next -- or the
    # (-1, 0) / (n-1, n) band of a folded 'reflect' coordinate: clamping is SciPy's value)
    if zero_mask is not None:
        out = np.array(out)
        out[zero_mask] = 0.0
    return _narrow(out, out_dtype)


def _rows_leave_window(height, width, xcenter, ycenter, list_fact, start, stop,
                       yd_min, yd_max):
    """Can any output row start..stop sample outside the row window
    ``[yd_min, yd_max)`` the reference crops to (``:289-301``: taken from the
    first and the last row only)?  Conservative (True when in doubt), cheap:
    along a row ``yd = yc + F(r) yu`` is extreme at the row ends, at the column
    nearest to the centre, or where ``F'(r) = 0``."""
    fact = np.asarray(list_fact, dtype=np.float64)
    rows = np.arange(start, stop + 1, dtype=np.float64)
    yu = rows - ycenter
    xs = [0.0 - xcenter, (width - 1) - xcenter]
    x_near = min(max(xcenter, 0.0), width - 1.0) - xcenter     # column nearest the centre
    cand = [np.sqrt(x * x + yu * yu) for x in xs + [x_near]]
    rmin, rmax = cand[2], np.maximum(cand[0], cand[1])
    if len(fact) > 2:
        deriv = fact[1:] * np.arange(1, len(fact))
        roots = np.roots(deriv[::-1]) if np.any(deriv[1:] != 0) else np.array([])
        for root in roots:
            if abs(root.imag) < 1e-9 * max(1.0, abs(root.real)) and root.real > 0:
                cand.append(np.clip(root.real, rmin, rmax))
    lo = np.full(rows.shape, np.inf)
    hi = np.full(rows.shape, -np.inf)
    for r in cand:
        yd = np.clip(ycenter + _radial_factor_1d(r, fact) * yu, 0, height - 1)
        lo, hi = np.minimum(lo, yd), np.maximum(hi, yd)
    if not (np.all(np.isfinite(lo)) and np.all(np.isfinite(hi))):
        return True
    return bool(lo.min() < yd_min or hi.max() > yd_max - 1)


def _chunk_outside_window(mat3D, xcenter, ycenter, list_fact, start, stop,
                          yd_min, yd_max):
    """``unwarp_chunk_slices_backward`` when some rows of the chunk sample
    outside the reference's row window (a strongly off-centre, non-monotone
    model).  The reference then samples the CROPPED slices with SciPy's
    'reflect' boundary (``:310-312`` -> ``_mapping`` ``:251``); restated here:
    explicit coordinates, reflected into the window like SciPy does, through
    the explicit-coordinate kernel.  Rare and not tuned."""
    (depth, height, width) = mat3D.shape
    xu = np.arange(width) - xcenter
    yu = np.arange(start, stop + 1) - ycenter
    xu_mat, yu_mat = np.meshgrid(xu, yu)
    ru = np.sqrt(xu_mat ** 2 + yu_mat ** 2)
    fmat = _radial_factor_1d(ru, list_fact)
    xd = np.float32(np.clip(xcenter + fmat * xu_mat, 0, width - 1))
    yd = np.float32(np.clip(ycenter + fmat * yu_mat, 0, height - 1))
    yd = yd - np.int16(yd_min)                     # float32, like :308-309
    yd = _fold_coordinate(yd, yd_max - yd_min, "reflect")  # float64
    # a mapped coordinate in (-1, 0) or (n-1, n) has both taps on the edge row: clamping
    # (mode 'nearest' of the explicit-coordinate kernel) gives SciPy's value
    xd = xd.astype(np.float64)
    out = [_map_coordinates(np.asarray(mat3D[i, yd_min:yd_max, :]), yd, xd, 1,
                            "nearest").reshape(yd.shape) for i in range(depth)]
    return np.asarray(out)


def unwarp_chunk_slices_backward(mat3D, xcenter, ycenter, list_fact,
                                 start_index, stop_index):
    """
    Generate a chunk of unwarped slices [:, start_index: stop_index, :] used
    for tomographic data (reference ``:255-313``).  ``stop_index`` is
    inclusive; the index checks (including the ``-1`` wart) are
    the reference's.

    Returns
    -------
    array_like
        3D array (depth, stop-start+1, width) of ``mat3D``'s dtype.
    """
    on_device = isinstance(mat3D, DeviceArray)
    if len(mat3D.shape) < 3:
        raise ValueError("Input must be a 3D data")
    (depth, height, width) = mat3D.shape
    index_list = np.arange(height, dtype=np.int16)
    if stop_index == -1:
        stop_index = height
    if (start_index not in index_list) or (stop_index not in index_list):
        raise ValueError("Selected index is out of the range")
    start_index, stop_index = int(start_index), int(stop_index)
    if stop_index < start_index:
        raise ValueError("Selected index is out of the range")
    yd1 = _row_yd(height, width, xcenter, ycenter, list_fact, start_index)
    yd2 = _row_yd(height, width, xcenter, ycenter, list_fact, stop_index)
    yd_min = int(np.int16(np.floor(np.amin(yd1))))
    yd_max = int(np.int16(np.ceil(np.amax(yd2)))) + 1
    nrows = stop_index - start_index + 1
    if yd_max <= yd_min:
        # the reference hands SciPy an empty slice here and SciPy reads past it
        raise ValueError("empty row window [%d, %d): the model maps the last row of the "
                         "chunk above the first one (the reference's result is undefined "
                         "here)" % (yd_min, yd_max))
    model = _cabi.make_radial(xcenter, ycenter, list_fact)
    is_f64 = (not on_device) and np.dtype(mat3D.dtype) == np.float64
    if is_f64 or _rows_leave_window(height, width, xcenter, ycenter, list_fact,
                                    start_index, stop_index, yd_min, yd_max):
        # (float64 stacks -- e.g. flat-field-normalised projections -- keep float64 in and out
        # like in the reference: per slice through the float64 sampler of the spline path)
        host = mat3D.to_host() if on_device else mat3D
        res = _chunk_outside_window(host, xcenter, ycenter, list_fact,
                                    start_index, stop_index, yd_min, yd_max)
        return DeviceArray.from_host(res) if on_device else res
    stream = _dev.current_stream()
    if on_device:
        dst = DeviceArray((depth, nrows, width))
        src_ptr = mat3D.ptr + yd_min * mat3D.pitch
        _stack_call(src_ptr, dst, depth, height, width, yd_min,
                    yd_max - yd_min, mat3D.pitch, mat3D.slice_stride,
                    start_index, nrows, 1, model, stream)
        return dst
    win_bytes = depth * (yd_max - yd_min) * width * 4
    if (np.dtype(mat3D.dtype) == np.float32 and depth >= 4
            and win_bytes >= config["stream_bytes"]):
        from . import streaming
        return streaming.unwarp_chunk_slices_backward_stream(
            mat3D, xcenter, ycenter, list_fact, start_index, stop_index,
            block_bytes=int(min(256 << 20, max(32 << 20, win_bytes // 8))))
    # (the output is allocated only now: a stack that streams never needs it whole on the device)
    dst = DeviceArray((depth, nrows, width))
    win, flags, out_dtype = _upload_native(mat3D[:, yd_min:yd_max, :], stream)
    _stack_call(win.ptr, dst, depth, height, width, yd_min, yd_max - yd_min,
                win.pitch, win.slice_stride, start_index, nrows, 1, model,
                stream, flags=flags)
    return _download_native(dst, out_dtype, stream)


_DTYPE_CODES = {np.dtype(np.float32): _cabi.DTYPE_F32,
                np.dtype(np.uint8): _cabi.DTYPE_U8,
                np.dtype(np.int8): _cabi.DTYPE_I8,
                np.dtype(np.uint16): _cabi.DTYPE_U16,
                np.dtype(np.int16): _cabi.DTYPE_I16}


def _host_pipeline_f32(entry, mat, models, order):
    """float32 host image -> pinned float32 host result through one of the banded
    host-buffer entries (``dcb_*_host_f32``: upload, kernels and download overlap
    in row bands; a pageable ``mat`` is staged band by band by the library)."""
    (height, width) = mat.shape
    src = mat if mat.flags.c_contiguous else np.ascontiguousarray(mat)
    opt = _opts(order)
    _dev.ensure_init()
    out = _dev.pinned_empty((height, width), np.float32)
    if src is mat:
        _dev.maybe_register(src)      # a frame buffer seen before: page-lock it in place
    args = [ctypes.byref(m) for m in models]
    _cabi.call(entry, _vp(src.ctypes.data), _vp(out.ctypes.data), height, width,
               width * 4, width * 4, *args, ctypes.byref(opt), config["bands"])
    return out


_STAGE_CHUNK = 4 << 20


def _h2d_raw(dptr, raw, stream):
    """Bytes of the C-contiguous host array ``raw`` -> device address ``dptr`` on ``stream``.
    Page-locked arrays are read by the DMA engine directly.  Pageable ones of 8 MiB and more
    are copied chunk by chunk into a page-locked block by the library's copy pool
    (non-temporal stores, ``dcb_host_copy_2d``) while the DMA engine moves the previous chunk
    -- the driver's own staging of pageable memory runs on one thread (4096^2 uint16: 2.3 ms
    of a 3 ms call).  Returns an object to keep alive until the stream has been synchronised
    (or None)."""
    n = raw.nbytes
    if n < (8 << 20) or _dev.is_pinned(raw):
        _cabi.call("dcb_h2d", _vp(dptr), _vp(raw.ctypes.data), n, _vp(stream.handle))
        return None
    bufs = [_dev.pinned_empty((_STAGE_CHUNK,), np.uint8) for _ in range(2)]
    evs = [_dev.Event(), _dev.Event()]
    for k, off in enumerate(range(0, n, _STAGE_CHUNK)):
        b, m = k & 1, min(_STAGE_CHUNK, n - off)
        if k >= 2:
            evs[b].sync()          # the DMA that last read this block is done
        _cabi.call("dcb_host_copy_2d", _vp(bufs[b].ctypes.data), m,
                   _vp(raw.ctypes.data + off), m, m, 1)
        _cabi.call("dcb_h2d", _vp(dptr + off), _vp(bufs[b].ctypes.data), m,
                   _vp(stream.handle))
        evs[b].record(stream)
    return bufs


def _upload_native(arr, stream):
    """Host array (H, W) or (D, H, W) of a supported dtype -> (float32
    DeviceArray of the same shape, kernel flags, dtype to return or None).
    Integer data crosses PCIe in its own dtype and is widened on the device
    (exactly: every such value is a float32) instead of by NumPy on one host
    core."""
    arr = np.asarray(arr)
    if arr.dtype == np.float32:
        return DeviceArray.from_host(arr, stream), 0, None
    if arr.dtype not in _INT_IMAGE_DTYPES:
        raise NotImplementedError(
            "dtype %s is not implemented on the CUDA path yet (float32, uint8, "
            "int8, uint16 and int16 are); there is no CPU fallback" % arr.dtype)
    raw = np.ascontiguousarray(arr)
    dev = DeviceArray(raw.shape)
    rows = int(np.prod(raw.shape[:-1], dtype=np.int64))
    sh = _vp(stream.handle)
    with _dev.borrowed(max(raw.nbytes, 16)) as draw:
        keep = _h2d_raw(draw.ptr, raw, stream)
        _cabi.call("dcb_unpack_hwc_to_planes_f32", _vp(draw.ptr),
                   _DTYPE_CODES[raw.dtype], _vp(dev.ptr), rows, raw.shape[-1],
                   1, dev.pitch, dev.pitch * rows, sh)
        stream.sync()     # `raw`, the staging block and the borrowed buffer are free again
        del keep
    return dev, _cabi.FLAG_ROUND_INT, raw.dtype


def _download_native(dev, dtype, stream):
    """float32 DeviceArray holding integer values (FLAG_ROUND_INT results) ->
    host array of ``dtype``, narrowed on the device; ``dtype`` None: float32."""
    if dtype is None:
        return dev.to_host(stream=stream)
    rows = int(np.prod(dev.shape[:-1], dtype=np.int64))
    out = _dev.pinned_empty(dev.shape, dtype)
    sh = _vp(stream.handle)
    with _dev.borrowed(max(out.nbytes, 16)) as draw:
        _cabi.call("dcb_pack_planes_f32_to_hwc", _vp(dev.ptr), _vp(draw.ptr),
                   _DTYPE_CODES[np.dtype(dtype)], rows, dev.shape[-1], 1,
                   dev.pitch, dev.pitch * rows, sh)
        _cabi.call("dcb_d2h", _vp(out.ctypes.data), _vp(draw.ptr), out.nbytes,
                   sh)
        stream.sync()
    return out


def _unwarp_frame_hwc(frame, xcenter, ycenter, list_fact, order):
    """A host (H, W, C) frame of any supported dtype through the Z-stack
    kernel with the image numerics (fp32-rounded coordinates).  The frame
    crosses PCIe in its own dtype; de-interleaving, widening and the way back
    happen on the device, and all channels share ONE remap launch in which the
    geometry is evaluated once per tile.  Used for colour images."""
    (height, width, chan) = frame.shape
    if frame.dtype not in _DTYPE_CODES:
        raise NotImplementedError(
            "dtype %s is not implemented on the CUDA path yet (float32, "
            "uint8, int8, uint16 and int16 are); there is no CPU fallback"
            % frame.dtype)
    code = _DTYPE_CODES[frame.dtype]
    flags = 0 if code == _cabi.DTYPE_F32 else _cabi.FLAG_ROUND_INT
    raw = np.ascontiguousarray(frame)
    model = _cabi.make_radial(xcenter, ycenter, list_fact)
    stream = _dev.current_stream()
    sh = _vp(stream.handle)
    planes = DeviceArray((chan, height, width))
    dst = DeviceArray((chan, height, width))
    out = _dev.pinned_empty(raw.shape, raw.dtype)
    with _dev.borrowed(max(raw.nbytes, 16)) as draw:
        keep = _h2d_raw(draw.ptr, raw, stream)
        if keep is None and not _dev.is_pinned(raw):
            stream.sync()
        _cabi.call("dcb_unpack_hwc_to_planes_f32", _vp(draw.ptr), code,
                   _vp(planes.ptr), height, width, chan, planes.pitch,
                   planes.slice_stride, sh)
        _stack_call(planes.ptr, dst, chan, height, width, 0, height,
                    planes.pitch, planes.slice_stride, 0, height, 1, model,
                    stream, order=order, flags=flags)
        _cabi.call("dcb_pack_planes_f32_to_hwc", _vp(dst.ptr), _vp(draw.ptr),
                   code, height, width, chan, dst.pitch, dst.slice_stride, sh)
        _cabi.call("dcb_d2h", _vp(out.ctypes.data), _vp(draw.ptr), raw.nbytes,
                   sh)
        stream.sync()
        del keep
    return out


def _generate_perspective_map(mat, list_coef):
    """
    Generate mapping indices between images (reference ``:444-459``).  Host
    NumPy: callers use it to precompute ``map_index`` once and pass it to
    :func:`correct_perspective_image` many times.
    """
    c1, c2, c3, c4, c5, c6, c7, c8 = list_coef
    (height, width) = mat.shape
    xu_mat, yu_mat = np.meshgrid(np.arange(width), np.arange(height))
    den = c7 * xu_mat + c8 * yu_mat + 1.0
    xd_mat = (c1 * xu_mat + c2 * yu_mat + c3) / den
    yd_mat = (c4 * xu_mat + c5 * yu_mat + c6) / den
    xd_mat = np.float32(np.clip(xd_mat, 0, width - 1))
    yd_mat = np.float32(np.clip(yd_mat, 0, height - 1))
    return np.reshape(yd_mat, (-1, 1)), np.reshape(xd_mat, (-1, 1))


def correct_perspective_image(mat, list_coef, order=1, mode="reflect",
                              map_index=None):
    """
    Apply perspective correction to an image (reference ``:462-492``).

    Parameters
    ----------
    mat : array_like or DeviceArray
        2D float32 array.
    list_coef : list of floats
        Coefficients of the backward-mapping matrix (eight).
    order, mode : see :func:`unwarp_image_backward`.
    map_index : tuple of array_like, optional
        Precomputed ``(yd, xd)`` indices; generated on the GPU if None.

    Returns
    -------
    array_like
        Corrected image.
    """
    if len(list_coef) != 8:
        raise ValueError("!!! Eight coefficients are required !!!")
    on_device = isinstance(mat, DeviceArray)
    if not on_device:
        mat = np.asarray(mat)
    (height, width) = mat.shape
    order = _check_order_mode(order, mode)
    if map_index is not None:
        yd, xd = map_index
        if on_device:
            # the coordinate arrays are host arrays in the reference's signature; the image
            # makes one round trip (this path is not tuned) and the result stays on the device
            out = _map_coordinates(mat.to_host(), np.asarray(yd), np.asarray(xd), order, mode)
            return DeviceArray.from_host(out.reshape((height, width)))
        out = _map_coordinates(mat, np.asarray(yd), np.asarray(xd), order,
                               mode)
        return out.reshape((height, width))
    model = _cabi.make_persp(list_coef)
    if _wants_spline(mat, order):
        return _spline.remap(mat, order, mode, _cabi.MAP_PERSP, persp=model)[0]
    if not on_device and mat.dtype == np.float32:
        # host in, host out: banded upload / compute / download pipeline
        return _host_pipeline_f32("dcb_correct_perspective_image_host_f32", mat,
                                  (model,), order)
    stream = _dev.current_stream()
    flags, out_dtype = 0, None
    if on_device:
        src = mat
    else:
        src, flags, out_dtype = _upload_native(mat, stream)
    dst = DeviceArray((height, width))
    opt = _opts(order, flags)
    _cabi.call("dcb_correct_perspective_image_f32", _vp(src.ptr),
               _vp(dst.ptr), height, width, src.pitch, dst.pitch,
               ctypes.byref(model), ctypes.byref(opt), _vp(stream.handle))
    return dst if on_device else _download_native(dst, out_dtype, stream)


def unwarp_image_backward_perspective(mat, xcenter, ycenter, list_fact,
                                      list_coef, order=1, mode="reflect"):
    """
    Radial unwarp followed by perspective correction in one call, the combined
    entry BASELINE.json names.  The reference has no such function; the result
    is defined as ``correct_perspective_image(unwarp_image_backward(mat, ...),
    list_coef)`` (``examples/readthedocs_demo/demo_05.py:127,147``) with the
    intermediate image rounded to the image dtype -- both passes run back to
    back on the GPU, the intermediate never leaves HBM.
    """
    if len(list_coef) != 8:
        raise ValueError("!!! Eight coefficients are required !!!")
    on_device = isinstance(mat, DeviceArray)
    if not on_device:
        mat = np.asarray(mat)
    (height, width) = mat.shape
    order = _check_order_mode(order, mode)
    radial = _cabi.make_radial(xcenter, ycenter, list_fact)
    persp = _cabi.make_persp(list_coef)
    if _wants_spline(mat, order):
        tmp = _spline.remap(mat, order, mode, _cabi.MAP_RADIAL, radial=radial)[0]
        return _spline.remap(tmp, order, mode, _cabi.MAP_PERSP, persp=persp)[0]
    if not on_device and mat.dtype == np.float32:
        # host in, host out: both passes band by band between the two PCIe directions
        return _host_pipeline_f32("dcb_unwarp_image_backward_perspective_host_f32", mat,
                                  (radial, persp), order)
    stream = _dev.current_stream()
    flags, out_dtype = 0, None
    if on_device:
        src = mat
    else:
        src, flags, out_dtype = _upload_native(mat, stream)
    tmp = DeviceArray((height, width))
    dst = DeviceArray((height, width))
    opt = _opts(order, flags)
    _cabi.call("dcb_unwarp_image_backward_perspective_f32", _vp(src.ptr),
               _vp(dst.ptr), _vp(tmp.ptr), height, width, src.pitch, dst.pitch,
               tmp.pitch, ctypes.byref(radial), ctypes.byref(persp),
               ctypes.byref(opt), _vp(stream.handle))
    if on_device:
        dst._keepalive = tmp      # until the stream has consumed it
        return dst
    return _download_native(dst, out_dtype, stream)


# ---------------------------------------------------------------------------
# point-list helpers (host side; tiny inputs)
# ---------------------------------------------------------------------------
def unwarp_line_forward(list_lines, xcenter, ycenter, list_fact):
    """
    Unwarp lines of dot-centroids using a forward model (reference ``:36-64``).

    Returns a list of 2D arrays of unwarped (y, x) coordinates.
    """
    fact = np.asarray(list_fact, dtype=np.float64)
    expo = np.arange(len(fact), dtype=np.int16)
    list_ulines = []
    for line in list_lines:
        line = np.asarray(line)
        uline = np.zeros_like(line)
        for j in range(len(line)):
            xd = line[j, 1] - xcenter
            yd = line[j, 0] - ycenter
            rd = np.sqrt(xd * xd + yd * yd)
            factor = np.sum(fact * np.power(rd, expo))
            uline[j, 1] = xcenter + factor * xd
            uline[j, 0] = ycenter + factor * yd
        list_ulines.append(uline)
    return list_ulines


def _func_diff(ru, rd, *list_fact):
    poly = np.sum(np.asarray([a * ru ** i for i, a in enumerate(list_fact)]))
    return (rd - ru * poly) ** 2


def unwarp_line_backward(list_lines, xcenter, ycenter, list_fact):
    """
    Unwarp lines of dot-centroids using a backward model (reference
    ``:72-108``): the undistorted radius of every dot is found by numerical
    minimisation of ``(rd - ru * F(ru))**2`` started at ``rd``.
    """
    list_ulines = []
    for line in list_lines:
        line = np.asarray(line)
        uline = np.zeros_like(line)
        for j in range(len(line)):
            xd = line[j, 1] - xcenter
            yd = line[j, 0] - ycenter
            rd = np.sqrt(xd * xd + yd * yd)
            res = optimize.minimize(_func_diff, rd,
                                    args=(rd,) + tuple(list_fact))
            ru = res.x[0]
            factor = ru / rd if rd != 0.0 else 0.0
            uline[j, 1] = xcenter + factor * xd
            uline[j, 0] = ycenter + factor * yd
        list_ulines.append(uline)
    return list_ulines


def _residual(list_ulines, xcenter, ycenter, horizontal):
    rows = []
    for line in list_ulines:
        line = np.asarray(line)
        y = line[:, 0] - ycenter
        x = line[:, 1] - xcenter
        if horizontal:
            (a, b) = np.polyfit(x, y, 1)
            dist = np.abs(a * x - y + b) / np.sqrt(a ** 2 + 1)
        else:
            (a, b) = np.polyfit(y, x, 1)
            dist = np.abs(a * y - x + b) / np.sqrt(a ** 2 + 1)
        radius = np.sqrt(x ** 2 + y ** 2)
        rows.extend(np.stack([radius, dist], axis=1))
    data = np.asarray(rows)
    return data[data[:, 0].argsort()]


def calc_residual_hor(list_ulines, xcenter, ycenter):
    """
    Distances of unwarped dots on each horizontal line to its fitted straight
    line (reference ``:316-351``).  Returns rows ``[radius, residual]`` sorted
    by radius.
    """
    return _residual(list_ulines, xcenter, ycenter, True)


def calc_residual_ver(list_ulines, xcenter, ycenter):
    """Same as :func:`calc_residual_hor` for vertical lines (``:354-388``)."""
    return _residual(list_ulines, xcenter, ycenter, False)


def check_distortion(list_data):
    """
    True if more than 15% of the dots have a residual above one pixel
    (reference ``:391-411``).
    """
    res = np.asarray(list_data[:, 1])
    return bool(1.0 * np.count_nonzero(res > 1.0) / len(res) > 0.15)


def correct_perspective_line(list_lines, list_coef):
    """
    Apply perspective correction to lines of (y, x) points (reference
    ``:414-441``).
    """
    if len(list_coef) != 8:
        raise ValueError("!!! Eight coefficients are required !!!")
    c1, c2, c3, c4, c5, c6, c7, c8 = list_coef
    list_clines = []
    for iline in list_lines:
        line = np.asarray(iline)
        x = line[:, 1]
        y = line[:, 0]
        den = c7 * x + c8 * y + 1.0
        xn = (c1 * x + c2 * y + c3) / den
        yn = (c4 * x + c5 * y + c6) / den
        list_clines.append(np.stack([yn, xn], axis=1))
    return list_clines
