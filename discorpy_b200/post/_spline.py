"""
Host side of the spline path (orders 2..5, and float64 images at any order):
``scipy.ndimage.map_coordinates``' float64 B-spline prefilter and
(order+1)^2-tap interpolation on the GPU (``csrc/spline.cuh``), reached through
``order=`` / ``mode=`` of the reference functions (``postprocessing.py:147``,
``:491``; ``util/utility.py:333, :338``).  No CPU fallback.
"""
import ctypes

import numpy as np

from .. import _cabi
from .. import device as _dev
from ..device import DeviceArray

_INT_DTYPES = (np.dtype(np.uint8), np.dtype(np.int8), np.dtype(np.uint16),
               np.dtype(np.int16))


def _vp(ptr):
    return ctypes.c_void_p(ptr)


def supported(dtype):
    dtype = np.dtype(dtype)
    return dtype in (np.dtype(np.float32), np.dtype(np.float64)) or dtype in _INT_DTYPES


def remap(mat, order, mode, map_kind, radial=None, persp=None, yd=None, xd=None):
    """Sample ``mat`` (2-D NumPy array: float32, float64, uint8/int8/uint16/int16,
    or a float32 :class:`DeviceArray`) with spline ``order`` and boundary
    ``mode`` through the radial map, the projective map or explicit coordinates.

    Returns ``(result, n_outside)``: an array of ``mat``'s dtype -- shape
    ``mat.shape`` for the two maps, ``(n,)`` for explicit coordinates (a
    ``DeviceArray`` for device input) -- and the number of explicit coordinates
    that lay outside the image and were clamped."""
    on_device = isinstance(mat, DeviceArray)
    (height, width) = mat.shape
    mode_code = _cabi.MODES[mode]
    if not on_device and not supported(np.asarray(mat).dtype):   # before any device work
        raise NotImplementedError(
            "dtype %s is not implemented on the CUDA path (float32, float64, "
            "uint8, int8, uint16 and int16 are); there is no CPU fallback"
            % np.asarray(mat).dtype)
    stream = _dev.current_stream()
    sh = _vp(stream.handle)
    flags, lo, hi, out_dtype = 0, 0.0, 0.0, None
    if on_device:
        src_f64, dst_f64 = 0, 0
        src_ptr, src_pitch, keep = mat.ptr, mat.pitch, mat
    else:
        mat = np.asarray(mat)
        if mat.dtype == np.float64:
            src_np, src_f64, dst_f64 = np.ascontiguousarray(mat), 1, 1
        elif mat.dtype == np.float32:
            src_np, src_f64, dst_f64 = np.ascontiguousarray(mat), 0, 0
        else:
            info = np.iinfo(mat.dtype)
            src_np = np.ascontiguousarray(mat, dtype=np.float32)   # exact
            src_f64, dst_f64, out_dtype = 0, 0, mat.dtype
            flags, lo, hi = _cabi.FLAG_ROUND_INT, float(info.min), float(info.max)
        keep = _dev.device_pool.take(max(src_np.nbytes, 16))
        src_ptr, src_pitch = keep.ptr, width * src_np.itemsize
        _cabi.call("dcb_h2d", _vp(src_ptr), _vp(src_np.ctypes.data), src_np.nbytes, sh)
        if not _dev.is_pinned(src_np):
            stream.sync()
    need = ctypes.c_size_t(0)
    _cabi.call("dcb_spline_workspace_bytes", height, width, order, mode_code,
               ctypes.byref(need))
    work = _dev.device_pool.take(max(need.value, 16))
    bufs = [work] if on_device else [work, keep]
    try:
        _cabi.call("dcb_spline_prefilter", _vp(src_ptr), src_f64, height, width,
                   src_pitch, order, mode_code, _vp(work.ptr), need.value, sh)
        esz = 8 if dst_f64 else 4
        n_out, n_oob = 0, 0
        if map_kind == _cabi.MAP_COORDS:
            kind = np.result_type(yd.dtype, xd.dtype)
            ctype = np.float32 if kind == np.float32 else np.float64
            yd = np.ascontiguousarray(yd, dtype=ctype).ravel()
            xd = np.ascontiguousarray(xd, dtype=ctype).ravel()
            if yd.size != xd.size:
                raise RuntimeError("invalid shape for coordinate array")
            n_out = yd.size
            csz = np.dtype(ctype).itemsize
            dy = _dev.device_pool.take(max(n_out * csz, 16))
            dx = _dev.device_pool.take(max(n_out * csz, 16))
            dflag = _dev.device_pool.take(16)
            dout = _dev.device_pool.take(max(n_out * esz, 16))
            bufs += [dy, dx, dflag, dout]
            _cabi.call("dcb_h2d", _vp(dy.ptr), _vp(yd.ctypes.data), n_out * csz, sh)
            _cabi.call("dcb_h2d", _vp(dx.ptr), _vp(xd.ctypes.data), n_out * csz, sh)
            _cabi.call("dcb_memset", _vp(dflag.ptr), 0, 16, sh)
            _cabi.call("dcb_spline_remap", _vp(work.ptr), height, width, order,
                       mode_code, _vp(dout.ptr), dst_f64, n_out * esz, map_kind,
                       None, None, _vp(dy.ptr), _vp(dx.ptr),
                       int(ctype is np.float64), n_out, _vp(dflag.ptr), flags, lo,
                       hi, sh)
            out = _dev.pinned_empty((n_out,), np.float64 if dst_f64 else np.float32)
            flag = np.zeros(4, dtype=np.uint32)
            _cabi.call("dcb_d2h", _vp(out.ctypes.data), _vp(dout.ptr), n_out * esz, sh)
            _cabi.call("dcb_d2h", _vp(flag.ctypes.data), _vp(dflag.ptr), 16, sh)
            stream.sync()
            n_oob = int(flag[0])
        else:
            rad = ctypes.byref(radial) if radial is not None else None
            per = ctypes.byref(persp) if persp is not None else None
            if on_device:
                dst = DeviceArray((height, width))
                _cabi.call("dcb_spline_remap", _vp(work.ptr), height, width, order,
                           mode_code, _vp(dst.ptr), 0, dst.pitch, map_kind, rad, per,
                           None, None, 0, 0, None, flags, lo, hi, sh)
                dst._keepalive = work            # until the stream has consumed it
                bufs = []
                # the workspace goes back to the pool when dst is collected
                import weakref
                weakref.finalize(dst, _dev.device_pool.give, work)
                return dst, 0
            dout = _dev.device_pool.take(max(height * width * esz, 16))
            bufs.append(dout)
            _cabi.call("dcb_spline_remap", _vp(work.ptr), height, width, order,
                       mode_code, _vp(dout.ptr), dst_f64, width * esz, map_kind, rad,
                       per, None, None, 0, 0, None, flags, lo, hi, sh)
            out = _dev.pinned_empty((height, width),
                                    np.float64 if dst_f64 else np.float32)
            _cabi.call("dcb_d2h", _vp(out.ctypes.data), _vp(dout.ptr),
                       height * width * esz, sh)
            stream.sync()
    finally:
        if bufs:
            stream.sync()
        for b in bufs:
            _dev.device_pool.give(b)
    if out_dtype is not None:
        out = out.astype(out_dtype)
    return out, n_oob
