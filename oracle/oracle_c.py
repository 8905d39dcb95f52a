"""
ctypes loader of ``oracle/liboracle_c.so`` (the C restatement in
``oracle_c.c``).  TEST INFRASTRUCTURE -- never imported by ``discorpy_b200``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_c.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        lib = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.orc_unwarp_stack_backward_f32.argtypes = (
            [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 8
            + [ctypes.c_double, ctypes.c_double, dp, ctypes.c_int,
               ctypes.c_int, ctypes.c_int])
        lib.orc_unwarp_stack_backward_f32.restype = None
        lib.orc_correct_perspective_f32.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, dp,
            ctypes.c_int, ctypes.c_int]
        lib.orc_correct_perspective_f32.restype = None
        _lib = lib
    return _lib


def num_threads():
    """Worker threads used by default: the cores this process may run on."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _coefs(vals):
    arr = (ctypes.c_double * len(vals))(*[float(v) for v in vals])
    return arr


def unwarp_stack_backward(stack, xc, yc, fact, row0=0, nrows=None,
                          coord_round=True, order=1, ylo=0, yhi=None,
                          nthreads=None):
    stack = np.ascontiguousarray(stack, dtype=np.float32)
    if stack.ndim == 2:
        stack = stack[None]
    d, h, w = stack.shape
    nrows = h - row0 if nrows is None else nrows
    yhi = h - 1 if yhi is None else yhi
    out = np.empty((d, nrows, w), dtype=np.float32)
    load().orc_unwarp_stack_backward_f32(
        stack.ctypes.data, out.ctypes.data, d, h, w, row0, nrows, ylo, yhi,
        int(bool(coord_round)), float(xc), float(yc), _coefs(fact), len(fact),
        order, num_threads() if nthreads is None else int(nthreads))
    return out


def unwarp_image_backward(mat, xc, yc, fact, order=1, nthreads=None):
    return unwarp_stack_backward(mat, xc, yc, fact, order=order,
                                 nthreads=nthreads)[0]


def correct_perspective_image(mat, coef, order=1, nthreads=None):
    mat = np.ascontiguousarray(mat, dtype=np.float32)
    h, w = mat.shape
    out = np.empty_like(mat)
    load().orc_correct_perspective_f32(mat.ctypes.data, out.ctypes.data, h, w,
                                       _coefs(coef), order,
                                       num_threads() if nthreads is None
                                       else int(nthreads))
    return out
