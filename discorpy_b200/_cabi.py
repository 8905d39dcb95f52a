"""
ctypes binding of ``libdiscorpy_b200.so`` (C ABI declared in
``include/discorpy_b200.h``).

This is the only place the Python host touches native code.  There is no CPU
fallback: if the shared library has not been built, or no sm_100 GPU is
present when a compute function is called, the call raises.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# DCB_LIB: load another build of the same library (A/B comparisons of kernel variants)
LIB_PATH = os.environ.get("DCB_LIB") or os.path.join(_HERE, "lib", "libdiscorpy_b200.so")

DCB_MAX_TERMS = 16
DCB_OK = 0
DCB_ERR_ARG = -1
DCB_ERR_CUDA = -2
DCB_ERR_UNSUPPORTED = -3
DCB_ERR_NO_DEVICE = -4

BLEND_EXACT, BLEND_LERP64, BLEND_LERP32 = 0, 1, 2
PATH_AUTO, PATH_DIRECT, PATH_TMA = 0, 1, 2
FLAG_ROUND_INT = 0x100      # DCB_FLAG_ROUND_INT
DTYPE_F32, DTYPE_U8, DTYPE_I8, DTYPE_U16, DTYPE_I16 = 0, 1, 2, 3, 4
#: enum dcb_mode, in the order the reference documents the modes (postprocessing.py:129-130)
MODES = {"reflect": 0, "grid-mirror": 1, "constant": 2, "grid-constant": 3,
         "nearest": 4, "mirror": 5, "grid-wrap": 6, "wrap": 7}
MAP_RADIAL, MAP_PERSP, MAP_COORDS = 0, 1, 2


class DcbError(RuntimeError):
    """A libdiscorpy_b200 call failed (status code in ``.status``)."""

    def __init__(self, status, message):
        super().__init__("libdiscorpy_b200: %s (status %d)" % (message, status))
        self.status = status


class Radial(ctypes.Structure):
    _fields_ = [("xc", ctypes.c_double), ("yc", ctypes.c_double),
                ("n", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("a", ctypes.c_double * DCB_MAX_TERMS)]


class Persp(ctypes.Structure):
    _fields_ = [("c", ctypes.c_double * 8)]


class Options(ctypes.Structure):
    _fields_ = [("order", ctypes.c_int32), ("blend", ctypes.c_int32),
                ("path", ctypes.c_int32), ("flags", ctypes.c_int32)]


_vp = ctypes.c_void_p
_sz = ctypes.c_size_t
_i = ctypes.c_int
_u64 = ctypes.c_uint64

# name -> argtypes; every function returns int except dcb_last_error
SIGNATURES = {
    "dcb_version": [],
    "dcb_device_count": [ctypes.POINTER(_i)],
    "dcb_init": [_i],
    "dcb_device_info": [_i, ctypes.POINTER(_i), ctypes.POINTER(_i),
                        ctypes.POINTER(_i), ctypes.POINTER(_sz),
                        ctypes.POINTER(_sz), ctypes.c_char_p, _i],
    "dcb_malloc": [ctypes.POINTER(_vp), _sz],
    "dcb_free": [_vp],
    "dcb_memset": [_vp, _i, _sz, _vp],
    "dcb_host_alloc": [ctypes.POINTER(_vp), _sz],
    "dcb_host_free": [_vp],
    "dcb_host_register": [_vp, _sz],
    "dcb_host_unregister": [_vp],
    "dcb_is_pinned": [_vp, ctypes.POINTER(_i)],
    "dcb_h2d": [_vp, _vp, _sz, _vp],
    "dcb_d2h": [_vp, _vp, _sz, _vp],
    "dcb_d2d": [_vp, _vp, _sz, _vp],
    "dcb_h2d_2d": [_vp, _sz, _vp, _sz, _sz, _sz, _vp],
    "dcb_d2h_2d": [_vp, _sz, _vp, _sz, _sz, _sz, _vp],
    "dcb_stream_create": [ctypes.POINTER(_vp)],
    "dcb_stream_destroy": [_vp],
    "dcb_stream_sync": [_vp],
    "dcb_device_sync": [],
    "dcb_event_create": [ctypes.POINTER(_vp)],
    "dcb_event_destroy": [_vp],
    "dcb_event_record": [_vp, _vp],
    "dcb_event_sync": [_vp],
    "dcb_stream_wait_event": [_vp, _vp],
    "dcb_event_elapsed_ms": [_vp, _vp, ctypes.POINTER(ctypes.c_float)],
    "dcb_ipc_export": [_vp, _vp],
    "dcb_ipc_open": [_vp, ctypes.POINTER(_vp)],
    "dcb_ipc_close": [_vp],
    "dcb_unwarp_image_backward_f32": [_vp, _vp, _i, _i, _sz, _sz,
                                      ctypes.POINTER(Radial),
                                      ctypes.POINTER(Options), _vp],
    "dcb_unwarp_image_backward_host_f32": [_vp, _vp, _i, _i, _sz, _sz,
                                           ctypes.POINTER(Radial),
                                           ctypes.POINTER(Options), _i],
    "dcb_host_copy_2d": [_vp, _sz, _vp, _sz, _sz, _i],
    "dcb_host_band_edges": [_i, _i, _i, ctypes.POINTER(ctypes.c_int),
                            ctypes.POINTER(ctypes.c_int)],
    "dcb_correct_perspective_image_host_f32": [_vp, _vp, _i, _i, _sz, _sz,
                                               ctypes.POINTER(Persp),
                                               ctypes.POINTER(Options), _i],
    "dcb_unwarp_image_backward_perspective_host_f32": [
        _vp, _vp, _i, _i, _sz, _sz, ctypes.POINTER(Radial),
        ctypes.POINTER(Persp), ctypes.POINTER(Options), _i],
    "dcb_unwarp_stack_backward_f32": [_vp, _vp, _i, _i, _i, _i, _i, _sz, _sz,
                                      _sz, _sz, _i, _i, _i,
                                      ctypes.POINTER(Radial),
                                      ctypes.POINTER(Options), _vp],
    "dcb_correct_perspective_image_f32": [_vp, _vp, _i, _i, _sz, _sz,
                                          ctypes.POINTER(Persp),
                                          ctypes.POINTER(Options), _vp],
    "dcb_map_coordinates_f32": [_vp, _vp, _i, _i, _sz, _vp, _vp, _i, _sz, _vp,
                                ctypes.POINTER(Options), _vp],
    "dcb_unwarp_image_backward_perspective_f32": [
        _vp, _vp, _vp, _i, _i, _sz, _sz, _sz, ctypes.POINTER(Radial),
        ctypes.POINTER(Persp), ctypes.POINTER(Options), _vp],
    "dcb_unpack_hwc_to_planes_f32": [_vp, _i, _vp, _i, _i, _i, _sz, _sz, _vp],
    "dcb_pack_planes_f32_to_hwc": [_vp, _vp, _i, _i, _i, _i, _sz, _sz, _vp],
    "dcb_fill_synthetic_f32": [_vp, _sz, _u64, _u64, _vp],
    "dcb_launch_count": [ctypes.POINTER(_u64)],
    "dcb_launch_count_reset": [],
    "dcb_last_plan": [ctypes.POINTER(_i)] * 5,
    "dcb_image_stats": [_i, ctypes.POINTER(_u64), _i],
    "dcb_image_timeline": [ctypes.POINTER(_u64), _i],
    "dcb_plan_cache_clear": [ctypes.POINTER(_u64)],
    "dcb_mg_unique_id": [_vp, ctypes.POINTER(_i)],
    "dcb_mg_init": [_vp, _i, _i],
    "dcb_mg_info": [ctypes.POINTER(_i), ctypes.POINTER(_i)],
    "dcb_mg_bcast": [_vp, _sz, _i, _vp],
    "dcb_mg_bcast_host": [_vp, _sz, _i],
    "dcb_mg_allgather": [_vp, _vp, _sz, _vp],
    "dcb_mg_allreduce_max_f64": [ctypes.POINTER(ctypes.c_double), _i],
    "dcb_mg_barrier": [],
    "dcb_mg_finalize": [],
    "dcb_selftest_sqrt": [_sz, _u64, ctypes.POINTER(_u64)],
    "dcb_selftest_sqrt_fast": [_sz, _u64, ctypes.POINTER(_u64), ctypes.POINTER(_u64)],
    "dcb_selftest_tma": [_vp, _i, _i, _i, _sz, _sz, _i, _i, _i, _i, _i, _vp,
                         ctypes.POINTER(_i)],
    "dcb_microbench": [_i, ctypes.POINTER(ctypes.c_double)],
    "dcb_unwarp_image_forward_f32": [_vp, _vp, _i, _i, _sz, _sz, ctypes.POINTER(Radial), _vp, _vp],
    "dcb_spline_workspace_bytes": [_i, _i, _i, _i, ctypes.POINTER(_sz)],
    "dcb_spline_prefilter": [_vp, _i, _i, _i, _sz, _i, _i, _vp, _sz, _vp],
    "dcb_spline_remap": [_vp, _i, _i, _i, _i, _vp, _i, _sz, _i,
                         ctypes.POINTER(Radial), ctypes.POINTER(Persp), _vp, _vp,
                         _i, _sz, _vp, _i, ctypes.c_double, ctypes.c_double, _vp],
}

_lib = None
_lock = threading.Lock()


def load():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise DcbError(
                DCB_ERR_UNSUPPORTED,
                "%s is missing -- build it with `python -c 'import "
                "__graft_entry__ as g; g.build()'` or `make -C discorpy_b200/"
                "csrc`; there is no CPU fallback" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = ctypes.c_int
        lib.dcb_last_error.argtypes = []
        lib.dcb_last_error.restype = ctypes.c_char_p
        _lib = lib
    return _lib


def check(status):
    if status != DCB_OK:
        msg = load().dcb_last_error()
        raise DcbError(status, msg.decode("utf-8", "replace") if msg else "?")


def call(name, *args):
    """Call ``name`` and raise :class:`DcbError` on a non-zero status."""
    check(getattr(load(), name)(*args))


def make_radial(xcenter, ycenter, list_fact):
    n = len(list_fact)
    if n < 1:
        raise ValueError("list_fact must hold at least one coefficient")
    if n > DCB_MAX_TERMS:
        raise NotImplementedError(
            "polynomials with more than %d coefficients are not supported by "
            "the CUDA path (got %d)" % (DCB_MAX_TERMS, n))
    m = Radial()
    m.xc = float(xcenter)
    m.yc = float(ycenter)
    m.n = n
    for i, a in enumerate(list_fact):
        m.a[i] = float(a)
    return m


def make_persp(list_coef):
    m = Persp()
    for i, c in enumerate(list_coef):
        m.c[i] = float(c)
    return m


def make_options(order=1, blend=BLEND_EXACT, path=PATH_AUTO, flags=None):
    """``flags``: ``FLAG_ROUND_INT`` for integer images.  A/B builds of the
    library (-DDCB_AB) also read an experimental kernel variant from the low
    byte, settable through ``DCB_FLAGS``."""
    env = int(os.environ.get("DCB_FLAGS", "0") or 0) & 0xff
    return Options(int(order), int(blend), int(path), int(flags or 0) | env)
