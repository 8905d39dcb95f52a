"""
Device-side plumbing for the Python host: device selection, a small pool of
device buffers, pinned NumPy arrays, streams and events -- all through the C
ABI (no torch, no cuda-python).
"""
import ctypes
import os
import threading
import weakref

import numpy as np

from . import _cabi

_state = threading.local()
_init_lock = threading.Lock()
_initialised = set()


def _default_device():
    for key in ("DCB_DEVICE", "LOCAL_RANK"):
        val = os.environ.get(key)
        if val not in (None, ""):
            try:
                return int(val)
            except ValueError:
                pass
    return 0


def set_device(index):
    """Bind the calling thread (and by default later threads) to a GPU."""
    _cabi.call("dcb_init", int(index))
    if getattr(_state, "device", None) != int(index):
        # the per-thread stream belongs to the device it was created on
        _state.streams = getattr(_state, "streams", {})
        _state.stream = _state.streams.get(int(index))
    _state.device = int(index)
    with _init_lock:
        _initialised.add(int(index))
    os.environ["DCB_DEVICE"] = str(int(index))


def ensure_init():
    """Initialise the library on first use; raise if there is no usable GPU."""
    dev = getattr(_state, "device", None)
    if dev is None:
        set_device(_default_device())
        dev = _state.device
    return dev


def bind_host_to_device(index=None):
    """Pin the calling process to the CPU cores next to GPU ``index`` (NVML's CPU affinity
    of the device) so that the pinned staging buffers allocated afterwards are placed on
    that GPU's NUMA node.  With one process per GPU on a multi-socket host this keeps the
    host<->device copies off the inter-socket link.  Returns the core set, or None when
    NVML (``pynvml``) or the information is unavailable -- never an error."""
    index = ensure_init() if index is None else int(index)
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cores = {64 * w + b for w, word in enumerate(words) for b in range(64)
                 if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cores &= allowed
        if not cores:
            return None
        os.sched_setaffinity(0, cores)
        return cores
    except Exception:
        return None


def device_count():
    n = ctypes.c_int(0)
    try:
        _cabi.call("dcb_device_count", ctypes.byref(n))
    except _cabi.DcbError:
        return 0
    return n.value


def device_info(index=None):
    index = ensure_init() if index is None else index
    sm, maj, mnr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    tot, free = ctypes.c_size_t(), ctypes.c_size_t()
    name = ctypes.create_string_buffer(128)
    _cabi.call("dcb_device_info", index, ctypes.byref(sm), ctypes.byref(maj),
               ctypes.byref(mnr), ctypes.byref(tot), ctypes.byref(free), name,
               128)
    return dict(index=index, name=name.value.decode(), sm_count=sm.value,
                cc=(maj.value, mnr.value), total_mem=tot.value,
                free_mem=free.value)


# --------------------------------------------------------------------------
# device buffers
# --------------------------------------------------------------------------
class DeviceBuffer:
    """Owning handle of ``nbytes`` of device memory."""

    def __init__(self, nbytes):
        self.device = ensure_init()
        ptr = ctypes.c_void_p()
        _cabi.call("dcb_malloc", ctypes.byref(ptr), int(nbytes))
        self.ptr = ptr.value
        self.nbytes = int(nbytes)

    def free(self):
        if self.ptr:
            _cabi.load().dcb_free(ctypes.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _pool_key():
    """Buffers are recycled only on the device they were allocated on."""
    return getattr(_state, "device", None)


class _Pool:
    """Size-bucketed free list per device; avoids a cudaMalloc/cudaFree per call.

    A buffer handed back may still be read or written by work queued on the stream of the thread
    that used it; before it goes to another thread (another stream) the pool drains the device
    (``sync`` hook), same-thread reuse is ordered by the stream itself."""

    def __init__(self, factory, max_bytes, keyfn=lambda: None, sync=None):
        self._factory = factory
        self._keyfn = keyfn
        self._sync = sync
        self._free = {}
        self._held = 0
        self._max = max_bytes
        self._lock = threading.Lock()

    @staticmethod
    def _bucket(nbytes):
        nbytes = max(int(nbytes), 256)
        if nbytes <= (1 << 20):
            return 1 << (nbytes - 1).bit_length()
        step = 1 << 20
        return (nbytes + step - 1) // step * step

    def take(self, nbytes):
        b = self._bucket(nbytes)
        key = (self._keyfn(), b)
        me = threading.get_ident()
        got = None
        with self._lock:
            lst = self._free.get(key)
            if lst:
                self._held -= b
                got = lst.pop()
        if got is None:
            return self._factory(b)
        buf, owner = got
        if owner != me and self._sync is not None:
            self._sync()        # another thread's stream may still be using it
        return buf

    def give(self, buf):
        b = buf.nbytes
        key = (getattr(buf, "device", None), b)
        with self._lock:
            if self._held + b <= self._max:
                self._free.setdefault(key, []).append((buf, threading.get_ident()))
                self._held += b
                return
        buf.free()

    def clear(self):
        with self._lock:
            bufs = [b for lst in self._free.values() for b, _ in lst]
            self._free.clear()
            self._held = 0
        for b in bufs:
            b.free()


_pool_bytes = int(os.environ.get("DCB_POOL_BYTES", str(8 << 30)))
device_pool = _Pool(DeviceBuffer, _pool_bytes, keyfn=_pool_key,
                    sync=lambda: _cabi.call("dcb_device_sync"))


def device_pool_clear():
    """Free the device buffers the pool is holding (large benchmarks between phases)."""
    device_pool.clear()


class borrowed:
    """``with borrowed(nbytes) as buf:`` -- a pooled device buffer."""

    def __init__(self, nbytes):
        self.nbytes = nbytes

    def __enter__(self):
        self.buf = device_pool.take(self.nbytes)
        return self.buf

    def __exit__(self, *exc):
        device_pool.give(self.buf)
        return False


# --------------------------------------------------------------------------
# pinned host arrays
# --------------------------------------------------------------------------
class _PinnedBlock:
    def __init__(self, nbytes):
        ensure_init()
        ptr = ctypes.c_void_p()
        _cabi.call("dcb_host_alloc", ctypes.byref(ptr), int(nbytes))
        self.ptr = ptr.value
        self.nbytes = int(nbytes)

    def free(self):
        if self.ptr:
            _cabi.load().dcb_host_free(ctypes.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


pinned_pool = _Pool(_PinnedBlock, int(os.environ.get("DCB_PINNED_POOL_BYTES",
                                                     str(4 << 30))))


def pinned_empty(shape, dtype=np.float32):
    """A NumPy array in page-locked host memory (full-rate async DMA).  The
    memory goes back to a pool when the array and all its views are collected."""
    dtype = np.dtype(dtype)
    shape = (shape,) if np.isscalar(shape) else tuple(int(s) for s in shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    block = pinned_pool.take(max(nbytes, 1))
    raw = (ctypes.c_char * max(nbytes, 1)).from_address(block.ptr)
    arr = np.frombuffer(raw, dtype=dtype, count=nbytes // dtype.itemsize)
    arr = arr.reshape(shape)
    weakref.finalize(raw, pinned_pool.give, block)
    return arr


# Page-locking a caller's array in place.  A pageable source costs the host pipeline a staging
# copy (64 MiB through a pool of host threads: ~1.7 ms on top of the 1.8 ms the pinned call
# takes); page-locking 64 MiB costs ~10x that once, after which the DMA engine reads the caller's
# pages directly.  So an array is registered the SECOND time the same buffer is seen (loops that
# refill a preallocated frame buffer -- the usual acquisition pattern -- run at the pinned rate,
# one-shot calls never pay for a registration), and unregistered when the array object dies.
_seen_once = {}
_registered = {}
_reg_lock = threading.Lock()
_REG_MIN_BYTES = 8 << 20
_REG_MAX_BYTES = int(os.environ.get("DCB_REGISTER_BYTES", str(2 << 30)))


def _unregister(addr):
    with _reg_lock:
        if _registered.pop(addr, None) is not None:
            try:
                _cabi.load().dcb_host_unregister(ctypes.c_void_p(addr))
            except Exception:
                pass


def maybe_register(array):
    """See above; ``array``: C-contiguous ndarray about to be read by a host-buffer entry point.
    Returns True when its pages are (now) page-locked."""
    if os.environ.get("DCB_REGISTER", "1") == "0":
        return False
    nbytes = array.nbytes
    if nbytes < _REG_MIN_BYTES or not array.flags.c_contiguous or not array.flags.writeable:
        return False
    addr = array.ctypes.data
    with _reg_lock:
        if addr in _registered:
            return _registered[addr] >= nbytes
        if _seen_once.get(addr) != nbytes:
            if len(_seen_once) > 64:
                _seen_once.clear()
            _seen_once[addr] = nbytes
            return False
        if sum(_registered.values()) + nbytes > _REG_MAX_BYTES:
            return False
        del _seen_once[addr]
        try:
            _cabi.call("dcb_host_register", ctypes.c_void_p(addr), nbytes)
        except _cabi.DcbError:
            return False                # e.g. pages that cannot be locked: keep staging
        _registered[addr] = nbytes
    try:
        weakref.finalize(array, _unregister, addr)
    except TypeError:
        _unregister(addr)
        return False
    return True


def pinned_copy(array):
    out = pinned_empty(np.shape(array), np.asarray(array).dtype)
    out[...] = array
    return out


def is_pinned(array):
    flag = ctypes.c_int(0)
    _cabi.call("dcb_is_pinned", ctypes.c_void_p(array.ctypes.data),
               ctypes.byref(flag))
    return bool(flag.value)


# --------------------------------------------------------------------------
# streams / events
# --------------------------------------------------------------------------
class Stream:
    def __init__(self):
        ensure_init()
        h = ctypes.c_void_p()
        _cabi.call("dcb_stream_create", ctypes.byref(h))
        self.handle = h.value

    def sync(self):
        _cabi.call("dcb_stream_sync", ctypes.c_void_p(self.handle))

    def wait(self, event):
        """Work submitted to this stream from now on waits for ``event``."""
        _cabi.call("dcb_stream_wait_event", ctypes.c_void_p(self.handle),
                   ctypes.c_void_p(event.handle))

    def __del__(self):
        try:
            if self.handle:
                _cabi.load().dcb_stream_destroy(ctypes.c_void_p(self.handle))
                self.handle = None
        except Exception:
            pass


class Event:
    def __init__(self):
        ensure_init()
        h = ctypes.c_void_p()
        _cabi.call("dcb_event_create", ctypes.byref(h))
        self.handle = h.value

    def record(self, stream=None):
        _cabi.call("dcb_event_record", ctypes.c_void_p(self.handle),
                   ctypes.c_void_p(stream.handle if stream else None))

    def sync(self):
        _cabi.call("dcb_event_sync", ctypes.c_void_p(self.handle))

    def elapsed_ms(self, later):
        ms = ctypes.c_float()
        _cabi.call("dcb_event_elapsed_ms", ctypes.c_void_p(self.handle),
                   ctypes.c_void_p(later.handle), ctypes.byref(ms))
        return ms.value

    def __del__(self):
        try:
            if self.handle:
                _cabi.load().dcb_event_destroy(ctypes.c_void_p(self.handle))
                self.handle = None
        except Exception:
            pass


def current_stream():
    """One non-blocking stream per host thread."""
    s = getattr(_state, "stream", None)
    if s is None:
        s = _state.stream = Stream()
        streams = getattr(_state, "streams", None)
        if streams is None:
            streams = _state.streams = {}
        streams[getattr(_state, "device", None)] = s
    return s


def synchronize():
    _cabi.call("dcb_device_sync")


def launch_count():
    n = ctypes.c_uint64()
    _cabi.call("dcb_launch_count", ctypes.byref(n))
    return n.value


def last_plan():
    vals = [ctypes.c_int() for _ in range(5)]
    _cabi.call("dcb_last_plan", *[ctypes.byref(v) for v in vals])
    return dict(zip(("path", "box_w", "box_h", "grid", "smem_bytes"),
                    [v.value for v in vals]))


def image_stats(enable=True, reset=False):
    """Counters of the single-image kernel's patch path (``dcb_image_stats``); ``enable``
    switches the counting on or off, the dict holds the counts since the last reset
    (``rows_*`` come from the plans built while counting was on)."""
    out = (ctypes.c_uint64 * 8)()
    _cabi.call("dcb_image_stats", 1 if enable else 0, out, 1 if reset else 0)
    return {"rows_blend_redo": int(out[3]), "tiles_odd": int(out[4]), "rows_patch": int(out[5]),
            "rows_partial": int(out[6]), "rows": int(out[7]), "rows_tile_not_eligible": int(out[0]),
            "rows_binade": int(out[1]), "rows_no_segment_verified": int(out[2])}


def plan_cache_clear():
    """Drop the cached plans of the single-image kernel; returns how many plans this process has
    built so far."""
    n = ctypes.c_uint64()
    _cabi.call("dcb_plan_cache_clear", ctypes.byref(n))
    return n.value


# --------------------------------------------------------------------------
# device-resident float32 arrays
# --------------------------------------------------------------------------
def _pitch_for(width):
    """Row pitch in bytes: dense when that is already a multiple of 16 (the
    TMA requirement), otherwise padded up to one."""
    return (int(width) * 4 + 15) // 16 * 16


class DeviceArray:
    """A float32 image (H, W) or stack (D, H, W) resident in HBM.

    Rows are ``pitch`` bytes apart (multiple of 16), slices ``slice_stride``
    bytes apart.  The memory comes from the buffer pool and returns to it when
    the object is collected.
    """

    def __init__(self, shape, pitch=None):
        shape = tuple(int(s) for s in shape)
        if len(shape) not in (2, 3):
            raise ValueError("DeviceArray is 2-D or 3-D")
        self.shape = shape
        self.dtype = np.dtype(np.float32)
        h, w = shape[-2], shape[-1]
        self.pitch = _pitch_for(w) if pitch is None else int(pitch)
        self.slice_stride = self.pitch * h
        depth = shape[0] if len(shape) == 3 else 1
        self.nbytes = self.slice_stride * depth
        self._buf = device_pool.take(max(self.nbytes, 16))
        self.ptr = self._buf.ptr
        self._fin = weakref.finalize(self, device_pool.give, self._buf)

    @property
    def ndim(self):
        return len(self.shape)

    @classmethod
    def from_host(cls, array, stream=None):
        array = np.ascontiguousarray(array, dtype=np.float32)
        out = cls(array.shape)
        out.copy_from_host(array, stream)
        return out

    def copy_from_host(self, array, stream=None):
        array = np.ascontiguousarray(array, dtype=np.float32)
        if tuple(array.shape) != self.shape:
            raise ValueError("shape mismatch %s vs %s" % (array.shape,
                                                          self.shape))
        stream = stream or current_stream()
        w = self.shape[-1]
        rows = int(np.prod(self.shape[:-1], dtype=np.int64))
        if self.pitch == w * 4:
            _cabi.call("dcb_h2d", ctypes.c_void_p(self.ptr),
                       ctypes.c_void_p(array.ctypes.data), array.nbytes,
                       ctypes.c_void_p(stream.handle))
        else:
            _cabi.call("dcb_h2d_2d", ctypes.c_void_p(self.ptr), self.pitch,
                       ctypes.c_void_p(array.ctypes.data), w * 4, w * 4, rows,
                       ctypes.c_void_p(stream.handle))
        if not is_pinned(array):
            stream.sync()       # pageable source: do not outlive the caller's array
        return self

    def to_host(self, out=None, stream=None):
        stream = stream or current_stream()
        if out is None:
            out = pinned_empty(self.shape, np.float32)
        w = self.shape[-1]
        rows = int(np.prod(self.shape[:-1], dtype=np.int64))
        if self.pitch == w * 4:
            _cabi.call("dcb_d2h", ctypes.c_void_p(out.ctypes.data),
                       ctypes.c_void_p(self.ptr), out.nbytes,
                       ctypes.c_void_p(stream.handle))
        else:
            _cabi.call("dcb_d2h_2d", ctypes.c_void_p(out.ctypes.data), w * 4,
                       ctypes.c_void_p(self.ptr), self.pitch, w * 4, rows,
                       ctypes.c_void_p(stream.handle))
        stream.sync()
        return out

    def fill(self, value=0.0, stream=None):
        """Set every byte to zero (the only fill the path needs: padding of uneven shards)."""
        if value != 0:
            raise ValueError("DeviceArray.fill supports 0 only")
        stream = stream or current_stream()
        _cabi.call("dcb_memset", ctypes.c_void_p(self.ptr), 0, self.nbytes,
                   ctypes.c_void_p(stream.handle))
        return self

    def copy_rows_from(self, src, src_row, nrows, dst_row=0, stream=None):
        """Device-to-device copy of ``nrows`` rows of the 2-D array ``src`` (same width and
        pitch) starting at ``src_row`` into rows ``dst_row...`` of this 2-D array."""
        if self.ndim != 2 or src.ndim != 2 or src.shape[1] != self.shape[1] or src.pitch != self.pitch:
            raise ValueError("copy_rows_from needs two 2-D arrays of equal width and pitch")
        if src_row < 0 or dst_row < 0 or src_row + nrows > src.shape[0] or dst_row + nrows > self.shape[0]:
            raise ValueError("row range outside the arrays")
        stream = stream or current_stream()
        _cabi.call("dcb_d2d", ctypes.c_void_p(self.ptr + dst_row * self.pitch),
                   ctypes.c_void_p(src.ptr + src_row * src.pitch), nrows * self.pitch,
                   ctypes.c_void_p(stream.handle))
        return self

    def fill_synthetic(self, seed, offset=0, stream=None):
        """Fill with the stateless splitmix64 stream (bench inputs)."""
        if self.pitch != self.shape[-1] * 4:
            raise ValueError("fill_synthetic needs a dense array (W % 4 == 0)")
        stream = stream or current_stream()
        _cabi.call("dcb_fill_synthetic_f32", ctypes.c_void_p(self.ptr),
                   self.nbytes // 4, int(seed), int(offset),
                   ctypes.c_void_p(stream.handle))
        return self


def synthetic_host(n, seed, offset=0):
    """NumPy restatement of dcb_fill_synthetic_f32 (spot checks in tests)."""
    idx = (np.arange(n, dtype=np.uint64) + np.uint64(offset)) ^ np.uint64(seed)
    with np.errstate(over="ignore"):
        x = idx + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return ((x >> np.uint64(40)).astype(np.float32)
            * np.float32(1.0 / 16777216.0))
