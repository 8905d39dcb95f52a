"""
Multi-GPU sharding of the stack path: one process per GPU.

The path shards without any data exchange: every slice of a
(D, H, W) stack -- or every image of a batch -- is independent and the ranks
share only the <= 256-byte parameter block (centre + polynomial, or the eight
perspective coefficients).  So the only collective is ONE broadcast of that
block from rank 0; the slices themselves never cross a link.

The one real exchange step of 8(e) is optional: a caller who wants ONE
unwarped sinogram (``unwarp_slice_backward``: row ``index`` of every slice, a
(D, W) array) on a single GPU while the stack is sharded over N.  Two ways:

* :class:`SinogramWindow` + :func:`unwarp_slice_backward_sharded` -- the fused
  form: the owner's (D, W) buffer is mapped into every rank (CUDA IPC over
  NVLink / NVSwitch peer access) and each rank's remap kernel stores its D/N
  rows straight into it; the only synchronisation is one barrier;
* :func:`gather_rows` -- the plain collective (all-gather of the per-rank
  rows), kept as the baseline the fused form is measured against.

The collectives go through a *communicator*: :class:`NcclComm` is the product
one -- NCCL behind the C ABI (``dcb_mg_*``, csrc/mg.cu: ``ncclCommInitRank``, the
unique id handed from rank 0 to the others over a TCP connection on
``MASTER_ADDR``), no PyTorch anywhere in this package.  Any object with the
same five methods works in its place; the CPU tests drive the host logic of
this module (packing, shard arithmetic, padding and compaction of uneven
shards) with a gloo-backed one (tests/gloo_comm.py).
"""
import os
import socket
import struct
import time

import numpy as np

PARAM_SLOTS = 32   # xc, yc, n, 16 coefficients, 8 perspective coefficients, pad


def shard_range(total, rank, world_size):
    """Contiguous block of ``total`` slices owned by ``rank``: sizes differ by
    at most one, the first ``total % world_size`` ranks take the extra slice."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank %r / world size %r" % (rank, world_size))
    base, extra = divmod(int(total), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def pack_params(xcenter=0.0, ycenter=0.0, list_fact=(), list_coef=()):
    """Flatten the model parameters into a fixed float64 vector."""
    if len(list_fact) > 16:
        raise ValueError("at most 16 polynomial coefficients")
    if len(list_coef) not in (0, 8):
        raise ValueError("!!! Eight coefficients are required !!!")
    vec = np.zeros(PARAM_SLOTS, dtype=np.float64)
    vec[0], vec[1], vec[2] = float(xcenter), float(ycenter), len(list_fact)
    vec[3:3 + len(list_fact)] = np.asarray(list_fact, dtype=np.float64)
    vec[19] = len(list_coef)
    vec[20:20 + len(list_coef)] = np.asarray(list_coef, dtype=np.float64)
    return vec


def unpack_params(vec):
    vec = np.asarray(vec, dtype=np.float64)
    n = int(vec[2])
    ncoef = int(vec[19])
    return dict(xcenter=float(vec[0]), ycenter=float(vec[1]),
                list_fact=[float(v) for v in vec[3:3 + n]],
                list_coef=[float(v) for v in vec[20:20 + ncoef]])


# ---------------------------------------------------------------------------
# communicators
# ---------------------------------------------------------------------------
_default_comm = None


def set_default_comm(comm):
    """Make ``comm`` the communicator the functions of this module use when none is passed."""
    global _default_comm
    _default_comm = comm
    return comm


def default_comm():
    return _default_comm


def _comm_of(comm):
    return comm if comm is not None else _default_comm


def _exchange_unique_id(rank, world, make_id, addr, port, timeout=120.0):
    """Rank 0 creates the id and serves it to the world - 1 other ranks over TCP; the others
    fetch it.  The payload is the 128-byte NCCL unique id, nothing else crosses this socket."""
    if rank == 0:
        uid = make_id()
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind((addr, port))
        srv.listen(world)
        srv.settimeout(timeout)
        try:
            for _ in range(world - 1):
                conn, _peer = srv.accept()
                with conn:
                    conn.sendall(struct.pack("<I", len(uid)) + uid)
        finally:
            srv.close()
        return uid
    deadline = time.monotonic() + timeout
    while True:
        try:
            with socket.create_connection((addr, port), timeout=5.0) as conn:
                head = b""
                while len(head) < 4:
                    head += conn.recv(4 - len(head)) or _raise_closed()
                (n,) = struct.unpack("<I", head)
                uid = b""
                while len(uid) < n:
                    uid += conn.recv(n - len(uid)) or _raise_closed()
                return uid
        except (ConnectionRefusedError, ConnectionResetError, socket.timeout, OSError):
            if time.monotonic() > deadline:
                raise
            time.sleep(0.05)


def _raise_closed():
    raise ConnectionResetError("rendezvous connection closed early")


class NcclComm:
    """NCCL communicator of this process through the C ABI (``dcb_mg_*``).

    ``NcclComm.from_env()`` reads what ``torchrun`` / any launcher exports (``RANK``,
    ``WORLD_SIZE``, ``LOCAL_RANK``, ``MASTER_ADDR``, ``MASTER_PORT``), binds the local device and
    joins; the unique id travels over TCP port ``MASTER_PORT + 1`` (``DCB_MG_PORT`` overrides)."""

    def __init__(self, rank, world, unique_id):
        import ctypes
        from . import _cabi
        self.rank, self.world = int(rank), int(world)
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        _cabi.call("dcb_mg_init", buf, self.world, self.rank)
        self._open = True

    @staticmethod
    def unique_id():
        import ctypes
        from . import _cabi
        buf = ctypes.create_string_buffer(128)
        ver = ctypes.c_int()
        _cabi.call("dcb_mg_unique_id", buf, ctypes.byref(ver))
        return buf.raw

    @classmethod
    def from_env(cls, set_device=True):
        from . import device as _dev
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        local = int(os.environ.get("LOCAL_RANK", str(rank)))
        if set_device:
            _dev.set_device(local)
        addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
        port = int(os.environ.get("DCB_MG_PORT", str(int(os.environ.get("MASTER_PORT", "29500")) + 1)))
        uid = _exchange_unique_id(rank, world, cls.unique_id, addr, port) if world > 1 \
            else cls.unique_id()
        return set_default_comm(cls(rank, world, uid))

    # -- the five methods a communicator has --------------------------------------------
    def bcast_bytes(self, payload, nbytes, src=0):
        import ctypes
        from . import _cabi
        buf = ctypes.create_string_buffer(int(nbytes))
        if self.rank == src:
            raw = bytes(payload)
            if len(raw) != nbytes:
                raise ValueError("payload is %d bytes, expected %d" % (len(raw), nbytes))
            buf.raw = raw
        _cabi.call("dcb_mg_bcast_host", buf, int(nbytes), int(src))
        return buf.raw

    def allgather_rows(self, padded):
        """``padded``: this rank's (most, width) float32 block, NumPy or DeviceArray; returns
        the (world * most, width) concatenation in the same kind."""
        import ctypes
        from . import _cabi, device as _dev
        is_np = isinstance(padded, np.ndarray)
        send = _dev.DeviceArray.from_host(np.ascontiguousarray(padded, dtype=np.float32)) \
            if is_np else padded
        most, width = send.shape
        if send.pitch != width * 4:
            raise ValueError("all-gather needs densely packed rows")
        recv = _dev.DeviceArray((self.world * most, width))
        if recv.pitch != width * 4:
            raise ValueError("all-gather needs a width whose rows pack densely (multiple of 4)")
        st = _dev.current_stream()
        _cabi.call("dcb_mg_allgather", ctypes.c_void_p(send.ptr), ctypes.c_void_p(recv.ptr),
                   most * width * 4, ctypes.c_void_p(st.handle))
        if is_np:
            return recv.to_host()
        recv._keep = send
        return recv

    def allreduce_max(self, values):
        import ctypes
        from . import _cabi
        vals = [float(v) for v in values]
        arr = (ctypes.c_double * len(vals))(*vals)
        _cabi.call("dcb_mg_allreduce_max_f64", arr, len(vals))
        return [float(v) for v in arr]

    def barrier(self):
        from . import _cabi
        _cabi.call("dcb_mg_barrier")

    def close(self):
        from . import _cabi
        global _default_comm
        if self._open:
            _cabi.call("dcb_mg_finalize")
            self._open = False
        if _default_comm is self:
            _default_comm = None


def broadcast_params(params=None, src=0, comm=None):
    """Broadcast the packed parameter vector from rank ``src`` to every rank of ``comm`` (the
    default communicator when None) and return it unpacked.  Ranks other than ``src`` may pass
    ``None``.  Without a communicator (single process) the parameters are returned unchanged."""
    comm = _comm_of(comm)
    if comm is None or comm.world == 1:
        return unpack_params(pack_params(**params))
    payload = pack_params(**params).tobytes() if comm.rank == src else b""
    got = comm.bcast_bytes(payload, PARAM_SLOTS * 8, src=src)
    return unpack_params(np.frombuffer(got, dtype=np.float64))


def unwarp_stack_sharded(stack_shard, params, rows=None):
    """Unwarp this rank's slices (a NumPy array or DeviceArray holding only the
    shard) with parameters that came from :func:`broadcast_params`.  ``rows`` =
    (start, stop) inclusive selects output rows, default all."""
    from .post import postprocessing as post
    height = stack_shard.shape[1]
    start, stop = (0, height - 1) if rows is None else rows
    return post.unwarp_chunk_slices_backward(
        stack_shard, params["xcenter"], params["ycenter"],
        params["list_fact"], start, stop)


# ---------------------------------------------------------------------------
# the optional exchange step: one sinogram assembled from all ranks (8e)
# ---------------------------------------------------------------------------
def broadcast_bytes(payload, nbytes, src=0, comm=None):
    """Broadcast ``nbytes`` opaque bytes (``payload`` on rank ``src``, ignored elsewhere);
    returns them as ``bytes``."""
    comm = _comm_of(comm)
    if comm is None or comm.world == 1:
        return bytes(payload)
    return comm.bcast_bytes(payload, nbytes, src=src)


def gather_rows(local_rows, depth, comm=None):
    """All-gather the per-rank row blocks of a (depth, width) float32 array whose rows are
    sharded by :func:`shard_range` -- the plain collective form of the sinogram exchange.
    ``local_rows``: this rank's (hi - lo, width) block, a NumPy array or a DeviceArray; the full
    array comes back in the same kind on every rank.  Shards may differ by one row: blocks are
    padded to the largest for the collective and compacted afterwards."""
    from . import device as _dev
    comm = _comm_of(comm)
    rank, world = (0, 1) if comm is None else (comm.rank, comm.world)
    lo, hi = shard_range(depth, rank, world)
    if tuple(local_rows.shape)[0] != hi - lo:
        raise ValueError("rank %d owns rows [%d, %d) but was given %d rows"
                         % (rank, lo, hi, local_rows.shape[0]))
    if world == 1:
        return local_rows
    is_np = isinstance(local_rows, np.ndarray)
    width = int(local_rows.shape[1])
    most = -(-int(depth) // world)                      # rows of the largest shard
    if is_np:
        padded = np.zeros((most, width), dtype=np.float32)
        padded[:hi - lo] = local_rows
    else:
        padded = _dev.DeviceArray((most, width))
        padded.fill(0.0)
        if hi > lo:
            padded.copy_rows_from(local_rows, 0, hi - lo)
    blocks = comm.allgather_rows(padded)
    if is_np:
        full = np.empty((int(depth), width), dtype=np.float32)
        for r in range(world):
            rlo, rhi = shard_range(depth, r, world)
            full[rlo:rhi] = blocks[r * most:r * most + (rhi - rlo)]
        return full
    full = _dev.DeviceArray((int(depth), width))
    for r in range(world):
        rlo, rhi = shard_range(depth, r, world)
        if rhi > rlo:
            full.copy_rows_from(blocks, r * most, rhi - rlo, dst_row=rlo)
    return full


class _Rows:
    """Rows [lo, hi) of a window as a kernel destination."""

    def __init__(self, ptr, pitch, shape):
        self.ptr, self.pitch, self.shape = ptr, pitch, shape


class SinogramWindow:
    """A (depth, width) float32 sinogram in the HBM of rank ``owner`` that every
    rank's kernels can write: the owner exports its buffer (``dcb_ipc_export``),
    the 64-byte handle is broadcast, the other ranks map it (``dcb_ipc_open``,
    peer access over NVLink / NVSwitch).  Collective: every rank constructs it,
    calls :func:`unwarp_slice_backward_sharded` any number of times with
    :meth:`fence` after each, and :meth:`close` at the end.

    ``array`` is the owner's DeviceArray (None elsewhere); ``rows`` this rank's
    (lo, hi) of :func:`shard_range`.
    """

    HEADER = 64 + 8     # IPC handle + row pitch

    def __init__(self, depth, width, owner=0, comm=None):
        import ctypes
        from . import _cabi, device as _dev
        self.comm = _comm_of(comm)
        rank, world = (0, 1) if self.comm is None else (self.comm.rank, self.comm.world)
        self.depth, self.width, self.owner = int(depth), int(width), int(owner)
        self.rank, self.world = rank, world
        self.rows = shard_range(depth, rank, world)
        self.array = None
        self._peer = None
        payload = b""
        if rank == owner:
            self.array = _dev.DeviceArray((self.depth, self.width))
            self.pitch = self.array.pitch
            self._base = self.array.ptr
            if world > 1:
                handle = ctypes.create_string_buffer(64)
                _cabi.call("dcb_ipc_export", ctypes.c_void_p(self._base), handle)
                payload = handle.raw + np.int64(self.pitch).tobytes()
        if world > 1:
            got = self.comm.bcast_bytes(payload, self.HEADER, src=owner)
            if rank != owner:
                self.pitch = int(np.frombuffer(got[64:72], dtype=np.int64)[0])
                peer = ctypes.c_void_p()
                _cabi.call("dcb_ipc_open", ctypes.create_string_buffer(got[:64], 64),
                           ctypes.byref(peer))
                self._peer = peer.value
                self._base = peer.value

    def my_rows(self):
        """This rank's rows of the window, as a destination for the kernel."""
        lo, hi = self.rows
        return _Rows(self._base + lo * self.pitch, self.pitch, (hi - lo, self.width))

    def fence(self):
        """Every rank's stores have landed in the owner's buffer when this
        returns: the local stream is drained (stores to peer memory are
        performed before the kernel completes), then one barrier."""
        from . import device as _dev
        _dev.current_stream().sync()
        if self.world > 1:
            self.comm.barrier()

    def close(self):
        import ctypes
        from . import _cabi
        if self._peer is not None:
            _cabi.call("dcb_ipc_close", ctypes.c_void_p(self._peer))
            self._peer = None
        if self.world > 1:
            self.comm.barrier()      # the owner keeps the buffer until every mapping is gone


def unwarp_slice_backward_sharded(stack_shard, params, index, window):
    """``unwarp_slice_backward`` (reference ``postprocessing.py:188-229``) over
    a stack sharded by :func:`shard_range`: this rank's slices give rows
    ``window.rows`` of the (D, W) sinogram, written by the remap kernel itself
    into the owner's buffer.  Asynchronous; ``window.fence()`` completes it."""
    from .post import postprocessing as post
    lo, hi = window.rows
    if stack_shard.shape[0] != hi - lo or stack_shard.shape[2] != window.width:
        raise ValueError("shard of shape %s does not match rows [%d, %d) of a %d-wide window"
                         % (tuple(stack_shard.shape), lo, hi, window.width))
    if hi == lo:
        return
    post._unwarp_slice_into(stack_shard, params["xcenter"], params["ycenter"],
                            params["list_fact"], index, window.my_rows())
