"""
Multi-GPU sharding of the stack path: one process per GPU.

The path shards without any data exchange: every slice of a
(D, H, W) stack -- or every image of a batch -- is independent and the ranks
share only the <= 160-byte parameter block (centre + polynomial, or the eight
perspective coefficients).  So the only collective is ONE broadcast of that
block from rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests); the
slices themselves never cross a link.

The one real exchange step of 8(e) is optional: a caller who wants ONE
unwarped sinogram (``unwarp_slice_backward``: row ``index`` of every slice, a
(D, W) array) on a single GPU while the stack is sharded over N.  Two ways:

* :class:`SinogramWindow` + :func:`unwarp_slice_backward_sharded` -- the fused
  form: the owner's (D, W) buffer is mapped into every rank (CUDA IPC over
  NVLink / NVSwitch peer access) and each rank's remap kernel stores its D/N
  rows straight into it; the only synchronisation is one barrier;
* :func:`gather_rows` -- the plain collective (all-gather of the per-rank
  rows: NCCL on GPUs, gloo on CPU), kept as the baseline the fused form is
  measured against and as what the CPU tests can run.

``torch.distributed`` is used for that plumbing only and is imported lazily;
single-GPU use of the package never imports torch.
"""
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


def broadcast_params(params=None, src=0, device=None):
    """Broadcast the packed parameter vector from rank ``src`` to every rank
    of the default ``torch.distributed`` process group and return it unpacked.
    Ranks other than ``src`` may pass ``None``.  With no initialised process
    group (single process) the parameters are returned unchanged."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return unpack_params(pack_params(**params))
    is_src = dist.get_rank() == src
    vec = pack_params(**params) if is_src else np.zeros(PARAM_SLOTS)
    if device is None:
        device = ("cuda:%d" % torch.cuda.current_device()
                  if dist.get_backend() == "nccl" else "cpu")
    t = torch.from_numpy(vec).to(device)
    dist.broadcast(t, src=src)
    return unpack_params(t.cpu().numpy())


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
def _group():
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None, 0, 1
    return dist, dist.get_rank(), dist.get_world_size()


def _comm_device(dist):
    import torch
    return ("cuda:%d" % torch.cuda.current_device()
            if dist.get_backend() == "nccl" else "cpu")


def broadcast_bytes(payload, nbytes, src=0):
    """Broadcast ``nbytes`` opaque bytes (``payload`` on rank ``src``, ignored
    elsewhere) over the default process group; returns them as ``bytes``."""
    dist, rank, _ = _group()
    if dist is None:
        return bytes(payload)
    import torch
    buf = np.zeros(nbytes, dtype=np.uint8)
    if rank == src:
        raw = np.frombuffer(bytes(payload), dtype=np.uint8)
        if raw.size != nbytes:
            raise ValueError("payload is %d bytes, expected %d" % (raw.size, nbytes))
        buf[:] = raw
    t = torch.from_numpy(buf).to(_comm_device(dist))
    dist.broadcast(t, src=src)
    return t.cpu().numpy().tobytes()


def gather_rows(local_rows, depth):
    """All-gather the per-rank row blocks of a (depth, width) array whose rows
    are sharded by :func:`shard_range` -- the plain collective form of the
    sinogram exchange.  ``local_rows``: this rank's (hi - lo, width) block, a
    NumPy array (gloo) or a torch CUDA tensor (NCCL); the full array comes back
    in the same kind on every rank.  Shards may differ by one row: blocks are
    padded to the largest for the collective and compacted afterwards."""
    dist, rank, world = _group()
    lo, hi = shard_range(depth, rank, world)
    if tuple(local_rows.shape)[0] != hi - lo:
        raise ValueError("rank %d owns rows [%d, %d) but was given %d rows"
                         % (rank, lo, hi, local_rows.shape[0]))
    if dist is None:
        return local_rows
    import torch
    is_np = isinstance(local_rows, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local_rows)) if is_np else local_rows
    t = t.to(_comm_device(dist))
    width = t.shape[1]
    most = -(-int(depth) // world)                      # rows of the largest shard
    padded = torch.zeros((most, width), dtype=t.dtype, device=t.device)
    padded[:hi - lo] = t
    blocks = torch.empty((world * most, width), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(blocks, padded)
    full = torch.empty((int(depth), width), dtype=t.dtype, device=t.device)
    for r in range(world):
        rlo, rhi = shard_range(depth, r, world)
        full[rlo:rhi] = blocks[r * most:r * most + (rhi - rlo)]
    return full.cpu().numpy() if is_np else full


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

    def __init__(self, depth, width, owner=0):
        import ctypes
        from . import _cabi, device as _dev
        dist, rank, world = _group()
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
            got = broadcast_bytes(payload, self.HEADER, src=owner)
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
        dist, _, world = _group()
        if world > 1:
            dist.barrier()

    def close(self):
        import ctypes
        from . import _cabi
        if self._peer is not None:
            _cabi.call("dcb_ipc_close", ctypes.c_void_p(self._peer))
            self._peer = None
        dist, _, world = _group()
        if world > 1:
            dist.barrier()      # the owner keeps the buffer until every mapping is gone


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
