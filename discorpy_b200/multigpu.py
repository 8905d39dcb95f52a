"""
Multi-GPU sharding of the stack path: one process per GPU.

The path shards without any data exchange: every slice of a
(D, H, W) stack -- or every image of a batch -- is independent and the ranks
share only the <= 160-byte parameter block (centre + polynomial, or the eight
perspective coefficients).  So the only collective is ONE broadcast of that
block from rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests); the
slices themselves never cross a link.

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
