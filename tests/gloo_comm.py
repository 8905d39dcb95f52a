"""A communicator with the five methods of discorpy_b200.multigpu.NcclComm on top of
torch.distributed's gloo backend, so that the N > 1 host logic of the package (parameter packing,
shard arithmetic, padding / compaction of uneven shards, the IPC-handle broadcast) runs on CPU
ranks.  Test infrastructure: the package itself never imports torch."""
import numpy as np
import torch
import torch.distributed as dist


class GlooComm:
    def __init__(self):
        if not dist.is_initialized():
            dist.init_process_group("gloo")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def bcast_bytes(self, payload, nbytes, src=0):
        buf = np.zeros(nbytes, dtype=np.uint8)
        if self.rank == src:
            raw = np.frombuffer(bytes(payload), dtype=np.uint8)
            if raw.size != nbytes:
                raise ValueError("payload is %d bytes, expected %d" % (raw.size, nbytes))
            buf[:] = raw
        t = torch.from_numpy(buf)
        dist.broadcast(t, src=src)
        return t.numpy().tobytes()

    def allgather_rows(self, padded):
        t = torch.from_numpy(np.ascontiguousarray(padded, dtype=np.float32))
        out = torch.empty((self.world * t.shape[0], t.shape[1]), dtype=t.dtype)
        dist.all_gather_into_tensor(out, t)
        return out.numpy()

    def allreduce_max(self, values):
        t = torch.tensor([float(v) for v in values], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def barrier(self):
        dist.barrier()

    def close(self):
        dist.destroy_process_group()
