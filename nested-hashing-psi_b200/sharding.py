"""Bin sharding of the BatchedFHEPIE evaluation over the GPUs of one box.

result[bin] depends only on pt[:, bin, :], mask[bin] and the replicated query + relinearisation key
(the loop at BatchedFHEHIPPIE.cpp:91 carries nothing from one bin to the next), so the b bins are
split into contiguous, balanced blocks, one per rank.  No collective is needed while evaluating;
the only cross-GPU traffic is the final gather of the b result ciphertexts to the rank that
serialises the response (BatchedFHEPSIServer.cpp:143-152).  torch.distributed is the plumbing:
NCCL over NVLink on the GPUs, gloo in the CPU tests.
"""
import numpy as np


def bin_shard(b, rank, world):
    """Contiguous balanced block of bins owned by `rank`: sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank/world out of range")
    return (rank * b) // world, ((rank + 1) * b) // world


class _CudaView:
    """Zero-copy torch view of library-owned device memory through __cuda_array_interface__."""

    def __init__(self, ptr, n_i64):
        self.__cuda_array_interface__ = {"shape": (n_i64,), "typestr": "<i8", "data": (ptr, False), "version": 3,
                                         "strides": None}


class ShardedPIE:
    """Host logic of the bin-sharded evaluation for one rank.

    evaluate(pt_local, mask_local, idx, minus) is injected by the caller: on the GPU box it is the
    C-ABI path (CryptoContext.db_load_limbs / query_set / run), in the CPU tests it is the oracle.
    """

    def __init__(self, b, rank, world):
        self.b, self.rank, self.world = b, rank, world
        self.begin, self.end = bin_shard(b, rank, world)
        self.sizes = [bin_shard(b, r, world)[1] - bin_shard(b, r, world)[0] for r in range(world)]
        self.max_local = max(self.sizes)

    def local_db(self, pt, mask):
        """Slices of the full database this rank keeps resident: pt [K][b][E][L][N], mask [b][L][N]."""
        return (np.ascontiguousarray(pt[:, self.begin:self.end]), np.ascontiguousarray(mask[self.begin:self.end]))

    def gather(self, local, dst=0, group=None):
        """Response gather.  local: torch int64 tensor [b_local, 2, L, N] on this rank's device (or CPU
        under gloo).  Returns the full [b, 2, L, N] tensor on `dst`, None elsewhere."""
        import torch
        import torch.distributed as dist

        tail = tuple(local.shape[1:])
        if local.shape[0] != self.end - self.begin:
            raise ValueError("local result does not match this rank's bin block")
        if self.world == 1:
            return local
        padded = local
        if local.shape[0] < self.max_local:
            padded = torch.zeros((self.max_local,) + tail, dtype=local.dtype, device=local.device)
            padded[:local.shape[0]] = local
        padded = padded.contiguous()
        bufs = None
        if self.rank == dst:
            bufs = [torch.empty_like(padded) for _ in range(self.world)]
        dist.gather(padded, bufs, dst=dst, group=group)
        if self.rank != dst:
            return None
        return torch.cat([bufs[r][:self.sizes[r]] for r in range(self.world)], dim=0)

    @staticmethod
    def device_result_tensor(ctx):
        """torch view (int64, flat) of the library's result buffer on the context's device."""
        import torch

        ptr, nbytes = ctx.result_device_ptr()
        return torch.as_tensor(_CudaView(ptr, nbytes // 8), device="cuda:%d" % ctx.device)
