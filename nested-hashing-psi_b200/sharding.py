"""Bin sharding of the BatchedFHEPIE evaluation over the GPUs of one box.

result[bin] depends only on pt[:, bin, :], mask[bin] and the replicated query + relinearisation key
(the loop at BatchedFHEHIPPIE.cpp:91 carries nothing from one bin to the next), so the b bins are
split into contiguous, balanced blocks, one per rank.  No collective is needed while evaluating;
the only cross-GPU traffic is the final gather of the b result ciphertexts to the rank that
serialises the response (BatchedFHEPSIServer.cpp:143-152).  torch.distributed is the plumbing:
NCCL over NVLink on the GPUs, gloo in the CPU tests.

The query (K*E + 1 ciphertexts, the same for every bin) has to reach every GPU.  QueryDistributor moves
it over PCIe ONCE: rank r copies the r-th 1/G of the index ciphertexts from pinned host memory and the
slices are exchanged with one all-gather over NVLink; G full uploads would each be PCIe-bound at the
single-GPU time.
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
        if world > b:
            # an empty bin block would leave that rank outside the evaluation while the others wait for it in the
            # NCCL gather; the caller shrinks the process group instead (b bins keep at most b GPUs busy)
            raise ValueError("more ranks (%d) than bins (%d): run at most one rank per bin" % (world, b))
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


def query_slice(n_words, rank, world):
    """[begin, end) of the index-ciphertext words rank uploads itself; n_words must divide evenly."""
    if n_words % world:
        raise ValueError("index ciphertext words (%d) do not split evenly over %d ranks" % (n_words, world))
    step = n_words // world
    return rank * step, (rank + 1) * step


class QueryDistributor:
    """Sliced H2D + all-gather of one query into every rank's landing buffers.

    landings: list of (idx, minus) torch int64 tensor pairs (flat) that receive the full query on this rank, used
    in turn: the library's two device landing buffers on the GPU box (for_context), plain CPU tensors under gloo
    in the tests.  on_filled(i) is called after the copies of pair i are enqueued.

    chunks > 1 cuts the index words into that many contiguous chunks; every rank uploads its 1/world of EACH chunk
    and the chunks are all-gathered one by one, so that on a GPU the upload of chunk j+1 (copy stream) runs under the
    all-gather of chunk j (measured at 2 GPUs: query-in stage 1.14 ms with one chunk = upload, then all-gather)."""

    def __init__(self, landings, rank, world, group=None, next_landing=None, on_filled=None, chunks=1):
        import torch

        self.landings, self.rank, self.world, self.group = list(landings), rank, world, group
        self.next_landing, self.on_filled, self.turn = next_landing, on_filled, 0
        idx0 = self.landings[0][0]
        n = idx0.numel()
        if n % world:
            raise ValueError("index ciphertext words (%d) do not split evenly over %d ranks" % (n, world))
        self.chunks = chunks if chunks > 1 and world > 1 and n % (chunks * world) == 0 else 1
        self.chunk_words = n // self.chunks
        self.piece_words = self.chunk_words // world
        self.begin, self.end = query_slice(n, rank, world)    # the rank's words when chunks == 1
        # NCCL writes straight into the landing buffer; the own pieces are staged separately (one staging buffer per
        # landing buffer) so that the collective's input never aliases its output
        self.slices = [torch.empty(self.chunks * self.piece_words, dtype=idx0.dtype, device=idx0.device) for _ in self.landings]
        self.up_stream, self.events = None, None
        if idx0.is_cuda and self.chunks > 1:
            self.up_stream = torch.cuda.Stream(device=idx0.device)
            self.events = [torch.cuda.Event() for _ in range(self.chunks)]

    def own_ranges(self):
        """[begin, end) word ranges of the host query this rank reads itself."""
        return [(j * self.chunk_words + self.rank * self.piece_words, j * self.chunk_words + (self.rank + 1) * self.piece_words)
                for j in range(self.chunks)]

    @classmethod
    def for_context(cls, ctx, rank, world, group=None, chunks=2):
        """Distributor over the two device landing buffers of a CryptoContext."""
        import torch

        dev = "cuda:%d" % ctx.device
        pairs = []
        for w in (0, 1):
            pi, ni, pm, nm = ctx.query_landing_ptrs(w)
            pairs.append((torch.as_tensor(_CudaView(pi, ni // 8), device=dev), torch.as_tensor(_CudaView(pm, nm // 8), device=dev)))
        return cls(pairs, rank, world, group, next_landing=ctx.query_next_landing, on_filled=ctx.query_uploaded, chunks=chunks)

    def distribute(self, host_idx, host_minus):
        """host_idx / host_minus: flat int64 tensors holding the whole query in (pinned) host memory; every rank
        passes the same query.  Enqueued on the current torch stream; returns the landing pair that was filled."""
        import torch
        import torch.distributed as dist

        i = self.next_landing() if self.next_landing else self.turn % len(self.landings)
        self.turn += 1
        idx, minus = self.landings[i]
        staging = self.slices[i]
        ranges = self.own_ranges()
        if self.up_stream is not None:
            # uploads on the copy stream, collectives on the caller's stream, chunk by chunk
            main = torch.cuda.current_stream(idx.device)
            self.up_stream.wait_stream(main)     # earlier use of the staging buffer and of the landing pair
            with torch.cuda.stream(self.up_stream):
                for j, (b0, b1) in enumerate(ranges):
                    staging[j * self.piece_words:(j + 1) * self.piece_words].copy_(host_idx[b0:b1], non_blocking=True)
                    self.events[j].record(self.up_stream)
                minus.copy_(host_minus, non_blocking=True)
            for j in range(self.chunks):
                main.wait_event(self.events[j])
                dist.all_gather_into_tensor(idx[j * self.chunk_words:(j + 1) * self.chunk_words],
                                            staging[j * self.piece_words:(j + 1) * self.piece_words], group=self.group)
            main.wait_stream(self.up_stream)     # the minusCompareElement copy
        else:
            for j, (b0, b1) in enumerate(ranges):
                staging[j * self.piece_words:(j + 1) * self.piece_words].copy_(host_idx[b0:b1], non_blocking=True)
            minus.copy_(host_minus, non_blocking=True)
            for j in range(self.chunks):
                piece = staging[j * self.piece_words:(j + 1) * self.piece_words]
                dst = idx[j * self.chunk_words:(j + 1) * self.chunk_words]
                if self.world == 1:
                    dst.copy_(piece, non_blocking=True)
                else:
                    dist.all_gather_into_tensor(dst, piece, group=self.group)
        if self.on_filled:
            self.on_filled(i)
        return idx, minus
