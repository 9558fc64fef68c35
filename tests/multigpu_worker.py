"""One rank of the 2-GPU parity run (launched by tests/test_gpu_multigpu.py through torch.distributed.run).

Every rank keeps its own block of bins resident, receives the query through QueryDistributor (1/world of the
index ciphertexts from its own pinned host copy, the rest by NCCL all-gather into the library's landing buffers),
evaluates through the C ABI and rank 0 gathers the result ciphertexts (NCCL) and checks them bit for bit
against the oracle's single-process evaluation."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import psi_b200 as P  # noqa: E402
import scenario as sc  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    params = P.params_generate(1024, 4296540161, 3)
    o = Oracle(params)
    rng = np.random.default_rng(17)  # same seed on every rank: same database, same query
    K, b, E = 2, 5, 6
    pt = sc.random_pt(rng, params, (K, b, E))
    mask = sc.random_pt(rng, params, (b,))
    idx = sc.random_ct(rng, params, (K, E))
    minus = sc.random_ct(rng, params)
    _, evk_b, evk_a = o.keygen(3)

    shard = P.ShardedPIE(b, rank, world)
    pt_l, mask_l = shard.local_db(pt, mask)
    cc = P.CryptoContext(params, device=local)
    cc.InsertEvalMultKey(evk_b, evk_a)
    cc.db_load_limbs(pt_l, mask_l)

    # query: pinned host copy whose foreign slices are poisoned, so only the all-gather can complete it
    host_idx = torch.from_numpy(idx.view(np.int64).reshape(-1).copy()).pin_memory()
    host_minus = torch.from_numpy(minus.view(np.int64).reshape(-1).copy()).pin_memory()
    qd = P.QueryDistributor.for_context(cc, rank, world)
    keep = torch.zeros_like(host_idx, dtype=torch.bool)
    for b0, b1 in qd.own_ranges():
        keep[b0:b1] = True
    host_idx[~keep] = -1
    qd.distribute(host_idx, host_minus)
    torch.cuda.current_stream().synchronize()
    cc.query_commit()
    cc.run()
    cc.sync()
    local_res = P.ShardedPIE.device_result_tensor(cc).view(shard.end - shard.begin, 2, params.L, params.N)
    full = shard.gather(local_res, dst=0)
    if rank == 0:
        want = o.run(pt, mask, idx, minus, evk_b, evk_a)
        got = full.cpu().numpy().view(np.uint64)
        assert got.shape == want.shape and np.array_equal(got, want), "sharded result differs from the oracle"
        print("MULTIGPU_PARITY_OK world=%d" % world, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
