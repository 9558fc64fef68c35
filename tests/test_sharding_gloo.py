"""Multi-GPU host logic on the CPU: bin sharding + response gather over torch.distributed (gloo,
world_size 2 and 3).  Each rank evaluates its own bins (with the oracle standing in for the device path),
rank 0 gathers and must obtain exactly the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import psi_b200 as P


def test_bin_shard_partition():
    for b in (1, 2, 7, 14, 47, 75, 148):
        for world in (1, 2, 3, 4, 8):
            blocks = [P.bin_shard(b, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == b
            for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
                assert a1 == b0
            sizes = [e - s for s, e in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        P.bin_shard(4, 4, 4)


def test_query_slice_partition():
    for world in (1, 2, 4, 8):
        n = 2 * 47 * 2 * 4 * 16384
        cuts = [P.query_slice(n, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
    with pytest.raises(ValueError):
        P.query_slice(10, 0, 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmpdir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import scenario as sc
    from oracle.oracle import Oracle
    from oracle.params_ref import RefParams
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params = RefParams(256, 4296540161, L=2).to_struct()
    o = Oracle(params)
    rng = np.random.default_rng(5)          # same seed on every rank: the query is replicated
    K, b, E = 2, 5, 3
    sk, evk_b, evk_a = o.keygen(2)
    pt = sc.random_pt(rng, params, (K, b, E))
    mask = sc.random_pt(rng, params, (b,))
    idx = sc.random_ct(rng, params, (K, E))
    minus = sc.random_ct(rng, params)
    # query distribution: every rank moves only its own 1/world of the index ciphertexts "over PCIe", the rest
    # arrives by all-gather; the other slices of this rank's host copy are poisoned to prove they are never read
    idx_flat = torch.from_numpy(idx.view(np.int64).reshape(-1))
    minus_flat = torch.from_numpy(minus.view(np.int64).reshape(-1))
    landing_idx, landing_minus = torch.zeros_like(idx_flat), torch.zeros_like(minus_flat)
    for chunks in (1, 2):     # one slice per rank / the chunked form whose uploads overlap the collectives on a GPU
        landing_idx.zero_()
        landing_minus.zero_()
        qd = P.QueryDistributor([(landing_idx, landing_minus)], rank, world, chunks=chunks)
        assert qd.chunks == (chunks if idx_flat.numel() % (chunks * world) == 0 else 1)
        host = torch.full_like(idx_flat, -1)
        for b0, b1 in qd.own_ranges():
            host[b0:b1] = idx_flat[b0:b1]
        if qd.chunks == 1:
            assert qd.own_ranges() == [(qd.begin, qd.end)]
        qd.distribute(host, minus_flat)
        assert torch.equal(landing_idx, idx_flat) and torch.equal(landing_minus, minus_flat)
    idx = landing_idx.numpy().view(np.uint64).reshape(idx.shape)
    minus = landing_minus.numpy().view(np.uint64).reshape(minus.shape)

    shard = P.ShardedPIE(b, rank, world)
    pt_l, mask_l = shard.local_db(pt, mask)            # this rank keeps only its bins resident
    local = o.run(pt_l, mask_l, idx, minus, evk_b, evk_a)
    assert local.shape[0] == shard.end - shard.begin
    full = shard.gather(torch.from_numpy(local.view(np.int64)), dst=0)
    if rank == 0:
        want = o.run(pt, mask, idx, minus, evk_b, evk_a)
        got = full.numpy().view(np.uint64)
        assert got.shape == want.shape and np.array_equal(got, want)
        open(os.path.join(tmpdir, "ok_%d" % world), "w").write("ok")
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gather_matches_single_process(world, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / ("ok_%d" % world)).exists()
