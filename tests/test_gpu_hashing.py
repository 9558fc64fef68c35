"""Device-side nested cuckoo table build and device-resident constructor path (SURVEY 8f "next" #3) against
the host implementation (host/hashing.cpp, itself pinned by tests/test_hashing.py): bit-identical tables,
including the random-walk evictions, and bit-identical encoded databases."""
import numpy as np
import pytest

import psi_b200 as P

pytestmark = pytest.mark.gpu
T32 = 4296540161


def host_cells(h, k, e, K, E, b, items, eviction_seed):
    hct = P.HierarchicalCuckooHashTable(h, e, E, 0, k, K, True, True, b, evictionSeed=eviction_seed)
    hct.insertAll(items)
    return hct, hct.cells()


@pytest.mark.parametrize("k,e,K,E,b,n,seed", [
    (2, 37, 2, 6, 8, 1900, 1),        # comfortable load
    (2, 64, 2, 8, 3, 1000, 2),        # 3-slot bins: positions overflow, evictions happen
    (3, 16, 3, 7, 2, 380, 3),         # three simple and three cuckoo functions, 2-slot bins
    (2, 4949, 2, 14, 14, 1 << 20, 4), # BASELINE configs[1] table (Parameters1.txt:11)
    (1, 5, 2, 40, 40, 9000, 5),       # bins wider than one warp
])
def test_device_table_equals_host_table(k, e, K, E, b, n, seed):
    rng = np.random.default_rng(seed)
    items = rng.integers(1, 1 << 32, n, dtype=np.uint64)
    items[: n // 50] = items[n // 2: n // 2 + n // 50]          # some duplicates: lookUp must drop them
    h = P.TabulationHashing(1000 + seed, k + K)
    cc = P.CryptoContext(P.params_generate(1024, T32, 2))
    _, want = host_cells(h, k, e, K, E, b, items, 77 + seed)
    got = cc.hct_build_device(h, k, e, K, E, b, items, evictionSeed=77 + seed)
    assert np.array_equal(got, want)
    assert (got != 0).sum() > 0


def test_device_table_insertion_failure():
    cc = P.CryptoContext(P.params_generate(1024, T32, 2))
    h = P.TabulationHashing(5, 4)
    with pytest.raises(RuntimeError, match="Cuckoo hashing error"):
        cc.hct_build_device(h, 2, 1, 2, 2, 1, np.arange(1, 50, dtype=np.uint64))


def test_device_constructor_path_equals_host_constructor():
    """psi_db_build_from_items == BatchedFHEHIPPIE(ctx, pk, host table): same plaintext DB limbs, same masks."""
    d = P.RandomDataInput(3000, 40, 21, 99, 32)
    k, e, K, E, b = 2, 50, 2, 7, 9
    h = P.TabulationHashing(4242, k + K)
    params = P.params_generate(2048, T32, 2)
    cc_host, cc_dev = P.CryptoContext(params), P.CryptoContext(params)
    hct = P.HierarchicalCuckooHashTable(h, e, E, 0, k, K, True, True, b, evictionSeed=5)
    hct.insertAll(d.serverSet)
    P.BatchedFHEHIPPIE(cc_host, P.PublicKey(), hct, shuffleSeed=11, maskSeed=12)
    cc_dev.db_build_from_items(h, k, e, K, E, b, d.serverSet, evictionSeed=5, shuffleSeed=11, maskSeed=12)
    pt_h, mask_h = cc_host.db_get_limbs()
    pt_d, mask_d = cc_dev.db_get_limbs()
    assert np.array_equal(pt_d, pt_h) and np.array_equal(mask_d, mask_h)
    with pytest.raises(ValueError):
        cc_dev.db_build_from_items(h, k, e, K, E, b, np.array([T32], dtype=np.uint64))   # item >= t


def test_operator_from_server_set_end_to_end():
    """BatchedFHEHIPPIE.fromServerSet (offline phase on the device) -> query -> run -> decrypted intersection."""
    import scenario as sc
    d = P.RandomDataInput(3000, 48, 25, 777, 32)
    s = sc.table_scenario(2048, T32, 3, 2, 40, 2, 7, 9, d.serverSet, d.clientSet)
    cc = P.CryptoContext(s.params)
    cc.InsertEvalMultKey(s.evk_b, s.evk_a)
    pie = P.BatchedFHEHIPPIE.fromServerSet(cc, P.PublicKey(), s.hash, 40, 7, 2, 2, 9, d.serverSet)
    pie.setIndex(s.idx)
    pie.setMinusCompareElement(s.minus)
    pie.run()
    dec, budget = sc.decrypt_results(s, pie.getResultList())
    assert budget > 5
    assert np.array_equal(np.sort(P.extract_intersection(s.client_cells, dec)), np.sort(np.intersect1d(d.clientSet, d.serverSet)))
