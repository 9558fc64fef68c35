"""Host-side nested cuckoo hashing (nested-hashing-psi_b200/host/hashing.cpp) against an independent
pure-Python restatement of the reference's hashing layer:
  TabulationHashing.cpp:16-54, HashUtils.cpp:34-37, CuckooHashTable.cpp:72-114,135-167,
  HierarchicalCuckooHashTable.cpp:55-73, RandomDataInput.cpp:10-66."""
import numpy as np
import pytest

import psi_b200 as P


class MT19937:
    """std::mt19937 (32-bit Mersenne twister), pure Python."""

    def __init__(self, seed):
        self.mt = [0] * 624
        self.mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            self.mt[i] = (1812433253 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.idx = 624

    def __call__(self):
        if self.idx >= 624:
            mt = self.mt
            for i in range(624):
                y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7FFFFFFF)
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.idx = 0
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF


def libstdcxx_uniform_u64(gen):
    """std::uniform_int_distribution<uint64_t>()(std::mt19937&) in libstdc++: two engine words, HIGH first."""
    hi = gen()
    return (hi << 32) | gen()


def tabulation_tables(seed, nhf):
    gen = MT19937(seed)
    return [[[libstdcxx_uniform_u64(gen) for _ in range(256)] for _ in range(16)] for _ in range(nhf)]


def tab_hash(tables, hf, x):
    res = 0
    for i in range(16):
        res ^= tables[hf][i][(x >> (8 * i)) & 0xFF] if i < 8 else tables[hf][i][0]
    return res


def test_tabulation_hash_matches_restatement():
    seed, nhf = 12223222, 4       # the seed of tests/TestBatchedFHEPIE.cpp:96
    tables = tabulation_tables(seed, nhf)
    rng = np.random.default_rng(0)
    items = rng.integers(1, 1 << 48, 200, dtype=np.uint64)
    h = P.TabulationHashing(seed, nhf)
    for hf in range(nhf):
        for size in (10, 47, 4949):
            got = P.hash_index(h, items, hf, size)
            want = [tab_hash(tables, hf, int(x)) % size for x in items]
            assert [int(v) for v in got] == want


def test_hierarchical_table_invariants():
    """Every item sits in every simple table (HierarchicalCuckooHashTable.cpp:59-72), at one of its K cuckoo
    positions of the inner table its simple hash selects; nothing else is stored."""
    k, e, K, E, b = 2, 37, 2, 6, 8
    rng = np.random.default_rng(3)
    items = np.unique(rng.integers(1, 1 << 32, 2500, dtype=np.uint64))
    h = P.TabulationHashing(987654321, k + K)
    hct = P.HierarchicalCuckooHashTable(h, e, E, 0, k, K, True, True, b)
    hct.insertAll(items)
    cells = hct.cells()                                   # [k][e][K][b][E]
    assert cells.shape == (k, e, K, b, E)
    for st in range(k):
        stored = cells[st][cells[st] != 0]
        assert np.array_equal(np.sort(stored), items)     # each item exactly once per simple table
        outer = P.hash_index(h, items, st, e).astype(np.int64)
        pos = [P.hash_index(h, items, k + hf, E).astype(np.int64) for hf in range(K)]
        for idx, x in enumerate(items):
            found = [(hf, bin_) for hf in range(K) for bin_ in range(b) if cells[st, outer[idx], hf, bin_, pos[hf][idx]] == x]
            assert len(found) == 1
    # bins fill from row 0 upwards before the PIE constructor shuffles them (CuckooHashTable.cpp:88-95)
    occ = cells != 0
    assert not (occ[..., 1:, :] & ~occ[..., :-1, :]).any()


def test_insertion_failure_and_duplicates():
    h = P.TabulationHashing(5, 4)
    hct = P.HierarchicalCuckooHashTable(h, 1, 2, 0, 2, 2, True, True, 1)   # capacity 4 per simple table
    with pytest.raises(RuntimeError, match="Cuckoo hashing error"):
        hct.insertAll(np.arange(1, 50, dtype=np.uint64))
    hct = P.HierarchicalCuckooHashTable(h, 4, 8, 0, 2, 2, True, True, 4)
    hct.insertAll(np.array([7, 7, 7, 9], dtype=np.uint64))              # lookUp first: duplicates are dropped
    c = hct.cells()
    assert (c == 7).sum() == 2 and (c == 9).sum() == 2


def test_random_data_input_structure():
    """RandomDataInput.cpp:31-66: intersection = first I draws of the server stream = last I client items;
    the server set does not depend on the client set size."""
    d = P.RandomDataInput(5000, 300, 151, 123456789, 32)
    assert np.array_equal(d.intersectionSet, d.serverSet[:151])
    assert np.array_equal(d.clientSet[-151:], d.serverSet[:151])
    assert d.serverSet.max() < (1 << 32)
    d2 = P.RandomDataInput(5000, 100, 51, 123456789, 32)
    assert np.array_equal(d2.serverSet, d.serverSet)
    assert np.array_equal(d2.clientSet[:49], d.clientSet[:49])
    with pytest.raises(ValueError):
        P.RandomDataInput(10, 20, 5)
    both = np.intersect1d(d.clientSet, d.serverSet)
    assert set(d.intersectionSet) <= set(both)


def test_client_query_slots():
    """BatchedFHEPSIClient.cpp:114-152: one-hot rows, -x, +1 for empty client slots."""
    k, e, K, E = 2, 16, 2, 5
    h = P.TabulationHashing(42, k + K)
    items = np.array([11, 222, 3333, 44444], dtype=np.uint64)
    cells = P.client_table(h, k, e, K, items)
    assert np.array_equal(np.sort(cells[cells != 0]), items)
    idx, minus = P.build_query_slots(h, cells, K, E)
    flat = cells.reshape(-1)
    for s, x in enumerate(flat):
        if x == 0:
            assert minus[s] == 1 and not idx[:, :, s].any()
        else:
            assert minus[s] == -int(x)
            for hf in range(K):
                pos = int(P.hash_index(h, [x], k + hf, E)[0])
                assert idx[hf, pos, s] == 1 and idx[hf, :, s].sum() == 1
    dec = np.ones((3, k * e), dtype=np.int64)
    s0 = int(np.nonzero(flat == 222)[0][0])
    dec[1, s0] = 0
    assert list(P.extract_intersection(cells, dec)) == [222]
