"""Full-size checks on the B200 (BASELINE.json configs 2^20 and 2^24 server items vs 2^10 / 2^12 client items):
the real pipeline RandomDataInput -> nested cuckoo table -> BatchedFHEHIPPIE ctor (GPU encode) -> encrypted
query -> run() -> results, verified by
  * bit-exact limbs against the oracle on randomly sampled bins (the oracle needs ~0.1 s per bin),
  * the size-independent acceptance property of the protocol: the decrypted intersection equals the planted
    one (PSIClient::intersectionMatches, PSIClient.hpp:142-164), with noise budget to spare."""
import numpy as np
import pytest

import psi_b200 as P
from oracle.oracle import Oracle

import scenario as sc

pytestmark = pytest.mark.gpu
T32 = 4296540161

CONFIGS = {
    # Parameters1.txt:11 and :17
    "2^20": dict(S=1 << 20, C=1 << 10, I=513, k=2, e=4949, K=2, b=14, E=14),
    "2^24": dict(S=1 << 24, C=1 << 10, I=513, k=2, e=4949, K=2, b=47, E=47),
    # Parameters1.txt:67 (BASELINE configs[4]: 2^24 server items vs 2^12 client items, three simple hash functions)
    "2^24_vs_2^12": dict(S=1 << 24, C=1 << 12, I=2049, k=3, e=1791, K=2, b=75, E=75),
    # BASELINE configs[3]: 2^22 server items (b = E from tools/param_sweep.py, profiles/r01_param_sweep_2p22.md)
    "2^22": dict(S=1 << 22, C=1 << 10, I=513, k=2, e=4949, K=2, b=26, E=26),
    # BASELINE configs[0]: TestBatchedFHEPIE's context (depth 2, ring dimension 8192, TestBatchedFHEPIE.cpp:14-26) at
    # 2^16 server items vs 2^8 client items (bench.py workload 2^16_vs_2^8)
    "2^16_vs_2^8": dict(S=1 << 16, C=1 << 8, I=129, k=2, e=1900, K=2, b=8, E=8, N=8192, depth=2, L=3),
}


@pytest.mark.parametrize("name", ["2^20", "2^24", "2^24_vs_2^12", "2^22", "2^16_vs_2^8"])
def test_full_protocol(name):
    w = CONFIGS[name]
    k, e, K, b, E = w["k"], w["e"], w["K"], w["b"], w["E"]
    # client rules (BatchedFHEPSIClient.cpp:50-73): N=16384, depth 3; configs[0] runs in the unit test's context
    params = P.params_generate(w.get("N", 16384), T32, w.get("depth") or P.depth_for_E(E))
    assert (params.N, params.L) == (w.get("N", 16384), w.get("L", 4))
    o = Oracle(params)
    data = P.RandomDataInput(w["S"], w["C"], w["I"], 123456789, 32)
    hashf = P.TabulationHashing(987654321, k + K)
    hct = P.HierarchicalCuckooHashTable(hashf, e, E, 0, k, K, True, True, b)
    hct.insertAll(data.serverSet)
    cc = P.CryptoContext(params)
    sk, evk_b, evk_a = o.keygen(99)
    cc.InsertEvalMultKey(evk_b, evk_a)
    pie = P.BatchedFHEHIPPIE(cc, P.PublicKey(), hct, keepSlots=True)
    assert (pie.K, pie.b, pie.E, pie.batchSize) == (K, b, E, k * e)

    client_cells = P.client_table(hashf, k, e, K, data.clientSet)
    idx_slots, minus_slots = P.build_query_slots(hashf, client_cells, K, E)
    idx = np.empty((K, E, 2, params.L, params.N), dtype=np.uint64)
    for hf in range(K):
        for pos in range(E):
            idx[hf, pos] = o.encrypt(sk, idx_slots[hf, pos], 5000 + hf * E + pos)
    minus = o.encrypt(sk, minus_slots, 4999)
    pie.setIndex(idx)
    pie.setMinusCompareElement(minus)
    pie.run()
    got = pie.getResultList()
    assert got.shape == (b, 2, params.L, params.N)

    # (1) sampled bins, bit-exact
    slots, mask_slots = pie.slots()
    rng = np.random.default_rng(7)
    for bin_ in sorted(rng.choice(b, size=2, replace=False)):
        pt_bin = sc.encode_db(o, slots[:, bin_:bin_ + 1])
        mask_bin = sc.encode_masks(o, mask_slots[bin_:bin_ + 1])
        want = o.run(pt_bin, mask_bin, idx, minus, evk_b, evk_a, nthreads=8)
        assert np.array_equal(got[bin_], want[0]), "bin %d differs from the oracle" % bin_

    # (2) decrypted intersection == planted intersection
    dec = np.empty((b, params.N), dtype=np.int64)
    budget = 1e9
    for bin_ in range(b):
        dec[bin_], amb, nb = o.decrypt(sk, got[bin_])
        assert amb == 0
        budget = min(budget, nb)
    inter = np.sort(P.extract_intersection(client_cells, dec))
    # the planted intersection, plus the client-only 32-bit items that hit the 2^24-item server set by
    # chance (probability 2^-8 each, so about two of the 511 at 2^24): the TRUE intersection
    truth = np.intersect1d(data.clientSet, data.serverSet)
    assert set(data.intersectionSet) <= set(truth)
    assert np.array_equal(inter, truth)
    assert budget > 10, budget
    # masked non-matches look random: zero appears only at matches
    zeros = int((dec[:, :k * e] == 0).sum())
    assert zeros >= w["I"]


def test_sampled_bins_at_2p28_shape():
    """Parameters1.txt:23 (2^28 server items vs 2^10: b = E = 176, 61 952 plaintexts = 32 GB resident): the shapes no
    smaller case reaches -- E > 128 (the inner product folds its lazy accumulators mid-way), 176 bins per launch --
    with a synthetic table of that shape (hashing 2^28 items adds nothing the 2^24 cases do not cover).  Sampled
    bins must be bit-exact against the oracle; the checksum of all result ciphertexts must not depend on how the
    bins were grouped in phase 2."""
    K, b, E, nslots = 2, 176, 176, 2 * 4949
    params = P.params_generate(16384, T32, 3)
    o = Oracle(params)
    cc = P.CryptoContext(params)
    rng = np.random.default_rng(28)
    sk, evk_b, evk_a = o.keygen(28)
    cc.InsertEvalMultKey(evk_b, evk_a)
    slots = rng.integers(1, T32, (K, b, E, nslots), dtype=np.int64)
    mask_slots = rng.integers(1, T32, (b, nslots), dtype=np.int64)
    cc.db_encode_slots(slots, mask_slots)
    idx = sc.random_ct(rng, params, (K, E))
    minus = sc.random_ct(rng, params)
    cc.query_set(idx, minus)
    cc.run()
    got = cc.result_get()
    for bin_ in (0, 97, 175):
        pt_bin = sc.encode_db(o, slots[:, bin_:bin_ + 1])
        mask_bin = sc.encode_masks(o, mask_slots[bin_:bin_ + 1])
        want = o.run(pt_bin, mask_bin, idx, minus, evk_b, evk_a, nthreads=8)
        assert np.array_equal(got[bin_], want[0]), "bin %d differs from the oracle" % bin_
    cc.set_tuning(phase2_groups=1)
    cc.run()
    assert np.array_equal(cc.result_get(), got)
    cc.set_tuning(phase2_groups=0)
