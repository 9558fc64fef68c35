"""Single-process multi-device evaluation (psi_multi_*), the sharded database calls it is built from, the
scatter-gather limb ingestion (psi_multi_query_set_limbs / psi_multi_result_get_limbs) and the selectable
plaintext lift, all through the C ABI, bit-exact against the oracle.

A device may be listed more than once, so the sharding logic (bin blocks, sliced upload + device-to-device
exchange, per-device download into one buffer) is exercised on a single-GPU box too; with two or more GPUs the
same tests also run on distinct devices."""
import numpy as np
import pytest
import torch

import psi_b200 as P
from oracle.oracle import Oracle
from oracle.params_ref import RefParams

import scenario as sc

pytestmark = pytest.mark.gpu
T32 = 4296540161


def device_lists():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    lists = [[0], [0, 0], [0, 0, 0]]
    if n >= 2:
        lists += [[0, 1], [1, 0, 1]]
    if n >= 4:
        lists += [[0, 1, 2, 3]]
    return lists


@pytest.fixture(scope="module")
def small():
    params = sc.make_params(1024, T32, L=2)
    o = Oracle(params)
    rng = np.random.default_rng(21)
    K, b, E = 2, 5, 6
    s = dict(params=params, o=o, K=K, b=b, E=E)
    s["pt"] = sc.random_pt(rng, params, (K, b, E))
    s["mask"] = sc.random_pt(rng, params, (b,))
    s["idx"] = sc.random_ct(rng, params, (K, E))
    s["minus"] = sc.random_ct(rng, params)
    s["idx2"] = sc.random_ct(rng, params, (K, E))
    s["minus2"] = sc.random_ct(rng, params)
    _, s["evk_b"], s["evk_a"] = o.keygen(3)
    s["want"] = o.run(s["pt"], s["mask"], s["idx"], s["minus"], s["evk_b"], s["evk_a"])
    s["want2"] = o.run(s["pt"], s["mask"], s["idx2"], s["minus2"], s["evk_b"], s["evk_a"])
    return s


@pytest.mark.parametrize("devices", device_lists(), ids=lambda d: "dev" + "".join(map(str, d)))
def test_multi_run_matches_oracle(small, devices):
    s = small
    mc = P.MultiContext(s["params"], devices)
    mc.InsertEvalMultKey(s["evk_b"], s["evk_a"])
    mc.db_load_limbs(s["pt"], s["mask"])
    ranges = mc.bin_ranges()
    assert ranges[0][0] == 0 and ranges[-1][1] == s["b"]
    assert all(r[1] == n[0] for r, n in zip(ranges, ranges[1:])) and all(r[1] > r[0] for r in ranges)
    mc.query_set(s["idx"], s["minus"])
    mc.run()
    got = mc.result_get()
    assert np.array_equal(got, s["want"])
    # 1 inner-product launch + 5 fused ct x ct kernels per bin group (two concurrent groups from 4 resident bins on)
    assert mc.run_launch_count() == sum(1 + 5 * (2 if r1 - r0 >= 4 else 1) for r0, r1 in ranges)


@pytest.mark.parametrize("devices", device_lists()[:3] + device_lists()[3:4], ids=lambda d: "dev" + "".join(map(str, d)))
def test_multi_back_to_back_queries_without_sync(small, devices):
    """Two queries enqueued back to back (no host synchronisation in between): the second upload / exchange lands
    in the other landing buffer while the first query is evaluated, results are double-buffered."""
    s = small
    mc = P.MultiContext(s["params"], devices)
    mc.InsertEvalMultKey(s["evk_b"], s["evk_a"])
    mc.db_load_limbs(s["pt"], s["mask"])
    out1 = np.zeros_like(s["want"])
    out2 = np.zeros_like(s["want"])
    for rep in range(3):
        mc.query_set(s["idx"], s["minus"])
        mc.run()
        mc.result_get(out1, sync=False)
        mc.query_set(s["idx2"], s["minus2"])
        mc.run()
        mc.result_get(out2, sync=False)
        mc.sync()
        assert np.array_equal(out1, s["want"]), rep
        assert np.array_equal(out2, s["want2"]), rep


@pytest.mark.parametrize("devices", [[0], [0, 0, 0]] + device_lists()[3:4], ids=lambda d: "dev" + "".join(map(str, d)))
def test_multi_limb_vector_ingestion(small, devices):
    """The query arrives as K*E*2*L separately allocated limb vectors (what a deserialised OpenFHE query holds) and
    the results leave as b*2*L vectors."""
    s = small
    L, N = s["params"].L, s["params"].N
    mc = P.MultiContext(s["params"], devices)
    mc.InsertEvalMultKey(s["evk_b"], s["evk_a"])
    mc.db_load_limbs(s["pt"], s["mask"])
    for idx, minus, want in ((s["idx"], s["minus"], s["want"]), (s["idx2"], s["minus2"], s["want2"])):
        iv = [idx.reshape(-1, N)[i].copy() for i in range(s["K"] * s["E"] * 2 * L)]
        mv = [minus.reshape(-1, N)[i].copy() for i in range(2 * L)]
        mc.query_set_limbs(iv, mv)
        del iv, mv    # the vectors may be freed as soon as the call returns
        mc.run()
        vecs = mc.result_get_limbs()
        assert len(vecs) == s["b"] * 2 * L
        assert np.array_equal(np.stack(vecs).reshape(want.shape), want)


@pytest.mark.parametrize("devices", [[0], [0, 0]] + device_lists()[3:4], ids=lambda d: "dev" + "".join(map(str, d)))
def test_multi_query_run_limbs(small, devices):
    """psi_multi_query_run_limbs (what the adapter's run() calls): one device takes the streamed single-query path, a
    device list the three-call sequence; two different queries in a row, same object."""
    s = small
    L, N = s["params"].L, s["params"].N
    mc = P.MultiContext(s["params"], devices)
    mc.InsertEvalMultKey(s["evk_b"], s["evk_a"])
    mc.db_load_limbs(s["pt"], s["mask"])
    for idx, minus, want in ((s["idx"], s["minus"], s["want"]), (s["idx2"], s["minus2"], s["want2"])):
        iv = [idx.reshape(-1, N)[i].copy() for i in range(s["K"] * s["E"] * 2 * L)]
        mv = [minus.reshape(-1, N)[i].copy() for i in range(2 * L)]
        vecs = mc.query_run_limbs(iv, mv)
        del iv, mv
        assert np.array_equal(np.stack(vecs).reshape(want.shape), want)
    # the plain calls still work on the same object afterwards
    mc.query_set(s["idx"], s["minus"])
    mc.run()
    assert np.array_equal(mc.result_get(), s["want"])
    mc.close()


def test_single_context_limb_vector_ingestion(small):
    """psi_query_upload_limbs / psi_result_get_limbs on one psi_ctx, two queries in a row (the pinned pool is reused)."""
    s = small
    L, N = s["params"].L, s["params"].N
    cc = P.CryptoContext(s["params"])
    cc.InsertEvalMultKey(s["evk_b"], s["evk_a"])
    cc.db_load_limbs(s["pt"], s["mask"])
    cc.set_host_threads(3)
    for idx, minus, want in ((s["idx"], s["minus"], s["want"]), (s["idx2"], s["minus2"], s["want2"])):
        iv = [idx.reshape(-1, N)[i].copy() for i in range(s["K"] * s["E"] * 2 * L)]
        mv = [minus.reshape(-1, N)[i].copy() for i in range(2 * L)]
        cc.query_upload_limbs(iv, mv)
        del iv, mv
        cc.query_commit()
        cc.run()
        vecs = cc.result_get_limbs()
        assert np.array_equal(np.stack(vecs).reshape(want.shape), want)


def test_streamed_single_query_from_limb_vectors(small):
    """psi_query_run_streamed_limbs: the streamed single query with the host gather / scatter of the limb vectors inside
    the overlap; two queries in a row (pools and landing buffers are reused), then a plain run() on the same context."""
    s = small
    L, N = s["params"].L, s["params"].N
    cc = P.CryptoContext(s["params"])
    cc.InsertEvalMultKey(s["evk_b"], s["evk_a"])
    cc.db_load_limbs(s["pt"], s["mask"])
    cc.set_host_threads(3)
    for idx, minus, want in ((s["idx"], s["minus"], s["want"]), (s["idx2"], s["minus2"], s["want2"]), (s["idx"], s["minus"], s["want"])):
        iv = [idx.reshape(-1, N)[i].copy() for i in range(s["K"] * s["E"] * 2 * L)]
        mv = [minus.reshape(-1, N)[i].copy() for i in range(2 * L)]
        vecs = cc.query_run_streamed_limbs(iv, mv)
        del iv, mv
        assert np.array_equal(np.stack(vecs).reshape(want.shape), want)
    cc.query_set(s["idx2"], s["minus2"])
    cc.run()
    assert np.array_equal(cc.result_get(), s["want2"])
    # K = 1 (mask only), K = 3, one bin / one position, ragged; then a HYBRID context (unfused chain, no operand split)
    rng = np.random.default_rng(78)
    for variant in (dict(), dict(ks_technique=1)):
        params = sc.make_params(1024, T32, L=2) if not variant else RefParams(1024, T32, depth=2, L=2, **variant).to_struct()
        o = Oracle(params)
        cv = P.CryptoContext(params)
        sk, evk_b, evk_a = o.keygen(3)
        cv.InsertEvalMultKey(evk_b, evk_a)
        for K, b, E in ((1, 3, 2), (3, 2, 9), (2, 1, 1), (2, 7, 13)):
            pt, mask = sc.random_pt(rng, params, (K, b, E)), sc.random_pt(rng, params, (b,))
            idx, minus = sc.random_ct(rng, params, (K, E)), sc.random_ct(rng, params)
            cv.db_load_limbs(pt, mask)
            want = o.run(pt, mask, idx, minus, evk_b, evk_a)
            iv = [idx.reshape(-1, N)[i].copy() for i in range(K * E * 2 * L)]
            mv = [minus.reshape(-1, N)[i].copy() for i in range(2 * L)]
            vecs = cv.query_run_streamed_limbs(iv, mv)
            assert np.array_equal(np.stack(vecs).reshape(want.shape), want), (variant, K, b, E)
        cv.close()


def test_streamed_single_query_matches_run(small):
    """psi_query_run_streamed (upload slices -> partial inner products -> bin groups -> downloads, all overlapped inside
    ONE query) returns the limbs psi_query_set + psi_run + psi_result_get return; also K = 1, K = 3, ragged E and b."""
    s = small
    cc = P.CryptoContext(s["params"])
    cc.InsertEvalMultKey(s["evk_b"], s["evk_a"])
    cc.db_load_limbs(s["pt"], s["mask"])
    for rep in range(2):
        assert np.array_equal(cc.query_run_streamed(s["idx"], s["minus"]), s["want"])
        assert np.array_equal(cc.query_run_streamed(s["idx2"], s["minus2"]), s["want2"])
    # interleaved with the two-step path (landing buffers and result buffers alternate in both)
    cc.query_set(s["idx"], s["minus"])
    cc.run()
    assert np.array_equal(cc.result_get(), s["want"])
    assert np.array_equal(cc.query_run_streamed(s["idx2"], s["minus2"]), s["want2"])
    rng = np.random.default_rng(77)
    o, params = s["o"], s["params"]
    for K, b, E in ((1, 3, 2), (3, 2, 9), (2, 1, 1), (2, 7, 13)):
        pt, mask = sc.random_pt(rng, params, (K, b, E)), sc.random_pt(rng, params, (b,))
        idx, minus = sc.random_ct(rng, params, (K, E)), sc.random_ct(rng, params)
        cc.db_load_limbs(pt, mask)
        want = o.run(pt, mask, idx, minus, s["evk_b"], s["evk_a"])
        assert np.array_equal(cc.query_run_streamed(idx, minus), want), (K, b, E)


def test_sharded_db_calls_match_unsharded(small):
    """psi_db_*_shard: the resident shard equals the corresponding bins of the unsharded database."""
    s = small
    K, b, E = s["K"], s["b"], s["E"]
    rng = np.random.default_rng(5)
    n = 700
    slots = rng.integers(-(T32 // 2), T32 // 2, (K, b, E, n), dtype=np.int64)
    mask_slots = rng.integers(1, T32, (b, n), dtype=np.int64)
    full = P.CryptoContext(s["params"])
    full.db_encode_slots(slots, mask_slots)
    pt_full, mask_full = full.db_get_limbs()
    for b0, b1 in ((0, 2), (2, 5), (4, 5)):
        cc = P.CryptoContext(s["params"])
        cc.db_encode_slots_shard(slots, mask_slots, b0, b1)
        pt, mask = cc.db_get_limbs()
        assert np.array_equal(pt, pt_full[:, b0:b1]) and np.array_equal(mask, mask_full[b0:b1])
        cc.db_load_limbs_shard(pt_full, mask_full, b0, b1)
        pt, mask = cc.db_get_limbs()
        assert np.array_equal(pt, pt_full[:, b0:b1]) and np.array_equal(mask, mask_full[b0:b1])
        ptb, maskb = cc.db_get_bin_limbs(b1 - b0 - 1)
        assert np.array_equal(ptb, pt_full[:, b1 - 1]) and np.array_equal(maskb, mask_full[b1 - 1])
    with pytest.raises(P.PsiError):
        P.CryptoContext(s["params"]).db_load_limbs_shard(pt_full, mask_full, 3, 3)
    bad = pt_full.copy()
    bad[1, 2, 3, 1, 17] = s["params"].q[1]          # not a canonical residue
    with pytest.raises(P.PsiError) as ei:
        P.CryptoContext(s["params"]).db_load_limbs(bad, mask_full)
    assert ei.value.status == P.capi.PSI_ERR_INVALID


def test_device_build_shards_and_multi(small):
    """psi_db_build_from_items on a device list == the single-device build with the same seeds."""
    params = s_params = small["params"]
    k, e, K, E, b = 2, 300, 2, 4, 5
    d = P.RandomDataInput(3000, 16, 9, 77, 32)
    h = P.TabulationHashing(4242, k + K)
    one = P.CryptoContext(s_params)
    one.db_build_from_items(h, k, e, K, E, b, d.serverSet, evictionSeed=5, shuffleSeed=11, maskSeed=12)
    pt_full, mask_full = one.db_get_limbs()
    cc = P.CryptoContext(params)
    cc.db_build_from_items_shard(h, k, e, K, E, b, d.serverSet, 1, 4, evictionSeed=5, shuffleSeed=11, maskSeed=12)
    pt, mask = cc.db_get_limbs()
    assert np.array_equal(pt, pt_full[:, 1:4]) and np.array_equal(mask, mask_full[1:4])
    with pytest.raises(ValueError):     # a shard cannot draw its own random shuffle
        cc.db_build_from_items_shard(h, k, e, K, E, b, d.serverSet, 1, 4, shuffleSeed=None)

    o = small["o"]
    rng = np.random.default_rng(8)
    idx, minus = sc.random_ct(rng, params, (K, E)), sc.random_ct(rng, params)
    want = o.run(pt_full, mask_full, idx, minus, small["evk_b"], small["evk_a"])
    for devices in ([0, 0], device_lists()[-1]):
        mc = P.MultiContext(params, devices)
        mc.InsertEvalMultKey(small["evk_b"], small["evk_a"])
        mc.db_build_from_items(h, k, e, K, E, b, d.serverSet, evictionSeed=5, shuffleSeed=11, maskSeed=12)
        mc.query_set(idx, minus)
        mc.run()
        assert np.array_equal(mc.result_get(), want)
    # random seeds: drawn ONCE for all devices (the shards must belong to one database) — decrypt-level check is in
    # test_pie_operator_over_device_list; here: the call succeeds and two builds differ
    mc.db_build_from_items(h, k, e, K, E, b, d.serverSet, evictionSeed=5)
    mc.query_set(idx, minus)
    mc.run()
    assert not np.array_equal(mc.result_get(), want)


@pytest.mark.parametrize("lift", [0, 1])
def test_encode_lift_switch(small, lift):
    """PSI_ENCODE_LIFT_PLAIN / _CENTRED against the oracle's two readings; both decode to the same slots."""
    params, o = small["params"], Oracle(small["params"])
    o.set_encode_lift(lift)
    rng = np.random.default_rng(31)
    K, b, E, n = 2, 2, 3, 900
    slots = rng.integers(-(T32 // 2), T32 // 2, (K, b, E, n), dtype=np.int64)
    mask_slots = rng.integers(1, T32, (b, n), dtype=np.int64)
    for ctx in (P.CryptoContext(params), P.MultiContext(params, [0, 0])):
        ctx.set_encode_lift(lift)
        ctx.db_encode_slots(slots, mask_slots)
        if isinstance(ctx, P.CryptoContext):
            pt, mask = ctx.db_get_limbs()
            assert np.array_equal(pt, sc.encode_db(o, slots))
            assert np.array_equal(mask, sc.encode_masks(o, mask_slots))
    plain = Oracle(params)
    if lift == 1:
        assert not np.array_equal(plain.encode(slots[0, 0, 0]), o.encode(slots[0, 0, 0]))
    with pytest.raises(P.PsiError):
        P.CryptoContext(params).set_encode_lift(7)


def test_pie_operator_over_device_list():
    """The reference-shaped operator on a device list, random (default) shuffle and mask seeds: the decrypted
    intersection is the true one."""
    N, L = 1024, 3
    k, e, K, E, b = 2, 200, 2, 5, 6
    rng = np.random.default_rng(3)
    server = rng.choice(np.arange(1, 1 << 20, dtype=np.uint64), size=2500, replace=False)
    client = np.concatenate([server[:40], rng.integers(1 << 21, 1 << 22, 60, dtype=np.uint64)])
    s = sc.table_scenario(N, T32, L, k, e, K, E, b, server, client)
    devices = device_lists()[-1] if len(device_lists()) > 3 else [0, 0]
    mc = P.MultiContext(s.params, devices)
    mc.InsertEvalMultKey(s.evk_b, s.evk_a)
    pie = P.BatchedFHEHIPPIE(mc, P.PublicKey(), s.hct)
    pie.setIndex(s.idx)
    pie.setMinusCompareElement(s.minus)
    pie.run()
    res = pie.getResultList()
    dec, budget = sc.decrypt_results(s, res)
    inter = np.sort(P.extract_intersection(s.client_cells, dec))
    assert np.array_equal(inter, np.sort(np.intersect1d(server, client)))
    assert budget > 5
