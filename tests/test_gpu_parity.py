"""GPU parity tests proper (run on the B200 box: pytest -m gpu).  Every test calls the product through
the C ABI (ctypes -> libpsi_b200.so) and compares with the ORACLE on the same seeded inputs.
Bar: bit-exact (integer residues)."""
import os

import numpy as np
import pytest

import psi_b200 as P
from oracle.oracle import Oracle
from oracle.params_ref import RefParams

import scenario as sc

pytestmark = pytest.mark.gpu
T32 = 4296540161
T16 = 65537


def ctx_and_oracle(N, L, t=T32, Lp=None):
    params = RefParams(N, t, L=L, Lp=Lp).to_struct()
    return P.CryptoContext(params), Oracle(params), params


@pytest.mark.parametrize("N,L", [(16384, 4), (8192, 3), (4096, 2), (1024, 2), (256, 2), (64, 1)])
def test_ntt_every_modulus(N, L):
    """Forward / inverse negacyclic NTT for every modulus of the context (q_i, p_j, t)."""
    t = T32 if (T32 - 1) % (2 * N) == 0 else T16
    cc, o, params = ctx_and_oracle(N, L, t)
    rng = np.random.default_rng(N)
    nm = params.L + params.Lp + 1
    mods = np.repeat(np.arange(nm, dtype=np.uint32), 3)
    polys = np.empty((len(mods), N), dtype=np.uint64)
    for i, m in enumerate(mods):
        q = int(params.q[m]) if m < params.L else (int(params.p[m - params.L]) if m < nm - 1 else int(params.t))
        polys[i] = rng.integers(0, q, N, dtype=np.uint64)
        if i % 3 == 1:
            polys[i, ::2] = q - 1      # extreme residues
        if i % 3 == 2:
            polys[i] = 0
            polys[i, -1] = 1
    fwd = cc.debug_ntt(polys, mods)
    for i, m in enumerate(mods):
        assert np.array_equal(fwd[i], o.ntt(polys[i], int(m))), "forward, modulus %d" % m
    inv = cc.debug_ntt(polys, mods, inverse=True)
    for i, m in enumerate(mods):
        assert np.array_equal(inv[i], o.ntt(polys[i], int(m), inverse=True)), "inverse, modulus %d" % m
    assert np.array_equal(cc.debug_ntt(fwd, mods, inverse=True), polys)


@pytest.mark.parametrize("N,L", [(16384, 4), (8192, 3), (1024, 2), (256, 1), (16384, 6), (2048, 5), (1024, 7), (4096, 8)])
def test_mul_ctct_relin(N, L):
    """EvalMult(ct,ct) + relinearisation, random (worst-case: uniformly distributed) operands."""
    cc, o, params = ctx_and_oracle(N, L)
    rng = np.random.default_rng(7 * N + L)
    sk, evk_b, evk_a = o.keygen(11)
    cc.InsertEvalMultKey(evk_b, evk_a)
    for trial in range(2):
        ct1, ct2 = sc.random_ct(rng, params), sc.random_ct(rng, params)
        if trial == 1:  # structured edge values
            ct1[0, :, :8] = 0
            ct2[1, :, :8] = 0
            for l in range(L):
                ct1[1, l, 8:16] = int(params.q[l]) - 1
                ct2[0, l, 8:16] = int(params.q[l]) - 1
        got = cc.debug_mul_ctct(ct1, ct2)
        want = o.mul_ctct(ct1, ct2, evk_b, evk_a)
        assert np.array_equal(got, want)
    # operand order matters (the two operands are extended differently)
    assert not np.array_equal(cc.debug_mul_ctct(ct2, ct1), want) or np.array_equal(ct1, ct2)


@pytest.mark.parametrize("N,L,Lp", [(2048, 1, 2), (1024, 2, 3), (16384, 3, 4), (2048, 4, 5), (1024, 5, 6), (2048, 6, 7)])
def test_mul_ctct_auxiliary_basis_one_limb_larger(N, L, Lp):
    """sizeP = sizeQ + 1 (the shape of the plain HPS auxiliary basis): every (L, Lp) instantiation of the
    fused column kernels, real ciphertexts, limbs and decrypted product."""
    cc, o, params = ctx_and_oracle(N, L, Lp=Lp)
    assert (params.L, params.Lp) == (L, Lp)
    rng = np.random.default_rng(N + 10 * L)
    sk, evk_b, evk_a = o.keygen(3)
    cc.InsertEvalMultKey(evk_b, evk_a)
    t = int(params.t)
    m1 = rng.integers(-(t // 2), t // 2, params.N, dtype=np.int64)
    m2 = rng.integers(-(t // 2), t // 2, params.N, dtype=np.int64)
    ct1, ct2 = o.encrypt(sk, m1, 1), o.encrypt(sk, m2, 2)
    got = cc.debug_mul_ctct(ct1, ct2)
    assert np.array_equal(got, o.mul_ctct(ct1, ct2, evk_b, evk_a))
    if L >= 2:   # one 60-bit limb leaves no room for a 33-bit plaintext product
        dec, amb, _ = o.decrypt(sk, got)
        want = np.array([(int(x) * int(y)) % t for x, y in zip(m1, m2)], dtype=np.int64)
        want = np.where(want > t // 2, want - t, want)
        assert amb == 0 and np.array_equal(dec, want)
    r1, r2 = sc.random_ct(rng, params), sc.random_ct(rng, params)
    assert np.array_equal(cc.debug_mul_ctct(r1, r2), o.mul_ctct(r1, r2, evk_b, evk_a))


def test_mul_ctct_real_ciphertexts_decrypt():
    cc, o, params = ctx_and_oracle(2048, 3)
    rng = np.random.default_rng(5)
    sk, evk_b, evk_a = o.keygen(2)
    cc.InsertEvalMultKey(evk_b, evk_a)
    t = int(params.t)
    m1 = rng.integers(-(t // 2), t // 2, params.N, dtype=np.int64)
    m2 = rng.integers(-(t // 2), t // 2, params.N, dtype=np.int64)
    ct1, ct2 = o.encrypt(sk, m1, 1), o.encrypt(sk, m2, 2)
    got = cc.debug_mul_ctct(ct1, ct2)
    assert np.array_equal(got, o.mul_ctct(ct1, ct2, evk_b, evk_a))
    dec, amb, budget = o.decrypt(sk, got)
    want = np.array([(int(x) * int(y)) % t for x, y in zip(m1, m2)], dtype=np.int64)
    want = np.where(want > t // 2, want - t, want)
    assert amb == 0 and np.array_equal(dec, want)


CASES = [
    # N, L, K, b, E   -- ragged / edge shapes the loops of run() must survive
    (1024, 2, 2, 5, 7),
    (1024, 2, 1, 3, 4),     # K = 1: no ct x ct at all, mask only
    (1024, 2, 3, 2, 3),     # K = 3: two chained ct x ct (depth 2)
    (256, 1, 2, 1, 1),      # single bin, single position, single limb
    (2048, 3, 2, 9, 2),
    (512, 2, 2, 2, 300),    # E > 128: the lazy 128-bit accumulator is folded mid-way
    (8192, 3, 2, 4, 5),
    (16384, 6, 2, 2, 3),    # depth-5 context (E >= 500 in the client's rule): sizeQ = sizeP = 6
    (2048, 5, 3, 2, 2),
]


@pytest.mark.parametrize("N,L,K,b,E", CASES)
def test_run_random_limbs(N, L, K, b, E):
    """run() on uniformly random limbs: database, masks, query, key — bit-exact result limbs."""
    cc, o, params = ctx_and_oracle(N, L)
    rng = np.random.default_rng(N + 31 * K + 7 * b + E)
    sk, evk_b, evk_a = o.keygen(4)
    cc.InsertEvalMultKey(evk_b, evk_a)
    pt = sc.random_pt(rng, params, (K, b, E))
    mask = sc.random_pt(rng, params, (b,))
    idx = sc.random_ct(rng, params, (K, E))
    minus = sc.random_ct(rng, params)
    cc.db_load_limbs(pt, mask)
    cc.query_set(idx, minus)
    cc.run()
    got = cc.result_get()
    want = o.run(pt, mask, idx, minus, evk_b, evk_a, nthreads=8)
    assert np.array_equal(got, want)
    assert cc.run_launch_count() >= 1
    # idempotent: a second run on the same query gives the same limbs
    cc.run()
    assert np.array_equal(cc.result_get(), want)


def test_encode_slots_matches_oracle():
    """MakePackedPlaintext + SetFormat(EVALUATION) on the device, incl. negative values, short
    vectors (nslots < N) and the value range check."""
    cc, o, params = ctx_and_oracle(2048, 2)
    rng = np.random.default_rng(8)
    t = int(params.t)
    K, b, E, n = 2, 3, 4, 1500
    slots = rng.integers(-(t - 1), t, (K, b, E, n), dtype=np.int64)
    slots[0, 0, 0, :4] = [0, t - 1, -(t - 1), 1]
    mask_slots = rng.integers(1, t, (b, n), dtype=np.int64)
    cc.db_encode_slots(slots, mask_slots)
    pt, mask = cc.db_get_limbs()
    assert np.array_equal(pt, sc.encode_db(o, slots))
    assert np.array_equal(mask, sc.encode_masks(o, mask_slots))
    bad = slots.copy()
    bad[1, 2, 3, 7] = t
    with pytest.raises(ValueError):
        cc.db_encode_slots(bad, mask_slots)


def test_golden_small_case():
    """Committed fixture (tests/golden/small_case.npz, written by make_golden.py)."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "small_case.npz"))
    params = RefParams(256, T32, L=2).to_struct()
    cc = P.CryptoContext(params)
    cc.InsertEvalMultKey(z["evk_b"], z["evk_a"])
    cc.db_encode_slots(z["slots"], z["mask_slots"])
    pt, mask = cc.db_get_limbs()
    assert np.array_equal(pt, z["pt"]) and np.array_equal(mask, z["mask"])
    cc.query_set(z["idx"], z["minus"])
    cc.run()
    assert np.array_equal(cc.result_get(), z["out"])


def test_pie_operator_end_to_end():
    """The reference-shaped operator: HierarchicalCuckooHashTable -> BatchedFHEHIPPIE(ctx, pk, hct) ->
    setIndex / setMinusCompareElement / run / getResultList; limbs equal the oracle's evaluation of the
    same (shuffled) table, and the decrypted intersection equals the planted one."""
    d = P.RandomDataInput(3000, 48, 25, 777, 32)
    s = sc.table_scenario(2048, T32, 3, 2, 40, 2, 7, 9, d.serverSet, d.clientSet)
    cc = P.CryptoContext(s.params)
    cc.InsertEvalMultKey(s.evk_b, s.evk_a)
    pie = P.BatchedFHEHIPPIE(cc, P.PublicKey(), s.hct, keepSlots=True)
    slots, mask_slots = pie.slots()
    # ctor semantics: slot s of plaintext (hf, bin, pos) is the (shuffled) cell of outer table s
    cells = s.hct.cells()  # after the in-place shuffle
    k, e, K, b, E = cells.shape
    assert np.array_equal(slots, cells.reshape(k * e, K, b, E).transpose(1, 2, 3, 0).astype(np.int64))
    assert mask_slots.min() >= 1 and mask_slots.max() < int(s.params.t)
    pie.setIndex([[s.idx[hf, pos] for pos in range(E)] for hf in range(K)])
    pie.setMinusCompareElement(s.minus)
    pie.run()
    got = pie.getResultList()
    pt, mask = sc.encode_db(s.oracle, slots), sc.encode_masks(s.oracle, mask_slots)
    want = s.oracle.run(pt, mask, s.idx, s.minus, s.evk_b, s.evk_a, nthreads=8)
    assert np.array_equal(got, want)
    dec, budget = sc.decrypt_results(s, got)
    assert budget > 5
    inter = np.sort(P.extract_intersection(s.client_cells, dec))
    assert np.array_equal(inter, np.sort(d.intersectionSet))


def test_reference_known_answer_on_gpu():
    """tests/TestBatchedFHEPIE.cpp scenario through the GPU operator: "Matches" exactly twice."""
    from test_oracle import reference_test_scenario
    elems, client_elem = reference_test_scenario()
    s = sc.table_scenario(8192, T32, None, 2, 1, 2, 10, 20, elems, [client_elem], hash_seed=12223222, depth=2)
    o = s.oracle
    s.idx_slots[:] = 0
    for hf in range(2):
        pos = int(P.hash_index(s.hash, [client_elem], 2 + hf, 10)[0])
        s.idx_slots[hf, pos, :] = 1
    s.minus_slots[:] = -int(client_elem)
    for hf in range(2):
        for pos in range(10):
            s.idx[hf, pos] = o.encrypt(s.sk, s.idx_slots[hf, pos], 1000 + hf * 10 + pos)
    s.minus = o.encrypt(s.sk, s.minus_slots, 999)
    cc = P.CryptoContext(s.params)
    cc.InsertEvalMultKey(s.evk_b, s.evk_a)
    pie = P.BatchedFHEHIPPIE(cc, P.PublicKey(), s.hct)
    pie.setIndex(s.idx)
    pie.setMinusCompareElement(s.minus)
    pie.run()
    dec, budget = sc.decrypt_results(s, pie.getResultList())
    assert int((dec[:, :2] == 0).sum()) == 2 and budget > 5


def test_error_behaviour():
    """Call-order and argument errors surface as exceptions, as the reference throws
    (std::invalid_argument, BatchedFHEHIPPIE.cpp:13-21)."""
    cc, o, params = ctx_and_oracle(256, 2)
    rng = np.random.default_rng(1)
    with pytest.raises(P.PsiError):
        cc._dims = (2, 1, 1)
        cc.query_set(sc.random_ct(rng, params, (2, 1)), sc.random_ct(rng, params))   # no database yet
    pt = sc.random_pt(rng, params, (2, 1, 1))
    cc.db_load_limbs(pt, sc.random_pt(rng, params, (1,)))
    with pytest.raises(P.PsiError):
        cc.run()                                                                       # no query
    cc.query_set(sc.random_ct(rng, params, (2, 1)), sc.random_ct(rng, params))
    with pytest.raises(P.PsiError):
        cc.run()                                                                       # K = 2 needs the relin key
    h = P.TabulationHashing(1, 4)
    hct = P.HierarchicalCuckooHashTable(h, 4, 3, 2, 2, 2, True, True, 2)               # stash = 2
    with pytest.raises(ValueError, match="stash"):
        P.BatchedFHEHIPPIE(cc, P.PublicKey(), hct)
    with pytest.raises(ValueError, match="combined"):
        P.HierarchicalCuckooHashTable(h, 4, 3, 0, 2, 2, False, True, 2)


def test_pipelined_queries_double_buffered_results():
    """psi_query_upload / psi_query_commit / double-buffered results: two different queries in flight give
    each its own result (the bench's pipelined e2e path), and re-running the first after the second too."""
    cc, o, params = ctx_and_oracle(1024, 2)
    rng = np.random.default_rng(77)
    sk, evk_b, evk_a = o.keygen(6)
    cc.InsertEvalMultKey(evk_b, evk_a)
    K, b, E = 2, 3, 4
    pt = sc.random_pt(rng, params, (K, b, E))
    mask = sc.random_pt(rng, params, (b,))
    cc.db_load_limbs(pt, mask)
    queries = [(sc.random_ct(rng, params, (K, E)), sc.random_ct(rng, params)) for _ in range(3)]
    want = [o.run(pt, mask, q[0], q[1], evk_b, evk_a) for q in queries]
    outs = [np.empty((b, 2, params.L, params.N), dtype=np.uint64) for _ in queries]
    for i, (idx, minus) in enumerate(queries):
        idx = np.ascontiguousarray(idx)
        minus = np.ascontiguousarray(minus)
        cc.query_upload_ptr(idx.ctypes.data, minus.ctypes.data)   # default stream: ordered, pageable host memory
        cc.query_commit()
        cc.run()
        cc.result_get(outs[i], sync=False)                          # read back only after everything is queued
    cc.sync()
    for i in range(3):
        assert np.array_equal(outs[i], want[i]), "query %d" % i


def test_two_landing_buffers_queue_uploads():
    """Two landing buffers: query i+1 may be uploaded before query i is committed; commits consume the uploads in
    order; a third pending upload and a commit without an upload are PSI_ERR_STATE."""
    cc, o, params = ctx_and_oracle(1024, 2)
    rng = np.random.default_rng(78)
    sk, evk_b, evk_a = o.keygen(6)
    cc.InsertEvalMultKey(evk_b, evk_a)
    K, b, E = 2, 2, 3
    pt = sc.random_pt(rng, params, (K, b, E))
    mask = sc.random_pt(rng, params, (b,))
    cc.db_load_limbs(pt, mask)
    queries = [(np.ascontiguousarray(sc.random_ct(rng, params, (K, E))), np.ascontiguousarray(sc.random_ct(rng, params)))
               for _ in range(3)]
    want = [o.run(pt, mask, q[0], q[1], evk_b, evk_a) for q in queries]
    with pytest.raises(P.PsiError) as err:
        cc.query_commit()
    assert err.value.status == P.capi.PSI_ERR_STATE
    assert cc.query_next_landing() == 0
    cc.query_upload_ptr(queries[0][0].ctypes.data, queries[0][1].ctypes.data)
    assert cc.query_next_landing() == 1
    cc.query_upload_ptr(queries[1][0].ctypes.data, queries[1][1].ctypes.data)
    with pytest.raises(P.PsiError) as err:
        cc.query_upload_ptr(queries[2][0].ctypes.data, queries[2][1].ctypes.data)
    assert err.value.status == P.capi.PSI_ERR_STATE
    outs = []
    for i in range(3):
        cc.query_commit()
        if i == 0:   # buffer 0 is free again: the third query goes there while the second is still pending
            assert cc.query_next_landing() == 0
            cc.query_upload_ptr(queries[2][0].ctypes.data, queries[2][1].ctypes.data)
        cc.run()
        outs.append(cc.result_get())
    for i in range(3):
        assert np.array_equal(outs[i], want[i]), "query %d" % i


def test_cpp_known_answer_program():
    """tests/cpp/TestBatchedFHEPIE.cpp: the reference's own test program re-targeted at the C++ drop-in class;
    success criterion of the reference: the string "Matches" printed exactly twice (TestBatchedFHEPIE.cpp:73)."""
    import subprocess
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp")
    subprocess.check_call(["make", "-C", here, "-s"])
    import torch
    second = 1 if torch.cuda.device_count() > 1 else 0
    # no argument: one psi_ctx; with a device list: the single-process multi-device evaluator (two distinct GPUs
    # when the box has them, else two contexts on GPU 0)
    for extra in ([], ["0"], ["0,%d" % second], ["0,%d,0" % second]):
        out = subprocess.run([os.path.join(here, "TestBatchedFHEPIE")] + extra, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        assert out.stdout.count("Matches\n") == 2
        assert "Test should output matches twice" in out.stdout


def test_run_randomised_shapes():
    """Seeded sweep over ragged shapes and limb counts: every draw is a full run() compared limb for limb."""
    rng = np.random.default_rng(20261018)
    for trial in range(10):
        N = int(rng.choice([256, 512, 1024, 2048, 4096]))
        L = int(rng.integers(1, 6))
        K = int(rng.choice([1, 2, 2, 2, 3]))
        b = int(rng.integers(1, 12))
        E = int(rng.integers(1, 20))
        cc, o, params = ctx_and_oracle(N, L)
        sk, evk_b, evk_a = o.keygen(100 + trial)
        cc.InsertEvalMultKey(evk_b, evk_a)
        pt = sc.random_pt(rng, params, (K, b, E))
        mask = sc.random_pt(rng, params, (b,))
        idx = sc.random_ct(rng, params, (K, E))
        minus = sc.random_ct(rng, params)
        cc.db_load_limbs(pt, mask)
        got_pt, got_mask = cc.db_get_limbs()               # tile-major split-30 storage round trip
        assert np.array_equal(got_pt, pt) and np.array_equal(got_mask, mask)
        cc.query_set(idx, minus)
        cc.run()
        want = o.run(pt, mask, idx, minus, evk_b, evk_a, nthreads=8)
        assert np.array_equal(cc.result_get(), want), (trial, N, L, K, b, E)
        cc.close()


def test_one_context_many_databases_and_streams():
    """run() is replayed as a CUDA graph captured per (phases, result buffer, grouping, buffer addresses): the same
    context is re-loaded with databases of other shapes (larger, smaller, K = 1, back to the first), run on
    different streams, with the phases alone and together -- every result against the oracle."""
    import torch

    cc, o, params = ctx_and_oracle(1024, 3)
    rng = np.random.default_rng(77)
    sk, evk_b, evk_a = o.keygen(6)
    cc.InsertEvalMultKey(evk_b, evk_a)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    shapes = [(2, 5, 7), (2, 9, 3), (2, 2, 4), (1, 4, 5), (3, 3, 2), (2, 5, 7)]
    for i, (K, b, E) in enumerate(shapes):
        pt = sc.random_pt(rng, params, (K, b, E))
        mask = sc.random_pt(rng, params, (b,))
        cc.db_load_limbs(pt, mask)
        for rep in range(3):
            sp = streams[(i + rep) & 1].cuda_stream
            idx = sc.random_ct(rng, params, (K, E))
            minus = sc.random_ct(rng, params)
            cc.query_set(idx, minus, sp)
            if rep == 1:           # the two phases as separate calls
                cc.run(sp, phases=1)
                cc.run(sp, phases=2)
            else:
                cc.run(sp)
            got = cc.result_get(stream=sp)
            assert np.array_equal(got, o.run(pt, mask, idx, minus, evk_b, evk_a, nthreads=8)), (K, b, E, rep)
        if i == 2:
            cc.set_tuning(phase2_groups=1)
        if i == 4:
            cc.set_tuning(phase2_groups=0)
    cc.close()


def test_run_is_repeatable_under_load():
    """Race hunting without a sanitizer (compute-sanitizer is closed on the GPU pool): the full-size ring dimension,
    enough bins to fill the machine several times over, the same query evaluated 150 times back to back - every
    result must be bit-identical to the first, and the first to the oracle on a sampled bin.  A missing barrier
    between register passes, a ring slot released too early or a stale twiddle tile shows up as a flaky limb."""
    import zlib
    cc, o, params = ctx_and_oracle(16384, 4)
    rng = np.random.default_rng(4242)
    _, evk_b, evk_a = o.keygen(12)
    cc.InsertEvalMultKey(evk_b, evk_a)
    K, b, E = 2, 40, 11
    pt = sc.random_pt(rng, params, (K, b, E))
    mask = sc.random_pt(rng, params, (b,))
    idx = sc.random_ct(rng, params, (K, E))
    minus = sc.random_ct(rng, params)
    cc.db_load_limbs(pt, mask)
    cc.query_set(idx, minus)
    cc.run()
    first = cc.result_get()
    bin_ = 23
    want = o.run(np.ascontiguousarray(pt[:, bin_:bin_ + 1]), np.ascontiguousarray(mask[bin_:bin_ + 1]), idx, minus, evk_b,
                 evk_a, nthreads=4)
    assert np.array_equal(first[bin_], want[0])
    ref = zlib.crc32(first.tobytes())
    out = np.empty_like(first)
    for it in range(150):
        cc.run()
        cc.result_get(out)
        assert zlib.crc32(out.tobytes()) == ref, "run %d differs from the first" % it


def test_three_stream_pipeline_returns_every_query_its_own_result():
    """The choreography bench.py times as e2e (upload stream / compute stream / download stream, two landing
    buffers, two result buffers, events between them) with DIFFERENT queries in flight: every query must get the
    oracle's result for its own ciphertexts."""
    import torch
    cc, o, params = ctx_and_oracle(2048, 3)
    rng = np.random.default_rng(99)
    _, evk_b, evk_a = o.keygen(5)
    cc.InsertEvalMultKey(evk_b, evk_a)
    K, b, E, nq = 2, 6, 5, 7
    pt = sc.random_pt(rng, params, (K, b, E))
    mask = sc.random_pt(rng, params, (b,))
    cc.db_load_limbs(pt, mask)
    ct_words = 2 * params.L * params.N
    queries, hosts, outs = [], [], []
    for _ in range(nq):
        idx, minus = sc.random_ct(rng, params, (K, E)), sc.random_ct(rng, params)
        queries.append((idx, minus))
        h = torch.empty((K * E + 1) * ct_words, dtype=torch.int64).pin_memory()
        h[:K * E * ct_words] = torch.from_numpy(np.ascontiguousarray(idx).view(np.int64).reshape(-1))
        h[K * E * ct_words:] = torch.from_numpy(np.ascontiguousarray(minus).view(np.int64).reshape(-1))
        hosts.append(h)
        outs.append(torch.empty(b * ct_words, dtype=torch.int64).pin_memory())
    s_in, s_run, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    ev_commit, ev_d2h = [None, None], [None, None]
    for i in range(nq):
        if ev_commit[i & 1] is not None:
            s_in.wait_event(ev_commit[i & 1])          # landing buffer i & 1 is free once query i-2 was committed
        base = hosts[i].data_ptr()
        cc.query_upload_ptr(base, base + K * E * ct_words * 8, s_in.cuda_stream)
        ev_up = torch.cuda.Event()
        ev_up.record(s_in)
        s_run.wait_event(ev_up)
        cc.query_commit(s_run.cuda_stream)
        ev_commit[i & 1] = torch.cuda.Event()
        ev_commit[i & 1].record(s_run)
        if ev_d2h[i & 1] is not None:
            s_run.wait_event(ev_d2h[i & 1])            # result buffer i & 1 was read back
        cc.run(s_run.cuda_stream)
        ev_run = torch.cuda.Event()
        ev_run.record(s_run)
        s_out.wait_event(ev_run)
        cc.result_get_ptr(outs[i].data_ptr(), s_out.cuda_stream)
        ev_d2h[i & 1] = torch.cuda.Event()
        ev_d2h[i & 1].record(s_out)
    torch.cuda.synchronize()
    for i, (idx, minus) in enumerate(queries):
        want = o.run(pt, mask, idx, minus, evk_b, evk_a)
        got = outs[i].numpy().view(np.uint64).reshape(want.shape)
        assert np.array_equal(got, want), "query %d" % i
