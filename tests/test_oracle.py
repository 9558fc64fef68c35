"""CPU tests of the ORACLE (oracle/psi_oracle.c): the checker must itself be pinned before the GPU
path is compared with it.  Parity with OpenFHE limbs is UNPINNED (no OpenFHE here, no golden
ciphertexts in the reference); what is pinned:
  * the transforms against their mathematical definition,
  * BFV multiplication against an exact big-integer model (oracle/bfv_exact.py),
  * the decrypted semantics of the reference's own test (tests/TestBatchedFHEPIE.cpp: "Matches" twice),
  * committed golden digests (tests/golden/) so the checker cannot drift silently.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import psi_b200 as P
from oracle.bfv_exact import ExactBFV
from oracle.oracle import Oracle
from oracle.params_ref import RefParams

import scenario as sc

T32 = 4296540161  # 2^32 + 2^20 + 2^19 + 1  (BatchedFHEPSIClient.cpp:29, TestBatchedFHEPIE.cpp:16)


def small_oracle(N=64, L=2, **variant):
    return Oracle(RefParams(N, T32, L=L, **variant).to_struct())


def test_ntt_matches_definition():
    """forward output index j holds a(psi^(2*bitrev(j)+1))  (natural in, bit-reversed out)."""
    o = small_oracle(32, 2)
    rng = np.random.default_rng(0)
    N = o.N
    logN = N.bit_length() - 1
    for m in range(o.L + o.Lp + 1):
        if m < o.L:
            q, psi = int(o.params.q[m]), int(o.params.psi_q[m])
        elif m < o.L + o.Lp:
            q, psi = int(o.params.p[m - o.L]), int(o.params.psi_p[m - o.L])
        else:
            q, psi = int(o.params.t), int(o.params.psi_t)
        a = rng.integers(0, q, N, dtype=np.uint64)
        got = o.ntt(a, m)
        for j in range(N):
            r = int(format(j, "0%db" % logN)[::-1], 2)
            x = pow(psi, 2 * r + 1, q)
            want = sum(int(a[i]) * pow(x, i, q) for i in range(N)) % q
            assert int(got[j]) == want
        assert np.array_equal(o.ntt(got, m, inverse=True), a)


def test_ntt_is_negacyclic_convolution():
    o = small_oracle(64, 2)
    rng = np.random.default_rng(1)
    q = int(o.params.q[0])
    a = rng.integers(0, q, o.N, dtype=np.uint64)
    b = rng.integers(0, q, o.N, dtype=np.uint64)
    fa, fb = o.ntt(a, 0), o.ntt(b, 0)
    prod = np.array([int(x) * int(y) % q for x, y in zip(fa, fb)], dtype=np.uint64)
    got = o.ntt(prod, 0, inverse=True)
    from oracle.bfv_exact import negacyclic_mul
    want = [x % q for x in negacyclic_mul([int(x) for x in a], [int(x) for x in b])]
    assert [int(x) for x in got] == want


def test_pack_is_a_ring_homomorphism():
    """unpack(pack(a) * pack(b)) == a * b slot-wise, unpack(pack(a)) == a, negatives are centred."""
    o = small_oracle(64, 2)
    rng = np.random.default_rng(2)
    t = int(o.t)
    a = rng.integers(-(t // 2), t // 2, o.N, dtype=np.int64)
    b = rng.integers(-(t // 2), t // 2, o.N, dtype=np.int64)
    pa, pb = o.pack(a), o.pack(b)
    assert np.array_equal(o.unpack(pa), a)
    from oracle.bfv_exact import negacyclic_mul
    prod = np.array([x % t for x in negacyclic_mul([int(x) for x in pa], [int(x) for x in pb])], dtype=np.uint64)
    want = np.array([(int(x) * int(y)) % t for x, y in zip(a, b)], dtype=np.int64)
    want = np.where(want > t // 2, want - t, want)
    assert np.array_equal(o.unpack(prod), want)
    with pytest.raises(ValueError):
        o.pack(np.array([t], dtype=np.int64))  # PackedEncoding::Encode rejects |v| >= t
    short = o.unpack(o.pack(np.array([5, -7], dtype=np.int64)))
    assert short[0] == 5 and short[1] == -7 and not short[2:].any()


def test_encrypt_decrypt_roundtrip():
    o = small_oracle(128, 3)
    rng = np.random.default_rng(3)
    sk, _, _ = o.keygen(5)
    m = rng.integers(-(int(o.t) // 2), int(o.t) // 2, o.N, dtype=np.int64)
    ct = o.encrypt(sk, m, 17)
    got, amb, budget = o.decrypt(sk, ct)
    assert amb == 0 and np.array_equal(got, m)
    assert budget >= 60  # fresh ciphertext (the estimate saturates at 64 bits of fixed-point precision)


VARIANTS = [dict(), dict(mult_technique=0), dict(ks_technique=1), dict(mult_technique=0, ks_technique=1, fp_contract=1),
            dict(fp_contract=1)]


@pytest.mark.parametrize("variant", VARIANTS, ids=lambda v: "-".join("%s%d" % (k[:2], x) for k, x in v.items()) or "default")
@pytest.mark.parametrize("N,L", [(32, 2), (64, 3), (64, 5)])
def test_mul_core_against_exact_bfv(N, L, variant):
    """HPSPOVERQ / HPS tensor + scale-and-round vs the exact definition round(t/Q * tensor): the RNS
    procedure may only add a noise-sized term (bounded by ~ t * N * (L+1), from rounding P/Q * ct2),
    never a wrap-around; and both decrypt (with s, s^2) to the slot-wise product.  Relinearisation (BV or HYBRID)
    must leave the message alone; separate / fused evaluation of the double sums only moves roundings."""
    o = small_oracle(N, L, **variant)
    ex = ExactBFV(o)
    rng = np.random.default_rng(4)
    sk, evk_b, evk_a = o.keygen(9)
    t = int(o.t)
    m1 = rng.integers(-(t // 2), t // 2, N, dtype=np.int64)
    m2 = rng.integers(-(t // 2), t // 2, N, dtype=np.int64)
    ct1, ct2 = o.encrypt(sk, m1, 21), o.encrypt(sk, m2, 22)
    res = o.mul_core(ct1, ct2)                       # [3][L][N] COEFFICIENT
    exact = ex.mul(ct1, ct2)
    from oracle.bfv_exact import crt_reconstruct, centre
    worst = 0
    for c in range(3):
        v, Q = crt_reconstruct(res[c], ex.q)
        diff = centre([(x - y) % Q for x, y in zip(v, exact[c])], Q)
        worst = max(worst, max(abs(d) for d in diff))
    assert worst <= t * N * (L + 1), worst
    want = np.array([(int(x) * int(y)) % t for x, y in zip(m1, m2)], dtype=np.int64)
    want = np.where(want > t // 2, want - t, want)
    # 3-component decrypt of the RNS result (EVALUATION limbs for orc_decrypt)
    res_eval = np.stack([np.stack([o.ntt(res[c][l], l) for l in range(L)]) for c in range(3)])
    got3, amb, _ = o.decrypt(sk, res_eval)
    assert amb == 0 and np.array_equal(got3, want)
    # exact model decrypts to the same message
    m_exact, _ = ex.decrypt_int(exact, sk)
    assert np.array_equal(o.unpack(np.array(m_exact, dtype=np.uint64)), want)
    # relinearised 2-component result
    got2, amb, budget = o.decrypt(sk, o.relin(res, evk_b, evk_a))
    assert amb == 0 and np.array_equal(got2, want) and budget > 10
    assert np.array_equal(o.mul_ctct(ct1, ct2, evk_b, evk_a), o.relin(res, evk_b, evk_a))


def test_mul_ctpt_and_mac_semantics():
    o = small_oracle(64, 2)
    rng = np.random.default_rng(5)
    sk, _, _ = o.keygen(3)
    t = int(o.t)
    E = 5
    sel = 3
    items = rng.integers(1, t // 2, (E, o.N), dtype=np.int64)
    x = items[sel] .copy()
    idx = np.stack([o.encrypt(sk, np.full(o.N, 1 if pos == sel else 0, dtype=np.int64), 40 + pos) for pos in range(E)])
    pt = np.stack([o.encode(items[pos]) for pos in range(E)])
    minus = o.encrypt(sk, -x, 50)
    acc = o.mac_bin(idx, pt, minus)
    got, amb, _ = o.decrypt(sk, acc)
    assert amb == 0 and not got.any()          # selected item minus itself
    minus2 = o.encrypt(sk, -(x + 1), 51)
    got, _, _ = o.decrypt(sk, o.mac_bin(idx, pt, minus2))
    assert (got == -1).all()
    r = rng.integers(1, t // 2, o.N, dtype=np.int64)
    got, _, _ = o.decrypt(sk, o.mul_ctpt(o.mac_bin(idx, pt, minus2), o.encode(r)))
    assert np.array_equal(got, -r)


def reference_test_scenario():
    """tests/TestBatchedFHEPIE.cpp:54-99 — 100 random non-zero elements mod t, client element = #50,
    k=2, K=2, e=1, E=10, b=20, hash seed 12223222 with 4 hash functions, depth 2 (N chosen by the
    library: 8192, sizeQ 3 by this repo's restatement of the parameter generator)."""
    rng = np.random.default_rng(122333444455555 % (2**32))
    elems = rng.integers(1, T32, 100, dtype=np.uint64)
    return elems, elems[50]


def test_reference_known_answer_matches_twice():
    """"Test should output matches twice" (TestBatchedFHEPIE.cpp:73,145-146): over the b decrypted
    results the value 0 appears exactly twice in the first two slots (once per outer table, the element
    sits in one bin of each), and the intersection is exactly the client element."""
    elems, client_elem = reference_test_scenario()
    s = sc.table_scenario(8192, T32, None, 2, 1, 2, 10, 20, elems, [client_elem], hash_seed=12223222, depth=2)
    assert s.params.N == 8192 and s.params.L == 3
    # the test fills both slots with the same element (TestBatchedFHEPIE.cpp:111-113,132-134)
    s.idx_slots[:] = 0
    for hf in range(2):
        pos = int(P.hash_index(s.hash, [client_elem], 2 + hf, 10)[0])
        s.idx_slots[hf, pos, :] = 1
    s.minus_slots[:] = -int(client_elem)
    o = s.oracle
    for hf in range(2):
        for pos in range(10):
            s.idx[hf, pos] = o.encrypt(s.sk, s.idx_slots[hf, pos], 1000 + hf * 10 + pos)
    s.minus = o.encrypt(s.sk, s.minus_slots, 999)
    slots, mask_slots, pt, mask = sc.oracle_db(s)
    out = o.run(pt, mask, s.idx, s.minus, s.evk_b, s.evk_a, nthreads=4)
    dec, budget = sc.decrypt_results(s, out)
    assert budget > 5
    matches = int((dec[:, :2] == 0).sum())
    assert matches == 2
    # every other value is a non-zero multiple produced by the mask
    assert (dec[:, :2] != 0).sum() == 2 * 20 - 2


def test_small_protocol_intersection():
    """End-to-end acceptance on a small random instance (PSIClient::intersectionMatches,
    PSIClient.hpp:142-164): decrypted intersection == planted intersection."""
    d = P.RandomDataInput(2000, 40, 21, 4242, 32)
    s = sc.table_scenario(1024, T32, 3, 2, 32, 2, 6, 8, d.serverSet, d.clientSet)
    slots, mask_slots, pt, mask = sc.oracle_db(s)
    out = s.oracle.run(pt, mask, s.idx, s.minus, s.evk_b, s.evk_a, nthreads=4)
    dec, budget = sc.decrypt_results(s, out)
    got = np.sort(P.extract_intersection(s.client_cells, dec))
    assert np.array_equal(got, np.sort(d.intersectionSet))
    assert budget > 5


GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_digests.json")


def test_golden_digests():
    """Digests written by tests/golden/make_golden.py: the checker's outputs for fixed seeds."""
    from golden.make_golden import compute_digests
    with open(GOLDEN) as f:
        want = json.load(f)
    got = compute_digests()
    assert got == want
