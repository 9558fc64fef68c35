"""Risk-register variants of the BFV context on the GPU: MultiplicationTechnique HPS next to
HPSPOVERQ, KeySwitchTechnique HYBRID next to BV, separate / fused evaluation of the double sums.  The installed
OpenFHE decides which one the reference runs (CMakeLists.txt:10); every combination must be bit-exact against the
oracle's restatement of the same variant and must decrypt to the slot-wise product."""
import numpy as np
import pytest

import psi_b200 as P
from oracle.oracle import Oracle
from oracle.params_ref import RefParams

import scenario as sc

pytestmark = pytest.mark.gpu
T32 = 4296540161

VARIANTS = [
    dict(mult_technique=0),
    dict(ks_technique=1),
    dict(fp_contract=1),
    dict(mult_technique=0, ks_technique=1),
    dict(mult_technique=0, ks_technique=1, fp_contract=1),
]
IDS = lambda v: "-".join("%s%d" % (k[:2], x) for k, x in v.items())


def make(N, L, depth, variant):
    params = RefParams(N, T32, depth=depth, L=L, **variant).to_struct()
    return P.CryptoContext(params), Oracle(params), params


@pytest.mark.parametrize("variant", VARIANTS, ids=IDS)
@pytest.mark.parametrize("N,L,depth", [(16384, 4, 3), (2048, 3, 2), (1024, 5, 5), (1024, 2, 1), (4096, 7, 3)])
def test_mul_ctct_variants(N, L, depth, variant):
    if variant.get("mult_technique") == 0 and L + 1 > 8:
        pytest.skip("HPS needs sizeQ + 1 <= PSI_MAX_LIMBS")
    cc, o, params = make(N, L, depth, variant)
    rng = np.random.default_rng(N + L)
    sk, evk_b, evk_a = o.keygen(5)
    cc.InsertEvalMultKey(evk_b, evk_a)
    # uniformly random limbs (worst case for every rounding)
    r1, r2 = sc.random_ct(rng, params), sc.random_ct(rng, params)
    assert np.array_equal(cc.debug_mul_ctct(r1, r2), o.mul_ctct(r1, r2, evk_b, evk_a))
    # real ciphertexts: limbs and decrypted product
    t = int(params.t)
    m1 = rng.integers(-(t // 2), t // 2, params.N, dtype=np.int64)
    m2 = rng.integers(-(t // 2), t // 2, params.N, dtype=np.int64)
    ct1, ct2 = o.encrypt(sk, m1, 1), o.encrypt(sk, m2, 2)
    got = cc.debug_mul_ctct(ct1, ct2)
    assert np.array_equal(got, o.mul_ctct(ct1, ct2, evk_b, evk_a))
    if L >= 2:
        dec, amb, _ = o.decrypt(sk, got)
        want = np.array([(int(x) * int(y)) % t for x, y in zip(m1, m2)], dtype=np.int64)
        want = np.where(want > t // 2, want - t, want)
        assert amb == 0 and np.array_equal(dec, want)


@pytest.mark.parametrize("variant", VARIANTS, ids=IDS)
@pytest.mark.parametrize("N,L,K,b,E", [(1024, 3, 2, 5, 7), (2048, 4, 3, 2, 3), (16384, 4, 2, 3, 4)])
def test_run_variants(N, L, K, b, E, variant):
    """The whole run() (inner products, ct x ct chain, mask) under each variant."""
    cc, o, params = make(N, L, K + 1, variant)
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
    assert np.array_equal(cc.result_get(), o.run(pt, mask, idx, minus, evk_b, evk_a, nthreads=8))


def test_variants_differ():
    """The switches are live: HPS and HPSPOVERQ (and BV / HYBRID) produce different limbs for the same operands
    (the messages agree, the roundings do not) -- a context that silently ignored its variant would pass the
    parity tests above against an oracle that ignored it the same way."""
    rng = np.random.default_rng(1)
    outs = []
    for variant in [dict(), dict(mult_technique=0), dict(ks_technique=1)]:
        cc, o, params = make(1024, 3, 2, variant)
        sk, evk_b, evk_a = o.keygen(5)
        cc.InsertEvalMultKey(evk_b, evk_a)
        m = np.arange(params.N, dtype=np.int64)
        ct1, ct2 = o.encrypt(sk, m, 1), o.encrypt(sk, m, 2)
        outs.append(cc.debug_mul_ctct(ct1, ct2))
    assert not np.array_equal(outs[0], outs[1]) and not np.array_equal(outs[0], outs[2])
