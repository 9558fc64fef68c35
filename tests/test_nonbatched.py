"""CPU tests of the ORACLE's non-batched FHEHIPPIE restatement (oracle/psi_oracle.c: orc_nb_run and the OpenFHE
operations under it; reference FHEHIPPIE.cpp:9-77, SimpleFHEPSIClient.cpp:79-153,245-262).

Limb parity with OpenFHE is unpinned (no OpenFHE here).  Pinned on the CPU: the EVALUATION-format permutation against
the coefficient-domain definition of the automorphism, rotation / EvalSum / EvalMerge slot semantics after decryption,
and the reference's decode rule for one PIE: the client item is in the server's cell iff a decrypted slot < b is 0."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from oracle.params_ref import RefParams

T32 = 4296540161


def small_oracle(N=64, L=2, depth=None):
    return Oracle(RefParams(N, T32, L=L, depth=depth).to_struct())


def test_evaluation_permutation_is_the_automorphism():
    """NTT(a(X^g)) == AutomorphismTransform_g(NTT(a)) for every limb and several indices."""
    o = small_oracle(64, 2)
    rng = np.random.default_rng(0)
    for g in (5, 25, 3, 2 * o.N - 1, o.find_automorphism_index(-3)):
        for l in range(o.L):
            q = int(o.params.q[l])
            a = rng.integers(0, q, o.N, dtype=np.uint64)
            want = o.ntt(o.automorphism_coeff(a, g, l), l)
            full = np.zeros((o.L, o.N), dtype=np.uint64)
            full[l] = o.ntt(a, l)
            got = o.automorphism_eval(full, g)[l]
            assert np.array_equal(got, want), (g, l)


def test_automorphism_indices():
    o = small_oracle(64, 2)
    m = 2 * o.N
    assert o.find_automorphism_index(1) == 5 and o.find_automorphism_index(3) == 125 % m
    assert o.find_automorphism_index(-1) * 5 % m == 1
    assert o.find_automorphism_index(-4) * pow(5, 4, m) % m == 1
    # EvalSum_2n / GenerateIndices_2n: ceil(log2 bs) indices, squares of 5
    assert o.eval_sum_indices(1) == []
    assert o.eval_sum_indices(2) == [5]
    assert o.eval_sum_indices(5) == [5, 25, 625 % m]
    assert o.eval_sum_indices(8) == [5, 25, 625 % m]
    assert o.eval_sum_indices(o.N) == [pow(5, 2 ** i, m) for i in range(5)] + [m - 1]


def test_rotation_semantics_and_key_switch():
    """EvalAtIndex(ct, i) rotates the packed slots left by i (first half-row), and decrypts under the SAME secret."""
    o = small_oracle(128, 2)
    sk, _, _ = o.keygen(5)
    half = o.N // 2
    v = np.arange(1, o.N + 1, dtype=np.int64)
    ct = o.encrypt(sk, v, 77)
    for i in (1, 3, -2):
        g = o.find_automorphism_index(i)
        kb, ka = o.auto_keygen(sk, 9, [g])
        rot = o.eval_automorphism(ct, g, kb[0], ka[0])
        dec, amb, budget = o.decrypt(sk, rot)
        assert amb == 0 and budget > 5
        want = np.concatenate([np.roll(v[:half], -i), np.roll(v[half:], -i)])
        assert np.array_equal(dec, want), i


def nb_scenario(o, K, b, rng, x, items, key_seed=3):
    """One PIE the way the reference builds it.  items [K][b][E] (0 = empty cell), E == b; client element x."""
    E = b
    sk, _, _ = o.keygen(key_seed)
    t = int(o.t)
    # FHEHIPPIE ctor (FHEHIPPIE.cpp:41-58): plainVec = the E positions of (hf, bin) + a 1 for the minus element
    pt = np.empty((K, b, o.L, o.N), dtype=np.uint64)
    for hf in range(K):
        for bin_ in range(b):
            pt[hf, bin_] = o.encode(np.concatenate([items[hf, bin_], [1]]).astype(np.int64))
    mask_slots = rng.integers(1, t, size=(K, b), dtype=np.int64)
    mask = np.stack([o.encode(mask_slots[hf]) for hf in range(K)])
    merge_pt = o.encode(np.array([1], dtype=np.int64))
    # client (SimpleFHEPSIClient.cpp:79-90,118-153): keys, one-hot index + minus element in slot E
    sum_idx = o.eval_sum_indices(E + 1)
    rot_idx = [o.find_automorphism_index(-(i + 1)) for i in range(E)]
    key_index = list(dict.fromkeys(sum_idx + rot_idx))
    key_b, key_a = o.auto_keygen(sk, 1234, key_index)
    return sk, pt, mask_slots, mask, merge_pt, key_index, key_b, key_a


def query_for(o, sk, K, E, x, positions, seed=500):
    idx = np.empty((K, 2, o.L, o.N), dtype=np.uint64)
    for hf in range(K):
        v = np.zeros(E + 1, dtype=np.int64)
        v[positions[hf]] = 1
        v[E] = -int(x)
        idx[hf] = o.encrypt(sk, v, seed + hf)
    return idx


@pytest.mark.parametrize("b", [3, 5])
def test_nb_run_decodes_like_the_reference_client(b):
    """Result hf, slot bin = (item[hf][bin][pos_hf] - x) * r: zero exactly where the server cell holds x."""
    o = small_oracle(256, 3)
    rng = np.random.default_rng(b)
    K, E = 2, b
    t = int(o.t)
    items = rng.integers(2, 2 ** 32, size=(K, b, E), dtype=np.int64)
    x = 123456789
    positions = [1, E - 1]
    items[1, 2, positions[1]] = x  # the match: hash function 1, bin 2
    sk, pt, mask_slots, mask, merge_pt, key_index, key_b, key_a = nb_scenario(o, K, b, rng, x, items)
    idx = query_for(o, sk, K, E, x, positions)
    out = o.nb_run(idx, pt, merge_pt, mask, key_index, key_b, key_a)
    for hf in range(K):
        dec, amb, budget = o.decrypt(sk, out[hf])
        assert amb == 0 and budget > 5, (hf, budget)
        for bin_ in range(b):
            want = (int(items[hf, bin_, positions[hf]]) - x) * int(mask_slots[hf, bin_]) % t
            want = want - t if want > t // 2 else want
            assert int(dec[bin_]) == want, (hf, bin_)
        zero_bins = [bin_ for bin_ in range(b) if dec[bin_] == 0]
        assert zero_bins == ([2] if hf == 1 else [])


def test_nb_run_reports_missing_key():
    o = small_oracle(64, 2)
    rng = np.random.default_rng(1)
    K, b = 1, 3
    items = rng.integers(2, 2 ** 32, size=(K, b, b), dtype=np.int64)
    sk, pt, _, mask, merge_pt, key_index, key_b, key_a = nb_scenario(o, K, b, rng, 5, items)
    idx = query_for(o, sk, K, b, 5, [0])
    with pytest.raises(KeyError):
        o.nb_run(idx, pt, merge_pt, mask, key_index[:1], key_b[:1], key_a[:1])


def test_hybrid_rotation_semantics():
    """The same rotation semantics with KeySwitchTechnique HYBRID (digits of alpha limbs, special primes, ApproxModDown)."""
    o = Oracle(RefParams(128, T32, L=3, ks_technique=1).to_struct())
    sk, _, _ = o.keygen(5)
    half = o.N // 2
    v = np.arange(1, o.N + 1, dtype=np.int64)
    ct = o.encrypt(sk, v, 77)
    for i in (1, -2):
        g = o.find_automorphism_index(i)
        kb, ka = o.auto_keygen(sk, 9, [g], key_seed=5)
        assert kb.shape == (1,) + o.evk_shape()
        rot = o.eval_automorphism(ct, g, kb[0], ka[0])
        dec, amb, budget = o.decrypt(sk, rot)
        assert amb == 0 and budget > 5
        want = np.concatenate([np.roll(v[:half], -i), np.roll(v[half:], -i)])
        assert np.array_equal(dec, want), i


def test_nb_run_hybrid_decodes():
    o = Oracle(RefParams(256, T32, L=3, ks_technique=1).to_struct())
    rng = np.random.default_rng(8)
    K, b = 2, 3
    E, t = b, int(o.t)
    items = rng.integers(2, 2 ** 32, size=(K, b, E), dtype=np.int64)
    x, positions = 987654321, [2, 0]
    items[0, 1, positions[0]] = x
    sk, _, _ = o.keygen(3)
    pt = np.stack([np.stack([o.encode(np.concatenate([items[hf, bin_], [1]]).astype(np.int64)) for bin_ in range(b)]) for hf in range(K)])
    mask_slots = rng.integers(1, t, size=(K, b), dtype=np.int64)
    mask = np.stack([o.encode(mask_slots[hf]) for hf in range(K)])
    merge_pt = o.encode(np.array([1], dtype=np.int64))
    key_index = list(dict.fromkeys(o.eval_sum_indices(E + 1) + [o.find_automorphism_index(-(i + 1)) for i in range(E)]))
    key_b, key_a = o.auto_keygen(sk, 1234, key_index, key_seed=3)
    idx = query_for(o, sk, K, E, x, positions)
    out = o.nb_run(idx, pt, merge_pt, mask, key_index, key_b, key_a)
    for hf in range(K):
        dec, amb, budget = o.decrypt(sk, out[hf])
        assert amb == 0 and budget > 5
        for bin_ in range(b):
            want = (int(items[hf, bin_, positions[hf]]) - x) * int(mask_slots[hf, bin_]) % t
            assert int(dec[bin_]) == (want - t if want > t // 2 else want)
        assert [bin_ for bin_ in range(b) if dec[bin_] == 0] == ([1] if hf == 0 else [])


def test_golden_nonbatched_case():
    """Committed fixture (tests/golden/nonbatched_case.npz, written by make_golden.py): the checker cannot drift silently."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "nonbatched_case.npz"))
    o = Oracle(RefParams(256, T32, L=2).to_struct())
    for p in range(z["idx"].shape[0]):
        assert np.array_equal(o.encode(z["slots"][p, 0, 0]), z["pt"][p, 0, 0])
        got = o.nb_run(z["idx"][p], z["pt"][p], z["merge"], z["mask"][p], z["key_index"], z["key_b"], z["key_a"])
        assert np.array_equal(got, z["out"][p])


def test_collection_constructor_logic_without_a_device():
    """FHEHIPPIE ctor bookkeeping of the Python mirror (FHEHIPPIE.cpp:9-59): errors, bin permutation, trailing 1, masks."""
    import types
    import psi_b200 as P
    coll = P.FHEHIPPIECollection(types.SimpleNamespace(t=T32), P.PublicKey(), seed=7)
    with pytest.raises(ValueError, match="size of a cuckoo bin"):
        coll.addPIE(np.ones((2, 3, 4), dtype=np.int64))
    with pytest.raises(ValueError, match="stash"):
        coll.addPIE(np.ones((2, 3, 3), dtype=np.int64), stash_size=2)
    rng = np.random.default_rng(0)
    table = rng.integers(1, 2 ** 32, size=(2, 5, 5), dtype=np.int64)
    pie = coll.addPIE(table)
    slots, masks, perm = coll._slots[0], coll._masks[0], coll._perm[0]
    assert slots.shape == (2, 5, 6) and (slots[:, :, 5] == 1).all()
    for hf in range(2):   # the bins are permuted as whole rows, the same permutation for every hash function
        assert sorted(map(tuple, slots[hf, :, :5])) == sorted(map(tuple, table[hf]))
    order = [next(j for j in range(5) if (slots[0, j, :5] == table[0, i]).all()) for i in range(5)]
    assert [next(j for j in range(5) if (slots[1, j, :5] == table[1, i]).all()) for i in range(5)] == order
    assert masks.shape == (2, 5) and (masks >= 1).all() and (masks < T32).all()
    assert sorted(perm) == [0, 1]
    with pytest.raises(ValueError, match="same table shape"):
        coll.addPIE(np.ones((2, 4, 4), dtype=np.int64))
    with pytest.raises(ValueError, match="run\\(\\) has not been called"):
        pie.getResultList()


def test_cpp_constructor_bookkeeping_without_a_device():
    """tests/cpp/TestFHEPIECtor.cpp: the C++ mirror's constructor (host/FHEHIPPIE.hpp) - the reference's two
    invalid_argument cases, bin permutation, trailing 1, mask range, seeded reproducibility; no device involved."""
    import os
    import subprocess
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp")
    subprocess.check_call(["make", "-C", here, "-s", "TestFHEPIECtor"])
    r = subprocess.run([os.path.join(here, "TestFHEPIECtor")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "constructor bookkeeping ok" in r.stdout
