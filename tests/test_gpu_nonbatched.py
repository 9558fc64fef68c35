"""Non-batched FHEHIPPIE on the GPU (psi_nb_*; reference FHEHIPPIE.cpp:9-77, SimpleFHEPSIServer.cpp:126-160) against
the oracle's restatement (orc_nb_run), bit-exact, through the C ABI; and the reference-shaped collection end to end
with the client's decode rule (SimpleFHEPSIClient.cpp:245-262)."""
import numpy as np
import pytest

import psi_b200 as P
from oracle.oracle import Oracle
from oracle.params_ref import RefParams

import scenario as sc

pytestmark = pytest.mark.gpu
T32 = 4296540161


def make(N, L, depth=None):
    params = RefParams(N, T32, depth=depth, L=L).to_struct()
    return P.CryptoContext(params), Oracle(params), params


def keys_for(o, sk, b, seed=77):
    """The client's key set (SimpleFHEPSIClient.cpp:79-90): EvalSumKeyGen for batch size E + 1, rotations -1 .. -E."""
    idx = list(dict.fromkeys(o.eval_sum_indices(b + 1) + [o.find_automorphism_index(-(i + 1)) for i in range(b)]))
    kb, ka = o.auto_keygen(sk, seed, idx)
    return idx, kb, ka


@pytest.mark.parametrize("N,L,n_pie,K,b", [(1024, 2, 3, 2, 3), (2048, 3, 2, 3, 5), (4096, 4, 1, 1, 7), (16384, 4, 2, 2, 6),
                                           (1024, 2, 2, 2, 1), (8192, 3, 1, 2, 9)])
def test_nb_run_random_limbs(N, L, n_pie, K, b):
    """Uniformly random residues everywhere (worst case for every reduction and digit lift)."""
    cc, o, params = make(N, L)
    rng = np.random.default_rng(N + b)
    pt = sc.random_pt(rng, params, (n_pie, K, b))
    mask = sc.random_pt(rng, params, (n_pie, K))
    merge = sc.random_pt(rng, params)
    idx = sc.random_ct(rng, params, (n_pie, K))
    key_index = list(dict.fromkeys(o.eval_sum_indices(b) + [o.find_automorphism_index(-i) for i in range(1, b)])) or [5]
    key_b = sc.random_pt(rng, params, (len(key_index), L))
    key_a = sc.random_pt(rng, params, (len(key_index), L))
    cc.InsertEvalAutomorphismKeys(key_index, key_b, key_a)
    cc.nb_db_load_limbs(pt, mask, merge)
    got = cc.nb_run(idx)
    for p in range(n_pie):
        want = o.nb_run(idx[p], pt[p], merge, mask[p], key_index, key_b, key_a)
        assert np.array_equal(got[p], want), p
    # a sub-range of the collection gives the same ciphertexts
    if n_pie > 1:
        assert np.array_equal(cc.nb_run(idx[1:], 1, n_pie), got[1:])
    assert cc.nb_launch_count() > 0


def test_nb_device_encode_matches_oracle():
    cc, o, params = make(2048, 3)
    rng = np.random.default_rng(4)
    n_pie, K, b = 2, 2, 4
    t = int(params.t)
    slots = rng.integers(-(t // 2), t // 2, (n_pie, K, b, b + 1), dtype=np.int64)
    masks = rng.integers(1, t, (n_pie, K, b), dtype=np.int64)
    cc.nb_db_encode_slots(slots, masks)
    pt, mask, merge = cc.nb_db_get_limbs()
    for p in range(n_pie):
        for hf in range(K):
            assert np.array_equal(mask[p, hf], o.encode(masks[p, hf]))
            for bin_ in range(b):
                assert np.array_equal(pt[p, hf, bin_], o.encode(slots[p, hf, bin_]))
    assert np.array_equal(merge, o.encode(np.array([1], dtype=np.int64)))


def test_nb_collection_end_to_end():
    """Real keys and ciphertexts; the reference-shaped collection; decode like the reference client."""
    cc, o, params = make(4096, 4)   # the client asks for depth 3 (SimpleFHEPSIClient.cpp:65): 4 limbs; BV key switching
    rng = np.random.default_rng(9)  # with 60-bit digits and three plaintext products leave no budget at 3 limbs
    n_pie, K, b = 3, 2, 5
    E, t = b, int(params.t)
    sk, _, _ = o.keygen(3)
    key_index, key_b, key_a = keys_for(o, sk, b)
    cc.InsertEvalAutomorphismKeys(key_index, key_b, key_a)
    tables = rng.integers(2, 2 ** 32, size=(n_pie, K, b, E), dtype=np.int64)
    tables[0, 1, 3, 2] = 0                      # an empty cell
    xs = [int(rng.integers(2, 2 ** 32)) for _ in range(n_pie)]
    pos = rng.integers(0, E, size=(n_pie, K))
    tables[1, 0, 4, pos[1, 0]] = xs[1]          # PIE 1 holds its client's element; PIEs 0 and 2 do not
    coll = P.FHEHIPPIECollection(cc, P.PublicKey(), seed=5)
    for p in range(n_pie):
        pie = coll.addPIE(tables[p])
        ims = []
        for hf in range(K):
            v = np.zeros(E + 1, dtype=np.int64)
            v[pos[p, hf]] = 1
            v[E] = -xs[p]
            ims.append(o.encrypt(sk, v, 100 + p * K + hf))
        pie.setIndex(ims)
    coll.runAll()
    # bit-exact against the oracle on the database the device encoded
    pt, mask, merge = cc.nb_db_get_limbs()
    found = []
    for p in range(n_pie):
        res = coll.myPIEs[p].getResultList()
        want = o.nb_run(np.stack(coll.myPIEs[p].indexMatrix), pt[p], merge, mask[p], key_index, key_b, key_a)
        assert np.array_equal(res[coll._perm[p]], want)
        hit = False
        for hf in range(K):
            dec, amb, budget = o.decrypt(sk, res[hf])
            assert amb == 0 and budget > 5
            hit = hit or bool((dec[:b] == 0).any())   # receiveResult: SetLength(maxItemsPerPosition), any 0
        found.append(hit)
    assert found == [False, True, False]
    # one PIE on its own gives the same ciphertexts as inside runAll
    before = coll.myPIEs[2].getResultList().copy()
    coll.myPIEs[2].run()
    assert np.array_equal(coll.myPIEs[2].getResultList(), before)


def test_nb_errors_like_the_reference():
    cc, o, params = make(1024, 2)
    coll = P.FHEHIPPIECollection(cc, P.PublicKey(), seed=1)
    with pytest.raises(ValueError, match="size of a cuckoo bin"):
        coll.addPIE(np.ones((2, 3, 4), dtype=np.int64))
    with pytest.raises(ValueError, match="stash"):
        coll.addPIE(np.ones((2, 3, 3), dtype=np.int64), stash_size=1)
    rng = np.random.default_rng(0)
    pt = sc.random_pt(rng, params, (1, 1, 3))
    cc.nb_db_load_limbs(pt, sc.random_pt(rng, params, (1, 1)), sc.random_pt(rng, params))
    idx = sc.random_ct(rng, params, (1, 1))
    with pytest.raises(P.PsiError):             # no keys at all
        cc.nb_run(idx)
    cc.InsertEvalAutomorphismKeys([5], sc.random_pt(rng, params, (1, 2)), sc.random_pt(rng, params, (1, 2)))
    with pytest.raises(P.PsiError, match="automorphism key for index"):   # OpenFHE throws on a missing key
        cc.nb_run(idx)


def test_cpp_known_answer_program_nonbatched():
    """tests/cpp/TestFHEPIE.cpp: the reference's own non-batched test program (tests/TestFHEPIE.cpp: 15000 elements,
    3 hash functions, 100 x 100 table, depth 3) re-targeted at the C++ drop-in class psi::FHEHIPPIE; "Matches" exactly
    once, healthy noise budget, and limbs identical to the oracle's FHEHIPPIE::run on the device-encoded database."""
    import os
    import subprocess
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp")
    subprocess.check_call(["make", "-C", here, "-s"])
    out = subprocess.run([os.path.join(here, "TestFHEPIE"), "check"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("Matches\n") == 1
    assert "noise budget ok: yes" in out.stdout and "limb parity with the oracle: identical" in out.stdout


def test_nb_randomised_shapes():
    """Seeded sweep over ring dimensions, limb counts and table shapes (incl. b a power of two, b = 2, K = 1, extreme
    residues 0 / q-1 in the operands): every draw is a whole collection compared limb for limb with the oracle."""
    rng = np.random.default_rng(20261019)
    for trial in range(12):
        N = int(rng.choice([256, 512, 1024, 2048, 4096]))
        L = int(rng.integers(1, 7))
        n_pie, K, b = int(rng.integers(1, 4)), int(rng.integers(1, 4)), int(rng.choice([2, 3, 4, 6, 8, 11]))
        cc, o, params = make(N, L)
        pt = sc.random_pt(rng, params, (n_pie, K, b))
        mask = sc.random_pt(rng, params, (n_pie, K))
        merge = sc.random_pt(rng, params)
        idx = sc.random_ct(rng, params, (n_pie, K))
        if trial % 3 == 0:   # extreme residues: the digit lift's centring threshold and the reductions' upper ends
            for l in range(L):
                q = int(params.q[l])
                idx[..., l, ::2] = q - 1
                idx[..., l, 1::4] = 0
                idx[..., l, 3::8] = (q - 1) // 2
                idx[..., l, 7::8] = (q - 1) // 2 + 1
                pt[..., l, ::3] = q - 1
        key_index = list(dict.fromkeys(o.eval_sum_indices(b) + [o.find_automorphism_index(-i) for i in range(1, b)]))
        key_b = sc.random_pt(rng, params, (len(key_index), L))
        key_a = sc.random_pt(rng, params, (len(key_index), L))
        cc.InsertEvalAutomorphismKeys(key_index, key_b, key_a)
        cc.nb_db_load_limbs(pt, mask, merge)
        got = cc.nb_run(idx)
        for p in range(n_pie):
            want = o.nb_run(idx[p], pt[p], merge, mask[p], key_index, key_b, key_a)
            assert np.array_equal(got[p], want), (trial, N, L, n_pie, K, b, p)
        cc.close()


@pytest.mark.parametrize("devices", [(0, 0), (0, 1, 0)])
def test_nb_collection_over_several_devices(devices):
    """psi_multi_nb_*: the PIEs of a collection sharded over a device list (a device may repeat: two contexts on one
    GPU; distinct GPUs when the box has them), one host thread per device; same ciphertexts as one context."""
    import torch
    devices = tuple(d if d < torch.cuda.device_count() else 0 for d in devices)
    params = RefParams(2048, T32, L=3).to_struct()
    o = Oracle(params)
    rng = np.random.default_rng(31)
    n_pie, K, b = 5, 2, 4
    t = int(params.t)
    slots = rng.integers(-(t // 2), t // 2, (n_pie, K, b, b + 1), dtype=np.int64)
    masks = rng.integers(1, t, (n_pie, K, b), dtype=np.int64)
    idx = sc.random_ct(rng, params, (n_pie, K))
    key_index = list(dict.fromkeys(o.eval_sum_indices(b) + [o.find_automorphism_index(-i) for i in range(1, b)]))
    key_b = sc.random_pt(rng, params, (len(key_index), 3))
    key_a = sc.random_pt(rng, params, (len(key_index), 3))
    one = P.CryptoContext(params)
    one.InsertEvalAutomorphismKeys(key_index, key_b, key_a)
    one.nb_db_encode_slots(slots, masks)
    want = one.nb_run(idx)
    pt, mask, merge = one.nb_db_get_limbs()
    assert np.array_equal(want[3], o.nb_run(idx[3], pt[3], merge, mask[3], key_index, key_b, key_a))
    mc = P.MultiContext(params, devices)
    mc.InsertEvalAutomorphismKeys(key_index, key_b, key_a)
    mc.nb_db_encode_slots(slots, masks)
    ranges = mc.nb_pie_ranges()
    assert ranges[0][0] == 0 and ranges[-1][1] == n_pie and all(a < b_ for a, b_ in ranges)
    assert np.array_equal(mc.nb_run(idx), want)
    mc.nb_db_load_limbs(pt, mask, merge)          # the limb form shards the same way
    assert np.array_equal(mc.nb_run(idx), want)
    with pytest.raises(ValueError):
        P.MultiContext(params, (0,) * 6).nb_db_encode_slots(slots, masks)   # more devices than PIEs
    mc.close()
    one.close()


@pytest.mark.parametrize("N,L,n_pie,K,b", [(1024, 3, 2, 2, 3), (2048, 4, 1, 2, 5), (16384, 4, 1, 1, 4), (512, 2, 2, 1, 2)])
def test_nb_run_hybrid_key_switching(N, L, n_pie, K, b):
    """KeySwitchTechnique HYBRID on the non-batched path (risk register): digits of alpha limbs over Q + pk,
    ApproxModDown; bit-exact against the oracle's restatement on uniformly random limbs, and a real rotation decrypts."""
    params = RefParams(N, T32, L=L, ks_technique=1).to_struct()
    cc, o = P.CryptoContext(params), Oracle(params)
    rng = np.random.default_rng(N + L)
    pt = sc.random_pt(rng, params, (n_pie, K, b))
    mask = sc.random_pt(rng, params, (n_pie, K))
    merge = sc.random_pt(rng, params)
    idx = sc.random_ct(rng, params, (n_pie, K))
    key_index = list(dict.fromkeys(o.eval_sum_indices(b) + [o.find_automorphism_index(-i) for i in range(1, b)]))
    shape = o.evk_shape()
    LE = shape[1]
    mods = [int(params.q[i]) for i in range(L)] + [int(params.pk[i]) for i in range(LE - L)]
    key_b = np.empty((len(key_index),) + shape, dtype=np.uint64)
    key_a = np.empty_like(key_b)
    for m, q in enumerate(mods):    # uniform residues of every limb of the extended basis
        key_b[:, :, m, :] = rng.integers(0, q, size=(len(key_index), shape[0], N), dtype=np.uint64)
        key_a[:, :, m, :] = rng.integers(0, q, size=(len(key_index), shape[0], N), dtype=np.uint64)
    cc.InsertEvalAutomorphismKeys(key_index, key_b, key_a)
    cc.nb_db_load_limbs(pt, mask, merge)
    got = cc.nb_run(idx)
    for p in range(n_pie):
        assert np.array_equal(got[p], o.nb_run(idx[p], pt[p], merge, mask[p], key_index, key_b, key_a)), p


def test_golden_nonbatched_case_on_the_gpu():
    """The committed fixture through the C ABI: device-encoded database == fixture plaintexts, results == fixture results."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nonbatched_case.npz"))
    cc, o, params = make(256, 2)
    cc.InsertEvalAutomorphismKeys(z["key_index"], z["key_b"], z["key_a"])
    cc.nb_db_encode_slots(z["slots"], z["mask_slots"])
    pt, mask, merge = cc.nb_db_get_limbs()
    assert np.array_equal(pt, z["pt"]) and np.array_equal(mask, z["mask"]) and np.array_equal(merge, z["merge"])
    assert np.array_equal(cc.nb_run(z["idx"]), z["out"])


def test_nb_run_many_pies_exercises_the_chunk_pipeline():
    """200 PIEs -> four chunks: index / result buffers are reused (double-buffered) while uploads, evaluations and
    downloads of consecutive chunks overlap; every PIE against the oracle."""
    cc, o, params = make(1024, 2)
    rng = np.random.default_rng(77)
    n_pie, K, b = 200, 1, 3
    pt = sc.random_pt(rng, params, (n_pie, K, b))
    mask = sc.random_pt(rng, params, (n_pie, K))
    merge = sc.random_pt(rng, params)
    idx = sc.random_ct(rng, params, (n_pie, K))
    key_index = list(dict.fromkeys(o.eval_sum_indices(b) + [o.find_automorphism_index(-i) for i in range(1, b)]))
    key_b = sc.random_pt(rng, params, (len(key_index), 2))
    key_a = sc.random_pt(rng, params, (len(key_index), 2))
    cc.InsertEvalAutomorphismKeys(key_index, key_b, key_a)
    cc.nb_db_load_limbs(pt, mask, merge)
    got = cc.nb_run(idx)
    for p in range(n_pie):
        assert np.array_equal(got[p], o.nb_run(idx[p], pt[p], merge, mask[p], key_index, key_b, key_a)), p
    assert np.array_equal(cc.nb_run(idx[37:151], 37, 151), got[37:151])   # a range that starts and ends inside chunks
