"""Limb-level parity against a real OpenFHE, when a fixture produced by tools/dump_openfhe_limbs.cpp is present
(tests/golden/openfhe_fixture.bin).  OpenFHE is not in this image, so the fixture cannot be generated here and
these tests are SKIPPED: parity with OpenFHE's limbs stays unpinned until someone with
an OpenFHE install runs the harness and commits the file.  The file carries OpenFHE's own parameter tables, its
operands in EVALUATION form and the ciphertexts BatchedFHEHIPPIE::run produced from them."""
import ctypes
import os

import numpy as np
import pytest

import psi_b200 as P

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "openfhe_fixture.bin")
needs_fixture = pytest.mark.skipif(not os.path.exists(FIXTURE), reason="no OpenFHE fixture (tools/dump_openfhe_limbs.cpp)")


def load_fixture(path=FIXTURE):
    raw = open(path, "rb").read()
    assert raw[:8] == b"PSIOFHE1", "not an OpenFHE fixture"
    off = 8
    (size,) = np.frombuffer(raw, dtype="<u8", count=1, offset=off)
    off += 8
    assert int(size) == ctypes.sizeof(P.capi.PsiParams), "psi_params layout differs from the harness build"
    params = P.capi.PsiParams.from_buffer_copy(raw[off:off + int(size)])
    off += int(size)
    K, b, E, nslots = (int(v) for v in np.frombuffer(raw, dtype="<u8", count=4, offset=off))
    off += 32
    L, N = params.L, params.N

    def take(shape, dtype="<u8"):
        nonlocal off
        n = int(np.prod(shape))
        a = np.frombuffer(raw, dtype=dtype, count=n, offset=off).reshape(shape).copy()
        off += 8 * n
        return a
    fx = dict(params=params, K=K, b=b, E=E, nslots=nslots)
    fx["evk_b"], fx["evk_a"] = take((L, L, N)), take((L, L, N))
    fx["pt"], fx["mask"] = take((K, b, E, L, N)), take((b, L, N))
    fx["idx"], fx["minus"] = take((K, E, 2, L, N)), take((2, L, N))
    fx["result"] = take((b, 2, L, N))
    fx["slots"], fx["mask_slots"] = take((K, b, E, nslots), "<i8"), take((b, nslots), "<i8")
    if off < len(raw):   # optional section of the non-batched path (layout: tools/dump_openfhe_limbs.cpp)
        assert raw[off:off + 8] == b"PSINB001", "trailing bytes in the fixture"
        off += 8
        n_keys, batch, n_rot = (int(v) for v in np.frombuffer(raw, dtype="<u8", count=3, offset=off))
        off += 24
        nb = dict(batch=batch, key_index=take((n_keys,)))
        nb["key_b"], nb["key_a"] = take((n_keys, L, L, N)), take((n_keys, L, L, N))
        nb["ct_in"] = take((2, L, N))
        nb["rotations"] = []
        for _ in range(n_rot):
            rot = int(take((1,), "<i8")[0])
            nb["rotations"].append((rot, int(take((1,))[0]), take((2, L, N))))
        nb["sum_index"] = take((int(take((1,))[0]),))
        nb["sum_result"] = take((2, L, N))
        fx["nonbatched"] = nb
    assert off == len(raw), "trailing bytes in the fixture"
    return fx


def check_oracle_nonbatched_against(fx):
    """The recalled details of the non-batched path against what the host library did: FindAutomorphismIndex2n, the
    EvalSum index sequence, and the limbs of EvalAtIndex / EvalSum (key switch first, permutation after)."""
    from oracle.oracle import Oracle
    nb, o = fx["nonbatched"], Oracle(fx["params"])
    slot_of = {int(g): i for i, g in enumerate(nb["key_index"])}
    for rot, g, want in nb["rotations"]:
        assert o.find_automorphism_index(rot) == g
        assert np.array_equal(o.eval_automorphism(nb["ct_in"], g, nb["key_b"][slot_of[g]], nb["key_a"][slot_of[g]]), want), rot
    assert sorted(o.eval_sum_indices(nb["batch"])) == sorted(int(g) for g in nb["sum_index"])
    q = np.array([int(fx["params"].q[l]) for l in range(fx["params"].L)], dtype=np.uint64)[None, :, None]
    ct = nb["ct_in"]
    for g in o.eval_sum_indices(nb["batch"]):
        rot = o.eval_automorphism(ct, g, nb["key_b"][slot_of[g]], nb["key_a"][slot_of[g]])
        ct = (ct + rot) % q    # residues < 2^60: the sum does not wrap
    assert np.array_equal(ct, nb["sum_result"])


def check_oracle_against(fx):
    from oracle.oracle import Oracle
    o = Oracle(fx["params"])
    # packed encoding + SetFormat(EVALUATION)
    for hf in range(fx["K"]):
        for bin_ in range(fx["b"]):
            for pos in range(fx["E"]):
                assert np.array_equal(o.encode(fx["slots"][hf, bin_, pos]), fx["pt"][hf, bin_, pos])
    # the evaluation
    got = o.run(fx["pt"], fx["mask"], fx["idx"], fx["minus"], fx["evk_b"], fx["evk_a"])
    assert np.array_equal(got, fx["result"])


@needs_fixture
def test_oracle_matches_openfhe_limbs():
    fx = load_fixture()
    check_oracle_against(fx)
    if "nonbatched" in fx:
        check_oracle_nonbatched_against(fx)


def test_fixture_format_round_trip(tmp_path):
    """The replay harness itself, on a file written in the harness' layout from ORACLE data (this checks the
    reader and the comparison, not OpenFHE parity)."""
    import scenario as sc
    from oracle.oracle import Oracle
    params = sc.make_params(256, 4296540161, L=2)
    o = Oracle(params)
    rng = np.random.default_rng(3)
    K, b, E, nslots = 2, 2, 3, 20
    slots = rng.integers(0, 1 << 32, size=(K, b, E, nslots), dtype=np.int64)
    mask_slots = rng.integers(1, 4296540161, size=(b, nslots), dtype=np.int64)
    pt, mask = sc.encode_db(o, slots), sc.encode_masks(o, mask_slots)
    sk, evk_b, evk_a = o.keygen(4)
    idx = np.stack([np.stack([o.encrypt(sk, rng.integers(0, 2, size=nslots), 10 + hf * E + pos) for pos in range(E)])
                    for hf in range(K)])
    minus = o.encrypt(sk, -rng.integers(0, 1 << 32, size=nslots), 9)
    result = o.run(pt, mask, idx, minus, evk_b, evk_a)
    path = tmp_path / "fixture.bin"
    with open(path, "wb") as f:
        f.write(b"PSIOFHE1")
        raw_params = bytes(params)
        f.write(np.uint64(len(raw_params)).tobytes())
        f.write(raw_params)
        f.write(np.array([K, b, E, nslots], dtype="<u8").tobytes())
        for a in (evk_b, evk_a, pt, mask, idx, minus, result):
            f.write(np.ascontiguousarray(a, dtype="<u8").tobytes())
        for a in (slots, mask_slots):
            f.write(np.ascontiguousarray(a, dtype="<i8").tobytes())
        # the optional non-batched section, from oracle data as well
        batch, rotations = 5, [-1, -2, -3]
        sum_index = o.eval_sum_indices(batch)
        key_index = sorted(set(sum_index + [o.find_automorphism_index(r) for r in rotations]))
        key_b, key_a = o.auto_keygen(sk, 77, key_index)
        ct_in = o.encrypt(sk, np.arange(1, nslots + 1, dtype=np.int64), 5)
        f.write(b"PSINB001")
        f.write(np.array([len(key_index), batch, len(rotations)], dtype="<u8").tobytes())
        f.write(np.array(key_index, dtype="<u8").tobytes())
        for a in (key_b, key_a, ct_in):
            f.write(np.ascontiguousarray(a, dtype="<u8").tobytes())
        q = np.array([int(params.q[l]) for l in range(params.L)], dtype=np.uint64)[None, :, None]
        for r in rotations:
            g = o.find_automorphism_index(r)
            f.write(np.int64(r).tobytes())
            f.write(np.uint64(g).tobytes())
            f.write(o.eval_automorphism(ct_in, g, key_b[key_index.index(g)], key_a[key_index.index(g)]).tobytes())
        f.write(np.uint64(len(sum_index)).tobytes())
        f.write(np.array(sum_index, dtype="<u8").tobytes())
        ct = ct_in
        for g in sum_index:
            ct = (ct + o.eval_automorphism(ct, g, key_b[key_index.index(g)], key_a[key_index.index(g)])) % q
        f.write(np.ascontiguousarray(ct, dtype="<u8").tobytes())
    fx = load_fixture(str(path))
    assert (fx["K"], fx["b"], fx["E"], fx["nslots"]) == (K, b, E, nslots)
    check_oracle_against(fx)
    assert len(fx["nonbatched"]["rotations"]) == 3
    check_oracle_nonbatched_against(fx)


@needs_fixture
@pytest.mark.gpu
def test_gpu_matches_openfhe_limbs():
    fx = load_fixture()
    cc = P.CryptoContext(fx["params"])
    cc.InsertEvalMultKey(fx["evk_b"], fx["evk_a"])
    cc.db_load_limbs(fx["pt"], fx["mask"])
    cc.query_set(fx["idx"], fx["minus"])
    cc.run()
    assert np.array_equal(cc.result_get(), fx["result"])
    # the device encoder on OpenFHE's slot values
    cc.db_encode_slots(fx["slots"], fx["mask_slots"])
    pt, mask = cc.db_get_limbs()
    assert np.array_equal(pt, fx["pt"]) and np.array_equal(mask, fx["mask"])
