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
    assert off == len(raw), "trailing bytes in the fixture"
    return fx


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
    check_oracle_against(load_fixture())


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
    fx = load_fixture(str(path))
    assert (fx["K"], fx["b"], fx["E"], fx["nslots"]) == (K, b, E, nslots)
    check_oracle_against(fx)


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
