"""The product's C++ parameter generator (host/params.cpp) against the oracle's exact big-integer
definitions (oracle/params_ref.py), table by table, plus the properties the kernels rely on."""
import pytest

import psi_b200 as P
from oracle.params_ref import RefParams, struct_to_dict, PLAINTEXT_MODULUS, depth_for_E, is_prime

T32 = 4296540161


@pytest.mark.parametrize("N,t,depth", [(16384, T32, 3), (16384, T32, 5), (8192, T32, 2), (16384, 65537, 3),
                                       (16384, 1099579260929, 3), (16384, 281474981953537, 3), (1024, T32, 2)])
def test_generators_agree(N, t, depth):
    got = struct_to_dict(P.params_generate(N, t, depth))
    want = struct_to_dict(RefParams(N, t, depth).to_struct())
    for key in want:
        assert got[key] == want[key], key


@pytest.mark.parametrize("depth,L", [(3, 0), (2, 0), (5, 0), (1, 1), (3, 5)])
def test_generators_agree_on_variants(depth, L):
    """HPS tables, HYBRID digit count / special primes, fp mode: C++ generator == exact Python definitions."""
    got = struct_to_dict(P.params_generate(16384, T32, depth, L, mult_technique=0, ks_technique=1, fp_contract=1))
    want = struct_to_dict(RefParams(16384, T32, depth, L=L or None, mult_technique=0, ks_technique=1, fp_contract=1).to_struct())
    assert got == want
    assert got["mult_technique"] == 0 and got["ks_technique"] == 1 and got["fp_contract"] == 1
    assert got["ks_num_parts"] * got["Lk"] >= got["L"] > (got["ks_num_parts"] - 1) * got["Lk"]
    mods = got["q"] + got["p"] + got["pk"]
    assert len(set(mods)) == len(mods) and mods == sorted(mods, reverse=True)


def test_moduli_properties():
    p = P.params_generate(16384, T32, 3)
    mods = list(p.q[:p.L]) + list(p.p[:p.Lp])
    assert len(set(mods)) == len(mods)
    for q in mods:
        assert is_prime(q) and q % (2 * 16384) == 1 and (1 << 59) < q < (1 << 60)
    assert mods == sorted(mods, reverse=True)
    for q, psi in zip(mods, list(p.psi_q[:p.L]) + list(p.psi_p[:p.Lp])):
        assert pow(psi, 16384, q) == q - 1
    assert pow(p.psi_t, 16384, p.t) == p.t - 1


def test_client_parameter_rules():
    """BatchedFHEPSIClient.cpp:23-57."""
    assert P.PLAINTEXT_MODULUS == PLAINTEXT_MODULUS
    assert P.PLAINTEXT_MODULUS[32] == (1 << 32) + (1 << 20) + (1 << 19) + 1
    for E in (1, 47, 499, 500, 4999, 5000, 10**6):
        assert P.depth_for_E(E) == depth_for_E(E)
    assert [P.depth_for_E(E) for E in (499, 500, 4999, 5000)] == [3, 5, 5, 10]
    for bits, t in P.PLAINTEXT_MODULUS.items():
        assert is_prime(t) and t % 32768 == 1 and t.bit_length() == bits + 1
