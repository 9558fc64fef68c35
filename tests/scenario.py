"""Shared test scenario builder: sets -> nested cuckoo tables -> query slots -> ciphertexts.

Host objects come from the product library (psi_b200, host C++), cryptography from the ORACLE
(keygen / encrypt / decrypt are client-side operations the product does not ship)."""
import numpy as np

import psi_b200 as P
from oracle.oracle import Oracle
from oracle.params_ref import RefParams


class Scenario:
    pass


def make_params(N, t, depth=None, L=None):
    return RefParams(N, t, depth=depth, L=L).to_struct()


def table_scenario(N, t, L, k, e, K, E, b, server, client, hash_seed=987654321, key_seed=7, depth=None):
    """Full pipeline on the CPU side.  Returns a Scenario with everything a server run needs plus the
    secret key for the acceptance check."""
    s = Scenario()
    s.params = make_params(N, t, depth=depth, L=L)
    s.oracle = Oracle(s.params)
    s.k, s.e, s.K, s.E, s.b = k, e, K, E, b
    s.hash = P.TabulationHashing(hash_seed, k + K)
    s.hct = P.HierarchicalCuckooHashTable(s.hash, e, E, 0, k, K, True, True, b)
    s.hct.insertAll(np.asarray(server, dtype=np.uint64))
    s.client_cells = P.client_table(s.hash, k, e, K, np.asarray(client, dtype=np.uint64))
    s.idx_slots, s.minus_slots = P.build_query_slots(s.hash, s.client_cells, K, E)
    s.sk, s.evk_b, s.evk_a = s.oracle.keygen(key_seed)
    o = s.oracle
    s.idx = np.empty((K, E, 2, o.L, o.N), dtype=np.uint64)
    for hf in range(K):
        for pos in range(E):
            s.idx[hf, pos] = o.encrypt(s.sk, s.idx_slots[hf, pos], 1000 + hf * E + pos)
    s.minus = o.encrypt(s.sk, s.minus_slots, 999)
    return s


def oracle_db(s, shuffle_perm=None, mask_seed=11):
    """The reference constructor's transposition + encoding, done by the oracle
    (BatchedFHEHIPPIE.cpp:48-82).  Returns (slots, mask_slots, pt, mask)."""
    cells = s.hct.cells()  # [k][e][K][b][E]
    k, e, K, b, E = cells.shape
    n = k * e
    slots = np.ascontiguousarray(cells.reshape(n, K, b, E).transpose(1, 2, 3, 0)).astype(np.int64)
    rng = np.random.default_rng(mask_seed)
    mask_slots = rng.integers(1, int(s.oracle.t), size=(b, n), dtype=np.int64)
    return slots, mask_slots, encode_db(s.oracle, slots), encode_masks(s.oracle, mask_slots)


def encode_db(o, slots):
    K, b, E, n = slots.shape
    pt = np.empty((K, b, E, o.L, o.N), dtype=np.uint64)
    for hf in range(K):
        for bin_ in range(b):
            for pos in range(E):
                pt[hf, bin_, pos] = o.encode(slots[hf, bin_, pos])
    return pt


def encode_masks(o, mask_slots):
    return np.stack([o.encode(m) for m in mask_slots])


def random_ct(rng, params, shape=()):
    """Uniformly random residues with the ciphertext layout [...][2][L][N]."""
    L, N = params.L, params.N
    out = np.empty(tuple(shape) + (2, L, N), dtype=np.uint64)
    for l in range(L):
        out[..., l, :] = rng.integers(0, int(params.q[l]), size=tuple(shape) + (2, N), dtype=np.uint64)
    return out


def random_pt(rng, params, shape=()):
    L, N = params.L, params.N
    out = np.empty(tuple(shape) + (L, N), dtype=np.uint64)
    for l in range(L):
        out[..., l, :] = rng.integers(0, int(params.q[l]), size=tuple(shape) + (N,), dtype=np.uint64)
    return out


def decrypt_results(s, results):
    out = []
    budget = 1e9
    for r in results:
        slots, amb, nb = s.oracle.decrypt(s.sk, r)
        assert amb == 0
        budget = min(budget, nb)
        out.append(slots)
    return np.stack(out), budget
