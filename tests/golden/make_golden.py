"""Writes tests/golden/oracle_digests.json, tests/golden/small_case.npz and tests/golden/nonbatched_case.npz.

The reference holds no golden ciphertext vectors for this path (keys, noise, shuffle and masks are
random per run), and OpenFHE is not available to generate any.  These fixtures are therefore
outputs of THIS repo's oracle for fixed seeds; they pin the checker against silent drift and give
the GPU tests a committed input/output pair that does not depend on rebuilding the oracle.
Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

T32 = 4296540161


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def small_case():
    """K=2, b=3, E=4, N=256, L=2: inputs and the oracle's result limbs."""
    from oracle.oracle import Oracle
    from oracle.params_ref import RefParams
    import scenario as sc
    params = RefParams(256, T32, L=2).to_struct()
    o = Oracle(params)
    rng = np.random.default_rng(20261018)
    sk, evk_b, evk_a = o.keygen(1)
    K, b, E = 2, 3, 4
    slots = rng.integers(-(T32 // 2), T32 // 2, (K, b, E, 200), dtype=np.int64)
    mask_slots = rng.integers(1, T32, (b, 200), dtype=np.int64)
    pt = sc.encode_db(o, slots)
    mask = sc.encode_masks(o, mask_slots)
    idx = sc.random_ct(rng, params, (K, E))
    minus = sc.random_ct(rng, params)
    out = o.run(pt, mask, idx, minus, evk_b, evk_a)
    return dict(slots=slots, mask_slots=mask_slots, pt=pt, mask=mask, idx=idx, minus=minus, evk_b=evk_b,
                evk_a=evk_a, out=out)


def nonbatched_case():
    """Non-batched FHEHIPPIE (orc_nb_run): 2 PIEs, K=2, b=E=3, N=256, L=2, BV keys for EvalSum(3) and rotations -1, -2."""
    from oracle.oracle import Oracle
    from oracle.params_ref import RefParams
    import scenario as sc
    params = RefParams(256, T32, L=2).to_struct()
    o = Oracle(params)
    rng = np.random.default_rng(20261019)
    n_pie, K, b = 2, 2, 3
    sk, _, _ = o.keygen(4)
    slots = rng.integers(-(T32 // 2), T32 // 2, (n_pie, K, b, b + 1), dtype=np.int64)
    mask_slots = rng.integers(1, T32, (n_pie, K, b), dtype=np.int64)
    pt = np.stack([[[o.encode(slots[p, hf, bin_]) for bin_ in range(b)] for hf in range(K)] for p in range(n_pie)])
    mask = np.stack([[o.encode(mask_slots[p, hf]) for hf in range(K)] for p in range(n_pie)])
    merge = o.encode(np.array([1], dtype=np.int64))
    key_index = np.array(list(dict.fromkeys(o.eval_sum_indices(b) + [o.find_automorphism_index(-i) for i in range(1, b)])),
                         dtype=np.uint64)
    key_b, key_a = o.auto_keygen(sk, 99, [int(g) for g in key_index])
    idx = sc.random_ct(rng, params, (n_pie, K))
    out = np.stack([o.nb_run(idx[p], pt[p], merge, mask[p], key_index, key_b, key_a) for p in range(n_pie)])
    return dict(slots=slots, mask_slots=mask_slots, pt=pt, mask=mask, merge=merge, key_index=key_index, key_b=key_b, key_a=key_a,
                idx=idx, out=out)


def compute_digests():
    from oracle.oracle import Oracle
    from oracle.params_ref import RefParams
    import scenario as sc
    d = {}
    for N, L in ((256, 2), (1024, 3)):
        params = RefParams(N, T32, L=L).to_struct()
        o = Oracle(params)
        rng = np.random.default_rng(N + L)
        tag = "N%d_L%d" % (N, L)
        d[tag + "_q"] = [int(params.q[i]) for i in range(L)]
        d[tag + "_p"] = [int(params.p[i]) for i in range(params.Lp)]
        a = rng.integers(0, int(params.q[0]), N, dtype=np.uint64)
        d[tag + "_ntt"] = _digest(o.ntt(a, 0))
        d[tag + "_encode"] = _digest(o.encode(rng.integers(-1000, 1000, N // 2, dtype=np.int64)))
        sk, evk_b, evk_a = o.keygen(3)
        d[tag + "_keygen"] = _digest(np.concatenate([sk.ravel(), evk_b.ravel(), evk_a.ravel()]))
        ct1, ct2 = sc.random_ct(rng, params), sc.random_ct(rng, params)
        d[tag + "_mul_core"] = _digest(o.mul_core(ct1, ct2))
        d[tag + "_mul_ctct"] = _digest(o.mul_ctct(ct1, ct2, evk_b, evk_a))
    d["small_case_out"] = _digest(small_case()["out"])
    d["nonbatched_case_out"] = _digest(nonbatched_case()["out"])
    return d


if __name__ == "__main__":
    with open(os.path.join(HERE, "oracle_digests.json"), "w") as f:
        json.dump(compute_digests(), f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "small_case.npz"), **small_case())
    np.savez_compressed(os.path.join(HERE, "nonbatched_case.npz"), **nonbatched_case())
    print("written")
