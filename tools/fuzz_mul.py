#!/usr/bin/env python
"""Randomised parity sweep of EvalMult(ct,ct) + relinearise (fused and unfused kernels, every context variant) and of
whole run() calls against the oracle: many seeds, extreme residues (0, 1, q-1, q/2 +- 1, values that make the lazy
butterflies hit their bounds).  Not part of the test suite (minutes of oracle time); prints one JSON summary line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import psi_b200 as P  # noqa: E402
import scenario as sc  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from oracle.params_ref import RefParams  # noqa: E402

T32 = 4296540161


def extreme_ct(rng, params, mode):
    ct = sc.random_ct(rng, params)
    L, N = params.L, params.N
    for l in range(L):
        q = int(params.q[l])
        if mode == 0:
            ct[:, l, :] = q - 1
        elif mode == 1:
            ct[:, l, ::2] = q - 1
            ct[:, l, 1::2] = 0
        elif mode == 2:
            ct[:, l, :] = rng.choice(np.array([0, 1, q - 1, q // 2, q // 2 + 1, q - 2], dtype=np.uint64), size=(2, N))
        elif mode == 3:
            ct[0, l, :] = q - 1
            ct[1, l, :] = 1
        elif mode == 4:
            ct[:, l, : N // 2] = 0
    return ct


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 150.0
    t_end = time.time() + budget
    rng = np.random.default_rng(int(os.environ.get("FUZZ_SEED", "12345")))
    trials = bad = 0
    shapes = [(1024, 2), (1024, 4), (2048, 3), (4096, 4), (16384, 4), (8192, 3), (2048, 6), (1024, 7), (512, 2), (16384, 6)]
    variants = [dict(), dict(), dict(), dict(mult_technique=0), dict(ks_technique=1), dict(fp_contract=1), dict(mult_technique=0, ks_technique=1)]
    while time.time() < t_end:
        N, L = shapes[int(rng.integers(len(shapes)))]
        v = variants[int(rng.integers(len(variants)))]
        if v.get("mult_technique") == 0 and L + 1 > 8:
            continue
        params = RefParams(N, T32, depth=3, L=L, **v).to_struct()
        o, cc = Oracle(params), P.CryptoContext(params)
        sk, evk_b, evk_a = o.keygen(int(rng.integers(1 << 30)))
        cc.InsertEvalMultKey(evk_b, evk_a)
        for rep in range(6):
            m1, m2 = int(rng.integers(6)), int(rng.integers(6))
            a = extreme_ct(rng, params, m1) if m1 < 5 else sc.random_ct(rng, params)
            b = extreme_ct(rng, params, m2) if m2 < 5 else sc.random_ct(rng, params)
            ok = np.array_equal(cc.debug_mul_ctct(a, b), o.mul_ctct(a, b, evk_b, evk_a))
            trials += 1
            if not ok:
                bad += 1
                print(json.dumps({"MISMATCH": "mul_ctct", "N": N, "L": L, "variant": v, "modes": [m1, m2]}), flush=True)
        # one small run() with the same context
        K, b_, E = int(rng.choice([2, 2, 3])), int(rng.integers(1, 6)), int(rng.integers(1, 10))
        if L == 4 and N <= 4096 and rng.integers(2) == 0:
            b_, E = int(rng.integers(16, 27)), int(rng.integers(1, 4))  # >= 16 bins per launch: the four-group k_rows_relin shape
        pt, mask = sc.random_pt(rng, params, (K, b_, E)), sc.random_pt(rng, params, (b_,))
        idx, minus = sc.random_ct(rng, params, (K, E)), sc.random_ct(rng, params)
        cc.db_load_limbs(pt, mask)
        cc.query_set(idx, minus)
        cc.run()
        trials += 1
        if not np.array_equal(cc.result_get(), o.run(pt, mask, idx, minus, evk_b, evk_a, nthreads=8)):
            bad += 1
            print(json.dumps({"MISMATCH": "run", "N": N, "L": L, "variant": v, "shape": [K, b_, E]}), flush=True)
        # one non-batched collection (BV contexts only): fused key switch for N >= 1024, generic kernels below
        if not v.get("ks_technique"):
            n_pie, Kn, bn = int(rng.integers(1, 3)), int(rng.integers(1, 3)), int(rng.choice([2, 3, 4, 5, 8]))
            key_index = list(dict.fromkeys(o.eval_sum_indices(bn) + [o.find_automorphism_index(-i) for i in range(1, bn)]))
            kb, ka = sc.random_pt(rng, params, (len(key_index), L)), sc.random_pt(rng, params, (len(key_index), L))
            ptn, maskn, merge = sc.random_pt(rng, params, (n_pie, Kn, bn)), sc.random_pt(rng, params, (n_pie, Kn)), sc.random_pt(rng, params)
            idxn = sc.random_ct(rng, params, (n_pie, Kn))
            m1 = int(rng.integers(6))
            if m1 < 5:
                idxn[0, 0] = extreme_ct(rng, params, m1)
            cc.InsertEvalAutomorphismKeys(key_index, kb, ka)
            cc.nb_db_load_limbs(ptn, maskn, merge)
            got = cc.nb_run(idxn)
            trials += 1
            if not all(np.array_equal(got[p], o.nb_run(idxn[p], ptn[p], merge, maskn[p], key_index, kb, ka)) for p in range(n_pie)):
                bad += 1
                print(json.dumps({"MISMATCH": "nb_run", "N": N, "L": L, "shape": [n_pie, Kn, bn]}), flush=True)
        cc.close()
    print(json.dumps({"fuzz_trials": trials, "mismatches": bad, "seconds": budget}), flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
