"""Nested-cuckoo parameter sweep on the GPU (BASELINE.json configs[3]: bins, hash count at 2^22 server items).

For every (k = simple hash functions, e = simple table size) pair of Parameters1.txt's 1024-client rows and for
square (E = b) and rectangular (E = 2b) bins, finds the smallest bin size b for which the nested cuckoo
insertion of the whole server set succeeds with stash 0 (the batched PIE rejects a stash,
BatchedFHEHIPPIE.cpp:13) for several eviction seeds, builds the database with it on the device and times
`run()`.  The reference derives such rows offline (Performance-Evaluation/Parameters1.txt); 2^22 is not in
that file.

usage: python tools/param_sweep.py [--log2-server 22] [--seeds 3] [--steps 5] [--json out.json]
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

T32 = 4296540161


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2-server", type=int, default=22)
    ap.add_argument("--clients", type=int, default=1024)
    ap.add_argument("--seeds", type=int, default=3)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--K", type=int, default=2, help="inner cuckoo hash functions (reference default 2)")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()

    import torch
    import psi_b200 as P

    if not torch.cuda.is_available():
        raise SystemExit("param_sweep needs a CUDA device (no CPU fallback)")
    S, C, K = 1 << args.log2_server, args.clients, args.K
    # (k, e) of the rows with this client set size (Parameters1.txt:5/:11/:17 and :53/:59/:65 for C = 1024)
    simple = {32: [(2, 442), (3, 14)], 128: [(2, 1162), (3, 55)], 512: [(2, 3053), (3, 220)], 1024: [(2, 4949), (3, 443)],
              2048: [(2, 8022), (3, 890)], 4096: [(3, 1791)]}[C]
    data = P.RandomDataInput(S, C, C // 2 + 1, 123456789, 32)
    items = data.getServerSet()
    rows = []
    for k, e in simple:
        hashf = P.TabulationHashing(987654321, k + K)
        for shape, ratio in (("square", 1), ("rect", 2)):
            def fits(b):
                E = ratio * b
                cc = P.CryptoContext(P.params_generate(16384, T32, P.depth_for_E(E)))
                try:
                    for s in range(args.seeds):
                        cc.hct_build_device(hashf, k, e, K, E, b, items, evictionSeed=0x5EED + 7919 * s)
                    return True
                except RuntimeError:
                    return False
                finally:
                    del cc

            # capacity bound, then exponential + binary search for the smallest b that fits
            lo = max(1, int(math.sqrt(S / (e * K * ratio))))       # K*b*E >= S/e items per inner table
            hi = lo
            while not fits(hi):
                lo, hi = hi + 1, hi + max(2, hi // 4)
            while lo < hi:
                mid = (lo + hi) // 2
                if fits(mid):
                    hi = mid
                else:
                    lo = mid + 1
            b, E = lo, ratio * lo
            params = P.params_generate(16384, T32, P.depth_for_E(E))
            cc = P.CryptoContext(params)
            rng = np.random.default_rng(5)

            def limbs(lead):
                out = np.empty(tuple(lead) + (params.L, params.N), dtype=np.uint64)
                for l in range(params.L):
                    out[..., l, :] = rng.integers(0, int(params.q[l]), size=tuple(lead) + (params.N,), dtype=np.uint64)
                return out
            cc.InsertEvalMultKey(limbs((params.L,)), limbs((params.L,)))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            cc.db_build_from_items(hashf, k, e, K, E, b, items)
            torch.cuda.synchronize()
            t_build = time.perf_counter() - t0
            cc.query_set(limbs((K, E, 2)), limbs((2,)))
            for _ in range(2):
                cc.run()
            cc.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                cc.run()      # NULL stream = torch's default stream, where the events are recorded
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            load = S * k / float(k * e * K * b * E)
            rows.append(dict(S=S, C=C, k=k, e=e, K=K, b=b, E=E, shape=shape, slots=k * e, plaintexts=K * b * E,
                             db_gb=K * b * E * params.L * params.N * 8 / 1e9, table_load=load, build_s=t_build,
                             run_ms=ms, items_per_s=S / (ms * 1e-3), depth=P.depth_for_E(E), L=params.L))
            print(json.dumps(rows[-1]), flush=True)
            del cc
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)
    print("\n| k | e | K | bins b | positions E | shape | plaintexts | DB GB | table load | offline build s | run() ms | items/s |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        print("| %(k)d | %(e)d | %(K)d | %(b)d | %(E)d | %(shape)s | %(plaintexts)d | %(db_gb).2f | %(table_load).3f | "
              "%(build_s).2f | %(run_ms).3f | %(items_per_s).3e |" % r)


if __name__ == "__main__":
    main()
