#!/usr/bin/env python
"""Phase-2 probe at config B shapes (N = 16384, L = 4, b bins): prints the event-timed phase-2 time of the library
selected by PSI_B200_LIB; --once runs two evaluations only (for an ncu launch list around it); --digest prints a
checksum of the result limbs so that variants can be compared for bit-identity."""
import hashlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import psi_b200 as P  # noqa: E402

T32 = 4296540161


def limbs(rng, params, lead):
    out = np.empty(tuple(lead) + (params.L, params.N), dtype=np.uint64)
    for l in range(params.L):
        out[..., l, :] = rng.integers(0, int(params.q[l]), size=tuple(lead) + (params.N,), dtype=np.uint64)
    return out


def main():
    once = "--once" in sys.argv
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    b = int(args[0]) if args else 47
    groups = int(args[1]) if len(args) > 1 else 0
    E, K = 2, 2
    params = P.params_generate(16384, T32, 3, int(os.environ.get("P2_L", "0")))  # P2_L: override sizeQ (0 = depth 3's)
    L = params.L
    rng = np.random.default_rng(5)
    cc = P.CryptoContext(params)
    cc.InsertEvalMultKey(limbs(rng, params, (L,)), limbs(rng, params, (L,)))
    cc.db_load_limbs(limbs(rng, params, (K, b, E)), limbs(rng, params, (b,)))
    stream = torch.cuda.Stream()
    sp = stream.cuda_stream
    cc.query_set(limbs(rng, params, (K, E, 2)), limbs(rng, params, (2,)), sp)
    if groups:
        cc.set_tuning(phase2_groups=groups)
    cc.run(sp)
    got = cc.result_get(stream=sp)
    digest = hashlib.sha256(got.tobytes()).hexdigest()[:16]
    if once:
        cc.run(sp, phases=2)
        cc.sync(sp)
        print(json.dumps({"lib": os.environ.get("PSI_B200_LIB", "default"), "digest": digest}))
        return
    best = 1e9
    for rep in range(3):
        for _ in range(5):
            cc.run(sp, phases=2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(30):
            cc.run(sp, phases=2)
        e1.record(stream)
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) / 30)
    print(json.dumps({"lib": os.environ.get("PSI_B200_LIB", "default"), "b": b, "groups": groups, "p2_ms": round(best, 4),
                      "digest": digest}), flush=True)


if __name__ == "__main__":
    main()
