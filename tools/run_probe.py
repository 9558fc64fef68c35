#!/usr/bin/env python
"""Whole-run() probe at config B shapes (N = 16384, L = 4, K = 2, E = 47, b bins): event-timed psi_run and a digest of the
result limbs, for A/B comparisons of run-level scheduling switches inside one gpurun call.  usage: run_probe.py [bins]"""
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
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    K, E = 2, 47
    params = P.params_generate(16384, T32, 3)
    L = params.L
    rng = np.random.default_rng(5)
    cc = P.CryptoContext(params)
    cc.InsertEvalMultKey(limbs(rng, params, (L,)), limbs(rng, params, (L,)))
    cc.db_load_limbs(limbs(rng, params, (K, b, E)), limbs(rng, params, (b,)))
    stream = torch.cuda.Stream()
    sp = stream.cuda_stream
    cc.query_set(limbs(rng, params, (K, E, 2)), limbs(rng, params, (2,)), sp)
    cc.run(sp)
    digest = hashlib.sha256(cc.result_get(stream=sp).tobytes()).hexdigest()[:16]
    best = 1e9
    for rep in range(3):
        for _ in range(5):
            cc.run(sp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(30):
            cc.run(sp)
        e1.record(stream)
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) / 30)
    print(json.dumps({"b": b, "run_ms": round(best, 4), "overlap_a": bool(os.environ.get("PSI_RUN_OVERLAP_A")), "digest": digest}))


if __name__ == "__main__":
    main()
