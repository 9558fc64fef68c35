#!/usr/bin/env python
"""run() and phase-2 time against the number of concurrent bin groups, with the launch set replayed as a CUDA graph
(default) or launched directly (PSI_NO_GRAPH=1), for the resident-bin counts a sharded 2^24 query leaves per GPU.
One JSON line per (bins, graph, groups); results asserted bit-identical across settings."""
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
    bins = [int(a) for a in sys.argv[1:]] or [6, 12, 24, 47]
    E, K = 47, 2
    params = P.params_generate(16384, T32, 3)
    L = params.L
    rng = np.random.default_rng(5)
    evk = limbs(rng, params, (L,)), limbs(rng, params, (L,))
    idx, minus = limbs(rng, params, (K, E, 2)), limbs(rng, params, (2,))
    stream = torch.cuda.Stream()
    sp = stream.cuda_stream

    def timed(fn, steps=40):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) / steps)
        return best

    for b in bins:
        pt = limbs(rng, params, (K, b, E))
        mask = limbs(rng, params, (b,))
        ref = None
        for graph in (True, False):
            if graph:
                os.environ.pop("PSI_NO_GRAPH", None)
            else:
                os.environ["PSI_NO_GRAPH"] = "1"
            cc = P.CryptoContext(params)
            cc.InsertEvalMultKey(*evk)
            cc.db_load_limbs(pt, mask)
            cc.query_set(idx, minus, sp)
            for g in (1, 2, 3, 4, 6, 8):
                if g > b:
                    continue
                cc.set_tuning(phase2_groups=g)
                cc.run(sp)
                got = cc.result_get(stream=sp)
                if ref is None:
                    ref = got
                assert np.array_equal(ref, got), (b, graph, g)
                p2 = timed(lambda: cc.run(sp, phases=2))
                run = timed(lambda: cc.run(sp))
                print(json.dumps({"b": b, "graph": graph, "p2_groups": g, "p2_ms": round(p2, 4), "run_ms": round(run, 4),
                                  "launches": cc.run_launch_count()}), flush=True)
            cc.close()


if __name__ == "__main__":
    main()
