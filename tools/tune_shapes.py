#!/usr/bin/env python
"""Shape tuning on one B200: inner-product bin-block width (psi_debug_set_tuning mac_variant) and number of concurrent
phase-2 bin groups for the resident-bin counts a sharded 2^24 query leaves per GPU (47, 24, 12, 6) and the small-E
BASELINE shapes.  Prints one JSON line per (shape, setting); bit-identical results across settings are asserted."""
import json
import sys
import os

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
    shapes = [(47, 47), (24, 47), (12, 47), (6, 47), (5, 47), (14, 14), (26, 26), (2, 14), (75, 75)]
    if len(sys.argv) > 1:
        shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
    params = P.params_generate(16384, T32, 3)
    L, N, K = params.L, params.N, 2
    rng = np.random.default_rng(5)
    cc = P.CryptoContext(params)
    cc.InsertEvalMultKey(limbs(rng, params, (L,)), limbs(rng, params, (L,)))
    stream = torch.cuda.Stream()
    sp = stream.cuda_stream

    def timed(fn, steps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1) / steps

    for b, E in shapes:
        nslots = 9898
        slots = rng.integers(1, T32, (K, b, E, nslots), dtype=np.int64)
        mask_slots = rng.integers(1, T32, (b, nslots), dtype=np.int64)
        cc.db_encode_slots(slots, mask_slots)
        del slots
        cc.query_set(limbs(rng, params, (K, E, 2)), limbs(rng, params, (2,)), sp)
        cc.sync(sp)
        bytes1 = 8 * L * N * (K * b * E + 2 * K * E + 2 + 2 * K * b)
        ref = None
        for v in (2, 1, 6, 5):
            cc.set_tuning(mac_variant=v, phase2_groups=1)
            ms = timed(lambda: cc.run(sp, phases=1))
            cc.run(sp)
            got = cc.result_get(stream=sp)
            if ref is None:
                ref = got
            assert np.array_equal(ref, got), ("mac variant changes the result", b, E, v)
            print(json.dumps({"b": b, "E": E, "mac_bins_per_cta": 2 * (v if v < 3 else v - 4), "kernel": "one item per CTA" if v < 3 else "persistent, ring across items", "p1_ms": ms, "hbm_frac": bytes1 / (ms * 1e-3) / 6544e9}), flush=True)
        cc.set_tuning(mac_variant=0)
        for g in (2,):
            if g > b:
                continue
            cc.set_tuning(phase2_groups=g)
            ms2 = timed(lambda: cc.run(sp, phases=2))
            msr = timed(lambda: cc.run(sp))
            cc.run(sp)
            got = cc.result_get(stream=sp)
            assert np.array_equal(ref, got), ("phase-2 groups change the result", b, E, g)
            print(json.dumps({"b": b, "E": E, "p2_groups": g, "p2_ms": ms2, "run_ms": msr, "p2_us_per_bin": ms2 * 1e3 / b}), flush=True)
        cc.set_tuning(mac_variant=0, phase2_groups=0)


if __name__ == "__main__":
    main()
