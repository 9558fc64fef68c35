#!/usr/bin/env python
"""Single-process multi-device evaluator (psi_multi_*, the reference server's one-process shape,
BatchedFHEPSIServer.cpp:99-108) at config B: one query host -> host over device lists [0], [0,1], ... of ONE process.
  run_ms   psi_multi_run alone (query resident): host wall clock around run + psi_multi_sync, best of 3 x 20
  e2e_ms   psi_multi_query_set (pinned host query; 1/G per device over its own PCIe link + peer copies)
           -> psi_multi_run -> psi_multi_result_get (every device writes its bins into the one pinned result buffer)
           -> psi_multi_sync
Every device list must return the same limbs as the single device.  One JSON line per device list."""
import json
import os
import sys
import time

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
    ndev = torch.cuda.device_count()
    lists = [list(range(n)) for n in (1, 2, 4, 8) if n <= ndev]
    b = E = 47
    K, nslots = 2, 9898
    params = P.params_generate(16384, T32, 3)
    L, N = params.L, params.N
    rng = np.random.default_rng(5)
    evk = limbs(rng, params, (L,)), limbs(rng, params, (L,))
    slots = rng.integers(1, T32, (K, b, E, nslots), dtype=np.int64)
    mask_slots = rng.integers(1, T32, (b, nslots), dtype=np.int64)
    ctw = 2 * L * N
    q_host = torch.empty((K * E + 1) * ctw, dtype=torch.int64, pin_memory=True)
    q_np = q_host.numpy().view(np.uint64)
    q_np[:K * E * ctw] = limbs(rng, params, (K, E, 2)).reshape(-1)
    q_np[K * E * ctw:] = limbs(rng, params, (2,)).reshape(-1)
    idx_ptr = q_host.data_ptr()
    minus_ptr = idx_ptr + K * E * ctw * 8
    r_host = torch.empty(b * ctw, dtype=torch.int64, pin_memory=True)
    ref = None
    for devs in lists:
        mc = P.MultiContext(params, devs)
        mc.InsertEvalMultKey(*evk)
        mc.db_encode_slots(slots, mask_slots)

        def e2e():
            mc.query_set_ptr(idx_ptr, minus_ptr)
            mc.run()
            mc.result_get_ptr(r_host.data_ptr())
            mc.sync()

        def run_only():
            mc.run()
            mc.sync()

        def wall(fn, steps=20):
            for _ in range(3):
                fn()
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                for _ in range(steps):
                    fn()
                best = min(best, (time.perf_counter() - t0) * 1e3 / steps)
            return best

        e2e()
        got = r_host.numpy().copy()
        if ref is None:
            ref = got
        assert np.array_equal(ref, got), devs
        print(json.dumps({"devices": devs, "bins": mc.bin_ranges(), "run_ms": round(wall(run_only), 4), "e2e_ms": round(wall(e2e), 4),
                          "launches_per_run": mc.run_launch_count()}), flush=True)
        mc.close()


if __name__ == "__main__":
    main()
