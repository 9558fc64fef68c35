"""Timing of the non-batched FHEHIPPIE collection on one GPU next to the CPU port (oracle) — SURVEY 8f #4.
One step = psi_nb_run over a collection of P PIEs (K hash functions, b x b inner tables), host -> host.
Usage: python tools/nb_bench.py [--N 16384] [--L 4] [--pies 8] [--K 2] [--b 14] [--steps 5] [--cpu-pies 1] [--devices 0,1]
Prints one JSON line (also the per-PIE figures)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import psi_b200 as P  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from oracle.params_ref import RefParams  # noqa: E402
import scenario as sc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=16384)
    ap.add_argument("--L", type=int, default=4)
    ap.add_argument("--pies", type=int, default=8)
    ap.add_argument("--K", type=int, default=2)
    ap.add_argument("--b", type=int, default=14)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--cpu-pies", type=int, default=1)
    ap.add_argument("--devices", default="", help="comma-separated device list: the collection sharded over them (psi_multi_nb_*)")
    a = ap.parse_args()
    params = RefParams(a.N, 4296540161, L=a.L).to_struct()
    devices = [int(d) for d in a.devices.split(",") if d != ""]
    cc = P.MultiContext(params, devices) if devices else P.CryptoContext(params)
    o = Oracle(params)
    rng = np.random.default_rng(1)
    pt = sc.random_pt(rng, params, (a.pies, a.K, a.b))
    mask = sc.random_pt(rng, params, (a.pies, a.K))
    merge = sc.random_pt(rng, params)
    idx = sc.random_ct(rng, params, (a.pies, a.K))
    key_index = list(dict.fromkeys(o.eval_sum_indices(a.b) + [o.find_automorphism_index(-i) for i in range(1, a.b)]))
    key_b = sc.random_pt(rng, params, (len(key_index), a.L))
    key_a = sc.random_pt(rng, params, (len(key_index), a.L))
    cc.InsertEvalAutomorphismKeys(key_index, key_b, key_a)
    cc.nb_db_load_limbs(pt, mask, merge)
    got = cc.nb_run(idx)   # warm-up (allocations)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        got = cc.nb_run(idx)
    gpu_ms = (time.perf_counter() - t0) / a.steps * 1e3
    t0 = time.perf_counter()
    ok = True
    for p in range(min(a.cpu_pies, a.pies)):
        want = o.nb_run(idx[p], pt[p], merge, mask[p], key_index, key_b, key_a)
        ok = ok and bool(np.array_equal(got[p], want))
    cpu_ms_per_pie = (time.perf_counter() - t0) / max(1, min(a.cpu_pies, a.pies)) * 1e3
    n_sum = len(o.eval_sum_indices(a.b))
    keyswitches = a.pies * a.K * (a.b * n_sum + a.b - 1)
    print(json.dumps({
        "workload": "non-batched FHEHIPPIE collection: %d PIEs, K=%d, b=E=%d, N=%d, L=%d" % (a.pies, a.K, a.b, a.N, a.L),
        "devices": devices or [0], "gpu_ms_per_collection": gpu_ms, "gpu_ms_per_pie": gpu_ms / a.pies,
        "launches": None if devices else cc.nb_launch_count(),
        "key_switches": keyswitches, "gpu_key_switches_per_s": keyswitches / (gpu_ms * 1e-3),
        "limb_ntts": keyswitches * (a.L + a.L * a.L),
        "cpu_port_ms_per_pie_1_thread": cpu_ms_per_pie, "speedup_per_pie": cpu_ms_per_pie / (gpu_ms / a.pies),
        "parity_checked_pies": min(a.cpu_pies, a.pies), "parity_ok": ok,
        "timing": "host wall clock around psi_nb_run (pageable host buffers in and out, synchronous)"}))


if __name__ == "__main__":
    main()
