#!/usr/bin/env python
"""psi_query_run_streamed at config B: one query host -> host per setting of the upload slicing (PSI_STREAM_SLICES) and
the download group cuts (PSI_STREAM_CUTS), next to the unoverlapped serial path.  One JSON line per setting."""
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
    b = E = 47
    K = 2
    params = P.params_generate(16384, T32, 3)
    L, N = params.L, params.N
    rng = np.random.default_rng(5)
    cc = P.CryptoContext(params)
    cc.InsertEvalMultKey(limbs(rng, params, (L,)), limbs(rng, params, (L,)))
    nslots = 9898
    cc.db_encode_slots(rng.integers(1, T32, (K, b, E, nslots), dtype=np.int64), rng.integers(1, T32, (b, nslots), dtype=np.int64))
    ctw = 2 * L * N
    q_host = torch.empty((K * E + 1) * ctw, dtype=torch.int64, pin_memory=True)
    q_np = q_host.numpy().view(np.uint64)
    q_np[:K * E * ctw] = limbs(rng, params, (K, E, 2)).reshape(-1)
    q_np[K * E * ctw:] = limbs(rng, params, (2,)).reshape(-1)
    r_host = torch.empty(b * ctw, dtype=torch.int64, pin_memory=True)
    r_ref = torch.empty(b * ctw, dtype=torch.int64, pin_memory=True)
    idx_ptr = q_host.data_ptr()
    minus_ptr = idx_ptr + K * E * ctw * 8
    stream = torch.cuda.Stream()
    sp = stream.cuda_stream

    def timed(fn, steps=20):
        for _ in range(3):
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

    def serial():
        cc.query_upload_ptr(idx_ptr, minus_ptr, sp)
        cc.query_commit(sp)
        cc.run(sp)
        cc.result_get_ptr(r_ref.data_ptr(), sp)

    print(json.dumps({"setting": "serial (upload, run, download one after another)", "ms": round(timed(serial), 4)}), flush=True)
    torch.cuda.synchronize()
    settings = [(4, None), (6, None)]
    for cuts in ("0.25,0.50,0.75", "0.10,0.45,0.85", "0.08,0.40,0.80", "0.15,0.50,0.85", "0.12,0.42,0.78", "0.06,0.30,0.70", "0.10,0.35,0.70"):
        settings.append((4, cuts))
    os.environ["PSI_STREAM_TIMELINE"] = "1"
    for seq in (True, False):
        if seq:
            os.environ.pop("PSI_STREAM_CONCURRENT", None)
        else:
            os.environ["PSI_STREAM_CONCURRENT"] = "1"
        for _ in range(3):
            cc.query_run_streamed_ptr(idx_ptr, minus_ptr, r_host.data_ptr(), sp)
            torch.cuda.synchronize()
        ms = timed(lambda: cc.query_run_streamed_ptr(idx_ptr, minus_ptr, r_host.data_ptr(), sp)) if False else None
    os.environ.pop("PSI_STREAM_TIMELINE")
    os.environ.pop("PSI_STREAM_CONCURRENT", None)
    print(json.dumps({"setting": "groups one after another (default cuts)", "ms": round(timed(lambda: cc.query_run_streamed_ptr(idx_ptr, minus_ptr, r_host.data_ptr(), sp)), 4)}), flush=True)
    if "--timeline" in sys.argv:
        return
    os.environ["PSI_STREAM_CONCURRENT"] = "1"   # the sweep below: concurrent prioritised groups
    for sl, cuts in settings:
        os.environ["PSI_STREAM_SLICES"] = str(sl)
        if cuts:
            os.environ["PSI_STREAM_CUTS"] = cuts
        else:
            os.environ.pop("PSI_STREAM_CUTS", None)
        ms = timed(lambda: cc.query_run_streamed_ptr(idx_ptr, minus_ptr, r_host.data_ptr(), sp))
        torch.cuda.synchronize()
        assert torch.equal(r_host, r_ref), (sl, cuts)
        print(json.dumps({"slices_per_hf": sl, "cuts": cuts or "0.38,0.72,0.91", "ms": round(ms, 4)}), flush=True)


if __name__ == "__main__":
    main()
