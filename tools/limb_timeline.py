import os, sys, json, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import psi_b200 as P
T32 = 4296540161
def limbs(rng, params, lead):
    out = np.empty(tuple(lead) + (params.L, params.N), dtype=np.uint64)
    for l in range(params.L):
        out[..., l, :] = rng.integers(0, int(params.q[l]), size=tuple(lead) + (params.N,), dtype=np.uint64)
    return out
b = E = 47; K = 2
params = P.params_generate(16384, T32, 3); L, N = params.L, params.N
rng = np.random.default_rng(5)
cc = P.CryptoContext(params)
cc.InsertEvalMultKey(limbs(rng, params, (L,)), limbs(rng, params, (L,)))
cc.db_encode_slots(rng.integers(1, T32, (K, b, E, 9898), dtype=np.int64), rng.integers(1, T32, (b, 9898), dtype=np.int64))
q = limbs(rng, params, (K, E, 2)).reshape(-1, N); m = limbs(rng, params, (2,)).reshape(-1, N)
iv = [q[i].copy() for i in range(q.shape[0])]; mv = [m[i].copy() for i in range(m.shape[0])]
ov = [np.empty(N, dtype=np.uint64) for _ in range(b * 2 * L)]
ai, am, ao = P.MultiContext._ptr_array(iv), P.MultiContext._ptr_array(mv), P.MultiContext._ptr_array(ov)
stream = torch.cuda.Stream(); sp = stream.cuda_stream
for nt in (8, 16):
    cc.set_host_threads(min(nt, len(os.sched_getaffinity(0))))
    for _ in range(3): cc.query_run_streamed_limbs(ai, am, ao, sp)
    t0 = time.perf_counter()
    for _ in range(10): cc.query_run_streamed_limbs(ai, am, ao, sp)
    print(json.dumps({"host_threads": nt, "cores": len(os.sched_getaffinity(0)), "ms": round((time.perf_counter() - t0) * 100, 3)}), flush=True)
    os.environ["PSI_STREAM_TIMELINE"] = "1"
    cc.query_run_streamed_limbs(ai, am, ao, sp)
    os.environ.pop("PSI_STREAM_TIMELINE")
# raw host gather rate
pool = torch.empty(q.size, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64).reshape(-1, N)
t0 = time.perf_counter()
for i, v in enumerate(iv): pool[i] = v
print("single-thread numpy gather GB/s", q.nbytes / (time.perf_counter() - t0) / 1e9)
