#!/usr/bin/env python
"""bench.py — server-side BatchedFHEPIE evaluation throughput on B200 (BASELINE.json metric).

A "step" is one client query evaluated against the whole resident server database: one pass of
BatchedFHEHIPPIE::run() (reference: BatchedFHEHIPPIE.cpp:88-129, timed there as onlineComputation,
BatchedFHEPSIServer.cpp:99-106).

  value      server items matched / s, query and database already resident in HBM, CUDA-event time of
             K run() calls on the launching stream, max over ranks
  e2e        same metric through the reference-facing call sequence with HOST buffers: pinned-host
             query -> psi_query_upload/commit (H2D) -> psi_run -> psi_result_get (D2H); serial_* = one query at a
             time (the reference's one-query-per-session case), limb_vectors = the query held as K*E*2*L separately
             allocated pageable vectors (what OpenFHE holds) through psi_query_upload_limbs / psi_result_get_limbs
  roofline   the inner-product kernel (HBM-bound): algorithmic bytes / its own event-timed duration
  cpu_baseline  the CPU oracle (restatement of the OpenFHE path; the real reference cannot be built
             here) on the host cores: warm (plaintexts already in EVALUATION form) and cold
             (plaintext transforms inside run(), what the reference's one-query-per-process run() pays)
  parity_checked_bins  bins of THIS run's results (>= 1 per rank) re-computed by the oracle on rank 0, outside the
             timed region; the run aborts if a limb differs

N > 1 (torchrun, one rank per GPU): default --scaling strong = ONE 2^24-item query whose 47 bins are sharded over the
N GPUs (BASELINE configs[2]; the same query on every rank, bins [47 r/N, 47 (r+1)/N) on rank r); the weak-scaling
figures (every GPU a full 47-bin block of an N-times larger server set) ride in the same line under "weak".

python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload NAME] [--scaling strong|weak]
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T32 = 4296540161  # --bitSize 32 (BatchedFHEPSIClient.cpp:29)

# name -> nested-cuckoo parameters (Performance-Evaluation/Parameters1.txt rows; columns C S I k e b E)
WORKLOADS = {
    # "1024 16777216 513 2 4949 47 47"  (Parameters1.txt:17)  BASELINE configs[2]
    "2^24_vs_2^10": dict(S=1 << 24, C=1 << 10, I=513, k=2, e=4949, K=2, b=47, E=47, N=16384, bits=32),
    # "1024 1048576 513 2 4949 14 14"   (Parameters1.txt:11)  BASELINE configs[1]
    "2^20_vs_2^10": dict(S=1 << 20, C=1 << 10, I=513, k=2, e=4949, K=2, b=14, E=14, N=16384, bits=32),
    # "4096 16777216 2049 3 1791 75 75" (Parameters1.txt:67)  BASELINE configs[4] (the k = 2 rows for 4096 clients,
    # e = 13004, need 26008 slots and do not fit N = 16384)
    "2^24_vs_2^12": dict(S=1 << 24, C=1 << 12, I=2049, k=3, e=1791, K=2, b=75, E=75, N=16384, bits=32),
    # BASELINE configs[3]: 2^22 is not in Parameters1.txt; b = E found by tools/param_sweep.py (smallest bin size whose
    # insertion succeeds with stash 0), see profiles/r01_param_sweep_2p22.md
    "2^22_vs_2^10": dict(S=1 << 22, C=1 << 10, I=513, k=2, e=4949, K=2, b=26, E=26, N=16384, bits=32),
    # "1024 268435456 513 2 4949 176 176" (Parameters1.txt:23): the largest server set of the file, 32 GB of plaintexts
    "2^28_vs_2^10": dict(S=1 << 28, C=1 << 10, I=513, k=2, e=4949, K=2, b=176, E=176, N=16384, bits=32),
    # BASELINE configs[0]: the reference's own CPU-runnable case, TestBatchedFHEPIE's context (t = 2^32 + 2^20 + 2^19 + 1,
    # depth 2, ring dimension chosen by the library = 8192, TestBatchedFHEPIE.cpp:14-26) at 2^16 server items vs 2^8
    # client items; no Parameters1.txt row exists for 256 clients: e interpolated, b = E by insertion success
    "2^16_vs_2^8": dict(S=1 << 16, C=1 << 8, I=129, k=2, e=1900, K=2, b=8, E=8, N=8192, bits=32, depth=2),
}
DEFAULT_WORKLOAD = "2^24_vs_2^10"
METRIC = "server items matched/sec (BatchedFHEPIE run(), ms per client query in ms_per_step)"


def phase1_bytes(L, N, K, b, E):
    """plaintext DB read once + index cts read once + minus ct + accumulators written once."""
    return 8 * L * N * (K * b * E + 2 * K * E + 2 + 2 * K * b)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--no-weak", action="store_true", help="N > 1, strong scaling: skip the extra weak-scaling measurement")
    ap.add_argument("--cpu-sample-bins", type=int, default=0, help="bins per CPU-baseline repetition (0 = all)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stage-times", action="store_true", help="time upload / run / download of the e2e step separately")
    ap.add_argument("--no-limb-leg", action="store_true", help="skip the e2e leg with separately allocated limb vectors")
    ap.add_argument("--no-nb-leg", action="store_true", help="skip the non-batched FHEHIPPIE leg (SURVEY 8f #4)")
    ap.add_argument("--gather", default="host", choices=["host", "nccl"],
                    help="N > 1 response gather inside e2e: 'host' = every rank copies its own result ciphertexts to "
                         "pinned host memory over its own PCIe link (what a one-process server does with one pinned "
                         "buffer); 'nccl' = NCCL gather to rank 0 over NVLink, then one D2H on rank 0")
    ap.add_argument("--query-dist", default="auto", choices=["auto", "host", "allgather"],
                    help="N > 1, how the query reaches every GPU inside e2e: 'host' = every rank uploads the whole query "
                         "over its own PCIe link; 'allgather' = every rank uploads 1/N of the index ciphertexts and the "
                         "slices are exchanged with one NCCL all-gather over NVLink (auto: allgather when N > 1)")
    ap.add_argument("--host-build", action="store_true",
                    help="build the nested cuckoo table on the host (OpenMP) instead of on the GPU")
    ap.add_argument("--synthetic-db", action="store_true",
                    help="random slot values instead of hashing a real server set (same shapes, same timing)")
    return ap.parse_args()


def workload_depth(w):
    """Multiplicative depth of the context: the client's rule (BatchedFHEPSIClient.cpp:46-57) unless the workload
    pins it (configs[0] uses TestBatchedFHEPIE's depth 2)."""
    if "depth" in w:
        return w["depth"]
    E = w["E"]
    return 3 if E < 500 else (5 if E < 5000 else 10)


def random_limbs(rng, params, lead):
    L, N = params.L, params.N
    out = np.empty(tuple(lead) + (L, N), dtype=np.uint64)
    for l in range(L):
        out[..., l, :] = rng.integers(0, int(params.q[l]), size=tuple(lead) + (N,), dtype=np.uint64)
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent sampling (NVML) of SM clock + clock-event reasons during the timed regions."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.active = [], set(), False
        self.stop_flag = False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while not self.stop_flag and self.nv is not None:
            if self.active:
                try:
                    mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                    try:
                        mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.samples.append(mhz)
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.005)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}



def nonbatched_leg(cc, params, steps, with_cpu):
    """SURVEY 8f #4 beside the headline path: a collection of non-batched PIEs (FHEHIPPIE.cpp:61-77) on the same context,
    host -> host through psi_nb_run, one PIE re-computed by the CPU port (parity + the CPU figure).  Shape: BASELINE
    configs[1]'s inner tables (K = 2, b = E = 14), 8 PIEs of the k*e = 9898 the reference's server would hold."""
    import psi_b200 as P
    from oracle.oracle import Oracle
    n_pie, K, b = 8, 2, 14
    L, N = params.L, params.N
    rng = np.random.default_rng(2026)
    o = Oracle(params)
    key_index = list(dict.fromkeys(o.eval_sum_indices(b) + [o.find_automorphism_index(-i) for i in range(1, b)]))
    pt = random_limbs(rng, params, (n_pie, K, b)).reshape(n_pie, K, b, L, N)
    mask = random_limbs(rng, params, (n_pie, K)).reshape(n_pie, K, L, N)
    merge = random_limbs(rng, params, (1,)).reshape(L, N)
    idx = random_limbs(rng, params, (n_pie, K, 2)).reshape(n_pie, K, 2, L, N)
    key_b = random_limbs(rng, params, (len(key_index), L)).reshape(len(key_index), L, L, N)
    key_a = random_limbs(rng, params, (len(key_index), L)).reshape(len(key_index), L, L, N)
    cc.InsertEvalAutomorphismKeys(key_index, key_b, key_a)
    cc.nb_db_load_limbs(pt, mask, merge)
    got = cc.nb_run(idx)   # warm-up: work buffers
    t0 = time.perf_counter()
    for _ in range(steps):
        got = cc.nb_run(idx)
    ms = (time.perf_counter() - t0) * 1e3 / steps
    n_sum = len(o.eval_sum_indices(b))
    ks = n_pie * K * (b * n_sum + b - 1)
    leg = {"workload": "non-batched FHEHIPPIE collection: %d PIEs, K=%d, b=E=%d, N=%d, sizeQ=%d" % (n_pie, K, b, N, L),
           "ms_per_collection": ms, "ms_per_pie": ms / n_pie, "gpu_launches_per_collection": cc.nb_launch_count(),
           "key_switches": ks, "key_switches_per_s": ks / (ms * 1e-3), "limb_ntts_per_collection": ks * (L + L * L),
           "timing": "host wall clock around psi_nb_run: pageable host index ciphertexts in, result ciphertexts out",
           "cpu_port": None, "parity_checked_pies": 0, "parity_ok": None}
    if with_cpu:
        t0 = time.perf_counter()
        want = o.nb_run(idx[0], pt[0], merge, mask[0], key_index, key_b, key_a)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        ok = bool(np.array_equal(got[0], want))
        leg.update({"cpu_port": {"ms_per_pie": cpu_ms, "cores": 1, "kind": "port",
                                 "sample": "1 of %d PIEs (the reference pins its non-batched server to one OpenMP thread, "
                                           "SimpleFHEPSIServer.cpp:17)" % n_pie},
                    "parity_checked_pies": 1, "parity_ok": ok})
        if not ok:
            raise SystemExit("bench.py: non-batched leg: GPU result limbs differ from the oracle")
    return leg


def host_threads():
    """Host cores this process may use.  torch.distributed.run exports OMP_NUM_THREADS=1 for its workers; the CPU arm
    is not a worker of the GPU job, so the variable is cleared (before libgomp loads) and the count comes from the
    affinity mask."""
    os.environ.pop("OMP_NUM_THREADS", None)
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def ref_params(w):
    """psi_params from the ORACLE's own exact generator (oracle/params_ref.py): the reference arm never maps
    libpsi_b200.so."""
    from oracle.params_ref import RefParams
    return RefParams(w["N"], T32, depth=workload_depth(w)).to_struct()


def cpu_oracle_rate(w, params, n_bins, steps, threads, cold=False):
    """Oracle (CPU restatement, oracle/psi_oracle.c) on `n_bins` bins of the workload with random limbs.
    cold: the plaintexts are still packed coefficients and are lifted + transformed inside every run (the reference's
    first run()).  Returns (items_per_s, seconds_per_step, threads, bins)."""
    from oracle.oracle import Oracle, run_cold
    o = Oracle(params)
    rng = np.random.default_rng(99)
    K, E = w["K"], w["E"]
    # bounded sample: at most ~3 GB of host plaintexts (whole query for the BASELINE configs up to 2^24 vs 2^10)
    cap = max(min(threads, n_bins), int(3.2e9 // (K * E * params.L * params.N * 8)))
    n_bins = min(n_bins, w["b"], cap)
    idx = random_limbs(rng, params, (K, E, 2))
    minus = random_limbs(rng, params, (2,))
    evk_b = random_limbs(rng, params, (params.L,))
    evk_a = random_limbs(rng, params, (params.L,))
    if cold:
        ptc = rng.integers(0, int(params.t), size=(K, n_bins, E, params.N), dtype=np.uint64)
        mc = rng.integers(0, int(params.t), size=(n_bins, params.N), dtype=np.uint64)
        fn = lambda: run_cold(o, ptc, mc, idx, minus, evk_b, evk_a, nthreads=threads)  # noqa: E731
    else:
        pt = random_limbs(rng, params, (K, n_bins, E))
        mask = random_limbs(rng, params, (n_bins,))
        fn = lambda: o.run(pt, mask, idx, minus, evk_b, evk_a, nthreads=threads)  # noqa: E731
    fn()  # warm-up (page in, twiddles hot)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / steps
    items = w["S"] * n_bins / w["b"]
    return items / dt, dt, threads, n_bins


def run_reference(args, w, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The real
    reference (OpenFHE + libscapi + Boost 1.71) cannot be built in this image, so this is the oracle
    PORT of it, all host threads.  Same workload as the GPU arm at every N: with the default strong scaling the job
    is ONE query against the 2^24-item database whatever N is; with --scaling weak it is N such databases (the port
    evaluates one and the rate is what N of them would take one after the other, i.e. the same rate)."""
    if rank != 0:
        return
    threads = host_threads()
    params = ref_params(w)
    steps = max(args.steps, 1)
    rate, dt, threads, n_bins = cpu_oracle_rate(w, params, w["b"], steps, threads)   # the whole query per step
    cold_steps = max(1, min(steps, 3))
    rate_c, dt_c, _, n_bins_c = cpu_oracle_rate(w, params, w["b"], cold_steps, threads, cold=True)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "items/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 * w["b"] / n_bins,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, w, params, world),
        "cpu_baseline": {"value": rate, "unit": "items/s", "cores": threads, "kind": "port",
                         "sample": "%d of %d bins per step, warm plaintexts (already in EVALUATION form), "
                                   "OpenMP over bins" % (n_bins, w["b"]),
                         "cold": {"value": rate_c, "unit": "items/s", "ms_per_step": dt_c * 1e3 * w["b"] / n_bins_c,
                                  "steps": cold_steps,
                                  "sample": "%d of %d bins per step; K*b*E + b plaintexts lifted and transformed inside "
                                            "run(), the reference's one-query-per-process case" % (n_bins_c, w["b"])}},
        "e2e": {"value": rate, "unit": "items/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def bind_to_gpu_numa_node(index):
    """Multi-GPU hosts: run this rank on the CPUs of its GPU's NUMA node BEFORE any pinned buffer is allocated, so
    that the staging memory of the H2D/D2H copies is local to the GPU's PCIe root (first-touch placement); eight
    ranks staging through one socket's memory share that socket's bandwidth.  Returns the node or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.lower().split(":", 1)
        node = int(open("/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:], rest)).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def workload_config(args, w, params, world):
    return {"workload": "Server/Client BFV PIE %s: S=%d server items vs C=%d client items, nested cuckoo k=%d e=%d "
                        "K=%d b=%d E=%d (Parameters1.txt), N=%d, sizeQ=%d x 60-bit, sizeP=%d, t=%d, HPSPOVERQ + BV"
                        % (args.workload, w["S"], w["C"], w["k"], w["e"], w["K"], w["b"], w["E"], params.N, params.L,
                           params.Lp, params.t),
            "bins_per_gpu": w["b"] if (args.scaling == "weak" or world == 1) else "%d/%d" % (w["b"], world),
            "server_items_total": w["S"] * (world if args.scaling == "weak" else 1),
            "sharding": "bins" if world > 1 else "none",
            "l2_policy": "inputs larger than L2 (plaintext DB %.2f GB in total streams from HBM every step)"
                         % (8.0 * params.L * params.N * w["K"] * w["b"] * w["E"] / 1e9)}


# stdout carries exactly ONE JSON line: the real stdout is kept aside and file descriptor 1 is pointed at stderr for the
# rest of the run, so whatever a library prints there (NCCL's version line under NCCL_DEBUG=VERSION) cannot precede it
_JSON_OUT = None


def emit(line):
    print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)
    if _JSON_OUT is not None:   # the line is out: hand file descriptor 1 back
        sys.stdout.flush()
        os.dup2(_JSON_OUT.fileno(), 1)


def main():
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return

    import torch
    import torch.distributed as dist
    import psi_b200 as P

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU port)")
    if world > w["b"]:
        raise SystemExit("bench.py: %d ranks for %d bins" % (world, w["b"]))
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL logs (its version line under NCCL_DEBUG=VERSION / INFO) belong on stderr: stdout carries ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sampler = ClockSampler(local_rank)
    sampler.start()

    params = P.params_generate(w["N"], T32, workload_depth(w))
    L, N, K, E = params.L, params.N, w["K"], w["E"]
    ct_words = 2 * L * N
    cc = P.CryptoContext(params, device=local_rank)
    nslots = w["k"] * w["e"]
    hashf = P.TabulationHashing(987654321, w["k"] + K)               # hashSeed (CLI.cpp:68)

    def build_db(scaling):
        """Offline phase: server set -> nested cuckoo table -> BatchedFHEHIPPIE ctor (GPU build + encode).  Returns
        (b_local, first global bin)."""
        shard = P.ShardedPIE(w["b"], rank, world)
        sharded = scaling == "strong" and world > 1
        b0, b1 = (shard.begin, shard.end) if sharded else (0, w["b"])
        rng_db = np.random.default_rng(77 + (0 if sharded else rank))
        if args.synthetic_db:
            slots = rng_db.integers(1, T32, (K, b1 - b0, E, nslots), dtype=np.int64)
            mask_slots = rng_db.integers(1, T32, (b1 - b0, nslots), dtype=np.int64)
            cc.db_encode_slots(slots, mask_slots)
            return b1 - b0, b0
        # itemSeed (CLI.cpp:67); weak scaling: every GPU its own server set, strong: ONE set for all
        seed = 123456789 + (0 if sharded else rank)
        data = P.RandomDataInput(w["S"], w["C"], w["I"], seed, w["bits"])
        if args.host_build:
            hct = P.HierarchicalCuckooHashTable(hashf, w["e"], E, 0, w["k"], K, True, True, w["b"])
            hct.insertAll(data.serverSet)
            cells = hct.cells()
            k_, e_ = cells.shape[:2]
            slots = np.ascontiguousarray(cells.reshape(k_ * e_, K, w["b"], E).transpose(1, 2, 3, 0)).astype(np.int64)
            mask_slots = np.random.default_rng(5).integers(1, T32, (w["b"], nslots), dtype=np.int64)
            cc.db_encode_slots_shard(slots, mask_slots, b0, b1)
        else:
            # table build, shuffle, transposition and encode on the device; a shard keeps its own bins of the one
            # database (all ranks: same table, same shuffle seed, same mask seed)
            cc.db_build_from_items_shard(hashf, w["k"], w["e"], K, E, w["b"], data.serverSet, b0, b1, evictionSeed=0x5EED,
                                         shuffleSeed=0xB0B0 + seed, maskSeed=0xA11CE + seed)
        return b1 - b0, b0

    t_off = time.perf_counter()
    b_local, bin0 = build_db(args.scaling)
    offline_s = time.perf_counter() - t_off

    # ---- the query: K*E + 1 ciphertexts.  Uniform residues (what BFV ciphertexts look like); the server's work does
    # not depend on their content.  Strong scaling: the SAME query on every rank (it is one query).  Held in PINNED
    # host memory for the e2e legs.
    rng = np.random.default_rng(1234 + (rank if args.scaling == "weak" else 0))
    q_host = torch.empty((K * E + 1) * ct_words, dtype=torch.int64, pin_memory=True)
    q_np = q_host.numpy().view(np.uint64)
    q_np[:K * E * ct_words] = random_limbs(rng, params, (K, E, 2)).reshape(-1)
    q_np[K * E * ct_words:] = random_limbs(rng, params, (2,)).reshape(-1)
    idx_ptr = q_host.data_ptr()
    minus_ptr = idx_ptr + K * E * ct_words * 8
    evk_b, evk_a = random_limbs(rng, params, (L,)), random_limbs(rng, params, (L,))
    cc.InsertEvalMultKey(evk_b, evk_a)

    stream = torch.cuda.Stream()
    sp = stream.cuda_stream
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms

    def timed(fn, steps):
        """CUDA-event time (ms) of `steps` calls of fn on `stream`; barrier + synchronize both sides."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        e1.synchronize()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    def measure(b_loc, with_phases):
        """All timed legs for the database currently resident (b_loc bins on this rank)."""
        m = {}
        r_host = torch.empty(b_loc * ct_words, dtype=torch.int64, pin_memory=True)
        r_host2 = torch.empty(b_loc * ct_words, dtype=torch.int64, pin_memory=True)
        cc._dims = (K, b_loc, E)
        cc.query_set_ptr(idx_ptr, minus_ptr, sp)
        cc.sync(sp)
        # ---- warm-up, then the headline: K x run() with everything resident
        for _ in range(max(args.warmup, 3)):
            cc.run(sp)
        cc.sync(sp)
        m["launches_per_run"] = cc.run_launch_count()
        m["ms_step"] = timed(lambda: cc.run(sp), args.steps) / args.steps
        if with_phases:
            # per-phase (same stream, same residency): the roofline numerator/denominator come from these
            m["ms_p1"] = timed(lambda: cc.run(sp, phases=1), args.steps) / args.steps
            m["ms_p2"] = timed(lambda: cc.run(sp, phases=2), args.steps) / args.steps

        # ---- e2e: host query in, host results out, every step
        if world > 1 and args.gather == "nccl":
            gather_bufs = [torch.empty((b_loc, ct_words), dtype=torch.int64, device="cuda") for _ in range(world)] if rank == 0 else None
            R_host = torch.empty(world * b_loc * ct_words, dtype=torch.int64, pin_memory=True) if rank == 0 else None

        def fetch_result(st, host_buf):
            """D2H of the current result buffer on torch stream `st` (N > 1, --gather nccl: NCCL gather first)."""
            if world == 1 or args.gather == "host":
                cc.result_get_ptr(host_buf.data_ptr(), st.cuda_stream)
                return
            res_dev = P.ShardedPIE.device_result_tensor(cc).view(b_loc, ct_words)
            with torch.cuda.stream(st):
                dist.gather(res_dev, gather_bufs, dst=0)
                if rank == 0:
                    for r in range(world):
                        R_host[r * b_loc * ct_words:(r + 1) * b_loc * ct_words].copy_(gather_bufs[r].view(-1), non_blocking=True)

        query_dist = args.query_dist
        if query_dist == "auto":
            query_dist = "allgather" if world > 1 and (K * E * ct_words) % world == 0 else "host"
        qd = None
        if query_dist == "allgather":
            qd = P.QueryDistributor.for_context(cc, rank, world, chunks=int(os.environ.get("PSI_QD_CHUNKS", "2")))
            q_idx_host, q_minus_host = q_host[:K * E * ct_words], q_host[K * E * ct_words:]
        m["qd"] = qd is not None

        def upload_query(st):
            """This step's query from pinned host memory into the landing buffers, on torch stream `st`."""
            if qd is None:
                cc.query_upload_ptr(idx_ptr, minus_ptr, st.cuda_stream)
            else:
                with torch.cuda.stream(st):
                    qd.distribute(q_idx_host, q_minus_host)   # 1/N over PCIe + NCCL all-gather over NVLink

        def e2e_serial_step():
            upload_query(stream)
            cc.query_commit(sp)
            cc.run(sp)
            fetch_result(stream, r_host)

        for _ in range(3):
            e2e_serial_step()
        cc.sync(sp)
        m["ms_e2e_serial"] = timed(e2e_serial_step, args.steps) / args.steps
        m["last_result"] = r_host
        # the stages of the serial step on their own (each on the launching stream): what the pipeline can overlap
        def query_in():
            upload_query(stream)
            cc.query_commit(sp)

        m["stages"] = {
            "query_in_ms": timed(query_in, args.steps) / args.steps,        # H2D (+ all-gather at N > 1) + re-tiling
            "run_ms": timed(lambda: cc.run(sp), args.steps) / args.steps,
            "result_out_ms": timed(lambda: fetch_result(stream, r_host), args.steps) / args.steps,
        } if args.stage_times else None
        # the same single query with upload slices, evaluation and download groups overlapped inside the query
        m["ms_e2e_streamed"] = None
        if world == 1:
            def e2e_streamed_step():
                cc.query_run_streamed_ptr(idx_ptr, minus_ptr, r_host2.data_ptr(), sp)
            for _ in range(3):
                e2e_streamed_step()
            cc.sync(sp)
            m["ms_e2e_streamed"] = timed(e2e_streamed_step, args.steps) / args.steps
            if not torch.equal(r_host, r_host2):
                raise SystemExit("bench.py: streamed single-query path returned different results")

        def e2e_pipelined(steps, warm):
            """Elapsed ms of `steps` queries through the 3-stream pipeline (upload of query i+1 and download of result
            i-1 overlap run i).  The `warm` untimed queries before them run through the SAME pipeline, so the timed
            region is the pipeline in steady state: it opens right before the upload of the first timed query and
            closes when the result of the last one is in host memory, i.e. it contains every H2D, kernel and D2H of
            exactly `steps` queries."""
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev_commit = [None, None]               # commit events of the two landing buffers
            ev_d2h = [None, None]
            for i in range(warm + steps):
                if ev_commit[i & 1] is not None:
                    s_in.wait_event(ev_commit[i & 1])     # landing buffer i & 1 is free once the commit of query i-2 ran
                if i == warm:
                    e0.record(s_in)
                upload_query(s_in)
                ev_up = torch.cuda.Event()
                ev_up.record(s_in)
                stream.wait_event(ev_up)
                cc.query_commit(sp)
                ev_commit[i & 1] = torch.cuda.Event()
                ev_commit[i & 1].record(stream)
                if ev_d2h[i & 1] is not None:
                    stream.wait_event(ev_d2h[i & 1])       # run i reuses the result buffer of run i-2
                cc.run(sp)
                ev_run = torch.cuda.Event()
                ev_run.record(stream)
                s_out.wait_event(ev_run)
                fetch_result(s_out, r_host if (i & 1) == 0 else r_host2)
                ev_d2h[i & 1] = torch.cuda.Event()
                ev_d2h[i & 1].record(s_out)
            e1.record(s_out)
            e1.synchronize()
            barrier()
            return max_over_ranks(e0.elapsed_time(e1))

        # untimed passes through the same pipeline first: on a fresh box the first pipelined pass runs ~7 % slower
        # (pinned pages first touched by the copy engines, PCIe link and clocks settling); measured 2.25 vs 2.10 ms
        e2e_pipelined(max(args.steps, 10), 0)
        m["ms_e2e"] = e2e_pipelined(args.steps, max(args.warmup, 3)) / args.steps
        return m

    sampler.active = True
    m = measure(b_local, True)
    sampler.active = False

    # ---- e2e with the query held the way OpenFHE holds it: K*E*2*L + 2*L separately allocated PAGEABLE limb vectors
    # in, b*2*L vectors out (psi_query_upload_limbs / psi_result_get_limbs; N = 1 only — the multi-rank equivalent is
    # the single-process psi_multi_query_set_limbs, tests/test_gpu_multi.py)
    limb_leg = None
    if world == 1 and not args.no_limb_leg:
        nvec = K * E * 2 * L
        iv = [q_np[i * N:(i + 1) * N].copy() for i in range(nvec)]
        mv = [q_np[(nvec + i) * N:(nvec + i + 1) * N].copy() for i in range(2 * L)]
        ov = [np.empty(N, dtype=np.uint64) for _ in range(b_local * 2 * L)]
        ai, am, ao = P.MultiContext._ptr_array(iv), P.MultiContext._ptr_array(mv), P.MultiContext._ptr_array(ov)
        cc.set_host_threads(min(16, max(1, len(os.sched_getaffinity(0)))))

        def limb_step():
            cc.query_upload_limbs(ai, am, sp)
            cc.query_commit(sp)
            cc.run(sp)
            cc.result_get_limbs(ao, sp)    # synchronous: results are in the vectors on return

        for _ in range(3):
            limb_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            limb_step()
        ms_limb = (time.perf_counter() - t0) * 1e3 / args.steps
        same = np.array_equal(np.concatenate(ov), m["last_result"].numpy().view(np.uint64))
        # the same vectors through psi_query_run_streamed_limbs: gather, upload slices, evaluation, download groups and
        # scatter overlapped inside the one query
        for v in ov:
            v[:] = 0

        def limb_streamed_step():
            cc.query_run_streamed_limbs(ai, am, ao, sp)

        for _ in range(3):
            limb_streamed_step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            limb_streamed_step()
        ms_limb_streamed = (time.perf_counter() - t0) * 1e3 / args.steps
        same = same and np.array_equal(np.concatenate(ov), m["last_result"].numpy().view(np.uint64))
        limb_leg = {"serial_ms_per_step": ms_limb, "vs_pinned_serial": ms_limb / m["ms_e2e_serial"],
                    "streamed_ms_per_step": ms_limb_streamed,
                    "query_vectors": nvec + 2 * L, "result_vectors": len(ov), "bytes_per_vector": N * 8,
                    "host_threads": min(16, max(1, len(os.sched_getaffinity(0)))),
                    "timing": "host wall clock around upload_limbs -> commit -> run -> result_get_limbs (returns when the "
                              "result vectors are filled)", "matches_pinned_path": bool(same)}
        if not same:
            raise SystemExit("bench.py: limb-vector e2e leg returned different results than the pinned path")

    # ---- in-run parity: rank 0 re-computes >= 1 bin of every rank with the oracle (outside the timed regions)
    from oracle.oracle import Oracle
    check_local = [0] if world > 1 else sorted({0, b_local - 1})
    res_np = m["last_result"].numpy().view(np.uint64).reshape(b_local, 2, L, N)
    packs = []
    for lb in check_local:
        pt_bin, mask_bin = cc.db_get_bin_limbs(lb)
        packs.append((bin0 + lb, pt_bin, mask_bin, res_np[lb].copy()))
    checked, bad = [], []
    if world > 1:
        flat = torch.from_numpy(np.concatenate([np.concatenate([p[1].reshape(-1), p[2].reshape(-1), p[3].reshape(-1)])
                                                for p in packs]).view(np.int64)).cuda()
        bins_t = torch.tensor([p[0] for p in packs], device="cuda", dtype=torch.int64)
        gl = [torch.empty_like(flat) for _ in range(world)] if rank == 0 else None
        gb = [torch.empty_like(bins_t) for _ in range(world)] if rank == 0 else None
        dist.gather(flat, gl, dst=0)
        dist.gather(bins_t, gb, dst=0)
        if rank == 0:
            packs = []
            n_pt, n_m, n_r = K * E * L * N, L * N, 2 * L * N
            for r in range(world):
                a = gl[r].cpu().numpy().view(np.uint64)
                for i, gbin in enumerate(gb[r].cpu().tolist()):
                    o0 = i * (n_pt + n_m + n_r)
                    packs.append((gbin, a[o0:o0 + n_pt].reshape(K, E, L, N), a[o0 + n_pt:o0 + n_pt + n_m].reshape(L, N),
                                  a[o0 + n_pt + n_m:o0 + n_pt + n_m + n_r].reshape(2, L, N)))
    if rank == 0:
        o = Oracle(params)
        idx_np = q_np[:K * E * ct_words].reshape(K, E, 2, L, N)
        minus_np = q_np[K * E * ct_words:].reshape(2, L, N)
        for gbin, pt_bin, mask_bin, got in packs:
            want = o.run(np.ascontiguousarray(pt_bin[:, None]), np.ascontiguousarray(mask_bin[None]), idx_np, minus_np,
                         evk_b, evk_a, nthreads=min(8, host_threads()))
            (checked if np.array_equal(want[0], got) else bad).append(int(gbin))
        if bad:
            raise SystemExit("bench.py: PARITY FAILURE against the oracle in bins %s" % bad)

    # ---- N > 1, strong: the weak-scaling figures in the same line (every GPU a full block of b bins)
    weak = None
    if world > 1 and args.scaling == "strong" and not args.no_weak:
        b_w, _ = build_db("weak")
        mw = measure(b_w, False)
        weak = {"value": w["S"] * world / (mw["ms_step"] * 1e-3), "unit": "items/s", "ms_per_step": mw["ms_step"],
                "server_items_total": w["S"] * world, "bins_per_gpu": b_w,
                "e2e_ms_per_step": mw["ms_e2e"], "e2e_serial_ms_per_step": mw["ms_e2e_serial"],
                "e2e_value": w["S"] * world / (mw["ms_e2e"] * 1e-3),
                "note": "independent b-bin replicas, no data-path collective: near-linear by construction"}
    sampler.stop_flag = True

    ms_step, ms_p1, ms_p2 = m["ms_step"], m["ms_p1"], m["ms_p2"]
    total_items = w["S"] * (world if args.scaling == "weak" else 1)
    value = total_items / (ms_step * 1e-3)
    e2e_value = total_items / (m["ms_e2e"] * 1e-3)
    # whole-job H2D per step: the query crosses PCIe once per GPU ('host') or once in total plus the small minus
    # ciphertext per GPU ('allgather')
    h2d = (K * E + 1) * ct_words * 8 * world if not m["qd"] else (K * E + world) * ct_words * 8
    d2h = (w["b"] * (world if args.scaling == "weak" else 1)) * ct_words * 8  # every result ciphertext reaches host memory

    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_hbm, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak_hbm, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # ncu-measured DRAM traffic of the dominant kernel for exactly this workload (one launch), from the committed
    # capture under profiles/ (static: not re-measured in this run)
    traffic = None
    for name in ("r02_dram_traffic.json", "r01_dram_traffic.json"):
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", name))).get(args.workload)
            if tr and b_local == w["b"]:
                traffic = tr["dram_bytes_per_launch"].get("k_mac_tma")
                break
        except Exception:
            pass
    bytes1 = phase1_bytes(L, N, K, b_local, E)
    ach = bytes1 / (ms_p1 * 1e-3) / 1e9
    roofline = {"kernel": "k_mac_tma (inner product, phase 1)", "bound": "hbm", "achieved": ach, "peak": peak_hbm,
                "unit": "GB/s", "frac": ach / peak_hbm, "traffic": traffic,
                "traffic_source": "static: ncu --set full capture committed under profiles/" if traffic else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes1, "launch_ms": ms_p1}
    # phase 2 against the integer pipe: butterflies of the 88-NTT HPSPOVERQ + BV pipeline per second vs the
    # measured register-resident butterfly rate of this GPU (psi_bench_pipe_peak, same process, same clocks)
    import ctypes
    Lp = params.Lp
    ntt_per_mul = 4 * L + 2 * Lp + 2 * (L + Lp) + 3 * (L + Lp) + 2 * L + L * L
    logN = N.bit_length() - 1
    butterflies = b_local * (K - 1) * ntt_per_mul * (N // 2) * logN
    bf_peak, imad_peak = ctypes.c_double(), ctypes.c_double()
    P.capi.check(P.lib().psi_bench_pipe_peak(local_rank, 1, ctypes.byref(bf_peak)))
    P.capi.check(P.lib().psi_bench_pipe_peak(local_rank, 0, ctypes.byref(imad_peak)))
    roofline_int = {"kernels": "k_rows_inv + k_cols_extend + k_rows_tensor + k_cols_scale + k_rows_relin (phase 2)",
                    "bound": "integer pipe", "achieved": butterflies / (ms_p2 * 1e-3), "peak": bf_peak.value,
                    "unit": "butterflies/s", "frac": butterflies / (ms_p2 * 1e-3) / bf_peak.value,
                    "butterflies_per_launch_set": butterflies, "ntt_per_ct_mult": ntt_per_mul,
                    "imad_wide_peak_per_s": imad_peak.value,
                    "peak_source": "measured in this run: 64-bit Harvey/Shoup butterflies, operands in registers"}
    phases = {"inner_product_ms": ms_p1, "multiply_relin_mask_ms": ms_p2, "run_ms": ms_step,
              "launches_per_run": m["launches_per_run"]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        t0 = time.perf_counter()
        rate, dt, threads, nb = cpu_oracle_rate(w, params, args.cpu_sample_bins or w["b"], 3, threads)
        cpu = {"value": rate, "unit": "items/s", "cores": threads, "kind": "port",
               "sample": "%d of %d bins x 3 repetitions (%.1f s of CPU work), warm plaintexts, OpenMP over bins; "
                         "port = oracle/psi_oracle.c, the real reference (OpenFHE) cannot be built in this image"
                         % (nb, w["b"], time.perf_counter() - t0),
               "ms_per_query_extrapolated": dt * 1e3 * w["b"] / nb}
        # the same port on ONE host thread, two bins (SURVEY 8d asks for 1 thread and all threads)
        rate1, dt1, _, nb1 = cpu_oracle_rate(w, params, 2, 1, 1)
        cpu["single_thread"] = {"value": rate1, "unit": "items/s", "cores": 1, "sample": "%d of %d bins" % (nb1, w["b"]),
                                "ms_per_query_extrapolated": dt1 * 1e3 * w["b"] / nb1}
        # cold: plaintext lifts + transforms inside run() (the reference's first run())
        ratec, dtc, _, nbc = cpu_oracle_rate(w, params, args.cpu_sample_bins or w["b"], 1, threads, cold=True)
        cpu["cold"] = {"value": ratec, "unit": "items/s", "cores": threads, "sample": "%d of %d bins x 1" % (nbc, w["b"]),
                       "ms_per_query_extrapolated": dtc * 1e3 * w["b"] / nbc}

    nb_leg = None
    if rank == 0 and world == 1 and not args.no_nb_leg:
        nb_leg = nonbatched_leg(cc, params, min(args.steps, 5), not args.no_cpu_baseline)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "items/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(args, w, params, world),
            "roofline": roofline, "roofline_int": roofline_int, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "items/s", "ms_per_step": m["ms_e2e"], "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "serial_ms_per_step": m["ms_e2e_serial"], "stages": m.get("stages"),
                    "serial_streamed_ms_per_step": m["ms_e2e_streamed"],
                    "serial_value": total_items / (m["ms_e2e_serial"] * 1e-3),
                    "limb_vectors": limb_leg,
                    "path": "pinned host query -> psi_query_upload | psi_query_commit -> psi_run -> psi_result_get -> "
                            "pinned host, three streams: upload of query i+1 and download of result i-1 overlap "
                            "run i; the warm-up queries run through the same pipeline, the timed region holds every copy "
                            "and kernel of exactly `steps` queries (serial_* = one query at a time; serial_streamed = one query at a "
                            "time through psi_query_run_streamed: upload slices, partial inner products, bin groups and their "
                            "downloads overlapped inside the query)"
                            + ((" (N > 1: NCCL gather to rank 0 over NVLink, then one D2H)" if args.gather == "nccl" else
                                " (N > 1: every rank downloads its own bins over its own PCIe link)") if world > 1 else ""),
                    "gather": args.gather if world > 1 else None, "rank0_numa_node": numa_node,
                    "query_dist": ("1/N of the index ciphertexts per GPU over PCIe + NCCL all-gather over NVLink"
                                   if m["qd"] else "whole query over every GPU's own PCIe link") if world > 1 else None},
            "gpu_launches": m["launches_per_run"] * args.steps,
            "parity_checked_bins": len(checked), "parity_checked_bin_ids": checked, "parity_ok": not bad,
            "weak": weak, "nonbatched": nb_leg,
            "phases": phases, "clocks": sampler.summary(), "offline_build_s": offline_s,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
