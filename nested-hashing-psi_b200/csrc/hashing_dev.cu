// Nested (hierarchical) cuckoo table build on the device — SURVEY.md 8(f) "next" #3.
//
// Replaces, for the batched FHE path, the offline loop of the reference
//   HierarchicalCuckooHashTable::insertAll      src/Common/Hashing/HierarchicalCuckooHashTable.cpp:55-73
//     generateSimpleHashTable                   src/Common/Hashing/HashUtils.cpp:48-60
//     CuckooHashTable::insert / lookUp          src/Common/Hashing/CuckooHashTable.cpp:72-114,135-167
//     TabulationHashing::hashWithIndicator      src/Common/Hashing/TabulationHashing.cpp:45-54
// and is BIT-IDENTICAL to this repo's host build (host/hashing.cpp) for the same seeds: same insertion
// order inside every inner table (stable bucketing), same first-free-bin rule, same random-walk eviction
// (std::mt19937 per inner table, two engine words per draw, low word first).
//
// Mapping: the k*e inner cuckoo tables are independent (the reference's `#pragma omp parallel for`, :65);
// each is built by ONE WARP: lanes probe the b bins of a position in parallel (ballot), lane 0 places,
// evicts and runs the Mersenne twister (state in global scratch, initialised only if an eviction happens).
// Tables live in global memory in the psi_hct_get_cells layout [st][p][hf][bin][pos].
#include <cub/cub.cuh>

#include "psi_kernels.cuh"

namespace psi {

// ---- tabulation hashing ---------------------------------------------------------------------------
// T: [n_hf][16][256] u64.  64-bit items feed bytes 0..7; chunks 8..15 always index entry 0.
__device__ __forceinline__ u64 tab_hash(const u64* __restrict__ T, uint32_t hf, u64 x) {
    const u64* t = T + (size_t)hf * 16 * 256;
    u64 h = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) h ^= __ldg(t + i * 256 + ((x >> (8 * i)) & 0xff));
#pragma unroll
    for (int i = 8; i < 16; i++) h ^= __ldg(t + i * 256);
    return h;
}

__global__ void __launch_bounds__(256) k_outer_bucket(const u64* __restrict__ T, uint32_t hf, uint32_t e,
                                                      const u64* __restrict__ items, size_t n, uint32_t* __restrict__ bucket) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) bucket[i] = (uint32_t)(tab_hash(T, hf, items[i]) % e);
}

// start[p] = first index of bucket p in the sorted key array (start[e] = n)
__global__ void __launch_bounds__(256) k_bucket_starts(const uint32_t* __restrict__ sorted_keys, size_t n, uint32_t e,
                                                       uint32_t* __restrict__ start) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > e) return;
    size_t lo = 0, hi = n;  // lower_bound(p)
    while (lo < hi) {
        const size_t mid = (lo + hi) / 2;
        if (sorted_keys[mid] < p) lo = mid + 1;
        else hi = mid;
    }
    start[p] = (uint32_t)lo;
}

// ---- std::mt19937 (one instance per inner table, used by lane 0 only) -----------------------------
struct Mt19937 {
    uint32_t mt[624];
    uint32_t idx;
};
__device__ void mt_seed(Mt19937* s, uint32_t seed) {
    s->mt[0] = seed;
    for (uint32_t i = 1; i < 624; i++) s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + i;
    s->idx = 624;
}
__device__ uint32_t mt_next(Mt19937* s) {
    if (s->idx >= 624) {
        for (uint32_t i = 0; i < 624; i++) {
            const uint32_t y = (s->mt[i] & 0x80000000u) | (s->mt[(i + 1) % 624] & 0x7fffffffu);
            s->mt[i] = s->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        s->idx = 0;
    }
    uint32_t y = s->mt[s->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// ---- one warp builds one inner cuckoo table --------------------------------------------------------
// cells: [k][e][K][b][E] (zero-initialised).  bucketed / start: items of simple table `st` grouped by outer
// position in original order.  fail[0] is set when an insertion fails (CuckooHashTable.cpp:113).
__global__ void __launch_bounds__(128) k_cuckoo_build(const u64* __restrict__ T, uint32_t st, uint32_t k, uint32_t e,
                                                      uint32_t K, uint32_t b, uint32_t E, const u64* __restrict__ bucketed,
                                                      const uint32_t* __restrict__ start, u64 eviction_seed,
                                                      volatile u64* cells, Mt19937* rng_scratch, int* fail) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x & 31;
    if (warp >= e) return;
    const uint32_t p = warp;
    volatile u64* tbl = cells + ((size_t)st * e + p) * K * b * E;
    Mt19937* rng = rng_scratch + ((size_t)st * e + p);
    bool rng_ready = false;
    const uint32_t chunks = (b + 31) / 32;
    const u64 kNoCell = ~0ull;  // lanes beyond b: neither empty nor equal to any item (items are < 2^48)

    for (uint32_t it = start[p]; it < start[p + 1]; it++) {
        u64 value = bucketed[it];
        // lookUp: the element counts as present iff it appears before the first empty bin of one of its K positions
        bool present = false;
        for (uint32_t hf = 0; hf < K && !present; hf++) {
            const uint32_t pos = (uint32_t)(tab_hash(T, k + hf, value) % E);
            for (uint32_t c = 0; c < chunks; c++) {
                const uint32_t bin = c * 32 + lane;
                const u64 cur = bin < b ? tbl[((size_t)hf * b + bin) * E + pos] : kNoCell;
                const uint32_t m_eq = __ballot_sync(0xffffffffu, cur == value);
                const uint32_t m_zero = __ballot_sync(0xffffffffu, cur == 0);
                const uint32_t first_eq = m_eq ? __ffs(m_eq) : 99, first_zero = m_zero ? __ffs(m_zero) : 65;  // none: 99 > 65
                if (first_eq < first_zero) present = true;
                if (present || m_zero) break;
            }
        }
        if (present) continue;
        bool placed = false;
        for (uint32_t run = 0; run < 1000 && !placed; run++) {
            for (uint32_t hf = 0; hf < K && !placed; hf++) {
                const uint32_t pos = (uint32_t)(tab_hash(T, k + hf, value) % E);
                for (uint32_t c = 0; c < chunks && !placed; c++) {
                    const uint32_t bin = c * 32 + lane;
                    const u64 cur = bin < b ? tbl[((size_t)hf * b + bin) * E + pos] : kNoCell;
                    const uint32_t m_zero = __ballot_sync(0xffffffffu, cur == 0);
                    if (m_zero) {
                        const uint32_t free_bin = c * 32 + __ffs(m_zero) - 1;
                        if (lane == 0) tbl[((size_t)hf * b + free_bin) * E + pos] = value;
                        placed = true;
                    }
                }
                if (placed) break;
                // every bin at this position is taken: evict a uniformly chosen one (random walk)
                u64 next = 0;
                if (lane == 0) {
                    if (!rng_ready) {
                        mt_seed(rng, (uint32_t)(eviction_seed + 0x9e3779b9ull * ((u64)st * e + p + 1)));
                        rng_ready = true;
                    }
                    const u64 lo = mt_next(rng), hi = mt_next(rng);
                    const uint32_t victim = (uint32_t)((lo | (hi << 32)) % b);
                    volatile u64* cell = tbl + ((size_t)hf * b + victim) * E + pos;
                    next = *cell;
                    *cell = value;
                }
                rng_ready = __shfl_sync(0xffffffffu, (int)rng_ready, 0) != 0;
                value = __shfl_sync(0xffffffffu, next, 0);
                __syncwarp();
            }
        }
        __syncwarp();
        if (!placed && lane == 0) atomicExch(fail, 1);  // no stash on this path
        if (!placed) return;
    }
}

// ---- host-side driver ------------------------------------------------------------------------------
// items: device pointer (n u64).  T: device tabulation tables for k + K hash functions.  cells: device
// [k][e][K][b][E], overwritten.  Returns cudaErrorUnknown-free status; *failed = 1 if an insertion failed.
cudaError_t hct_build_device(cudaStream_t s, const u64* T, uint32_t k, uint32_t e, uint32_t K, uint32_t b, uint32_t E,
                             u64 eviction_seed, const u64* items, size_t n, u64* cells, int* failed) {
    cudaError_t err;
    const size_t n_cells = (size_t)k * e * K * b * E;
    if ((err = cudaMemsetAsync(cells, 0, n_cells * sizeof(u64), s)) != cudaSuccess) return err;
    if (n == 0) {  // an empty server set is an all-empty table (the host build accepts it too)
        *failed = 0;
        return cudaStreamSynchronize(s);
    }
    uint32_t *keys = nullptr, *keys_sorted = nullptr, *start = nullptr;
    u64* bucketed = nullptr;
    Mt19937* rng = nullptr;
    int* d_fail = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    int end_bit = 1;
    while ((1u << end_bit) < e) end_bit++;
#define HCT_CK(x) if ((err = (x)) != cudaSuccess) goto done
    HCT_CK(cudaMalloc(&keys, n * sizeof(uint32_t)));
    HCT_CK(cudaMalloc(&keys_sorted, n * sizeof(uint32_t)));
    HCT_CK(cudaMalloc(&bucketed, n * sizeof(u64)));
    HCT_CK(cudaMalloc(&start, ((size_t)e + 1) * sizeof(uint32_t)));
    HCT_CK(cudaMalloc(&rng, (size_t)k * e * sizeof(Mt19937)));
    HCT_CK(cudaMalloc(&d_fail, sizeof(int)));
    HCT_CK(cudaMemsetAsync(d_fail, 0, sizeof(int), s));
    HCT_CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_sorted, items, bucketed, (int)n, 0, end_bit, s));
    HCT_CK(cudaMalloc(&tmp, tmp_bytes));
    for (uint32_t st = 0; st < k; st++) {
        // generateSimpleHashTable as a STABLE sort by outer position: the order inside a bucket is the
        // original order, i.e. the reference's push_back order
        k_outer_bucket<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(T, st, e, items, n, keys);
        HCT_CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_sorted, items, bucketed, (int)n, 0, end_bit, s));
        k_bucket_starts<<<(e + 1 + 255) / 256, 256, 0, s>>>(keys_sorted, n, e, start);
        k_cuckoo_build<<<(e * 32 + 127) / 128, 128, 0, s>>>(T, st, k, e, K, b, E, bucketed, start, eviction_seed, cells, rng,
                                                            d_fail);
        HCT_CK(cudaGetLastError());
    }
    HCT_CK(cudaMemcpyAsync(failed, d_fail, sizeof(int), cudaMemcpyDeviceToHost, s));
    HCT_CK(cudaStreamSynchronize(s));
#undef HCT_CK
done:
    cudaFree(keys);
    cudaFree(keys_sorted);
    cudaFree(bucketed);
    cudaFree(start);
    cudaFree(rng);
    cudaFree(d_fail);
    cudaFree(tmp);
    return err;
}

// ---- ctor transposition straight from device cells --------------------------------------------------
// Packed-encoding front end for plaintexts p0 .. p0+n_pt-1 (p = (hf*b + bin)*E + pos), reading the slot
// values from the device tables instead of a host slot array (BatchedFHEHIPPIE.cpp:48-66):
//   slot s = outerHf*e + outerPos  holds  cells[s][hf][perm[s][hf][bin]][pos]   (perm = the ctor's bin shuffle)
// out: [n_pt][N] residues mod t in transform order.
__global__ void __launch_bounds__(256) k_cells_to_crt(uint32_t N, uint32_t n_pt, uint32_t p0, uint32_t nslots, uint32_t K,
                                                      uint32_t b, uint32_t E, uint32_t bin_begin, uint32_t b_local,
                                                      const u64* __restrict__ cells,
                                                      const uint16_t* __restrict__ perm, const uint32_t* __restrict__ to_crt,
                                                      u64* __restrict__ out) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)n_pt * N) return;
    const uint32_t p = p0 + (uint32_t)(tid / N), i = (uint32_t)(tid % N);
    // p counts the plaintexts of the shard: (hf * b_local + local bin) * E + pos
    const uint32_t pos = p % E, bin = bin_begin + (p / E) % b_local, hf = p / (E * b_local);
    const uint32_t s = to_crt[i];
    u64 r = 0;
    if (s < nslots) {
        const uint32_t src_bin = perm[((size_t)s * K + hf) * b + bin];
        r = cells[(((size_t)s * K + hf) * b + src_bin) * E + pos];  // items are < t: already a residue
    }
    out[tid] = r;
}
cudaError_t launch_cells_to_crt(const KCtx& k, uint32_t n_pt, uint32_t p0, uint32_t nslots, uint32_t K, uint32_t b, uint32_t E,
                                uint32_t bin_begin, uint32_t b_local, const u64* cells, const uint16_t* perm,
                                const uint32_t* to_crt, u64* out) {
    const size_t total = (size_t)n_pt * k.N;
    if (total == 0) return cudaSuccess;
    k_cells_to_crt<<<(unsigned)((total + 255) / 256), 256, 0, k.s>>>(k.N, n_pt, p0, nslots, K, b, E, bin_begin, b_local, cells, perm, to_crt, out);
    return cudaGetLastError();
}

}  // namespace psi
