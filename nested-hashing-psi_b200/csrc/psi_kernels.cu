// sm_100a kernels of the BatchedFHEPIE server evaluation (everything except the NTT, which
// lives in ntt.cu).  Each kernel names the OpenFHE operation at the reference call site it
// replaces; the CPU restatement it is checked against is oracle/psi_oracle.c.
#include "psi_kernels.cuh"
#include "async_copy.cuh"

#include <type_traits>
#include <vector>

namespace psi {

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------
// Phase 1 — the encrypted one-hot inner product.
// Replaces the EvalMult(ct,pt) / EvalAdd loop and the EvalAdd of minusCompareElement
// (/root/reference/.../BatchedFHEHIPPIE.cpp:101-116).
//
// HBM-bound by construction: every plaintext word is read exactly once (streaming, L1 bypass).
// To keep the integer pipe below the HBM time the products are formed WITHOUT carry chains:
// the plaintext DB and the index ciphertexts are stored "split-30" (word = hi30:lo30 in the two
// 32-bit halves, residues are < 2^60), so one 60x60-bit product is four IMAD.WIDE.U32 into three
// 64-bit partial sums  ll += x0*y0,  mid += x0*y1 + x1*y0,  hh += x1*y1  (each product < 2^60, so
// 8 positions fit before a fold).  Every 8 positions the partial sums are folded into a 128-bit
// running total, which is reduced once at the end: the canonical residue of the same sum OpenFHE
// forms term by term (ModMul / ModAdd), hence bit-identical.
// A CTA covers 128 coefficients (one row of one limb) x 4 bins; CTAs of the same row are adjacent in
// launch order, so the index words of that row come from DRAM once and from L2 afterwards.
// ------------------------------------------------------------------------------------------
// 32 x 32 + 64 -> 64 multiply-add.  Written in C, not inline PTX: with the inline-asm form ptxas splits every
// pair of accumulating mad.wide into two multiplies plus a 3-input add (twice the issue slots); the
// compiler's own code keeps IMAD.WIDE with its 64-bit addend.
__device__ __forceinline__ u64 madw(uint32_t a, uint32_t b, u64 c) { return c + (u64)a * b; }
__device__ __forceinline__ void fold30(u64& hi, u64& lo, u64 ll, u64 mid, u64 hh) {
    // (hi:lo) += ll + mid * 2^30 + hh * 2^60
    u128 t = (u128)ll + ((u128)mid << 30) + ((u128)hh << 60);
    t += ((u128)hi << 64) | lo;
    lo = (u64)t;
    hi = (u64)(t >> 64);
}

constexpr int kMacCoeffs = 128;  // coefficients per CTA tile (= one row of one limb)

// The operands are streamed by the copy engine: one producer thread issues cp.async.bulk (SASS UBLKCP)
// copies of whole position-chunks — CHUNK positions x 1 KiB per bin, CHUNK x 2 KiB of index words — into a
// STAGES-deep shared-memory ring guarded by mbarriers, and 256 consumer threads (128 coefficients x 2
// bin-lanes, BT bins each) do nothing but LDS + IMAD.WIDE.  The data in flight per SM is set by the rings, not
// by how many loads the compiler keeps in registers.
//
// Multiplier work is what limits the consumers (IMAD.WIDE issues at a quarter of the FP32 rate), so one
// 60 x 60-bit product is formed with THREE 32 x 32 products (Karatsuba on the 30-bit halves):
//   ll += x0*y0,  hh += x1*y1,  kk += (x0+x1)*(y0+y1),  middle = kk - ll - hh.
// ll and hh stay below 2^63 over 8 positions; kk may wrap, but the middle sum is < 2^64, so the wrapped
// 64-bit difference is exact.  Hence the fold into the 128-bit running total after every 8 positions.
//
// The bin-block width (2 * BT bins per CTA) is chosen from the number of resident bins (launch_mac): a block
// width that divides b_local wastes no bin-lane, and a wider block re-reads the index slice from L2 fewer
// times — what matters when a GPU holds 6 of the 47 bins of a sharded query.
constexpr int kMacFold = 8;  // positions between folds
constexpr uint32_t kMacShortRange = 16;  // position ranges up to this length take the persistent kernel (measured below)
constexpr int kMacLanes = 2;
constexpr int kMacConsumers = kMacCoeffs * kMacLanes;
template <int BT, int CHUNK>
__host__ __device__ constexpr size_t mac_stage_words() {
    return (size_t)CHUNK * kMacCoeffs * (2 + kMacLanes * BT);  // idx + pt words per stage
}

// canonical residue of hi * 2^64 + lo: hi through the Shoup pair of 2^64 mod q, lo through floor(2^64 / q)
__device__ __forceinline__ u64 reduce128(u64 hi, u64 lo, const u64 q, const u64 R, const u64 Rs, const u64 qrecip) {
    u64 r = mul_shoup_lazy(hi, R, Rs, q) + (lo - mulhi64(lo, qrecip) * q);  // [0, 2q) + [0, 2q)
    if (r >= 2 * q) r -= 2 * q;
    return r >= q ? r - q : r;
}
// canonical residue of (hi * 2^64 + lo) * 2^-64 for a sum of at most kMacMaxTerms products of canonical residues
// (+ one residue): T < (n + 1) q^2 and q < 2^60 give REDC(T) < T / 2^64 + q < (n / 16 + 2) q <= 8 q
constexpr uint32_t kMacMaxTerms = 96;
__device__ __forceinline__ u64 redc_canonical(u64 hi, u64 lo, const u64 q, const u64 qinv) {
    u64 r = mont_redc_lazy(hi, lo, q, qinv);
    if (r >= 4 * q) r -= 4 * q;
    if (r >= 2 * q) r -= 2 * q;
    return r >= q ? r - q : r;
}

// What one launch covers: hash functions [hf0, hf0 + gridDim.y), positions [pos0, pos1) of each.  flags: bit 0 =
// add the previous contents of acc (a later slice of the same inner product), bit 1 = add minusCompareElement
// (the last slice).  A whole query is one launch with pos0 = 0, pos1 = E, flags = 2; the streamed single-query
// path (psi_query_run_streamed) evaluates slices as their index ciphertexts arrive over PCIe.
struct MacRange {
    uint32_t hf0, pos0, pos1, flags;
};

template <int BT, int CHUNK, int STAGES>
__global__ void __launch_bounds__(kMacConsumers + 32, (BT == 1 ? 3 : 2))
    k_mac_tma(const DevTables* __restrict__ tab, uint32_t N, uint32_t L, uint32_t b, uint32_t E, MacRange rg,
              const u64* __restrict__ pt, const u64* __restrict__ idx, const u64* __restrict__ minus, u64* __restrict__ acc) {
    constexpr int kBins = kMacLanes * BT;
    constexpr size_t kStageWords = mac_stage_words<BT, CHUNK>();
    constexpr uint32_t kChunksPerFold = kMacFold / CHUNK;
    static_assert(kMacFold % CHUNK == 0, "a fold covers whole chunks");
    extern __shared__ __align__(128) u64 ring[];  // [stage][ idx: CHUNK x 2 x 128 | pt: bins x CHUNK x 128 ]
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
    const size_t LN = (size_t)L * N, T = LN / kMacCoeffs;
    // CTAs that share an index slice (same hf and tile, different bin blocks) are adjacent in launch
    // order, so the slice comes from DRAM once and from L2 for the other bin blocks
    const uint32_t nbb = (b + kBins - 1) / kBins;
    const uint32_t hf = rg.hf0 + blockIdx.y, tile = blockIdx.x / nbb, bin_blk0 = (blockIdx.x % nbb) * kBins;
    const uint32_t nbins = min((uint32_t)kBins, b - bin_blk0);
    const uint32_t npos = rg.pos1 - rg.pos0;
    const uint32_t nchunks = (npos + CHUNK - 1) / CHUNK;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kMacConsumers);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (threadIdx.x >= kMacConsumers) {
        // ---- producer: one elected thread keeps the ring full
        if (threadIdx.x == kMacConsumers) {
            const u64* isrc = idx + ((size_t)hf * T + tile) * E * 2 * kMacCoeffs;
            const u64* psrc = pt + ((size_t)hf * b + bin_blk0) * (size_t)E * LN + (size_t)tile * E * kMacCoeffs;
            for (uint32_t ch = 0; ch < nchunks; ch++) {
                const uint32_t s = ch % STAGES, round = ch / STAGES;
                if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
                const uint32_t p0 = rg.pos0 + ch * CHUNK, np = min((uint32_t)CHUNK, rg.pos1 - p0);
                u64* st = ring + (size_t)s * kStageWords;
                const uint32_t ibytes = np * 2 * kMacCoeffs * 8, pbytes = np * kMacCoeffs * 8;
                mbar_expect_tx(&full_bar[s], ibytes + nbins * pbytes);
                bulk_g2s(st, isrc + (size_t)p0 * 2 * kMacCoeffs, ibytes, &full_bar[s]);
                for (uint32_t j = 0; j < nbins; j++)
                    bulk_g2s(st + (size_t)CHUNK * kMacCoeffs * (2 + j), psrc + (size_t)j * E * LN + (size_t)p0 * kMacCoeffs,
                             pbytes, &full_bar[s]);
            }
        }
        return;
    }

    // ---- consumers
    const uint32_t w = threadIdx.x & (kMacCoeffs - 1), lane = threadIdx.x / kMacCoeffs;
    const size_t c = (size_t)tile * kMacCoeffs + w;
    u64 ll[BT][2], kk[BT][2], hh[BT][2], tlo[BT][2], thi[BT][2];
#pragma unroll
    for (int j = 0; j < BT; j++)
#pragma unroll
        for (int k = 0; k < 2; k++) ll[j][k] = kk[j][k] = hh[j][k] = tlo[j][k] = thi[j][k] = 0;

    uint32_t folds = 0;
    for (uint32_t ch = 0; ch < nchunks; ch++) {
        const uint32_t s = ch % STAGES, round = ch / STAGES;
        const uint32_t np = min((uint32_t)CHUNK, npos - ch * CHUNK);
        mbar_wait(&full_bar[s], round & 1);
        const uint2* si = reinterpret_cast<const uint2*>(ring + (size_t)s * kStageWords) + w;
        const uint2* sp = si + (size_t)CHUNK * kMacCoeffs * (2 + lane * BT);
        auto body = [&](int p) {
            const uint2 i0 = si[p * 2 * kMacCoeffs], i1 = si[p * 2 * kMacCoeffs + kMacCoeffs];
            const uint32_t s0 = i0.x + i0.y, s1 = i1.x + i1.y;
#pragma unroll
            for (int j = 0; j < BT; j++) {
                const uint2 y = sp[(j * CHUNK + p) * kMacCoeffs];
                const uint32_t sy = y.x + y.y;
                ll[j][0] = madw(i0.x, y.x, ll[j][0]);
                hh[j][0] = madw(i0.y, y.y, hh[j][0]);
                kk[j][0] = madw(s0, sy, kk[j][0]);
                ll[j][1] = madw(i1.x, y.x, ll[j][1]);
                hh[j][1] = madw(i1.y, y.y, hh[j][1]);
                kk[j][1] = madw(s1, sy, kk[j][1]);
            }
        };
        if (np == CHUNK) {
#pragma unroll
            for (int p = 0; p < CHUNK; p++) body(p);
        } else {
            for (uint32_t p = 0; p < np; p++) body((int)p);
        }
        mbar_arrive(&empty_bar[s]);  // this thread is done reading stage s
        if ((ch + 1) % kChunksPerFold != 0 && ch + 1 < nchunks) continue;
#pragma unroll
        for (int j = 0; j < BT; j++)
#pragma unroll
            for (int k = 0; k < 2; k++) {
                fold30(thi[j][k], tlo[j][k], ll[j][k], kk[j][k] - ll[j][k] - hh[j][k], hh[j][k]);
                ll[j][k] = kk[j][k] = hh[j][k] = 0;
            }
        if (++folds == kMacMaxTerms / kMacFold && ch + 1 < nchunks) {  // keep the total within redc_canonical's bound for any E
            folds = 0;
            const ModDev& mdr = tab->mods[c / N];
#pragma unroll
            for (int j = 0; j < BT; j++)
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    tlo[j][k] = reduce128(thi[j][k], tlo[j][k], mdr.q, mdr.Rmodq, mdr.Rmodq_s, mdr.mu_hi);
                    thi[j][k] = 0;
                }
        }
    }
    const ModDev& md = tab->mods[c / N];
    const u64 q = md.q, qinv = md.qinv;
    const bool add_old = rg.flags & 1u, add_minus = rg.flags & 2u;
    const u64 m0 = add_minus ? minus[c] : 0, m1 = add_minus ? minus[LN + c] : 0;
#pragma unroll
    for (int j = 0; j < BT; j++) {
        const uint32_t bin = bin_blk0 + lane * BT + j;
        if (bin < b) {
            u64* o = acc + (((size_t)hf * b + bin) * 2) * LN + c;
            u64 r0 = addmod(redc_canonical(thi[j][0], tlo[j][0], q, qinv), m0, q);
            u64 r1 = addmod(redc_canonical(thi[j][1], tlo[j][1], q, qinv), m1, q);
            if (add_old) {
                r0 = addmod(r0, o[0], q);
                r1 = addmod(r1, o[LN], q);
            }
            o[0] = r0;
            o[LN] = r1;
        }
    }
}

// Persistent form for SHORT position ranges (E = 14 of BASELINE configs[1]: two chunks per work item): every CTA walks
// over work items (bin block, row, hash function) with a stride of the grid and the producer's ring runs ACROSS items,
// so the first chunk of item i + 1 is in flight while the consumers reduce and store item i -- the per-item bubble
// (first-chunk latency, ~1.5 us against ~1 us of multiplies at E = 14) is what held the one-item kernel at 57-60 % of
// the HBM peak.  Same ring, same arithmetic, same results as k_mac_tma; at E = 47 the one-item kernel is faster.
template <int BT, int CHUNK, int STAGES>
__global__ void __launch_bounds__(kMacConsumers + 32, (BT == 1 ? 3 : 2))
    k_mac_persist(const DevTables* __restrict__ tab, uint32_t N, uint32_t L, uint32_t b, uint32_t E, MacRange rg, uint32_t nhf,
                  const u64* __restrict__ pt, const u64* __restrict__ idx, const u64* __restrict__ minus, u64* __restrict__ acc) {
    constexpr int kBins = kMacLanes * BT;
    constexpr size_t kStageWords = mac_stage_words<BT, CHUNK>();
    extern __shared__ __align__(128) u64 ring[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
    const size_t LN = (size_t)L * N, T = LN / kMacCoeffs;
    const uint32_t nbb = (b + kBins - 1) / kBins;
    const uint32_t npos = rg.pos1 - rg.pos0;
    const uint32_t nchunks = (npos + CHUNK - 1) / CHUNK;
    const uint32_t items_per_hf = (uint32_t)T * nbb, n_items = items_per_hf * nhf;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kMacConsumers);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (threadIdx.x >= kMacConsumers) {
        if (threadIdx.x == kMacConsumers) {
            uint32_t g = 0;  // chunks issued so far by this CTA, over all its items
            for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
                const uint32_t hf = rg.hf0 + item / items_per_hf, r = item % items_per_hf;
                const uint32_t tile = r / nbb, bin_blk0 = (r % nbb) * kBins;
                const uint32_t nbins = min((uint32_t)kBins, b - bin_blk0);
                const u64* isrc = idx + ((size_t)hf * T + tile) * E * 2 * kMacCoeffs;
                const u64* psrc = pt + ((size_t)hf * b + bin_blk0) * (size_t)E * LN + (size_t)tile * E * kMacCoeffs;
                for (uint32_t ch = 0; ch < nchunks; ch++, g++) {
                    const uint32_t s = g % STAGES, round = g / STAGES;
                    if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
                    const uint32_t p0 = rg.pos0 + ch * CHUNK, np = min((uint32_t)CHUNK, rg.pos1 - p0);
                    u64* st = ring + (size_t)s * kStageWords;
                    const uint32_t ibytes = np * 2 * kMacCoeffs * 8, pbytes = np * kMacCoeffs * 8;
                    mbar_expect_tx(&full_bar[s], ibytes + nbins * pbytes);
                    bulk_g2s(st, isrc + (size_t)p0 * 2 * kMacCoeffs, ibytes, &full_bar[s]);
                    for (uint32_t j = 0; j < nbins; j++)
                        bulk_g2s(st + (size_t)CHUNK * kMacCoeffs * (2 + j), psrc + (size_t)j * E * LN + (size_t)p0 * kMacCoeffs,
                                 pbytes, &full_bar[s]);
                }
            }
        }
        return;
    }

    const uint32_t w = threadIdx.x & (kMacCoeffs - 1), lane = threadIdx.x / kMacCoeffs;
    const bool add_old = rg.flags & 1u, add_minus = rg.flags & 2u;
    uint32_t g = 0;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t hf = rg.hf0 + item / items_per_hf, r = item % items_per_hf;
        const uint32_t tile = r / nbb, bin_blk0 = (r % nbb) * kBins;
        const size_t c = (size_t)tile * kMacCoeffs + w;
        u64 ll[BT][2], kk[BT][2], hh[BT][2], tlo[BT][2], thi[BT][2];
#pragma unroll
        for (int j = 0; j < BT; j++)
#pragma unroll
            for (int k = 0; k < 2; k++) ll[j][k] = kk[j][k] = hh[j][k] = tlo[j][k] = thi[j][k] = 0;
        for (uint32_t ch = 0; ch < nchunks; ch++, g++) {
            const uint32_t s = g % STAGES, round = g / STAGES;
            const uint32_t np = min((uint32_t)CHUNK, npos - ch * CHUNK);
            mbar_wait(&full_bar[s], round & 1);
            const uint2* si = reinterpret_cast<const uint2*>(ring + (size_t)s * kStageWords) + w;
            const uint2* sp = si + (size_t)CHUNK * kMacCoeffs * (2 + lane * BT);
            for (uint32_t p = 0; p < np; p++) {
                const uint2 i0 = si[p * 2 * kMacCoeffs], i1 = si[p * 2 * kMacCoeffs + kMacCoeffs];
                const uint32_t s0 = i0.x + i0.y, s1 = i1.x + i1.y;
#pragma unroll
                for (int j = 0; j < BT; j++) {
                    const uint2 y = sp[(j * CHUNK + p) * kMacCoeffs];
                    const uint32_t sy = y.x + y.y;
                    ll[j][0] = madw(i0.x, y.x, ll[j][0]);
                    hh[j][0] = madw(i0.y, y.y, hh[j][0]);
                    kk[j][0] = madw(s0, sy, kk[j][0]);
                    ll[j][1] = madw(i1.x, y.x, ll[j][1]);
                    hh[j][1] = madw(i1.y, y.y, hh[j][1]);
                    kk[j][1] = madw(s1, sy, kk[j][1]);
                }
            }
            mbar_arrive(&empty_bar[s]);
            // CHUNK = kMacFold positions per chunk: fold after every chunk (the launcher only takes this kernel for
            // ranges of at most kMacMaxTerms positions, so the 128-bit total needs no intermediate reduction)
#pragma unroll
            for (int j = 0; j < BT; j++)
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    fold30(thi[j][k], tlo[j][k], ll[j][k], kk[j][k] - ll[j][k] - hh[j][k], hh[j][k]);
                    ll[j][k] = kk[j][k] = hh[j][k] = 0;
                }
        }
        const ModDev& md = tab->mods[c / N];
        const u64 q = md.q, qinv = md.qinv;
        const u64 m0 = add_minus ? minus[c] : 0, m1 = add_minus ? minus[LN + c] : 0;
#pragma unroll
        for (int j = 0; j < BT; j++) {
            const uint32_t bin = bin_blk0 + lane * BT + j;
            if (bin < b) {
                u64* o = acc + (((size_t)hf * b + bin) * 2) * LN + c;
                u64 r0 = addmod(redc_canonical(thi[j][0], tlo[j][0], q, qinv), m0, q);
                u64 r1 = addmod(redc_canonical(thi[j][1], tlo[j][1], q, qinv), m1, q);
                if (add_old) {
                    r0 = addmod(r0, o[0], q);
                    r1 = addmod(r1, o[LN], q);
                }
                o[0] = r0;
                o[LN] = r1;
            }
        }
    }
}

// The instantiations the launcher chooses from: <bins per lane, positions per stage, stages>
//   <2, 8, 2>  4 bins per CTA, 96 KiB ring, 2 CTAs / SM: the full-database shape (b = 47: 12 blocks)
//   <1, 8, 2>  2 bins per CTA, 64 KiB ring, 3 CTAs / SM: few resident bins whose count 4 does not divide (a GPU's
//              5 or 6 of the 47 bins of a sharded query: 85.6 % of the HBM peak against 78 % with 4-bin blocks)
// (a 6-bin block, <3, 4, 3>, was measured too: never the fastest, 77-81 % at b = 5, 6, 47; profiles/r02_tune_shapes.md)
template <int BT, int CHUNK, int STAGES>
static cudaError_t mac_launch_t(const KCtx& k, uint32_t nhf, uint32_t b, uint32_t E, const MacRange& rg, const u64* pt,
                                const u64* idx, const u64* minus, u64* acc, bool init_only) {
    constexpr size_t smem = STAGES * mac_stage_words<BT, CHUNK>() * sizeof(u64);
    if (init_only)
        return cudaFuncSetAttribute(k_mac_tma<BT, CHUNK, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const size_t LN = (size_t)k.L * k.N;
    constexpr int kBins = kMacLanes * BT;
    dim3 grid(cdiv(LN, kMacCoeffs) * ((b + kBins - 1) / kBins), nhf);
    k_mac_tma<BT, CHUNK, STAGES><<<grid, kMacConsumers + 32, smem, k.s>>>(k.tab, k.N, k.L, b, E, rg, pt, idx, minus, acc);
    return cudaGetLastError();
}

template <int BT, int CHUNK, int STAGES>
static cudaError_t mac_launch_persist(const KCtx& k, uint32_t nhf, uint32_t b, uint32_t E, const MacRange& rg, const u64* pt,
                                      const u64* idx, const u64* minus, u64* acc, bool init_only) {
    static_assert(CHUNK == kMacFold, "the persistent kernel folds once per chunk");
    constexpr size_t smem = STAGES * mac_stage_words<BT, CHUNK>() * sizeof(u64);
    if (init_only)
        return cudaFuncSetAttribute(k_mac_persist<BT, CHUNK, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const size_t LN = (size_t)k.L * k.N;
    constexpr int kBins = kMacLanes * BT;
    const size_t n_items = cdiv(LN, kMacCoeffs) * ((b + kBins - 1) / kBins) * nhf;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t slots = (size_t)sms * (BT == 1 ? 3 : 2);
    const unsigned grid = (unsigned)(n_items < slots ? n_items : slots);
    k_mac_persist<BT, CHUNK, STAGES><<<grid, kMacConsumers + 32, smem, k.s>>>(k.tab, k.N, k.L, b, E, rg, nhf, pt, idx, minus, acc);
    return cudaGetLastError();
}

static int g_mac_force = 0;  // tuning / tests: 0 = choose by shape, 1 / 2 = force 2 / 4 bins per CTA, 3 / 4 = the same with a 4 x 4 ring
void mac_force_variant(int v) { g_mac_force = v; }
int mac_forced_variant() { return g_mac_force; }

// bin-block width for b resident bins: 2-bin blocks when 4-bin blocks would leave more than a tenth of the
// bin-lanes idle (measured: b = 5, 6, 14 faster with 2, b = 26, 47, 75 faster with 4)
// Kernel for npos positions per launch: a short range (E = 14 of BASELINE configs[1], E = 8 of configs[0], the
// 12-position slices of the streamed single query) is at most two chunks of eight, so the one-item kernel pays its
// first-chunk latency for every (bin block, row).  MEASURED (tools/tune_shapes.py, % of the HBM peak, one-item kernel ->
// persistent kernel with the ring across items): 14 bins x 14 positions 59.5 -> 62.6, 14 x 12 55.3 -> 58.6, 47 x 12
// 57.1 -> 63.1, 8 x 8 51.3 -> 57.9; from 26 positions on the one-item kernel wins (26 x 26: 73.3 against 71.4, 47 x 47:
// 88.1 against 83.6, 75 x 75: 92.4 against 85.1).  A ring of four stages of four positions (variants 3 / 4) never won.
static int mac_choose(uint32_t b, uint32_t npos) {
    if (g_mac_force && !(g_mac_force >= 5 && npos > kMacMaxTerms)) return g_mac_force;
    const uint32_t slots4 = ((b + 3) / 4) * 4, slots2 = ((b + 1) / 2) * 2;
    const int bins = (slots2 < slots4 && (slots4 - b) * 10 > slots4) ? 1 : 2;
    return npos <= kMacShortRange ? bins + 4 : bins;
}

cudaError_t mac_init_device() {
    const KCtx k{};
    const MacRange rg{};
    cudaError_t e = mac_launch_t<2, 8, 2>(k, 0, 0, 0, rg, nullptr, nullptr, nullptr, nullptr, true);
    if (e == cudaSuccess) e = mac_launch_t<1, 8, 2>(k, 0, 0, 0, rg, nullptr, nullptr, nullptr, nullptr, true);
    if (e == cudaSuccess) e = mac_launch_t<2, 4, 4>(k, 0, 0, 0, rg, nullptr, nullptr, nullptr, nullptr, true);
    if (e == cudaSuccess) e = mac_launch_t<1, 4, 4>(k, 0, 0, 0, rg, nullptr, nullptr, nullptr, nullptr, true);
    if (e == cudaSuccess) e = mac_launch_persist<1, 8, 2>(k, 0, 0, 0, rg, nullptr, nullptr, nullptr, nullptr, true);
    if (e == cudaSuccess) e = mac_launch_persist<2, 8, 2>(k, 0, 0, 0, rg, nullptr, nullptr, nullptr, nullptr, true);
    return e;
}

cudaError_t launch_mac_range(const KCtx& k, uint32_t hf0, uint32_t nhf, uint32_t b, uint32_t E, uint32_t pos0, uint32_t pos1,
                             uint32_t flags, const u64* pt, const u64* idx, const u64* minus, u64* acc) {
    const MacRange rg{hf0, pos0, pos1, flags};
    switch (mac_choose(b, pos1 - pos0)) {
        case 1: return mac_launch_t<1, 8, 2>(k, nhf, b, E, rg, pt, idx, minus, acc, false);
        case 3: return mac_launch_t<1, 4, 4>(k, nhf, b, E, rg, pt, idx, minus, acc, false);
        case 4: return mac_launch_t<2, 4, 4>(k, nhf, b, E, rg, pt, idx, minus, acc, false);
        case 5: return mac_launch_persist<1, 8, 2>(k, nhf, b, E, rg, pt, idx, minus, acc, false);
        case 6: return mac_launch_persist<2, 8, 2>(k, nhf, b, E, rg, pt, idx, minus, acc, false);
        default: return mac_launch_t<2, 8, 2>(k, nhf, b, E, rg, pt, idx, minus, acc, false);
    }
}

cudaError_t launch_mac(const KCtx& k, uint32_t K, uint32_t b, uint32_t E, const u64* pt, const u64* idx,
                       const u64* minus, u64* acc) {
    return launch_mac_range(k, 0, K, b, E, 0, E, 2u, pt, idx, minus, acc);
}

// Storage formats of the two operands of phase 1.  Words: canonical residue < 2^60 <-> split-30 word
// (hi30 in the upper 32 bits, lo30 in the lower).  Order: with T = L*N/128 tiles of 128 coefficients,
//   plaintext DB   pt_t [hf*b + bin][tile][pos][128]      from  [p = (hf*b+bin)*E + pos][L*N]
//   index cts      idx_t[hf][tile][pos][comp][128]        from  [hf][pos][comp][L*N]
__device__ __forceinline__ u64 to_split30(u64 v) { return ((v >> 30) << 32) | (v & 0x3fffffffull); }
__device__ __forceinline__ u64 from_split30(u64 v) { return ((v >> 32) << 30) | (v & 0x3fffffffull); }

// chunk of n plaintexts starting at plaintext index p0: flat [n][LN] canonical <-> tiled DB
// range_tab / bad: when given (caller-supplied limbs, psi_db_load_limbs), a residue >= q_l raises *bad instead of
// being silently truncated by to_split30
__global__ void __launch_bounds__(256) k_retile_pt(u64* __restrict__ flat, u64* __restrict__ tiled, size_t LN, uint32_t E,
                                                   size_t p0, size_t total, int to_tiled,
                                                   const DevTables* __restrict__ range_tab, int* __restrict__ bad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t p = p0 + i / LN, cidx = i % LN;
    const size_t g = p / E, pos = p % E, tile = cidx / kMacCoeffs, w = cidx % kMacCoeffs;
    const size_t T = LN / kMacCoeffs;
    const size_t d = ((g * T + tile) * E + pos) * kMacCoeffs + w;
    if (to_tiled) {
        const u64 v = flat[i];
        if (range_tab && v >= range_tab->mods[cidx / range_tab->N].q) *bad = 1;
        tiled[d] = to_split30(v);
    } else {
        flat[i] = from_split30(tiled[d]);
    }
}
cudaError_t launch_retile_pt(cudaStream_t s, u64* flat, u64* tiled, size_t LN, uint32_t E, size_t p0, size_t n,
                             bool to_tiled, const DevTables* range_tab, int* bad) {
    const size_t total = n * LN;
    if (total == 0) return cudaSuccess;
    k_retile_pt<<<cdiv(total, 256), 256, 0, s>>>(flat, tiled, LN, E, p0, total, to_tiled ? 1 : 0, range_tab, bad);
    return cudaGetLastError();
}
// The index words are stored in Montgomery form (times R = 2^64 mod q_l): the inner product of k_mac_tma then ends
// with ONE Montgomery reduction per output, REDC(sum idx*R*pt) = sum idx*pt, instead of a 128-bit Barrett; the
// multiplication by R rides on this bandwidth-bound re-tiling pass.
__global__ void __launch_bounds__(256) k_retile_idx(const DevTables* __restrict__ tab, uint32_t N, const u64* __restrict__ flat,
                                                    u64* __restrict__ tiled, size_t LN, uint32_t E, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t cidx = i % LN, r = i / LN;  // r = (hf*E + pos)*2 + comp
    const size_t comp = r & 1, pos = (r >> 1) % E, hf = (r >> 1) / E;
    const size_t T = LN / kMacCoeffs, tile = cidx / kMacCoeffs, w = cidx % kMacCoeffs;
    const ModDev& md = tab->mods[cidx / N];
    tiled[(((hf * T + tile) * E + pos) * 2 + comp) * kMacCoeffs + w] = to_split30(mul_shoup(flat[i], md.Rmodq, md.Rmodq_s, md.q));
}
// positions [pos0, pos1) of hash function hf only (the streamed single-query path re-tiles slices as they land)
__global__ void __launch_bounds__(256) k_retile_idx_range(const DevTables* __restrict__ tab, uint32_t N, const u64* __restrict__ flat,
                                                          u64* __restrict__ tiled, size_t LN, uint32_t E, uint32_t hf,
                                                          uint32_t pos0, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t cidx = i % LN, r = i / LN;  // r = (pos - pos0) * 2 + comp
    const size_t comp = r & 1, pos = pos0 + (r >> 1);
    const size_t T = LN / kMacCoeffs, tile = cidx / kMacCoeffs, w = cidx % kMacCoeffs;
    const ModDev& md = tab->mods[cidx / N];
    const u64 v = flat[(((size_t)hf * E + pos) * 2 + comp) * LN + cidx];
    tiled[((((size_t)hf * T + tile) * E + pos) * 2 + comp) * kMacCoeffs + w] = to_split30(mul_shoup(v, md.Rmodq, md.Rmodq_s, md.q));
}
cudaError_t launch_retile_idx_range(const KCtx& k, const u64* flat, u64* tiled, size_t LN, uint32_t E, uint32_t hf, uint32_t pos0,
                                    uint32_t pos1) {
    const size_t total = (size_t)(pos1 - pos0) * 2 * LN;
    if (total == 0) return cudaSuccess;
    k_retile_idx_range<<<cdiv(total, 256), 256, 0, k.s>>>(k.tab, k.N, flat, tiled, LN, E, hf, pos0, total);
    return cudaGetLastError();
}
cudaError_t launch_retile_idx(const KCtx& k, const u64* flat, u64* tiled, size_t LN, uint32_t K, uint32_t E) {
    const size_t total = (size_t)K * E * 2 * LN;
    if (total == 0) return cudaSuccess;
    k_retile_idx<<<cdiv(total, 256), 256, 0, k.s>>>(k.tab, k.N, flat, tiled, LN, E, total);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// EvalMult(ct,ct) — BFV-RNS multiplication, HPSPOVERQ variant (BatchedFHEHIPPIE.cpp:123).
// Coefficient-wise kernels: one thread owns one coefficient across all limbs.
// The double sequences (nu) use explicit round-to-nearest mul/add intrinsics so that no FMA
// contraction can change the rounding relative to the host library.
// ------------------------------------------------------------------------------------------

// DCRTPoly::SwitchCRTBasis for one coefficient: exact conversion basis A -> basis B.
template <bool Q_TO_P>
__device__ __forceinline__ void switch_basis(const DevTables* __restrict__ tab, int nA, int nB, const u64* x,
                                             u64* out) {
    u64 y[PSI_MAX_LIMBS];
    double nu = 0.5;
#pragma unroll
    for (int i = 0; i < PSI_MAX_LIMBS; i++) {
        if (i < nA) {
            const u64 a = Q_TO_P ? tab->mods[i].q : tab->mods[tab->L + i].q;
            const u64 c = Q_TO_P ? tab->QHatInvModq[i] : tab->PHatInvModp[i];
            const u64 cs = Q_TO_P ? tab->QHatInvModq_s[i] : tab->PHatInvModp_s[i];
            y[i] = mul_shoup(x[i], c, cs, a);
            const double inv = Q_TO_P ? tab->qInv[i] : tab->pInv[i];
            nu = nu_step(nu, __ull2double_rn(y[i]), inv, tab->fp_fma);
        }
    }
    const unsigned alpha = (unsigned)nu;
#pragma unroll
    for (int j = 0; j < PSI_MAX_LIMBS; j++) {
        if (j < nB) {
            const ModDev& mb = Q_TO_P ? tab->mods[tab->L + j] : tab->mods[j];
            u64 hi = 0, lo = 0;
#pragma unroll
            for (int i = 0; i < PSI_MAX_LIMBS; i++)
                if (i < nA) mac128(hi, lo, y[i], Q_TO_P ? tab->QHatModp[j][i] : tab->PHatModq[j][i]);
            const u64 v = barrett128(hi, lo, mb.q, mb.mu_hi, mb.mu_lo);
            out[j] = submod(v, Q_TO_P ? tab->alphaQModp[alpha][j] : tab->alphaPModq[alpha][j], mb.q);
        }
    }
}

// DCRTPoly::ExpandCRTBasis, first operand: coef [groups][L][N] (COEFFICIENT) -> P limbs written to
// ext [groups][L+Lp][N] at limb offset L (the Q limbs of ext keep the EVALUATION input).
__global__ void __launch_bounds__(256) k_expand_q_to_p(const DevTables* __restrict__ tab, uint32_t N, uint32_t groups,
                                                       const u64* __restrict__ coef, u64* __restrict__ ext) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)groups * N) return;
    const uint32_t g = tid / N, n = tid % N;
    const int L = tab->L, Lp = tab->Lp, LT = L + Lp;
    u64 x[PSI_MAX_LIMBS], y[PSI_MAX_LIMBS];
#pragma unroll
    for (int l = 0; l < PSI_MAX_LIMBS; l++)
        if (l < L) x[l] = coef[((size_t)g * L + l) * N + n];
    switch_basis<true>(tab, L, Lp, x, y);
#pragma unroll
    for (int j = 0; j < PSI_MAX_LIMBS; j++)
        if (j < Lp) ext[((size_t)g * LT + L + j) * N + n] = y[j];
}

// DCRTPoly::FastExpandCRTBasisPloverQ, second operand: coef [groups][L][N] (COEFFICIENT) ->
// ext [groups][L+Lp][N] (COEFFICIENT, all limbs).
__global__ void __launch_bounds__(256) k_fast_expand_poverq(const DevTables* __restrict__ tab, uint32_t N,
                                                            uint32_t groups, const u64* __restrict__ coef,
                                                            u64* __restrict__ ext) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)groups * N) return;
    const uint32_t g = tid / N, n = tid % N;
    const int L = tab->L, Lp = tab->Lp, LT = L + Lp;
    u64 y[PSI_MAX_LIMBS], pp[PSI_MAX_LIMBS], qq[PSI_MAX_LIMBS];
#pragma unroll
    for (int i = 0; i < PSI_MAX_LIMBS; i++)
        if (i < L)
            y[i] = mul_shoup(coef[((size_t)g * L + i) * N + n], tab->negPQHatInvModq[i], tab->negPQHatInvModq_s[i],
                             tab->mods[i].q);
#pragma unroll
    for (int j = 0; j < PSI_MAX_LIMBS; j++) {
        if (j < Lp) {
            const ModDev& mp = tab->mods[L + j];
            u64 hi = 0, lo = 0;
#pragma unroll
            for (int i = 0; i < PSI_MAX_LIMBS; i++)
                if (i < L) mac128(hi, lo, y[i], tab->qInvModp[i][j]);
            pp[j] = barrett128(hi, lo, mp.q, mp.mu_hi, mp.mu_lo);
        }
    }
    switch_basis<false>(tab, Lp, L, pp, qq);
#pragma unroll
    for (int l = 0; l < PSI_MAX_LIMBS; l++)
        if (l < L) ext[((size_t)g * LT + l) * N + n] = qq[l];
#pragma unroll
    for (int j = 0; j < PSI_MAX_LIMBS; j++)
        if (j < Lp) ext[((size_t)g * LT + L + j) * N + n] = pp[j];
}

// Tensor product in basis QP, EVALUATION: (c0 c0', c0 c1' + c1 c0', c1 c1').
// e1, e2: [B][2][LT][N], ten: [B][3][LT][N]
__global__ void __launch_bounds__(256) k_tensor(const DevTables* __restrict__ tab, uint32_t N, uint32_t B,
                                                const u64* __restrict__ e1, const u64* __restrict__ e2,
                                                u64* __restrict__ ten) {
    const size_t LTN = (size_t)(tab->L + tab->Lp) * N;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)B * LTN) return;
    const size_t bin = tid / LTN, c = tid % LTN;
    const ModDev& m = tab->mods[c / N];
    const u64 a0 = e1[(bin * 2) * LTN + c], a1 = e1[(bin * 2 + 1) * LTN + c];
    const u64 b0 = e2[(bin * 2) * LTN + c], b1 = e2[(bin * 2 + 1) * LTN + c];
    ten[(bin * 3) * LTN + c] = mulmod(a0, b0, m);
    u64 hi = 0, lo = 0;
    mac128(hi, lo, a0, b1);
    mac128(hi, lo, a1, b0);
    ten[(bin * 3 + 1) * LTN + c] = barrett128(hi, lo, m.q, m.mu_hi, m.mu_lo);
    ten[(bin * 3 + 2) * LTN + c] = mulmod(a1, b1, m);
}

// DCRTPoly::ScaleAndRound by t/P with output basis Q.  ten: [groups][LT][N] COEFFICIENT ->
// res: [groups][L][N] COEFFICIENT.
__global__ void __launch_bounds__(256) k_scale_round(const DevTables* __restrict__ tab, uint32_t N, uint32_t groups,
                                                     const u64* __restrict__ ten, u64* __restrict__ res) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)groups * N) return;
    const uint32_t g = tid / N, n = tid % N;
    const int L = tab->L, Lp = tab->Lp, LT = L + Lp;
    u64 xp[PSI_MAX_LIMBS];
    double nu = 0.5;
#pragma unroll
    for (int i = 0; i < PSI_MAX_LIMBS; i++) {
        if (i < Lp) {
            xp[i] = ten[((size_t)g * LT + L + i) * N + n];
            nu = nu_step(nu, tab->tQSfrac[i], __ull2double_rn(xp[i]), tab->fp_fma);
        }
    }
    const u64 alpha = __double2ull_rz(nu);
#pragma unroll
    for (int l = 0; l < PSI_MAX_LIMBS; l++) {
        if (l < L) {
            const ModDev& m = tab->mods[l];
            u64 hi = 0, lo = 0;
#pragma unroll
            for (int i = 0; i < PSI_MAX_LIMBS; i++)
                if (i < Lp) mac128(hi, lo, xp[i], tab->tQS[l][i]);
            mac128(hi, lo, ten[((size_t)g * LT + l) * N + n], tab->tQS[l][Lp]);
            const u64 v = barrett128(hi, lo, m.q, m.mu_hi, m.mu_lo);
            res[((size_t)g * L + l) * N + n] = addmod(v, alpha % m.q, m.q);
        }
    }
}

// DCRTPoly::CRTDecompose (BV, digit size 0): digit i = limb i of c2 (COEFFICIENT), switched to
// every modulus q_k with the centred lift of NativeVector::SwitchModulus.
// res: [B][3][L][N] (component 2 is read), dig: [B][L(i)][L(k)][N] COEFFICIENT.
__global__ void __launch_bounds__(256) k_relin_digits(const DevTables* __restrict__ tab, uint32_t N, uint32_t B,
                                                      const u64* __restrict__ res, u64* __restrict__ dig) {
    const int L = tab->L;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)B * L * N) return;
    const uint32_t n = tid % N, i = (tid / N) % L;
    const size_t bin = tid / ((size_t)N * L);
    const u64 v = res[((bin * 3 + 2) * L + i) * N + n];
    const u64 qi = tab->mods[i].q;
    const bool neg = v > ((qi - 1) >> 1);
#pragma unroll
    for (int k = 0; k < PSI_MAX_LIMBS; k++) {
        if (k < L) {
            const u64 qk = tab->mods[k].q;
            u64 r = v;
            if (k != (int)i) {
                // q_i and q_k are both just below 2^60, so v < 2 q_k; general moduli fall back to %
                r = (v < qk) ? v : ((v - qk < qk) ? v - qk : v % qk);
                if (neg) r = submod(r, tab->qModq[i][k], qk);
            }
            dig[((bin * L + i) * L + k) * N + n] = r;
        }
    }
}

// KeySwitchBV::EvalFastKeySwitchCore + cv[0] += ..., cv[1] += ... + optional EvalMult(ct, mask)
// (BatchedFHEHIPPIE.cpp:126).  res_eval: [B][3][L][N] with components 0,1 in EVALUATION;
// dig: [B][L][L][N] EVALUATION; evk_*: [L][L][N]; mask: [B][L][N] or null; out: [B][2][L][N].
__global__ void __launch_bounds__(256) k_relin_accum(const DevTables* __restrict__ tab, uint32_t N, uint32_t B,
                                                     const u64* __restrict__ res_eval, const u64* __restrict__ dig,
                                                     const u64* __restrict__ evk_b, const u64* __restrict__ evk_a,
                                                     const u64* __restrict__ mask, u64* __restrict__ out) {
    const int L = tab->L;
    const size_t LN = (size_t)L * N;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)B * LN) return;
    const size_t bin = tid / LN, c = tid % LN;  // c = k*N + n
    const ModDev& m = tab->mods[c / N];
    u64 h0 = 0, l0 = 0, h1 = 0, l1 = 0;
#pragma unroll
    for (int i = 0; i < PSI_MAX_LIMBS; i++) {
        if (i < L) {
            const u64 d = dig[(bin * L + i) * LN + c];
            mac128(h0, l0, d, evk_b[(size_t)i * LN + c]);
            mac128(h1, l1, d, evk_a[(size_t)i * LN + c]);
        }
    }
    u64 r0 = addmod(barrett128(h0, l0, m.q, m.mu_hi, m.mu_lo), res_eval[(bin * 3) * LN + c], m.q);
    u64 r1 = addmod(barrett128(h1, l1, m.q, m.mu_hi, m.mu_lo), res_eval[(bin * 3 + 1) * LN + c], m.q);
    if (mask) {
        const u64 mv = mask[bin * LN + c];
        r0 = mulmod(r0, mv, m);
        r1 = mulmod(r1, mv, m);
    }
    out[(bin * 2) * LN + c] = r0;
    out[(bin * 2 + 1) * LN + c] = r1;
}

// EvalMult(ct, pt) stand-alone (only used when K == 1, i.e. no ct x ct precedes the mask).
__global__ void __launch_bounds__(256) k_mul_ctpt(const DevTables* __restrict__ tab, uint32_t N, uint32_t B,
                                                  const u64* __restrict__ ct, const u64* __restrict__ pt,
                                                  u64* __restrict__ out) {
    const size_t LN = (size_t)tab->L * N;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)B * LN) return;
    const size_t bin = tid / LN, c = tid % LN;
    const ModDev& m = tab->mods[c / N];
    const u64 mv = pt[bin * LN + c];
    out[(bin * 2) * LN + c] = mulmod(ct[(bin * 2) * LN + c], mv, m);
    out[(bin * 2 + 1) * LN + c] = mulmod(ct[(bin * 2 + 1) * LN + c], mv, m);
}

// PackedEncoding::Encode front end: signed slot values -> residues mod t in CRT (transform) order.
__global__ void __launch_bounds__(256) k_slots_to_crt(const DevTables* __restrict__ tab, uint32_t N, uint32_t n_pt,
                                                      uint32_t nslots, const long long* __restrict__ slots,
                                                      const uint32_t* __restrict__ to_crt, u64* __restrict__ out) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)n_pt * N) return;
    const size_t p = tid / N;
    const uint32_t i = tid % N;
    const uint32_t s = to_crt[i];
    u64 r = 0;
    if (s < nslots) {
        const long long v = slots[p * nslots + s];
        const u64 a = (u64)(v < 0 ? -v : v);
        r = (v < 0 && a) ? tab->t - a : a;
    }
    out[tid] = r;
}

// PSI_MULT_HPS: DCRTPoly::ScaleAndRound by t/Q with output basis P (nu over the Q limbs), then the exact
// DCRTPoly::SwitchCRTBasis P -> Q.  ten: [groups][LT][N] COEFFICIENT -> res: [groups][L][N] COEFFICIENT.
__global__ void __launch_bounds__(256) k_scale_round_hps(const DevTables* __restrict__ tab, uint32_t N, uint32_t groups,
                                                         const u64* __restrict__ ten, u64* __restrict__ res) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)groups * N) return;
    const uint32_t g = tid / N, n = tid % N;
    const int L = tab->L, Lp = tab->Lp, LT = L + Lp;
    u64 xq[PSI_MAX_LIMBS], yp[PSI_MAX_LIMBS], out[PSI_MAX_LIMBS];
    double nu = 0.5;
#pragma unroll
    for (int i = 0; i < PSI_MAX_LIMBS; i++) {
        if (i < L) {
            xq[i] = ten[((size_t)g * LT + i) * N + n];
            nu = nu_step(nu, tab->tPSfrac[i], __ull2double_rn(xq[i]), tab->fp_fma);
        }
    }
    const u64 alpha = __double2ull_rz(nu);
#pragma unroll
    for (int j = 0; j < PSI_MAX_LIMBS; j++) {
        if (j < Lp) {
            const ModDev& m = tab->mods[L + j];
            u64 hi = 0, lo = 0;
#pragma unroll
            for (int i = 0; i < PSI_MAX_LIMBS; i++)
                if (i < L) mac128(hi, lo, xq[i], tab->tPS[j][i]);
            mac128(hi, lo, ten[((size_t)g * LT + L + j) * N + n], tab->tPS[j][L]);
            const u64 v = barrett128(hi, lo, m.q, m.mu_hi, m.mu_lo);
            yp[j] = addmod(v, alpha % m.q, m.q);
        }
    }
    switch_basis<false>(tab, Lp, L, yp, out);
#pragma unroll
    for (int l = 0; l < PSI_MAX_LIMBS; l++)
        if (l < L) res[((size_t)g * L + l) * N + n] = out[l];
}

// ---- HYBRID key switching (OpenFHE KeySwitchHYBRID as recalled; integers only, no floating point) ----------------
// c2 (COEFFICIENT, [B][3][L][N] component 2) is cut into ks_parts digits of ks_alpha consecutive limbs.  Digit j is
// lifted to every other limb of the extended basis Q + pk by ApproxSwitchCRTBasis (no rounding correction); its own
// limbs keep the coefficients.  dig: [B][parts][L+Lk][N] COEFFICIENT (the caller transforms it).
__global__ void __launch_bounds__(256) k_hybrid_modup(const DevTables* __restrict__ tab, uint32_t N, uint32_t B,
                                                      const u64* __restrict__ res, u64* __restrict__ dig) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)B * N) return;
    const size_t bin = tid / N;
    const uint32_t n = tid % N;
    const int L = tab->L, Lk = tab->Lk, LE = L + Lk, parts = tab->ks_parts, alpha = tab->ks_alpha;
    u64 x[PSI_MAX_LIMBS], y[PSI_MAX_LIMBS];
#pragma unroll
    for (int i = 0; i < PSI_MAX_LIMBS; i++)
        if (i < L) {
            x[i] = res[((bin * 3 + 2) * L + i) * N + n];
            y[i] = mul_shoup(x[i], tab->PartQHatInvModq[i], tab->PartQHatInvModq_s[i], tab->mods[i].q);
        }
    for (int j = 0; j < parts; j++) {
        const int lo = j * alpha, hi = min(L, lo + alpha);
        for (int m = 0; m < LE; m++) {
            u64 v;
            if (m >= lo && m < hi) {
                v = x[m];
            } else {
                const ModDev& md = tab->mods[m < L ? m : tab->L + tab->Lp + 1 + (m - L)];
                u64 h = 0, l = 0;
#pragma unroll
                for (int i = 0; i < PSI_MAX_LIMBS; i++)
                    if (i >= lo && i < hi) mac128(h, l, y[i], tab->PartQHatModt[i][m]);
                v = barrett128(h, l, md.q, md.mu_hi, md.mu_lo);
            }
            dig[((bin * parts + j) * LE + m) * N + n] = v;
        }
    }
}

// ext[bin][comp][m][n] = sum_j dig[bin][j][m][n] * evk_comp[j][m][n]  over the extended basis, EVALUATION
__global__ void __launch_bounds__(256) k_hybrid_inner(const DevTables* __restrict__ tab, uint32_t N, uint32_t B,
                                                      const u64* __restrict__ dig, const u64* __restrict__ evk_b,
                                                      const u64* __restrict__ evk_a, u64* __restrict__ ext) {
    const int L = tab->L, LE = L + tab->Lk, parts = tab->ks_parts;
    const size_t LEN = (size_t)LE * N;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)B * LEN) return;
    const size_t bin = tid / LEN, c = tid % LEN;
    const int m = (int)(c / N);
    const ModDev& md = tab->mods[m < L ? m : tab->L + tab->Lp + 1 + (m - L)];
    u64 h0 = 0, l0 = 0, h1 = 0, l1 = 0;
    for (int j = 0; j < parts; j++) {
        const u64 d = dig[(bin * parts + j) * LEN + c];
        mac128(h0, l0, d, evk_b[(size_t)j * LEN + c]);
        mac128(h1, l1, d, evk_a[(size_t)j * LEN + c]);
    }
    ext[(bin * 2) * LEN + c] = barrett128(h0, l0, md.q, md.mu_hi, md.mu_lo);
    ext[(bin * 2 + 1) * LEN + c] = barrett128(h1, l1, md.q, md.mu_hi, md.mu_lo);
}

// ApproxModDown, first half: the special-prime limbs of ext (COEFFICIENT after the caller's inverse transform) are
// switched to Q by ApproxSwitchCRTBasis.  sw: [B][2][L][N] COEFFICIENT.
__global__ void __launch_bounds__(256) k_hybrid_moddown(const DevTables* __restrict__ tab, uint32_t N, uint32_t B,
                                                        const u64* __restrict__ ext, u64* __restrict__ sw) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)B * 2 * N) return;
    const size_t g = tid / N;  // bin * 2 + comp
    const uint32_t n = tid % N;
    const int L = tab->L, Lk = tab->Lk, LE = L + Lk;
    u64 y[PSI_MAX_LIMBS];
#pragma unroll
    for (int u = 0; u < PSI_MAX_LIMBS; u++)
        if (u < Lk)
            y[u] = mul_shoup(ext[(g * LE + L + u) * N + n], tab->PkHatInvModpk[u], tab->PkHatInvModpk_s[u],
                             tab->mods[tab->L + tab->Lp + 1 + u].q);
#pragma unroll
    for (int i = 0; i < PSI_MAX_LIMBS; i++)
        if (i < L) {
            const ModDev& md = tab->mods[i];
            u64 h = 0, l = 0;
#pragma unroll
            for (int u = 0; u < PSI_MAX_LIMBS; u++)
                if (u < Lk) mac128(h, l, y[u], tab->PkHatModq[u][i]);
            sw[(g * L + i) * N + n] = barrett128(h, l, md.q, md.mu_hi, md.mu_lo);
        }
}

// ApproxModDown, second half, + (c0, c1) + optional mask: out = c + (ext_q - sw) * Pk^-1, all EVALUATION.
// res_eval: [B][3][L][N] (components 0, 1), ext: [B][2][L+Lk][N], sw: [B][2][L][N], out: [B][2][L][N].
__global__ void __launch_bounds__(256) k_hybrid_finish(const DevTables* __restrict__ tab, uint32_t N, uint32_t B,
                                                       const u64* __restrict__ res_eval, const u64* __restrict__ ext,
                                                       const u64* __restrict__ sw, const u64* __restrict__ mask,
                                                       u64* __restrict__ out) {
    const int L = tab->L, LE = L + tab->Lk;
    const size_t LN = (size_t)L * N;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)B * LN) return;
    const size_t bin = tid / LN, c = tid % LN;
    const int i = (int)(c / N);
    const ModDev& m = tab->mods[i];
#pragma unroll
    for (int comp = 0; comp < 2; comp++) {
        const u64 e = ext[((bin * 2 + comp) * LE) * N + c];
        const u64 d = submod(e, sw[(bin * 2 + comp) * LN + c], m.q);
        u64 r = addmod(res_eval[(bin * 3 + comp) * LN + c], mul_shoup(d, tab->PkInvModq[i], tab->PkInvModq_s[i], m.q), m.q);
        if (mask) r = mulmod(r, mask[bin * LN + c], m);
        out[(bin * 2 + comp) * LN + c] = r;
    }
}

// Centred lift of packed-encoding coefficients: crt [n][N] in [0, t) -> out [n][L][N], limb l holds c for
// c <= t/2 and q_l - (t - c) otherwise (the signed representative of c mod t, reduced mod q_l).
__global__ void __launch_bounds__(256) k_centre_lift(const DevTables* __restrict__ tab, uint32_t N, uint32_t n,
                                                     const u64* __restrict__ crt, u64* __restrict__ out) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)n * N) return;
    const size_t p = tid / N, i = tid % N;
    const u64 t = tab->t, v = crt[tid];
    const int L = tab->L;
    for (int l = 0; l < L; l++) out[(p * L + l) * N + i] = v > (t >> 1) ? v + (tab->mods[l].q - t) : v;
}

#define LAUNCH_1D(kernel, total, ...)                                 \
    kernel<<<cdiv((total), 256), 256, 0, k.s>>>(k.tab, k.N, __VA_ARGS__); \
    return cudaGetLastError();

cudaError_t launch_expand_q_to_p(const KCtx& k, uint32_t groups, const u64* coef, u64* ext) {
    LAUNCH_1D(k_expand_q_to_p, (size_t)groups * k.N, groups, coef, ext)
}
cudaError_t launch_fast_expand_poverq(const KCtx& k, uint32_t groups, const u64* coef, u64* ext) {
    LAUNCH_1D(k_fast_expand_poverq, (size_t)groups * k.N, groups, coef, ext)
}
cudaError_t launch_tensor(const KCtx& k, uint32_t B, const u64* e1, const u64* e2, u64* ten) {
    LAUNCH_1D(k_tensor, (size_t)B * (k.L + k.Lp) * k.N, B, e1, e2, ten)
}
cudaError_t launch_scale_round(const KCtx& k, uint32_t groups, const u64* ten, u64* res) {
    LAUNCH_1D(k_scale_round, (size_t)groups * k.N, groups, ten, res)
}
cudaError_t launch_scale_round_hps(const KCtx& k, uint32_t groups, const u64* ten, u64* res) {
    LAUNCH_1D(k_scale_round_hps, (size_t)groups * k.N, groups, ten, res)
}
cudaError_t launch_hybrid_modup(const KCtx& k, uint32_t B, const u64* res, u64* dig) {
    LAUNCH_1D(k_hybrid_modup, (size_t)B * k.N, B, res, dig)
}
cudaError_t launch_hybrid_inner(const KCtx& k, uint32_t B, const u64* dig, const u64* evk_b, const u64* evk_a, u64* ext) {
    // the extended basis has L + Lk limbs; Lk comes from the tables
    k_hybrid_inner<<<cdiv((size_t)B * (k.L + k.Lk) * k.N, 256), 256, 0, k.s>>>(k.tab, k.N, B, dig, evk_b, evk_a, ext);
    return cudaGetLastError();
}
cudaError_t launch_hybrid_moddown(const KCtx& k, uint32_t B, const u64* ext, u64* sw) {
    LAUNCH_1D(k_hybrid_moddown, (size_t)B * 2 * k.N, B, ext, sw)
}
cudaError_t launch_hybrid_finish(const KCtx& k, uint32_t B, const u64* res_eval, const u64* ext, const u64* sw, const u64* mask,
                                 u64* out) {
    LAUNCH_1D(k_hybrid_finish, (size_t)B * k.L * k.N, B, res_eval, ext, sw, mask, out)
}
cudaError_t launch_relin_digits(const KCtx& k, uint32_t B, const u64* res, u64* dig) {
    LAUNCH_1D(k_relin_digits, (size_t)B * k.L * k.N, B, res, dig)
}
cudaError_t launch_relin_accum(const KCtx& k, uint32_t B, const u64* res_eval, const u64* dig, const u64* evk_b,
                               const u64* evk_a, const u64* mask, u64* out) {
    LAUNCH_1D(k_relin_accum, (size_t)B * k.L * k.N, B, res_eval, dig, evk_b, evk_a, mask, out)
}
cudaError_t launch_mul_ctpt(const KCtx& k, uint32_t B, const u64* ct, const u64* pt, u64* out) {
    LAUNCH_1D(k_mul_ctpt, (size_t)B * k.L * k.N, B, ct, pt, out)
}
cudaError_t launch_centre_lift(const KCtx& k, uint32_t n, const u64* crt, u64* out) {
    LAUNCH_1D(k_centre_lift, (size_t)n * k.N, n, crt, out)
}
cudaError_t launch_slots_to_crt(const KCtx& k, uint32_t n_pt, uint32_t nslots, const long long* slots,
                                const uint32_t* to_crt, u64* out) {
    LAUNCH_1D(k_slots_to_crt, (size_t)n_pt * k.N, n_pt, nslots, slots, to_crt, out)
}

// ------------------------------------------------------------------------------------------
// Integer-pipe micro-benchmarks: the denominators of the integer roofline are MEASURED, not assumed.
//   kind 0: IMAD.WIDE.U32 (32x32 + 64 -> 64), the instruction the lazy inner product is made of
//   kind 1: Harvey/Shoup lazy butterflies on 64-bit residues, the unit of work of the NTT
// Independent dependency chains per thread, no memory traffic.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_imad_peak(u64* out, uint32_t iters, uint32_t seed) {
    // eight accumulate chains; every multiplier is the high word of its own accumulator, so nothing is loop
    // invariant (with constant operands ptxas hoists the product and the loop degenerates into 64-bit adds)
    const uint32_t b = (blockIdx.x * 256u + threadIdx.x) * 40503u + 12345u + seed;
    u64 c[8];
#pragma unroll
    for (int u = 0; u < 8; u++) c[u] = out[(blockIdx.x * 256u + threadIdx.x + u) % 1024u] + seed;
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int u = 0; u < 8; u++) c[u] = madw((uint32_t)(c[u] >> 32), b, c[u]);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = c[0] ^ c[1] ^ c[2] ^ c[3] ^ c[4] ^ c[5] ^ c[6] ^ c[7];
}

__global__ void __launch_bounds__(256) k_butterfly_peak(u64* out, uint32_t iters, u64 q, u64 w, u64 ws) {
    const u64 q2 = 2 * q;
    u64 x[4], y[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        x[k] = (threadIdx.x * 2654435761ull + k * 977 + blockIdx.x) % q;
        y[k] = (threadIdx.x * 40503ull + k * 131 + 7 * blockIdx.x) % q;
    }
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            u64 u = x[k];
            if (u >= q2) u -= q2;
            const u64 v = mul_shoup_lazy(y[k], w, ws, q);
            x[k] = u + v;
            y[k] = u - v + q2;
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x[0] ^ x[1] ^ x[2] ^ x[3] ^ y[0] ^ y[1] ^ y[2] ^ y[3];
}

// kind 2: the hand-scheduled butterfly (mul_shoup_lazy_nq + approximate lazy correction) with
// per-thread twiddles, as the fused NTT kernels use it
__global__ void __launch_bounds__(256) k_butterfly2_peak(u64* out, uint32_t iters, u64 q, u64 w_in, u64 ws_in) {
    const u64 q2 = 2 * q, nq = 0 - q;
    u64 x[4], y[4];
    // per-thread twiddle so that nothing lands on the uniform datapath
    const u64 w = (w_in + threadIdx.x) % q;
    const u64 ws = (u64)(((unsigned __int128)w << 64) / q);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        x[k] = (threadIdx.x * 2654435761ull + k * 977 + blockIdx.x) % q;
        y[k] = (threadIdx.x * 40503ull + k * 131 + 7 * blockIdx.x) % q;
    }
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const u64 u = lazy_sub_hi(x[k], q2);
            const u64 v = mul_shoup_lazy_nq(y[k], w, ws, nq);
            x[k] = u + v;
            y[k] = u - v + q2;
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x[0] ^ x[1] ^ x[2] ^ x[3] ^ y[0] ^ y[1] ^ y[2] ^ y[3];
}

// kind 3: the same butterfly as one explicit PTX block
__global__ void __launch_bounds__(256) k_butterfly3_peak(u64* out, uint32_t iters, u64 q, u64 w_in, u64 ws_in) {
    const u64 q2 = 2 * q, nq = 0 - q;
    u64 x[4], y[4];
    const u64 w = (w_in + threadIdx.x) % q;
    const u64 ws = (u64)(((unsigned __int128)w << 64) / q);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        x[k] = (threadIdx.x * 2654435761ull + k * 977 + blockIdx.x) % q;
        y[k] = (threadIdx.x * 40503ull + k * 131 + 7 * blockIdx.x) % q;
    }
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) ct_butterfly_asm(x[k], y[k], w, ws, q2, nq);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x[0] ^ x[1] ^ x[2] ^ x[3] ^ y[0] ^ y[1] ^ y[2] ^ y[3];
}

// kinds 4 / 5: the data exchange between two radix passes in isolation.  16 lanes hold 16 u64 each (lane l holds
// elements 16 l .. 16 l + 15 of a 256-element block) and end up with the transpose (lane l holds elements l, l + 16, ...):
//   kind 4: through padded shared memory, as the fused kernels do it: 16 STS.64, barrier, 16 LDS.64
//   kind 5: through the register file: four butterfly steps with __shfl_xor, each moving half of the values
//           (2 x 8 SHFL per step and a select per moved word)
// A 64-bit add between exchanges keeps the values live.  Returns exchanges (16 values per thread) per second.
// kind 6: the Shoup QUOTIENT on the FP64 pipe (idea 1b) - a throughput and exactness probe, not used by any
// product kernel.  floor(y * w / q) is formed from an error-free double product instead of the 4-product mulhi64:
//   y = y1 * 2^31 + y0 (both halves exact doubles through the 2^52 bit trick), rho = w / q as a double-double scaled by 2^31:
//   P = y1 * floor(w 2^31 / q) (one exact 32 x 32 integer product), s = fma(d0, rho, d1 * a_lo) < 2^32 in doubles,
//   h = P + floor(s) - 1, the floor taken by a round-down add of 2^52 (the first version converted with F2I and ran at
//   5.0e11/s: the conversions sit on the quarter-rate XU pipe)
// The estimate is never above the true quotient and at most 2 below it, so the lazy product y w - h q lies in [0, 4q)
// (checked for every butterfly when `check` is set; violations are counted into out[0]).  What remains on the multiplier
// pipe are the two low 64-bit products.
__device__ __forceinline__ double u31_to_double(uint32_t v) { return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0; }
__global__ void __launch_bounds__(256) k_butterfly_fp64_peak(u64* out, uint32_t iters, u64 q, u64 w_in, uint32_t check) {
    const u64 q2 = 2 * q, q4 = 4 * q;
    const u64 w = (w_in + threadIdx.x) % q;
    // rho * 2^31 as a double-double, rho as a double (host-side constants in a product kernel; computed here once)
    const double qd = (double)q;  // q < 2^60 is not exact in a double: refine rho with one exact remainder step below
    const double rho = (double)w / qd;
    // exact double-double of w * 2^31 / q: hi = fl(.), lo from the integer remainder
    const unsigned __int128 num = (unsigned __int128)w << 31;
    const u64 ihi = (u64)(num / q);
    const u64 rem = (u64)(num % q);
    const uint32_t ihi32 = (uint32_t)ihi;                             // ihi < 2^31 * (w / q) < 2^31
    const double a_lo = (double)rem / qd;                             // in [0, 1): the fractional part of w 2^31 / q
    unsigned long long bad = 0;
    u64 x[4], y[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        x[k] = (threadIdx.x * 2654435761ull + k * 977 + blockIdx.x) % q;
        y[k] = (threadIdx.x * 40503ull + k * 131 + 7 * blockIdx.x) % q;
    }
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            u64 u = x[k];
            if (u >= q4) u -= q4;
            if (u >= q2) u -= q2;
            const u64 yy = y[k];  // < 2^62
            const uint32_t y1 = (uint32_t)(yy >> 31), y0 = (uint32_t)yy & 0x7fffffffu;
            // y * rho = y1 * (ihi + rem / q) + y0 * rho: the big part y1 * ihi is ONE exact 32 x 32 integer product, the
            // rest (< 2^32) is formed in doubles and floored by a round-down add of 2^52 (no F2I: the XU pipe is slow)
            const u64 P = (u64)y1 * ihi32;
            const double d1 = u31_to_double(y1), d0 = u31_to_double(y0);
            const double sfrac = __fma_rn(d0, rho, __dmul_rn(d1, a_lo));
            const u64 S = (u64)__double_as_longlong(__dadd_rd(sfrac, 4503599627370496.0)) & 0xfffffffffffffull;
            const u64 h = P + S - 1;
            const u64 v = yy * w - h * q;  // in [0, 4q) when h is the quotient or up to 2 below it
            if (check && v >= q4) bad++;
            x[k] = u + v;
            y[k] = u - v + q4;
            if (y[k] >= q4) y[k] -= q4;  // keep y below 2^62 for the split (the product kernels track this bound statically)
        }
    }
    u64 r = x[0] ^ x[1] ^ x[2] ^ x[3] ^ y[0] ^ y[1] ^ y[2] ^ y[3];
    if (check) r = bad;
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void __launch_bounds__(256) k_exchange_smem_peak(u64* out, uint32_t iters) {
    __shared__ u64 sm[256 * 17 + 16];
    u64 v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = threadIdx.x * 16 + k + blockIdx.x;
    const uint32_t lane16 = threadIdx.x & 15, grp = threadIdx.x >> 4;
    u64* base = sm + grp * (16 * 17);
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) base[lane16 * 17 + k] = v[k];   // conflict-free: lane stride 17 words
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = base[k * 17 + lane16] + i;
        __syncwarp();
    }
    u64 acc = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) acc ^= v[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(256) k_exchange_shfl_peak(u64* out, uint32_t iters) {
    u64 v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = threadIdx.x * 16 + k + blockIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int step = 0; step < 4; step++) {
            const int m = 8 >> step;  // lane distance and value distance of this step
            const bool upper = (lane & m) != 0;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (k & m) continue;
                // the lower lane keeps v[k] and receives the partner's v[k]; it sends v[k + m]; the upper lane mirrors
                const u64 send = upper ? v[k] : v[k + m];
                const u64 got = __shfl_xor_sync(0xffffffffu, send, m);
                if (upper)
                    v[k] = got;
                else
                    v[k + m] = got;
            }
        }
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] += i;
    }
    u64 acc = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) acc ^= v[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

cudaError_t pipe_peak(int device, int kind, double* per_second) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return e;
    // kind bits 4..7: resident 256-thread blocks per SM (0 = 8, i.e. 16 warps per scheduler); the fused kernels run at 3
    const unsigned per_sm = (kind >> 4) & 15 ? (kind >> 4) & 15 : 8;
    kind &= 15;
    const unsigned blocks = prop.multiProcessorCount * per_sm, threads = 256;
    const uint32_t iters = kind == 0 ? 4096 : 2048;
    u64* out = nullptr;
    if ((e = cudaMalloc(&out, (size_t)blocks * threads * sizeof(u64))) != cudaSuccess) return e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    float best = 1e30f;
    const u64 q = 1152921504606748673ull, w = 123456789123456789ull % q;
    const u64 ws = (u64)(((unsigned __int128)w << 64) / q);
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(t0);
        if (kind == 0)
            k_imad_peak<<<blocks, threads>>>(out, iters, rep);
        else if (kind == 1)
            k_butterfly_peak<<<blocks, threads>>>(out, iters, q, w, ws);
        else if (kind == 2)
            k_butterfly2_peak<<<blocks, threads>>>(out, iters, q, w, ws);
        else if (kind == 4)
            k_exchange_smem_peak<<<blocks, threads>>>(out, iters);
        else if (kind == 5)
            k_exchange_shfl_peak<<<blocks, threads>>>(out, iters);
        else if (kind == 6 || kind == 7)  // 7: exactness check pass (per_second then holds the number of violations)
            k_butterfly_fp64_peak<<<blocks, threads>>>(out, iters, q, w, kind == 7);
        else
            k_butterfly3_peak<<<blocks, threads>>>(out, iters, q, w, ws);
        cudaEventRecord(t1);
        if ((e = cudaEventSynchronize(t1)) != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, t0, t1);
        if (rep > 0 && ms < best) best = ms;
    }
    if (kind == 7 && e == cudaSuccess) {  // sum the violation counters
        std::vector<u64> host((size_t)blocks * threads);
        e = cudaMemcpy(host.data(), out, host.size() * sizeof(u64), cudaMemcpyDeviceToHost);
        double total = 0;
        for (u64 v : host) total += (double)v;
        cudaEventDestroy(t0);
        cudaEventDestroy(t1);
        cudaFree(out);
        *per_second = total;
        return e;
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(out);
    if (e != cudaSuccess) return e;
    const double per_thread = kind == 0 ? (double)iters * 64.0 : ((kind == 4 || kind == 5) ? (double)iters : (double)iters * 4.0);
    *per_second = (double)blocks * threads * per_thread / (best * 1e-3);
    return cudaSuccess;
}

}  // namespace psi
