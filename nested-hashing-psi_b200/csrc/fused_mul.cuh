// EvalMult(ct,ct) + relinearise + mask as FIVE fused sm_100a kernels (reference call sites:
// cryptoContext->EvalMult(multipliedResult, innerProductResult), BatchedFHEHIPPIE.cpp:123, and
// EvalMult(.., preCalcRandomMask[bin]), :126).
//
// The 2^14-point negacyclic NTT is split "four-step" style into a COLUMN pass (the log2(N)-7 stages
// whose butterfly stride is >= 128 coefficients) and a ROW pass (the 7 stages inside a 128-coefficient
// row).  A column tile (all rows x 8 columns) or a row tile (8 rows x 128 columns) of one limb is 8 KiB
// and lives in shared memory; every stage range is executed as radix-16 / radix-8 register passes.
// Because the coefficient-wise RNS operations (base extension, scale-and-round, digit lift) need all
// limbs of a coefficient but only that coefficient, they fuse with the column passes on either side;
// the slot-wise operations (tensor product, key-switch inner product, mask) fuse with the row passes:
//
//   k_rows_inv        EVAL operands            -> row-inverse halves                   (4L limb-polys / bin)
//   k_cols_extend     column-inverse, Q->P exact extension (1st operand) or P-over-Q fast extension
//                     (2nd operand), column-forward of the new limbs
//   k_rows_tensor     row-forward, tensor product (c0c0', c0c1'+c1c0', c1c1'), row-inverse
//   k_cols_scale      column-inverse, scale-and-round t/P back to Q, BV digit lift, column-forward
//   k_rows_relin      row-forward, sum_i digit_i * evk_i + (c0, c1), mask multiply
//
// Each coefficient crosses HBM/L2 ten times per ciphertext multiplication instead of ~19 with one
// kernel per OpenFHE call, and 88 limb-NTTs per bin run out of shared memory.  No tensor cores: this is
// 64-bit residue arithmetic on the integer pipes.  Between stages values are kept lazily reduced
// (Harvey); every value that feeds a double-precision rounding decision or leaves the pipeline is the
// canonical residue, so results are bit-identical to the unfused reference sequence (oracle/psi_oracle.c).
#pragma once
#include "psi_kernels.cuh"
#include "async_copy.cuh"

namespace psi {

constexpr uint32_t kLogCols = 7;     // row length 2^7 coefficients
constexpr uint32_t kRowTileLog = 3;  // 8 rows per row tile
constexpr uint32_t kColTileLog = 3;  // 8 columns per column tile
constexpr uint32_t kGroup = 64;      // threads cooperating on one shared-memory array
constexpr uint32_t kColGroups = 4;   // groups per CTA in the column kernels (arrays are dealt round-robin)

// padded shared-memory slot: one pad word per 16 coefficients keeps both the strided gathers and the
// 16-contiguous-per-thread pattern of the last radix pass off a single bank group
__device__ __forceinline__ uint32_t sl(uint32_t i) { return i + (i >> 4); }
__host__ __device__ constexpr uint32_t padded(uint32_t n) { return n + (n >> 4) + 1; }

// The arrays of a transform are dealt to 64-thread groups and a group only ever touches its own arrays, so
// the barrier between two register passes is a NAMED barrier over the group's two warps (ids 1..15; id 0 is
// __syncthreads), not a CTA-wide one: groups drift freely and idle groups do not wait at all.  Callers
// place a CTA-wide barrier where the next phase reads across arrays.
__device__ __forceinline__ void group_sync() {
    asm volatile("bar.sync %0, %1;" ::"r"(1u + threadIdx.x / kGroup), "r"(kGroup) : "memory");
}

// ---- lazy butterflies ------------------------------------------------------------------------
// ncu shows fmaheavy as the busiest pipe (every IMAD.WIDE occupies it for four cycles), so a variant of the Shoup
// product with an APPROXIMATE high product was built:
//   floor(y * w' / 2^64)  ~  y1 w'1 + hi32(y1 w'0) + hi32(y0 w'1)          (three IMAD.WIDE instead of four)
// which is the exact quotient or up to 2 below it, i.e. the lazy product lands in [0, 4q) instead of [0, 2q).
// All lazy values of the transforms therefore live below 2 * kLazy * q = 8q (q < 2^60: 8q < 2^63), corrections
// subtract kLazy * q, and whoever consumes a transform output either reduces exactly (shoup_canon, reduce_pow2q) or
// brings it below 2q first (lazy_below_2q).  PSI_BF_APPROX = 0 is the exact four-product form (values < 4q).
// MEASURED (round 2, config B, both builds bit-exact against the oracle): phase 2 takes 1.051 ms with the
// approximate product and 1.054 ms with the exact one (2 bin groups: 1.014 / 1.012) — the multiplier pipe is NOT
// what bounds these kernels, so the exact form stays the default and the approximate one an option (make EXTRA=-DPSI_BF_APPROX=1).
#ifndef PSI_BF_APPROX
#define PSI_BF_APPROX 0
#endif
#if PSI_BF_APPROX
constexpr u64 kLazy = 4;
__device__ __forceinline__ u64 bf_mul(u64 y, const ulonglong2 tw, u64 nq) {
    const uint32_t y0 = lo32(y), y1 = hi32(y), s0 = lo32(tw.y), s1 = hi32(tw.y);
    const u64 a = madw32(y1, s0, 0), b = madw32(y0, s1, 0);
    const u64 h = madw32(y1, s1, (u64)hi32(a)) + (u64)hi32(b);
    // low 64 bits of y * w + h * (2^64 - q): two IMAD.WIDE for the low words, four 32-bit IMADs into the high word
    const uint32_t w0 = lo32(tw.x), w1 = hi32(tw.x), h0 = lo32(h), h1 = hi32(h), n0 = lo32(nq), n1 = hi32(nq);
    u64 r = madw32(y0, w0, 0);
    r = madw32(h0, n0, r);
    uint32_t top = hi32(r);
    top += y0 * w1;
    top += y1 * w0;
    top += h0 * n1;
    top += h1 * n0;
    return ((u64)top << 32) | lo32(r);
}
#else
constexpr u64 kLazy = 2;
__device__ __forceinline__ u64 bf_mul(u64 y, const ulonglong2 tw, u64 nq) { return mul_shoup_lazy(y, tw.x, tw.y, 0 - nq); }
#endif
// a transform output (< 2 * kLazy * q + slack) brought below 2q + 2^32
__device__ __forceinline__ u64 lazy_below_2q(u64 x, u64 q) {
#if PSI_BF_APPROX
    x = lazy_sub_hi(x, 4 * q);
#endif
    return lazy_sub_hi(x, 2 * q);
}
// forward (Cooley-Tukey): inputs < 2 qc + slack, outputs < 2 qc + slack, qc = kLazy * q
__device__ __forceinline__ void ct_bf(u64& x, u64& y, const ulonglong2 tw, u64 qc, u64 nq) {
    const u64 u = lazy_sub_hi(x, qc);
    const u64 v = bf_mul(y, tw, nq);
    x = u + v;
    y = u - v + qc;
}
// inverse (Gentleman-Sande): inputs < qc + e (e grows by at most a factor two per stage from 2^32,
// far below q after 14 stages), x output < qc + 2e, y output < qc
__device__ __forceinline__ void gs_bf(u64& x, u64& y, const ulonglong2 tw, u64 qc, u64 nq) {
    const u64 s = lazy_sub_hi(x + y, qc);
    const u64 d = x - y + 2 * qc;
    x = s;
    y = bf_mul(d, tw, nq);
}

// One stage (local stage sig0 + r) of a radix-2^R register pass.  All loop bounds are template constants
// so that every index into v[] is a compile-time constant (the array must stay in registers).
template <int R, int r, bool INV, bool TWS = false>
__device__ __forceinline__ void reg_stage(u64 (&v)[1 << R], const ulonglong2* __restrict__ tw, uint32_t w0, u64 q2, u64 nq) {
    constexpr int half = 1 << (R - 1 - r);
#pragma unroll
    for (int j = 0; j < (1 << r); j++) {
        const ulonglong2 t = TWS ? tw[w0 + j] : __ldg(tw + w0 + j);
#pragma unroll
        for (int i = 0; i < half; i++) {
            if (INV)
                gs_bf(v[j * 2 * half + i], v[j * 2 * half + i + half], t, q2, nq);
            else
                ct_bf(v[j * 2 * half + i], v[j * 2 * half + i + half], t, q2, nq);
        }
    }
}
// TWS: `tw` is the row tile's twiddle table staged in shared memory (stage_row_twiddles): the 8 * 2^u
// twiddles of row stage u start at entry 8 * (2^u - 1)
template <int R, int rr, bool INV, bool TWS = false>
struct RegStages {
    static __device__ __forceinline__ void run(u64 (&v)[1 << R], const ulonglong2* __restrict__ tw, uint32_t m, uint32_t sig0,
                                               uint32_t delta, uint32_t tile_base, uint32_t grp, u64 q2, u64 nq) {
        constexpr int r = INV ? R - 1 - rr : rr;
        // TWS: packed row-tile table (layout: psi_api.cu build_tables).  The radix pass whose blocks are
        // single threads' contiguous coefficients (lt == 0) reads its own 2^R - 1 twiddles back to back; the
        // other pass reads the stage-major part shared by the threads of a row.
        uint32_t w0;
        if (TWS) {
            const bool own = (m - sig0 - R) == 0;
            if (!INV)
                w0 = own ? 120u + ((1u << R) - 1u) * grp + ((1u << r) - 1u) : ((8u << (sig0 + r - kRowTileLog)) - 8u) + (grp << r);
            else
                w0 = own ? ((1u << R) - 1u) * grp + ((1u << r) - 1u) : 960u + ((8u << (sig0 + r - kRowTileLog)) - 8u) + (grp << r);
        } else {
            w0 = (1u << (sig0 + r + delta)) + (tile_base >> (m - sig0 - r)) + (grp << r);
        }
        reg_stage<R, r, INV, TWS>(v, tw, w0, q2, nq);
        RegStages<R, rr + 1, INV, TWS>::run(v, tw, m, sig0, delta, tile_base, grp, q2, nq);
    }
};
template <int R, bool INV, bool TWS>
struct RegStages<R, R, INV, TWS> {
    static __device__ __forceinline__ void run(u64 (&)[1 << R], const ulonglong2* __restrict__, uint32_t, uint32_t, uint32_t,
                                               uint32_t, uint32_t, u64, u64) {}
};

// Radix-2^R pass over local stages [sig0, sig0 + R) of a local array of 2^m coefficients.
// Global twiddle index of local stage sig, local group g:  2^(sig + delta) + (tile_base >> (m - sig)) + g
template <int R, bool INV>
__device__ __forceinline__ void radix_pass(u64* __restrict__ sm, const ulonglong2* __restrict__ tw, uint32_t m,
                                           uint32_t sig0, uint32_t delta, uint32_t tile_base, u64 q, uint32_t tid) {
    const uint32_t lt = m - sig0 - R;  // log2 of the stride between a thread's coefficients
    const u64 q2 = kLazy * q, nq = 0 - q;  // correction amount of the lazy butterflies
    for (uint32_t blk = tid; blk < ((1u << m) >> R); blk += kGroup) {
        const uint32_t off = blk & ((1u << lt) - 1), grp = blk >> lt;
        const uint32_t base = (grp << (m - sig0)) + off;
        u64 v[1 << R];
#pragma unroll
        for (int k = 0; k < (1 << R); k++) v[k] = sm[sl(base + (k << lt))];
        RegStages<R, 0, INV>::run(v, tw, m, sig0, delta, tile_base, grp, q2, nq);
#pragma unroll
        for (int k = 0; k < (1 << R); k++) sm[sl(base + (k << lt))] = v[k];
    }
}

// how many stages the next register pass takes when `rem` remain: 7 -> 4+3, 6 -> 3+3, 5 -> 3+2
__device__ __forceinline__ uint32_t pass_width(uint32_t rem) { return (rem >= 7 || rem == 4) ? 4 : (rem >= 3 ? 3 : rem); }

template <bool INV>
__device__ __forceinline__ void one_pass(uint32_t w, uint32_t s, u64* sm, const ulonglong2* tw, uint32_t m, uint32_t delta,
                                         uint32_t tile_base, u64 q, uint32_t tid) {
    if (w == 4) radix_pass<4, INV>(sm, tw, m, s, delta, tile_base, q, tid);
    else if (w == 3) radix_pass<3, INV>(sm, tw, m, s, delta, tile_base, q, tid);
    else if (w == 2) radix_pass<2, INV>(sm, tw, m, s, delta, tile_base, q, tid);
    else radix_pass<1, INV>(sm, tw, m, s, delta, tile_base, q, tid);
}

// Stages [lo, hi) of n_arr arrays (ascending for the forward, descending for the inverse transform).
// Array `a` lives at smem + a * stride and uses modulus index mod_of(a); group g of ng handles arrays
// g, g + ng, ...  Every thread of the CTA must call (CTA-wide barrier after each register pass).
template <bool INV, typename ModOf>
__device__ __forceinline__ void transform(const DevTables* __restrict__ tab, u64* smem, uint32_t stride, uint32_t n_arr,
                                          ModOf mod_of, uint32_t g, uint32_t ng, uint32_t m, uint32_t lo, uint32_t hi,
                                          uint32_t delta, uint32_t tile_base, uint32_t tid) {
    uint32_t rem = hi - lo;
    uint32_t s = INV ? hi : lo;
    while (rem) {
        const uint32_t w = pass_width(rem);
        const uint32_t s0 = INV ? s - w : s;
        for (uint32_t a = g; a < n_arr; a += ng) {
            const ModDev& md = tab->mods[mod_of(a)];
            one_pass<INV>(w, s0, smem + a * stride, INV ? md.itw : md.ftw, m, delta, tile_base, md.q, tid);
        }
        s = INV ? s - w : s + w;
        rem -= w;
        if (g < n_arr) group_sync();
    }
}

// ---- compile-time variant: array size, stage range and radix are template constants, so every
// shared-memory offset is an immediate, the pass plan has no run-time dispatch, and twiddle addresses that
// do not depend on the thread (column passes) are computed once on the uniform datapath.
__host__ __device__ constexpr int ct_pass_width(int rem) { return (rem >= 7 || rem == 4) ? 4 : (rem >= 3 ? 3 : rem); }

template <int M, int SIG0, int R, bool INV, bool TWS = false>
__device__ __forceinline__ void radix_pass_ct(u64* __restrict__ sm, const ulonglong2* __restrict__ tw, uint32_t delta,
                                              uint32_t tile_base, u64 q, uint32_t tid) {
    constexpr int LT = M - SIG0 - R;
    constexpr uint32_t NBLK = (1u << M) >> R;
    const u64 q2 = kLazy * q, nq = 0 - q;  // correction amount of the lazy butterflies
#pragma unroll
    for (uint32_t it = 0; it < (NBLK + kGroup - 1) / kGroup; it++) {
        const uint32_t blk = tid + it * kGroup;
        if ((NBLK % kGroup) != 0 && blk >= NBLK) break;
        const uint32_t off = blk & ((1u << LT) - 1), grp = blk >> LT;
        // sl(base + k * 2^LT) == sl(base) + k * 2^LT + ((k * 2^LT) >> 4): the pad of the k-th coefficient
        // never carries into the next 16-block
        u64* p0 = sm + sl((grp << (M - SIG0)) + off);
        u64 v[1 << R];
#pragma unroll
        for (int k = 0; k < (1 << R); k++) v[k] = p0[(k << LT) + ((k << LT) >> 4)];
        RegStages<R, 0, INV, TWS>::run(v, tw, M, SIG0, delta, tile_base, grp, q2, nq);
#pragma unroll
        for (int k = 0; k < (1 << R); k++) p0[(k << LT) + ((k << LT) >> 4)] = v[k];
    }
}

template <bool INV, int M, int LO, int HI, bool TWS = false>
struct TransformCT {
    template <typename ModOf>
    static __device__ __forceinline__ void run(const DevTables* __restrict__ tab, u64* smem, uint32_t stride, uint32_t n_arr,
                                               ModOf mod_of, uint32_t g, uint32_t ng, uint32_t delta, uint32_t tile_base,
                                               uint32_t tid, const ulonglong2* tws = nullptr) {
        constexpr int W = ct_pass_width(HI - LO);
        constexpr int S0 = INV ? HI - W : LO;
        for (uint32_t a = g; a < n_arr; a += ng) {
            const ModDev& md = tab->mods[mod_of(a)];
            radix_pass_ct<M, S0, W, INV, TWS>(smem + a * stride, TWS ? tws : (INV ? md.itw : md.ftw), delta, tile_base, md.q,
                                              tid);
        }
        if (g < n_arr) group_sync();
        TransformCT<INV, M, INV ? LO : LO + W, INV ? HI - W : HI, TWS>::run(tab, smem, stride, n_arr, mod_of, g, ng, delta,
                                                                           tile_base, tid, tws);
    }
};
template <bool INV, int M, int LO, bool TWS>
struct TransformCT<INV, M, LO, LO, TWS> {
    template <typename ModOf>
    static __device__ __forceinline__ void run(const DevTables* __restrict__, u64*, uint32_t, uint32_t, ModOf, uint32_t, uint32_t,
                                               uint32_t, uint32_t, uint32_t, const ulonglong2* = nullptr) {}
};

// Row-tile twiddles through the copy engine: the 1016 {w, w'} pairs the 7 row stages of one tile need are
// pre-packed per tile in read order (ModDev::ftw_rows / itw_rows), so one 16 KiB bulk copy brings them.
constexpr uint32_t kRowTwEntries = 8u * 127u;
constexpr uint32_t kRowTwWords = 2u * kRowTwEntries;  // u64 words per staged table
__device__ __forceinline__ void stage_row_twiddles(ulonglong2* dst, const ulonglong2* __restrict__ packed, uint32_t tile,
                                                   uint64_t* bar) {
    bulk_g2s(dst, packed + (size_t)tile * kRowTwEntries, kRowTwEntries * 16u, bar);
}

// row tiles are always 2^10 coefficients, stages [3, 10): one compiled plan serves every ring dimension
template <bool INV, typename ModOf>
__device__ __forceinline__ void transform_rows(const DevTables* __restrict__ tab, u64* smem, uint32_t stride, uint32_t n_arr,
                                               ModOf mod_of, uint32_t g, uint32_t ng, uint32_t logN, uint32_t tile_base,
                                               uint32_t tid, const ulonglong2* tws) {
    constexpr int M = kLogCols + kRowTileLog;
    TransformCT<INV, M, kRowTileLog, M, true>::run(tab, smem, stride, n_arr, mod_of, g, ng, logN - M, tile_base, tid, tws);
}
// column tiles: LOGN_CT != 0 selects the compiled plan for that ring dimension, 0 the run-time plan
template <bool INV, int LOGN_CT, typename ModOf>
__device__ __forceinline__ void transform_cols(const DevTables* __restrict__ tab, u64* smem, uint32_t stride, uint32_t n_arr,
                                               ModOf mod_of, uint32_t g, uint32_t ng, uint32_t logN, uint32_t tid) {
    if constexpr (LOGN_CT != 0) {
        constexpr int LOGR = LOGN_CT - (int)kLogCols, M = LOGR + (int)kColTileLog;
        TransformCT<INV, M, 0, LOGR>::run(tab, smem, stride, n_arr, mod_of, g, ng, 0, 0, tid);
    } else {
        const uint32_t logR = logN - kLogCols;
        transform<INV>(tab, smem, stride, n_arr, mod_of, g, ng, logR + kColTileLog, 0, logR, 0, 0, tid);
    }
}

// ---- small modular helpers ----------------------------------------------------------------------
// canonical residue of a < 2^K * q by K exact compare-subtract steps
template <int K>
__device__ __forceinline__ u64 reduce_pow2q(u64 a, u64 q) {
#pragma unroll
    for (int k = K - 1; k >= 0; k--) {
        const u64 kq = q << k;
        if (a >= kq) a -= kq;
    }
    return a;
}
__device__ __forceinline__ u64 shoup_lazy(u64 x, u64 c, u64 cs, u64 q) { return mul_shoup_lazy(x, c, cs, q); }
__device__ __forceinline__ u64 shoup_canon(u64 x, u64 c, u64 cs, u64 q) {
    const u64 r = shoup_lazy(x, c, cs, q);
    return r >= q ? r - q : r;
}

// RNS inner products sum_i x_i * c_i mod q with at most 8 terms: the constants are stored in Montgomery form
// (times 2^64 mod q) and split-30, the residues x_i < 2^60 are split once, every product is THREE IMAD.WIDE.U32
// (Karatsuba on the 30-bit halves: ll += x0 c0, hh += x1 c1, kk += (x0 + x1)(c0 + c1)) into carry-free partial
// sums, and ONE Montgomery reduction returns the canonical residue of the sum.  ll, hh < 8 * 2^60; kk may wrap,
// but the middle sum kk - ll - hh is < 8 * 2^61 = 2^64, so the wrapped 64-bit difference is exact.
struct Acc3 {
    u64 ll = 0, kk = 0, hh = 0;
};
__device__ __forceinline__ void acc3_mad(Acc3& a, uint32_t x0, uint32_t x1, u64 c_split) {
    const uint32_t c0 = (uint32_t)c_split, c1 = (uint32_t)(c_split >> 32);
    a.ll = madw32(x0, c0, a.ll);
    a.hh = madw32(x1, c1, a.hh);
    a.kk = madw32(x0 + x1, c0 + c1, a.kk);
}
__device__ __forceinline__ u64 acc3_reduce(const Acc3& a, u64 q, u64 qinv) {
    const u128 t = (u128)a.ll + ((u128)(a.kk - a.ll - a.hh) << 30) + ((u128)a.hh << 60);
    const u64 r = mont_redc_lazy((u64)(t >> 64), (u64)t, q, qinv);
    return r >= q ? r - q : r;
}

// streamed-once tile data must not evict the twiddles from L1
__device__ __forceinline__ ulonglong2 ld_tile16(const u64* p) {
    ulonglong2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}

// ---- tile movers (16-byte global accesses) ------------------------------------------------------
// row tile: 2^(7+kRowTileLog) contiguous coefficients
// PSI_ASYNC_LOADS: the tile loads are per-thread cp.async copies (SASS LDGSTS) straight into the padded layout: every
// copy of a tile in flight at once, no staging registers, one wait before the barrier that publishes the tile.
// MEASURED (round 2, config B): same kernel times as the register version (LDG.128 + two STS.64, unroll 4) to three
// digits -- an experiment build of k_rows_inv without any tile load or store runs in 65.4 us against 73.0, i.e. the
// global traffic is 10 % of that kernel and far less of the others; see DESIGN.md 3.2 "what bounds phase 2".
#ifndef PSI_ASYNC_LOADS
#define PSI_ASYNC_LOADS 1
#endif
__device__ __forceinline__ void load_rows(u64* sm, const u64* __restrict__ poly_tile, uint32_t tid) {
    const uint32_t M2 = 1u << (kLogCols + kRowTileLog - 1);
#if PSI_ASYNC_LOADS
#pragma unroll
    for (uint32_t j = tid; j < M2; j += kGroup) {
        cp_async8(sm + sl(2 * j), poly_tile + 2 * j);
        cp_async8(sm + sl(2 * j) + 1, poly_tile + 2 * j + 1);
    }
#else
#pragma unroll 4
    for (uint32_t j = tid; j < M2; j += kGroup) {
        const ulonglong2 v = ld_tile16(poly_tile + 2 * j);
        sm[sl(2 * j)] = v.x;
        sm[sl(2 * j) + 1] = v.y;
    }
#endif
}
// every thread calls this between its loads and the barrier that publishes the tile
__device__ __forceinline__ void loads_wait() {
#if PSI_ASYNC_LOADS
    cp_async_wait_all();
#endif
}
__device__ __forceinline__ void store_rows(const u64* sm, u64* __restrict__ poly_tile, uint32_t tid) {
    const uint32_t M2 = 1u << (kLogCols + kRowTileLog - 1);
    ulonglong2* dst = reinterpret_cast<ulonglong2*>(poly_tile);
#pragma unroll 4
    for (uint32_t j = tid; j < M2; j += kGroup) dst[j] = make_ulonglong2(sm[sl(2 * j)], sm[sl(2 * j) + 1]);
}
// column tile: local j = r * 8 + cc  <->  global n = r * 128 + c0 + cc
__device__ __forceinline__ void load_cols(u64* sm, const u64* __restrict__ poly, uint32_t logR, uint32_t c0, uint32_t tid) {
    const uint32_t M2 = 1u << (logR + kColTileLog - 1);
#if PSI_ASYNC_LOADS
#pragma unroll 8
    for (uint32_t j = tid; j < M2; j += kGroup) {
        const uint32_t e = 2 * j;
        const u64* src = poly + ((e >> kColTileLog) << kLogCols) + c0 + (e & ((1u << kColTileLog) - 1));
        cp_async8(sm + sl(e), src);
        cp_async8(sm + sl(e) + 1, src + 1);
    }
#else
#pragma unroll 4
    for (uint32_t j = tid; j < M2; j += kGroup) {
        const uint32_t e = 2 * j;
        const ulonglong2 v = ld_tile16(poly + ((e >> kColTileLog) << kLogCols) + c0 + (e & ((1u << kColTileLog) - 1)));
        sm[sl(e)] = v.x;
        sm[sl(e) + 1] = v.y;
    }
#endif
}
__device__ __forceinline__ void store_cols(const u64* sm, u64* __restrict__ poly, uint32_t logR, uint32_t c0, uint32_t tid) {
    const uint32_t M2 = 1u << (logR + kColTileLog - 1);
#pragma unroll 4
    for (uint32_t j = tid; j < M2; j += kGroup) {
        const uint32_t e = 2 * j;
        *reinterpret_cast<ulonglong2*>(poly + ((e >> kColTileLog) << kLogCols) + c0 + (e & ((1u << kColTileLog) - 1))) =
            make_ulonglong2(sm[sl(e)], sm[sl(e) + 1]);
    }
}

// ---- (2) columns: inverse, basis extension, forward -----------------------------------------------
// grid (128/8, B, 4 = (operand, component) in launch order (1,0) (1,1) (0,0) (0,1)); kColGroups groups.
//   operand 0 (multipliedResult):  DCRTPoly::ExpandCRTBasis            -> e1p [B][2][Lp][N]
//   operand 1 (innerProductResult): DCRTPoly::FastExpandCRTBasisPloverQ -> e2h [B][2][LT][N]
// All modular sums are formed as sums of lazy Shoup products (< 2q each, at most 8 terms < 2^64) and
// reduced once: the canonical residue, identical to the 128-bit Barrett form of the reference.
template <int L, int LP, int LOGN_CT>
__global__ void __launch_bounds__(kColGroups* kGroup, 3)
    k_cols_extend(const DevTables* __restrict__ tab, uint32_t logN, const u64* __restrict__ ha, const u64* __restrict__ hb,
                  u64* __restrict__ e1p, u64* __restrict__ e2h, uint32_t z0) {
    extern __shared__ __align__(16) u64 smem[];
    constexpr int LT = L + LP;
    const uint32_t N = 1u << logN;
    const uint32_t logR = logN - kLogCols, m = logR + kColTileLog, M = 1u << m, P = padded(M);
    const uint32_t g = threadIdx.x / kGroup, tid = threadIdx.x % kGroup;
    const uint32_t c0 = blockIdx.x << kColTileLog;
    // longest work first: the CTAs of operand 1 (L + Lp output limbs) are launched before those of operand 0 (Lp),
    // so that the tail of the grid is made of the short ones
    // z0: first (operand, comp) slot of this launch (0: both operands or operand 1 only, 2: operand 0 only)
    const uint32_t zz = blockIdx.z + z0, operand = 1u - (zz >> 1), comp = zz & 1;
    const size_t bin = blockIdx.y;

    // column-inverse of the L input limbs
    const u64* src = (operand ? hb : ha) + ((bin * 2 + comp) * L) * (size_t)N;
    for (uint32_t a = g; a < L; a += kColGroups) load_cols(smem + a * P, src + (size_t)a * N, logR, c0, tid);
    loads_wait();
    __syncthreads();
    transform_cols<true, LOGN_CT>(tab, smem, P, L, [](uint32_t a) { return a; }, g, kColGroups, logN, tid);
    __syncthreads();  // the extension reads every limb of a coefficient

    // coefficient-wise extension; N^-1 of the inverse transform is folded into the first constant
    for (uint32_t j = threadIdx.x; j < M; j += blockDim.x) {
        u64 y[L];
        if (operand == 0) {
            double nu = 0.5;
#pragma unroll
            for (int i = 0; i < L; i++) {
                y[i] = shoup_canon(smem[i * P + sl(j)], tab->QHatInvNinv[i], tab->QHatInvNinv_s[i], tab->mods[i].q);
                nu = nu_step(nu, __ull2double_rn(y[i]), tab->qInv[i], tab->fp_fma);
            }
            const unsigned alpha = (unsigned)nu;
            uint32_t y0[L], y1[L];
#pragma unroll
            for (int i = 0; i < L; i++) {
                y0[i] = (uint32_t)y[i] & 0x3fffffffu;
                y1[i] = (uint32_t)(y[i] >> 30);
            }
#pragma unroll
            for (int jj = 0; jj < LP; jj++) {
                const ModDev& mp = tab->mods[L + jj];
                Acc3 acc;
#pragma unroll
                for (int i = 0; i < L; i++) acc3_mad(acc, y0[i], y1[i], tab->QHatModp_m[jj][i]);
                const u64 v = acc3_reduce(acc, mp.q, mp.qinv);
                smem[jj * P + sl(j)] = submod(v, tab->alphaQModp[alpha][jj], mp.q);
            }
        } else {
            u64 pp[LP], z[LP];
#pragma unroll
            for (int i = 0; i < L; i++)
                y[i] = shoup_canon(smem[i * P + sl(j)], tab->negPQHatInvNinv[i], tab->negPQHatInvNinv_s[i], tab->mods[i].q);
            double nu = 0.5;
            uint32_t y0[L], y1[L], z0[LP], z1[LP];
#pragma unroll
            for (int i = 0; i < L; i++) {
                y0[i] = (uint32_t)y[i] & 0x3fffffffu;
                y1[i] = (uint32_t)(y[i] >> 30);
            }
#pragma unroll
            for (int jj = 0; jj < LP; jj++) {
                const ModDev& mp = tab->mods[L + jj];
                Acc3 acc;
#pragma unroll
                for (int i = 0; i < L; i++) acc3_mad(acc, y0[i], y1[i], tab->qInvModp_m[i][jj]);
                pp[jj] = acc3_reduce(acc, mp.q, mp.qinv);
                // exact P -> Q (DCRTPoly::SwitchCRTBasis)
                z[jj] = shoup_canon(pp[jj], tab->PHatInvModp[jj], tab->PHatInvModp_s[jj], mp.q);
                nu = nu_step(nu, __ull2double_rn(z[jj]), tab->pInv[jj], tab->fp_fma);
                z0[jj] = (uint32_t)z[jj] & 0x3fffffffu;
                z1[jj] = (uint32_t)(z[jj] >> 30);
            }
            const unsigned alpha = (unsigned)nu;
#pragma unroll
            for (int i = 0; i < L; i++) {
                const ModDev& mq = tab->mods[i];
                Acc3 acc;
#pragma unroll
                for (int jj = 0; jj < LP; jj++) acc3_mad(acc, z0[jj], z1[jj], tab->PHatModq_m[i][jj]);
                const u64 v = acc3_reduce(acc, mq.q, mq.qinv);
                smem[i * P + sl(j)] = submod(v, tab->alphaPModq[alpha][i], mq.q);
            }
#pragma unroll
            for (int jj = 0; jj < LP; jj++) smem[(L + jj) * P + sl(j)] = pp[jj];
        }
    }
    __syncthreads();

    // column-forward of the produced limbs
    const uint32_t n_out = operand ? LT : LP;
    const uint32_t mod0 = operand ? 0 : L;
    transform_cols<false, LOGN_CT>(tab, smem, P, n_out, [mod0](uint32_t a) { return mod0 + a; }, g, kColGroups, logN, tid);
    for (uint32_t a = g; a < n_out; a += kColGroups) {
        u64* dst = operand ? e2h + ((bin * 2 + comp) * LT + a) * (size_t)N : e1p + ((bin * 2 + comp) * LP + a) * (size_t)N;
        store_cols(smem + a * P, dst, logR, c0, tid);
    }
}

// ---- (4) columns: inverse, scale-and-round, digit lift, forward -----------------------------------
// grid (128/8, B, 3 = component in launch order 2, 0, 1), kColGroups groups.  th: [B][3][LT][N]; rh: [B][2][L][N] (column-forward halves of
// c0, c1); dh: [B][L][L][N] (column-forward halves of the BV digits of c2)
template <int L, int LP, int LOGN_CT>
__global__ void __launch_bounds__(kColGroups* kGroup, 3)
    k_cols_scale(const DevTables* __restrict__ tab, uint32_t logN, const u64* __restrict__ th, u64* __restrict__ rh,
                 u64* __restrict__ dh) {
    extern __shared__ __align__(16) u64 smem[];
    constexpr int LT = L + LP;
    const uint32_t N = 1u << logN;
    const uint32_t logR = logN - kLogCols, m = logR + kColTileLog, M = 1u << m, P = padded(M);
    const uint32_t g = threadIdx.x / kGroup, tid = threadIdx.x % kGroup;
    // longest work first: component 2 (L digits = L * L forward limb transforms) is launched before components 0, 1
    const uint32_t c0 = blockIdx.x << kColTileLog, comp = (blockIdx.z + 2u) % 3u;
    const size_t bin = blockIdx.y;
    for (uint32_t a = g; a < LT; a += kColGroups)
        load_cols(smem + a * P, th + ((bin * 3 + comp) * LT + a) * (size_t)N, logR, c0, tid);
    loads_wait();
    __syncthreads();
    transform_cols<true, LOGN_CT>(tab, smem, P, LT, [](uint32_t a) { return a; }, g, kColGroups, logN, tid);
    __syncthreads();  // scale-and-round reads every limb of a coefficient

    // DCRTPoly::ScaleAndRound (t/P, output basis Q) on canonical coefficients
    for (uint32_t j = threadIdx.x; j < M; j += blockDim.x) {
        u64 xp[LP];
        double nu = 0.5;
#pragma unroll
        for (int i = 0; i < LP; i++) {
            const ModDev& mp = tab->mods[L + i];
            xp[i] = shoup_canon(smem[(L + i) * P + sl(j)], mp.ninvR, mp.ninvR_s, mp.q);
            nu = nu_step(nu, tab->tQSfrac[i], __ull2double_rn(xp[i]), tab->fp_fma);
        }
        const u64 alpha = __double2ull_rz(nu);  // < LP * 2^60
        uint32_t xp0[LP], xp1[LP];
#pragma unroll
        for (int i = 0; i < LP; i++) {
            xp0[i] = (uint32_t)xp[i] & 0x3fffffffu;
            xp1[i] = (uint32_t)(xp[i] >> 30);
        }
#pragma unroll
        for (int l = 0; l < L; l++) {
            const ModDev& mq = tab->mods[l];
            const u64 q = mq.q;
            const u64 xq = shoup_canon(smem[l * P + sl(j)], mq.ninvR, mq.ninvR_s, q);
            Acc3 acc;
            acc3_mad(acc, (uint32_t)xq & 0x3fffffffu, (uint32_t)(xq >> 30), tab->tQS_m[l][LP]);
#pragma unroll
            for (int i = 0; i < LP; i++) acc3_mad(acc, xp0[i], xp1[i], tab->tQS_m[l][i]);
            // alpha < 16 q for 60-bit moduli; smaller moduli take the division
            smem[l * P + sl(j)] = addmod(acc3_reduce(acc, q, mq.qinv), alpha < (q << 4) ? reduce_pow2q<4>(alpha, q) : alpha % q, q);
        }
    }
    __syncthreads();

    if (comp < 2) {
        transform_cols<false, LOGN_CT>(tab, smem, P, L, [](uint32_t a) { return a; }, g, kColGroups, logN, tid);
        for (uint32_t a = g; a < L; a += kColGroups)
            store_cols(smem + a * P, rh + ((bin * 2 + comp) * L + a) * (size_t)N, logR, c0, tid);
        return;
    }
    // DCRTPoly::CRTDecompose (BV, digit size 0): digit i = limb i of c2 switched to every q_k with the
    // centred lift of NativeVector::SwitchModulus; arrays L..2L-1 hold the L limbs of the current digit
    for (uint32_t i = 0; i < L; i++) {
        const u64 qi = tab->mods[i].q, half = (qi - 1) >> 1;
        for (uint32_t kk = g; kk < L; kk += kColGroups) {
            const u64 qk = tab->mods[kk].q, qiq = tab->qModq[i][kk];
            u64* dst = smem + (L + kk) * P;
            for (uint32_t j = tid; j < M; j += kGroup) {
                const u64 v = smem[i * P + sl(j)];
                u64 r = v;
                if (kk != i) {
                    r = (v < qk) ? v : ((v - qk < qk) ? v - qk : v % qk);
                    if (v > half) r = submod(r, qiq, qk);
                }
                dst[sl(j)] = r;
            }
        }
        // the lift, the transform and the store of digit limb kk all belong to group kk % kColGroups
        group_sync();
        transform_cols<false, LOGN_CT>(tab, smem + L * P, P, L, [](uint32_t a) { return a; }, g, kColGroups, logN, tid);
        for (uint32_t kk = g; kk < L; kk += kColGroups)
            store_cols(smem + (L + kk) * P, dh + ((bin * L + i) * L + kk) * (size_t)N, logR, c0, tid);
        group_sync();
    }
}

template <int L, int LP, int LOGN_CT>
cudaError_t launch_cols(const KCtx& k, uint32_t B, const u64* ha, const u64* hb, u64* e1p, u64* e2h, const u64* th,
                               u64* rh, u64* dh, int which) {
    const uint32_t logR = k.logN - kLogCols, col_tiles = (1u << kLogCols) >> kColTileLog;
    const size_t smem = (size_t)(L + LP) * padded(1u << (logR + kColTileLog)) * sizeof(u64);
    if (which < 0) {  // per-device set-up (psi_ctx_create): opt in to > 48 KiB of dynamic shared memory
        cudaError_t e;
        if ((e = cudaFuncSetAttribute(k_cols_extend<L, LP, LOGN_CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)) != cudaSuccess) return e;
        return cudaFuncSetAttribute(k_cols_scale<L, LP, LOGN_CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    }
    // which: 0 = extend operand a only, 1 = extend operand b only, 2 = extend both, 3 = scale
    if (which == 2)
        k_cols_extend<L, LP, LOGN_CT><<<dim3(col_tiles, B, 4), kColGroups * kGroup, smem, k.s>>>(k.tab, k.logN, ha, hb, e1p, e2h, 0);
    else if (which == 1)
        k_cols_extend<L, LP, LOGN_CT><<<dim3(col_tiles, B, 2), kColGroups * kGroup, smem, k.s>>>(k.tab, k.logN, ha, hb, e1p, e2h, 0);
    else if (which == 0)
        k_cols_extend<L, LP, LOGN_CT><<<dim3(col_tiles, B, 2), kColGroups * kGroup, smem, k.s>>>(k.tab, k.logN, ha, hb, e1p, e2h, 2);
    else
        k_cols_scale<L, LP, LOGN_CT><<<dim3(col_tiles, B, 3), kColGroups * kGroup, smem, k.s>>>(k.tab, k.logN, th, rh, dh);
    return cudaGetLastError();
}


// The (sizeQ, sizeP) instantiations of the column kernels are spread over four translation units
// (fused_cols_[a-d].cu) so that they compile in parallel; each returns cudaErrorInvalidValue when the
// context's limb counts are not among its cases.
#define PSI_COLS_ARGS const KCtx &k, uint32_t B, const u64 *ha, const u64 *hb, u64 *e1p, u64 *e2h, const u64 *th, u64 *rh, u64 *dh, int which
cudaError_t dispatch_cols_a(PSI_COLS_ARGS);
cudaError_t dispatch_cols_b(PSI_COLS_ARGS);
cudaError_t dispatch_cols_c(PSI_COLS_ARGS);
cudaError_t dispatch_cols_d(PSI_COLS_ARGS);
// compiled stage plans for the ring dimensions the reference uses (16384: client :73; 8192: the unit test)
#define PSI_COLS_CASE(l, lp)                                                                               \
    if (k.L == l && k.Lp == lp) {                                                                          \
        if (k.logN == 14) return launch_cols<l, lp, 14>(k, B, ha, hb, e1p, e2h, th, rh, dh, which);       \
        if (k.logN == 13) return launch_cols<l, lp, 13>(k, B, ha, hb, e1p, e2h, th, rh, dh, which);       \
        return launch_cols<l, lp, 0>(k, B, ha, hb, e1p, e2h, th, rh, dh, which);                           \
    }
// deeper contexts: compiled plan for N = 16384 only
#define PSI_COLS_CASE_BIG(l, lp)                                                                           \
    if (k.L == l && k.Lp == lp) {                                                                          \
        if (k.logN == 14) return launch_cols<l, lp, 14>(k, B, ha, hb, e1p, e2h, th, rh, dh, which);       \
        return launch_cols<l, lp, 0>(k, B, ha, hb, e1p, e2h, th, rh, dh, which);                           \
    }

}  // namespace psi
