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
#include "psi_kernels.cuh"

namespace psi {

constexpr uint32_t kLogCols = 7;    // row length 2^7 coefficients
constexpr uint32_t kRowTileLog = 3; // 8 rows per row tile
constexpr uint32_t kColTileLog = 3; // 8 columns per column tile
constexpr uint32_t kGroup = 64;     // threads cooperating on one shared-memory array

// padded shared-memory slot: one pad word per 16 coefficients keeps both the strided gathers and the
// 16-contiguous-per-thread pattern of the last radix pass off a single bank group
__device__ __forceinline__ uint32_t sl(uint32_t i) { return i + (i >> 4); }
__host__ __device__ constexpr uint32_t padded(uint32_t n) { return n + (n >> 4) + 1; }

// ---- lazy butterflies ------------------------------------------------------------------------
// forward (Cooley-Tukey): inputs < 4q + 2^32, outputs < 4q + 2^32
__device__ __forceinline__ void ct_bf(u64& x, u64& y, const ulonglong2 tw, u64 q2, u64 nq) {
    const u64 u = lazy_sub_hi(x, q2);
    const u64 v = mul_shoup_lazy_nq(y, tw.x, tw.y, nq);
    x = u + v;
    y = u - v + q2;
}
// inverse (Gentleman-Sande): inputs < 2q + e (e grows by at most a factor two per stage from 2^32,
// far below q after 14 stages), x output < 2q + 2e, y output < 2q
__device__ __forceinline__ void gs_bf(u64& x, u64& y, const ulonglong2 tw, u64 q2, u64 nq) {
    const u64 s = lazy_sub_hi(x + y, q2);
    const u64 d = x - y + 2 * q2;
    x = s;
    y = mul_shoup_lazy_nq(d, tw.x, tw.y, nq);
}

// Radix-2^R Cooley-Tukey pass over local stages [sig0, sig0 + R) of a local array of 2^m coefficients.
// Global twiddle index of local stage sig, local group g:  2^(sig + delta) + (tile_base >> (m - sig)) + g
template <int R>
__device__ __forceinline__ void fwd_pass(u64* __restrict__ sm, const ulonglong2* __restrict__ tw, uint32_t m, uint32_t sig0,
                                         uint32_t delta, uint32_t tile_base, u64 q, uint32_t tid) {
    const uint32_t tl = (1u << m) >> (sig0 + R);
    const u64 q2 = 2 * q, nq = 0 - q;
    for (uint32_t blk = tid; blk < ((1u << m) >> R); blk += kGroup) {
        const uint32_t off = blk & (tl - 1), grp = blk / tl;
        const uint32_t base = (grp << (m - sig0)) + off;
        u64 v[1 << R];
#pragma unroll
        for (int k = 0; k < (1 << R); k++) v[k] = sm[sl(base + k * tl)];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int half = 1 << (R - 1 - r);
            const uint32_t w0 = (1u << (sig0 + r + delta)) + (tile_base >> (m - sig0 - r)) + (grp << r);
#pragma unroll
            for (int j = 0; j < (1 << r); j++) {
                const ulonglong2 t = __ldg(tw + w0 + j);
#pragma unroll
                for (int i = 0; i < half; i++) ct_bf(v[j * 2 * half + i], v[j * 2 * half + i + half], t, q2, nq);
            }
        }
#pragma unroll
        for (int k = 0; k < (1 << R); k++) sm[sl(base + k * tl)] = v[k];
    }
}
template <int R>
__device__ __forceinline__ void inv_pass(u64* __restrict__ sm, const ulonglong2* __restrict__ tw, uint32_t m, uint32_t sig0,
                                         uint32_t delta, uint32_t tile_base, u64 q, uint32_t tid) {
    const uint32_t tl = (1u << m) >> (sig0 + R);
    const u64 q2 = 2 * q, nq = 0 - q;
    for (uint32_t blk = tid; blk < ((1u << m) >> R); blk += kGroup) {
        const uint32_t off = blk & (tl - 1), grp = blk / tl;
        const uint32_t base = (grp << (m - sig0)) + off;
        u64 v[1 << R];
#pragma unroll
        for (int k = 0; k < (1 << R); k++) v[k] = sm[sl(base + k * tl)];
#pragma unroll
        for (int r = R - 1; r >= 0; r--) {
            const int half = 1 << (R - 1 - r);
            const uint32_t w0 = (1u << (sig0 + r + delta)) + (tile_base >> (m - sig0 - r)) + (grp << r);
#pragma unroll
            for (int j = 0; j < (1 << r); j++) {
                const ulonglong2 t = __ldg(tw + w0 + j);
#pragma unroll
                for (int i = 0; i < half; i++) gs_bf(v[j * 2 * half + i], v[j * 2 * half + i + half], t, q2, nq);
            }
        }
#pragma unroll
        for (int k = 0; k < (1 << R); k++) sm[sl(base + k * tl)] = v[k];
    }
}

// how many stages the next register pass takes when `rem` remain: 7 -> 4+3, 6 -> 3+3, 5 -> 3+2
__device__ __forceinline__ uint32_t pass_width(uint32_t rem) { return (rem >= 7 || rem == 4) ? 4 : (rem >= 3 ? 3 : rem); }

// Forward stages [lo, hi) in ascending order.  `active` = this thread's group owns an array; all
// threads of the CTA must call (CTA-wide barriers between register passes).
__device__ __forceinline__ void fwd_range(bool active, u64* sm, const ulonglong2* tw, uint32_t m, uint32_t lo, uint32_t hi,
                                          uint32_t delta, uint32_t tile_base, u64 q, uint32_t tid) {
    uint32_t s = lo;
    while (s < hi) {
        const uint32_t w = pass_width(hi - s);
        if (active) {
            if (w == 4) fwd_pass<4>(sm, tw, m, s, delta, tile_base, q, tid);
            else if (w == 3) fwd_pass<3>(sm, tw, m, s, delta, tile_base, q, tid);
            else if (w == 2) fwd_pass<2>(sm, tw, m, s, delta, tile_base, q, tid);
            else fwd_pass<1>(sm, tw, m, s, delta, tile_base, q, tid);
        }
        s += w;
        __syncthreads();
    }
}
// Inverse stages [lo, hi) in descending order.
__device__ __forceinline__ void inv_range(bool active, u64* sm, const ulonglong2* tw, uint32_t m, uint32_t lo, uint32_t hi,
                                          uint32_t delta, uint32_t tile_base, u64 q, uint32_t tid) {
    uint32_t s = hi;
    while (s > lo) {
        const uint32_t w = pass_width(s - lo);
        if (active) {
            if (w == 4) inv_pass<4>(sm, tw, m, s - 4, delta, tile_base, q, tid);
            else if (w == 3) inv_pass<3>(sm, tw, m, s - 3, delta, tile_base, q, tid);
            else if (w == 2) inv_pass<2>(sm, tw, m, s - 2, delta, tile_base, q, tid);
            else inv_pass<1>(sm, tw, m, s - 1, delta, tile_base, q, tid);
        }
        s -= w;
        __syncthreads();
    }
}

// a mod q for a < 16q (alpha of ScaleAndRound is below sizeP * 2^60 and q is above 2^59... any q > a/16):
// four compare-subtract steps instead of a 64-bit division
__device__ __forceinline__ u64 reduce_lt16q(u64 a, u64 q) {
    if (a >= 8 * q) a -= 8 * q;
    if (a >= 4 * q) a -= 4 * q;
    if (a >= 2 * q) a -= 2 * q;
    if (a >= q) a -= q;
    return a;
}

// ---- tile movers -------------------------------------------------------------------------------
// row tile: 2^(7+kRowTileLog) contiguous coefficients starting at poly + tile_base
__device__ __forceinline__ void load_rows(u64* sm, const u64* __restrict__ poly_tile, uint32_t tid) {
    const uint32_t M = 1u << (kLogCols + kRowTileLog);
    for (uint32_t j = tid; j < M; j += kGroup) sm[sl(j)] = poly_tile[j];
}
__device__ __forceinline__ void store_rows(const u64* sm, u64* __restrict__ poly_tile, uint32_t tid) {
    const uint32_t M = 1u << (kLogCols + kRowTileLog);
    for (uint32_t j = tid; j < M; j += kGroup) poly_tile[j] = sm[sl(j)];
}
// column tile: local j = r * 8 + cc  <->  global n = r * 128 + c0 + cc
__device__ __forceinline__ void load_cols(u64* sm, const u64* __restrict__ poly, uint32_t logR, uint32_t c0, uint32_t tid) {
    const uint32_t M = 1u << (logR + kColTileLog);
    for (uint32_t j = tid; j < M; j += kGroup)
        sm[sl(j)] = poly[((j >> kColTileLog) << kLogCols) + c0 + (j & ((1u << kColTileLog) - 1))];
}
__device__ __forceinline__ void store_cols(const u64* sm, u64* __restrict__ poly, uint32_t logR, uint32_t c0, uint32_t tid) {
    const uint32_t M = 1u << (logR + kColTileLog);
    for (uint32_t j = tid; j < M; j += kGroup)
        poly[((j >> kColTileLog) << kLogCols) + c0 + (j & ((1u << kColTileLog) - 1))] = sm[sl(j)];
}

// ---- (1) rows, inverse: both operands, all 4L limb-polys of a bin --------------------------------
// grid (R/8, L, B), 4 groups: a.c0, a.c1, b.c0, b.c1 of limb blockIdx.y
__global__ void __launch_bounds__(4 * kGroup) k_rows_inv(const DevTables* __restrict__ tab, uint32_t logN,
                                                         const u64* __restrict__ a, const u64* __restrict__ b,
                                                         u64* __restrict__ ha, u64* __restrict__ hb) {
    extern __shared__ u64 smem[];
    const uint32_t N = 1u << logN, L = tab->L;
    const uint32_t m = kLogCols + kRowTileLog, M = 1u << m;
    const uint32_t g = threadIdx.x / kGroup, tid = threadIdx.x % kGroup;
    const uint32_t l = blockIdx.y, tile_base = blockIdx.x * M;
    const size_t bin = blockIdx.z;
    const size_t off = ((bin * 2 + (g & 1)) * L + l) * N + tile_base;
    const u64* src = (g < 2 ? a : b) + off;
    u64* dst = (g < 2 ? ha : hb) + off;
    u64* sm = smem + g * padded(M);
    const ModDev& md = tab->mods[l];
    load_rows(sm, src, tid);
    __syncthreads();
    inv_range(true, sm, md.itw, m, kRowTileLog, m, logN - m, tile_base, md.q, tid);
    store_rows(sm, dst, tid);
}

// ---- (2) columns: inverse, basis extension, forward -----------------------------------------------
// grid (128/8, 4 = operand*2 + component, B); LT groups.
//   operand 0 (multipliedResult):  DCRTPoly::ExpandCRTBasis            -> e1p [B][2][Lp][N]
//   operand 1 (innerProductResult): DCRTPoly::FastExpandCRTBasisPloverQ -> e2h [B][2][LT][N]
__global__ void __launch_bounds__(PSI_MAX_LIMBS * kGroup) k_cols_extend(const DevTables* __restrict__ tab,
                                                                           uint32_t logN, const u64* __restrict__ ha,
                                                                           const u64* __restrict__ hb,
                                                                           u64* __restrict__ e1p, u64* __restrict__ e2h) {
    extern __shared__ u64 smem[];
    const uint32_t N = 1u << logN, L = tab->L, Lp = tab->Lp, LT = L + Lp;
    const uint32_t logR = logN - kLogCols, m = logR + kColTileLog, M = 1u << m;
    const uint32_t g = threadIdx.x / kGroup, tid = threadIdx.x % kGroup;
    const uint32_t c0 = blockIdx.x << kColTileLog;
    const uint32_t operand = blockIdx.y >> 1, comp = blockIdx.y & 1;
    const size_t bin = blockIdx.z;
    const uint32_t P = padded(M);
    u64* sm = smem + g * P;

    // column-inverse of the L input limbs
    const u64* src = (operand ? hb : ha) + ((bin * 2 + comp) * L) * (size_t)N;
    const bool in_active = g < L;
    if (in_active) load_cols(sm, src + (size_t)g * N, logR, c0, tid);
    __syncthreads();
    inv_range(in_active, sm, tab->mods[in_active ? g : 0].itw, m, 0, logR, 0, 0, tab->mods[in_active ? g : 0].q, tid);

    // coefficient-wise extension; N^-1 of the inverse transform is folded into the first constant
    for (uint32_t j = threadIdx.x; j < M; j += blockDim.x) {
        u64 y[PSI_MAX_LIMBS];
        if (operand == 0) {
            double nu = 0.5;
#pragma unroll
            for (int i = 0; i < PSI_MAX_LIMBS; i++)
                if (i < (int)L) {
                    y[i] = mul_shoup(smem[i * P + sl(j)], tab->QHatInvNinv[i], tab->QHatInvNinv_s[i], tab->mods[i].q);
                    nu = __dadd_rn(nu, __dmul_rn(__ull2double_rn(y[i]), tab->qInv[i]));
                }
            const unsigned alpha = (unsigned)nu;
#pragma unroll
            for (int jj = 0; jj < PSI_MAX_LIMBS; jj++)
                if (jj < (int)Lp) {
                    const ModDev& mb = tab->mods[L + jj];
                    u64 hi = 0, lo = 0;
#pragma unroll
                    for (int i = 0; i < PSI_MAX_LIMBS; i++)
                        if (i < (int)L) mac128(hi, lo, y[i], tab->QHatModp[jj][i]);
                    const u64 v = barrett128(hi, lo, mb.q, mb.mu_hi, mb.mu_lo);
                    smem[jj * P + sl(j)] = submod(v, tab->alphaQModp[alpha][jj], mb.q);
                }
        } else {
            u64 pp[PSI_MAX_LIMBS];
#pragma unroll
            for (int i = 0; i < PSI_MAX_LIMBS; i++)
                if (i < (int)L)
                    y[i] = mul_shoup(smem[i * P + sl(j)], tab->negPQHatInvNinv[i], tab->negPQHatInvNinv_s[i], tab->mods[i].q);
            double nu = 0.5;
#pragma unroll
            for (int jj = 0; jj < PSI_MAX_LIMBS; jj++)
                if (jj < (int)Lp) {
                    const ModDev& mp = tab->mods[L + jj];
                    u64 hi = 0, lo = 0;
#pragma unroll
                    for (int i = 0; i < PSI_MAX_LIMBS; i++)
                        if (i < (int)L) mac128(hi, lo, y[i], tab->qInvModp[i][jj]);
                    pp[jj] = barrett128(hi, lo, mp.q, mp.mu_hi, mp.mu_lo);
                }
            // exact P -> Q (DCRTPoly::SwitchCRTBasis)
            u64 z[PSI_MAX_LIMBS];
#pragma unroll
            for (int jj = 0; jj < PSI_MAX_LIMBS; jj++)
                if (jj < (int)Lp) {
                    z[jj] = mul_shoup(pp[jj], tab->PHatInvModp[jj], tab->PHatInvModp_s[jj], tab->mods[L + jj].q);
                    nu = __dadd_rn(nu, __dmul_rn(__ull2double_rn(z[jj]), tab->pInv[jj]));
                }
            const unsigned alpha = (unsigned)nu;
#pragma unroll
            for (int i = 0; i < PSI_MAX_LIMBS; i++)
                if (i < (int)L) {
                    const ModDev& mq = tab->mods[i];
                    u64 hi = 0, lo = 0;
#pragma unroll
                    for (int jj = 0; jj < PSI_MAX_LIMBS; jj++)
                        if (jj < (int)Lp) mac128(hi, lo, z[jj], tab->PHatModq[i][jj]);
                    const u64 v = barrett128(hi, lo, mq.q, mq.mu_hi, mq.mu_lo);
                    smem[i * P + sl(j)] = submod(v, tab->alphaPModq[alpha][i], mq.q);
                }
#pragma unroll
            for (int jj = 0; jj < PSI_MAX_LIMBS; jj++)
                if (jj < (int)Lp) smem[(L + jj) * P + sl(j)] = pp[jj];
        }
    }
    __syncthreads();

    // column-forward of the produced limbs
    const uint32_t n_out = operand ? LT : Lp;
    const uint32_t mod_out = operand ? g : L + g;  // array g holds modulus index mod_out
    const bool out_active = g < n_out;
    const ModDev& mo = tab->mods[out_active ? mod_out : 0];
    fwd_range(out_active, sm, mo.ftw, m, 0, logR, 0, 0, mo.q, tid);
    if (out_active) {
        u64* dst = operand ? e2h + ((bin * 2 + comp) * LT + g) * (size_t)N : e1p + ((bin * 2 + comp) * Lp + g) * (size_t)N;
        store_cols(sm, dst, logR, c0, tid);
    }
}

// ---- (3) rows: forward, tensor, inverse -----------------------------------------------------------
// grid (R/8, LT, B), 4 groups (a0, a1, b0, b1 of limb blockIdx.y); a: [B][2][L][N] EVALUATION (Q limbs
// are used as given), e1p/e2h from (2); th: [B][3][LT][N] row-inverse halves
__global__ void __launch_bounds__(4 * kGroup) k_rows_tensor(const DevTables* __restrict__ tab, uint32_t logN,
                                                            const u64* __restrict__ a, const u64* __restrict__ e1p,
                                                            const u64* __restrict__ e2h, u64* __restrict__ th) {
    extern __shared__ u64 smem[];
    const uint32_t N = 1u << logN, L = tab->L, Lp = tab->Lp, LT = L + Lp;
    const uint32_t m = kLogCols + kRowTileLog, M = 1u << m, P = padded(M);
    const uint32_t g = threadIdx.x / kGroup, tid = threadIdx.x % kGroup;
    const uint32_t l = blockIdx.y, tile_base = blockIdx.x * M;
    const size_t bin = blockIdx.z;
    const ModDev& md = tab->mods[l];
    u64* sm = smem + g * P;
    const uint32_t comp = g & 1;
    bool transform = true;
    const u64* src;
    if (g < 2) {
        if (l < L) {
            src = a + ((bin * 2 + comp) * L + l) * (size_t)N;
            transform = false;
        } else {
            src = e1p + ((bin * 2 + comp) * Lp + (l - L)) * (size_t)N;
        }
    } else {
        src = e2h + ((bin * 2 + comp) * LT + l) * (size_t)N;
    }
    load_rows(sm, src + tile_base, tid);
    __syncthreads();
    fwd_range(transform, sm, md.ftw, m, kRowTileLog, m, logN - m, tile_base, md.q, tid);

    for (uint32_t j = threadIdx.x; j < M; j += blockDim.x) {
        const u64 a0 = smem[sl(j)], a1 = smem[P + sl(j)], b0 = smem[2 * P + sl(j)], b1 = smem[3 * P + sl(j)];
        smem[sl(j)] = barrett128(mulhi64(a0, b0), a0 * b0, md.q, md.mu_hi, md.mu_lo);
        u64 hi = 0, lo = 0;
        mac128(hi, lo, a0, b1);
        mac128(hi, lo, a1, b0);
        smem[P + sl(j)] = barrett128(hi, lo, md.q, md.mu_hi, md.mu_lo);
        smem[2 * P + sl(j)] = barrett128(mulhi64(a1, b1), a1 * b1, md.q, md.mu_hi, md.mu_lo);
    }
    __syncthreads();
    inv_range(g < 3, sm, md.itw, m, kRowTileLog, m, logN - m, tile_base, md.q, tid);
    if (g < 3) store_rows(sm, th + ((bin * 3 + g) * LT + l) * (size_t)N + tile_base, tid);
}

// ---- (4) columns: inverse, scale-and-round, digit lift, forward -----------------------------------
// grid (128/8, 3, B), LT groups.  th: [B][3][LT][N]; rh: [B][2][L][N] (column-forward halves of c0, c1);
// dh: [B][L][L][N] (column-forward halves of the BV digits of c2)
__global__ void __launch_bounds__(PSI_MAX_LIMBS * kGroup) k_cols_scale(const DevTables* __restrict__ tab,
                                                                          uint32_t logN, const u64* __restrict__ th,
                                                                          u64* __restrict__ rh, u64* __restrict__ dh) {
    extern __shared__ u64 smem[];
    const uint32_t N = 1u << logN, L = tab->L, Lp = tab->Lp, LT = L + Lp;
    const uint32_t logR = logN - kLogCols, m = logR + kColTileLog, M = 1u << m, P = padded(M);
    const uint32_t g = threadIdx.x / kGroup, tid = threadIdx.x % kGroup;
    const uint32_t c0 = blockIdx.x << kColTileLog, comp = blockIdx.y;
    const size_t bin = blockIdx.z;
    u64* sm = smem + g * P;
    const bool in_active = g < LT;
    const ModDev& mi = tab->mods[in_active ? g : 0];
    if (in_active) load_cols(sm, th + ((bin * 3 + comp) * LT + g) * (size_t)N, logR, c0, tid);
    __syncthreads();
    inv_range(in_active, sm, mi.itw, m, 0, logR, 0, 0, mi.q, tid);

    // DCRTPoly::ScaleAndRound (t/P, output basis Q) on canonical coefficients
    for (uint32_t j = threadIdx.x; j < M; j += blockDim.x) {
        u64 xp[PSI_MAX_LIMBS], xq[PSI_MAX_LIMBS];
        double nu = 0.5;
#pragma unroll
        for (int i = 0; i < PSI_MAX_LIMBS; i++)
            if (i < (int)Lp) {
                const ModDev& mp = tab->mods[L + i];
                xp[i] = mul_shoup(smem[(L + i) * P + sl(j)], mp.ninv, mp.ninv_s, mp.q);
                nu = __dadd_rn(nu, __dmul_rn(tab->tQSfrac[i], __ull2double_rn(xp[i])));
            }
#pragma unroll
        for (int l = 0; l < PSI_MAX_LIMBS; l++)
            if (l < (int)L) {
                const ModDev& mq = tab->mods[l];
                xq[l] = mul_shoup(smem[l * P + sl(j)], mq.ninv, mq.ninv_s, mq.q);
            }
        const u64 alpha = __double2ull_rz(nu);
#pragma unroll
        for (int l = 0; l < PSI_MAX_LIMBS; l++)
            if (l < (int)L) {
                const ModDev& mq = tab->mods[l];
                u64 hi = 0, lo = 0;
#pragma unroll
                for (int i = 0; i < PSI_MAX_LIMBS; i++)
                    if (i < (int)Lp) mac128(hi, lo, xp[i], tab->tQS[l][i]);
                mac128(hi, lo, xq[l], tab->tQS[l][Lp]);
                const u64 v = barrett128(hi, lo, mq.q, mq.mu_hi, mq.mu_lo);
                smem[l * P + sl(j)] = addmod(v, reduce_lt16q(alpha, mq.q), mq.q);
            }
    }
    __syncthreads();

    if (comp < 2) {
        const bool act = g < L;
        const ModDev& mo = tab->mods[act ? g : 0];
        fwd_range(act, sm, mo.ftw, m, 0, logR, 0, 0, mo.q, tid);
        if (act) store_cols(sm, rh + ((bin * 2 + comp) * L + g) * (size_t)N, logR, c0, tid);
        return;
    }
    // DCRTPoly::CRTDecompose (BV, digit size 0): digit i = limb i of c2 switched to every q_k with the
    // centred lift of NativeVector::SwitchModulus; arrays L..2L-1 hold the L limbs of the current digit
    const bool act = g >= L && g < 2 * L;
    const uint32_t kk = act ? g - L : 0;
    const ModDev& mo = tab->mods[kk];
    for (uint32_t i = 0; i < L; i++) {
        if (act) {
            const u64 qi = tab->mods[i].q, qk = mo.q, half = (qi - 1) >> 1, qiq = tab->qModq[i][kk];
            for (uint32_t j = tid; j < M; j += kGroup) {
                const u64 v = smem[i * P + sl(j)];
                u64 r = v;
                if (kk != i) {
                    r = (v < qk) ? v : ((v - qk < qk) ? v - qk : v % qk);
                    if (v > half) r = submod(r, qiq, qk);
                }
                sm[sl(j)] = r;
            }
        }
        __syncthreads();
        fwd_range(act, sm, mo.ftw, m, 0, logR, 0, 0, mo.q, tid);
        if (act) store_cols(sm, dh + ((bin * L + i) * L + kk) * (size_t)N, logR, c0, tid);
        __syncthreads();
    }
}

// ---- (5) rows: forward, key-switch inner product, add (c0, c1), mask ------------------------------
// grid (R/8, L, B), 2 + L groups (c0, c1, digit 0..L-1 of limb blockIdx.y)
__global__ void __launch_bounds__((2 + PSI_MAX_LIMBS) * kGroup) k_rows_relin(const DevTables* __restrict__ tab,
                                                                            uint32_t logN, const u64* __restrict__ rh,
                                                                            const u64* __restrict__ dh,
                                                                            const u64* __restrict__ evk_b,
                                                                            const u64* __restrict__ evk_a,
                                                                            const u64* __restrict__ mask,
                                                                            u64* __restrict__ out) {
    extern __shared__ u64 smem[];
    const uint32_t N = 1u << logN, L = tab->L;
    const uint32_t m = kLogCols + kRowTileLog, M = 1u << m, P = padded(M);
    const uint32_t g = threadIdx.x / kGroup, tid = threadIdx.x % kGroup;
    const uint32_t kk = blockIdx.y, tile_base = blockIdx.x * M;
    const size_t bin = blockIdx.z;
    const ModDev& md = tab->mods[kk];
    u64* sm = smem + g * P;
    const u64* src = g < 2 ? rh + ((bin * 2 + g) * L + kk) * (size_t)N : dh + ((bin * L + (g - 2)) * L + kk) * (size_t)N;
    load_rows(sm, src + tile_base, tid);
    __syncthreads();
    fwd_range(true, sm, md.ftw, m, kRowTileLog, m, logN - m, tile_base, md.q, tid);

    const size_t LN = (size_t)L * N;
    for (uint32_t j = threadIdx.x; j < M; j += blockDim.x) {
        const size_t n = (size_t)kk * N + tile_base + j;
        u64 h0 = 0, l0 = smem[sl(j)], h1 = 0, l1 = smem[P + sl(j)];
#pragma unroll
        for (int i = 0; i < PSI_MAX_LIMBS; i++)
            if (i < (int)L) {
                const u64 d = smem[(2 + i) * P + sl(j)];
                mac128(h0, l0, d, evk_b[(size_t)i * LN + n]);
                mac128(h1, l1, d, evk_a[(size_t)i * LN + n]);
            }
        u64 r0 = barrett128(h0, l0, md.q, md.mu_hi, md.mu_lo);
        u64 r1 = barrett128(h1, l1, md.q, md.mu_hi, md.mu_lo);
        if (mask) {
            const u64 mv = mask[bin * LN + n];
            r0 = mulmod(r0, mv, md);
            r1 = mulmod(r1, mv, md);
        }
        out[(bin * 2) * LN + n] = r0;
        out[(bin * 2 + 1) * LN + n] = r1;
    }
}

// ---- launcher --------------------------------------------------------------------------------------
// column kernels run one 64-thread group per limb of Q*P: at most PSI_MAX_LIMBS groups (512 threads)
bool fused_mul_supported(const KCtx& k) { return k.logN >= kLogCols + kRowTileLog && k.L + k.Lp <= PSI_MAX_LIMBS; }

cudaError_t launch_fused_mul(const KCtx& k, uint32_t B, const u64* a, const u64* b, u64* ha, u64* hb, u64* e1p,
                             u64* e2h, u64* th, u64* rh, u64* dh, const u64* evk_b, const u64* evk_a, const u64* mask,
                             u64* out) {
    if (B == 0) return cudaSuccess;
    const uint32_t L = k.L, Lp = k.Lp, LT = L + Lp;
    const uint32_t logR = k.logN - kLogCols;
    const uint32_t row_tiles = (1u << logR) >> kRowTileLog, col_tiles = (1u << kLogCols) >> kColTileLog;
    const size_t row_arr = padded(1u << (kLogCols + kRowTileLog)) * sizeof(u64);
    const size_t col_arr = padded(1u << (logR + kColTileLog)) * sizeof(u64);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e;
        if ((e = cudaFuncSetAttribute(k_cols_extend, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_cols_scale, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_rows_relin, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return e;
        attr_set = true;
    }
    k_rows_inv<<<dim3(row_tiles, L, B), 4 * kGroup, 4 * row_arr, k.s>>>(k.tab, k.logN, a, b, ha, hb);
    k_cols_extend<<<dim3(col_tiles, 4, B), LT * kGroup, LT * col_arr, k.s>>>(k.tab, k.logN, ha, hb, e1p, e2h);
    k_rows_tensor<<<dim3(row_tiles, LT, B), 4 * kGroup, 4 * row_arr, k.s>>>(k.tab, k.logN, a, e1p, e2h, th);
    k_cols_scale<<<dim3(col_tiles, 3, B), LT * kGroup, LT * col_arr, k.s>>>(k.tab, k.logN, th, rh, dh);
    k_rows_relin<<<dim3(row_tiles, L, B), (2 + L) * kGroup, (2 + L) * row_arr, k.s>>>(k.tab, k.logN, rh, dh, evk_b,
                                                                                       evk_a, mask, out);
    return cudaGetLastError();
}

}  // namespace psi
