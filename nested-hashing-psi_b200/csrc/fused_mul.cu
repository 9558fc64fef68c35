// Row kernels, launcher and dispatch of the fused EvalMult(ct,ct) pipeline; the device code shared with the
// column-kernel translation units lives in fused_mul.cuh (see the design notes at its top).
#include <cstdlib>

#include "fused_mul.cuh"

namespace psi {

// ---- (1) rows, inverse: both operands, all 4L limb-polys of a bin --------------------------------
// grid (R/8, L, B), 4 groups: a.c0, a.c1, b.c0, b.c1 of limb blockIdx.y.  ops: bit 0 = operand a, bit 1 = operand b
// (the streamed single-query path prepares operand a while the index ciphertexts of operand b are still uploading)
__global__ void __launch_bounds__(4 * kGroup, 3) k_rows_inv(const DevTables* __restrict__ tab, uint32_t logN,
                                                            const u64* __restrict__ a, const u64* __restrict__ b,
                                                            u64* __restrict__ ha, u64* __restrict__ hb, uint32_t ops) {
    extern __shared__ __align__(16) u64 smem[];
    const uint32_t N = 1u << logN, L = tab->L;
    const uint32_t m = kLogCols + kRowTileLog, M = 1u << m, P = padded(M);
    const uint32_t g = threadIdx.x / kGroup, tid = threadIdx.x % kGroup;
    const uint32_t l = blockIdx.y, tile_base = blockIdx.x * M;
    const size_t bin = blockIdx.z;
    const size_t off = ((bin * 2 + (g & 1)) * L + l) * N + tile_base;
    __shared__ __align__(8) uint64_t tw_bar;
    ulonglong2* tws = reinterpret_cast<ulonglong2*>(smem);  // [1016] inverse twiddles of this row tile
    u64* arr = smem + kRowTwWords;
    if (threadIdx.x == 0) {
        mbar_init(&tw_bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&tw_bar, kRowTwEntries * 16u);
        stage_row_twiddles(tws, tab->mods[l].itw_rows, blockIdx.x, &tw_bar);
    }
    const bool active = (ops >> (g >> 1)) & 1u;
    // experiment builds (results are garbage): bit 0 = no tile load, bit 1 = no transform, bit 2 = no store.  MEASURED at
    // config B: 73.0 us as is, 65.4 without load and store, 33.1 without the transform (tools/p2_variants.sh)
#ifndef PSI_DBG_SKIP
#define PSI_DBG_SKIP 0
#endif
    if (active && !(PSI_DBG_SKIP & 1)) load_rows(arr + g * P, (g < 2 ? a : b) + off, tid);
    loads_wait();
    __syncthreads();
    mbar_wait(&tw_bar, 0);
    // an inactive group owns no array (group index 4 = none) and takes part in no group barrier
    if (!(PSI_DBG_SKIP & 2))
        transform_rows<true>(tab, arr, P, 4, [l](uint32_t) { return l; }, active ? g : 4, 4, logN, tile_base, tid, tws);
    if (active && !(PSI_DBG_SKIP & 4)) store_rows(arr + g * P, (g < 2 ? ha : hb) + off, tid);
}

// ---- (3) rows: forward, tensor, inverse -----------------------------------------------------------
// grid (R/8, LT, B), 4 groups (a0, a1, b0, b1 of limb blockIdx.y); a: [B][2][L][N] EVALUATION (Q limbs
// are used as given), e1p/e2h from (2); th: [B][3][LT][N] row-inverse halves
__global__ void __launch_bounds__(4 * kGroup, 3) k_rows_tensor(const DevTables* __restrict__ tab, uint32_t logN,
                                                               const u64* __restrict__ a, const u64* __restrict__ e1p,
                                                               const u64* __restrict__ e2h, u64* __restrict__ th) {
    extern __shared__ __align__(16) u64 smem[];
    const uint32_t N = 1u << logN, L = tab->L, Lp = tab->Lp, LT = L + Lp;
    const uint32_t m = kLogCols + kRowTileLog, M = 1u << m, P = padded(M);
    const uint32_t g = threadIdx.x / kGroup, tid = threadIdx.x % kGroup;
    const uint32_t l = blockIdx.y, tile_base = blockIdx.x * M;
    const size_t bin = blockIdx.z;
    const ModDev& md = tab->mods[l];
    const uint32_t comp = g & 1;
    const u64* src;
    if (g < 2)
        src = l < L ? a + ((bin * 2 + comp) * L + l) * (size_t)N : e1p + ((bin * 2 + comp) * Lp + (l - L)) * (size_t)N;
    else
        src = e2h + ((bin * 2 + comp) * LT + l) * (size_t)N;
    __shared__ __align__(8) uint64_t tw_bar;
    ulonglong2* tws_f = reinterpret_cast<ulonglong2*>(smem);                  // forward twiddles of this row tile
    ulonglong2* tws_i = reinterpret_cast<ulonglong2*>(smem + kRowTwWords);    // inverse twiddles
    u64* arr = smem + 2 * kRowTwWords;
    if (threadIdx.x == 0) {
        mbar_init(&tw_bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&tw_bar, 2 * kRowTwEntries * 16u);
        stage_row_twiddles(tws_f, md.ftw_rows, blockIdx.x, &tw_bar);
        stage_row_twiddles(tws_i, md.itw_rows, blockIdx.x, &tw_bar);
    }
    load_rows(arr + g * P, src + tile_base, tid);
    loads_wait();
    __syncthreads();
    mbar_wait(&tw_bar, 0);
    // the Q limbs of the first operand are already in EVALUATION form: arrays 0, 1 are skipped for l < L
    const uint32_t first = l < L ? 2 : 0;
    transform_rows<false>(tab, arr + first * P, P, 4 - first, [l](uint32_t) { return l; }, g >= first ? g - first : 4,
                          4 - first, logN, tile_base, tid, tws_f);
    __syncthreads();  // the tensor product reads all four arrays

    // Tensor product with ONE Montgomery reduction per output: the factor R^-1 it leaves is undone for
    // free by k_cols_scale, whose N^-1 constant is N^-1 * R.  Operands are brought below 2q + 2^32 first
    // so that every 128-bit sum stays below q * 2^64; outputs are in [0, 2q), which the inverse row pass
    // accepts as is.
    const u64 q = md.q, qinv = md.qinv;
    for (uint32_t j = threadIdx.x; j < M; j += blockDim.x) {
        const u64 a0 = lazy_below_2q(arr[sl(j)], q), a1 = lazy_below_2q(arr[P + sl(j)], q);
        const u64 b0 = lazy_below_2q(arr[2 * P + sl(j)], q), b1 = lazy_below_2q(arr[3 * P + sl(j)], q);
        u64 hi, lo;
        mul128(hi, lo, a0, b0);
        arr[sl(j)] = mont_redc_lazy(hi, lo, q, qinv);
        mul128(hi, lo, a0, b1);
        mac128(hi, lo, a1, b0);
        arr[P + sl(j)] = mont_redc_lazy(hi, lo, q, qinv);
        mul128(hi, lo, a1, b1);
        arr[2 * P + sl(j)] = mont_redc_lazy(hi, lo, q, qinv);
    }
    __syncthreads();
    transform_rows<true>(tab, arr, P, 3, [l](uint32_t) { return l; }, g, 4, logN, tile_base, tid, tws_i);
    if (g < 3) store_rows(arr + g * P, th + ((bin * 3 + g) * LT + l) * (size_t)N + tile_base, tid);
}

// ---- (5) rows: forward, key-switch inner product, add (c0, c1), mask ------------------------------
// grid (R/8, L, B), 2 + L groups (c0, c1, digit 0..L-1 of limb blockIdx.y).  evk_bR / evk_aR / maskR are
// in Montgomery form (times R = 2^64), so each modular product is one 128-bit multiply-accumulate plus a
// Montgomery reduction that leaves no stray factor:  REDC(sum_i d_i * evkR_i) = sum_i d_i * evk_i.
// NG = groups per CTA: 2 + L (one array each) or 4 (arrays dealt round-robin); a template constant, because a run-time group
// count keeps a real loop around the unrolled register passes (measured: +2.6 % of phase 2 at the same shape)
template <int L, int NG>
__global__ void __launch_bounds__(NG * kGroup, NG * kGroup <= 256 ? 3 : (NG * kGroup <= 384 ? 2 : 1))
    k_rows_relin(const DevTables* __restrict__ tab, uint32_t logN, const u64* __restrict__ rh, const u64* __restrict__ dh,
                 const u64* __restrict__ evk_bR, const u64* __restrict__ evk_aR, const u64* __restrict__ maskR,
                 u64* __restrict__ out) {
    extern __shared__ __align__(16) u64 smem[];
    const uint32_t N = 1u << logN;
    const uint32_t m = kLogCols + kRowTileLog, M = 1u << m, P = padded(M);
    const uint32_t g = threadIdx.x / kGroup, tid = threadIdx.x % kGroup;
    const uint32_t kk = blockIdx.y, tile_base = blockIdx.x * M;
    const size_t bin = blockIdx.z;
    const ModDev& md = tab->mods[kk];
    constexpr uint32_t ng = NG;
    __shared__ __align__(8) uint64_t tw_bar;
    ulonglong2* tws = reinterpret_cast<ulonglong2*>(smem);  // forward twiddles of this row tile
    u64* arr = smem + kRowTwWords;
    if (threadIdx.x == 0) {
        mbar_init(&tw_bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&tw_bar, kRowTwEntries * 16u);
        stage_row_twiddles(tws, md.ftw_rows, blockIdx.x, &tw_bar);
    }
#pragma unroll
    for (uint32_t a = g; a < 2 + L; a += ng) {
        const u64* src = a < 2 ? rh + ((bin * 2 + a) * L + kk) * (size_t)N : dh + ((bin * L + (a - 2)) * L + kk) * (size_t)N;
        load_rows(arr + a * P, src + tile_base, tid);
    }
    loads_wait();
    __syncthreads();
    mbar_wait(&tw_bar, 0);
    transform_rows<false>(tab, arr, P, 2 + L, [kk](uint32_t) { return kk; }, g, ng, logN, tile_base, tid, tws);
    __syncthreads();  // the key-switch inner product reads all 2 + L arrays

    const size_t LN = (size_t)L * N;
    const u64 q = md.q, qinv = md.qinv;
    for (uint32_t j = threadIdx.x; j < M; j += blockDim.x) {
        const size_t n = (size_t)kk * N + tile_base + j;
        u64 h0 = 0, l0 = 0, h1 = 0, l1 = 0;
#pragma unroll
        for (int i = 0; i < L; i++) {
            const u64 d = lazy_below_2q(arr[(2 + i) * P + sl(j)], q);  // < 2q + 2^32: L <= 7 terms, 14 q^2 + small < q * 2^64 for q < 2^60
            mac128(h0, l0, d, evk_bR[(size_t)i * LN + n]);
            mac128(h1, l1, d, evk_aR[(size_t)i * LN + n]);
        }
        // + (c0, c1): forward-transform outputs < 2 kLazy q + 2^32 <= 8q + 2^32, reduction outputs < 2q
        u64 r0 = mont_redc_lazy(h0, l0, q, qinv) + arr[sl(j)];
        u64 r1 = mont_redc_lazy(h1, l1, q, qinv) + arr[P + sl(j)];
        if (maskR) {
            const u64 mv = maskR[bin * LN + n];
            u64 ph, pl;
            mul128(ph, pl, r0, mv);
            r0 = mont_redc_lazy(ph, pl, q, qinv);
            mul128(ph, pl, r1, mv);
            r1 = mont_redc_lazy(ph, pl, q, qinv);
            r0 = r0 >= q ? r0 - q : r0;
            r1 = r1 >= q ? r1 - q : r1;
        } else {
            r0 = reduce_pow2q<4>(r0, q);
            r1 = reduce_pow2q<4>(r1, q);
        }
        out[(bin * 2) * LN + n] = r0;
        out[(bin * 2 + 1) * LN + n] = r1;
    }
}

// constants into Montgomery form: dst = src * 2^64 mod q_l, polys laid out [groups][L][N]
__global__ void __launch_bounds__(256) k_to_montgomery(const DevTables* __restrict__ tab, uint32_t N, size_t total,
                                                       const u64* __restrict__ src, u64* __restrict__ dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const ModDev& md = tab->mods[(i / N) % tab->L];
    dst[i] = shoup_canon(src[i], md.Rmodq, md.Rmodq_s, md.q);
}
cudaError_t launch_to_montgomery(const KCtx& k, uint32_t groups, const u64* src, u64* dst) {
    const size_t total = (size_t)groups * k.L * k.N;
    if (total == 0) return cudaSuccess;
    k_to_montgomery<<<(unsigned)((total + 255) / 256), 256, 0, k.s>>>(k.tab, k.N, total, src, dst);
    return cudaGetLastError();
}

template <int L>
static cudaError_t relin_attr(int bytes) {
    cudaError_t e = cudaFuncSetAttribute(k_rows_relin<L, 2 + L>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess || L <= 2) return e;
    return cudaFuncSetAttribute(k_rows_relin<L, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

// Groups per k_rows_relin CTA: one per array (2 + L), or four that take the arrays round-robin (PSI_RELIN_GROUPS=4,
// or the policy below).  See profiles/r02_phase2_experiments.md 8 for the measurements.
static uint32_t relin_groups(uint32_t L, uint32_t B) {
    static const int forced = [] {
        const char* e = std::getenv("PSI_RELIN_GROUPS");
        return e ? std::atoi(e) : 0;
    }();
    (void)B;
    if (forced == 4) return 4;
    if (forced > 0) return 2 + L;
    // L = 4 (the BASELINE contexts): one group per array is faster at every bin count (47 bins: 1.0154 against 1.0227 ms).
    // L >= 5 (448 - 576 threads, one CTA per SM): four groups, phase 2 -1.3 % ... -1.8 % at 6 and 24 bins.
    return L >= 5 ? 4 : 2 + L;
}

// ---- launcher --------------------------------------------------------------------------------------
// The column kernels are instantiated for the limb counts BFVrns produces for this path: sizeQ 1..7 (depth 3
// gives 4, depth 5 — E >= 500, BatchedFHEPSIClient.cpp:50-53 — gives 6 at N = 16384), sizeP = sizeQ or
// sizeQ + 1; anything else takes the unfused kernels.  Bounds that hold up to 7 limbs: 8-term split-30 sums
// (16 products < 2^60 per partial sum), L * (2q + 2^32) < 2^64 for the key-switch Montgomery reduction.
bool fused_mul_supported(const KCtx& k) {
    return k.fused_ok && k.logN >= kLogCols + kRowTileLog && k.L >= 1 && k.L <= 7 && (k.Lp == k.L || (k.Lp == k.L + 1 && k.L <= 6));
}

static cudaError_t dispatch_cols(PSI_COLS_ARGS) {
    cudaError_t e = dispatch_cols_a(k, B, ha, hb, e1p, e2h, th, rh, dh, which);
    if (e == cudaErrorInvalidValue) e = dispatch_cols_b(k, B, ha, hb, e1p, e2h, th, rh, dh, which);
    if (e == cudaErrorInvalidValue) e = dispatch_cols_c(k, B, ha, hb, e1p, e2h, th, rh, dh, which);
    if (e == cudaErrorInvalidValue) e = dispatch_cols_d(k, B, ha, hb, e1p, e2h, th, rh, dh, which);
    return e;
}

// Function attributes are per device: psi_ctx_create calls this once for the context's device.
cudaError_t fused_mul_init_device(const KCtx& k) {
    if (!fused_mul_supported(k)) return cudaSuccess;
    return launch_fused_mul(k, 0xffffffffu, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                            nullptr, nullptr, nullptr, nullptr, 3);
}

cudaError_t launch_fused_mul(const KCtx& k, uint32_t B, const u64* a, const u64* b, u64* ha, u64* hb, u64* e1p,
                             u64* e2h, u64* th, u64* rh, u64* dh, const u64* evk_b, const u64* evk_a, const u64* mask,
                             u64* out, uint32_t ops) {
    const uint32_t L = k.L, LT = k.L + k.Lp;
    const uint32_t logR = k.logN - kLogCols;
    const uint32_t row_tiles = (1u << logR) >> kRowTileLog;
    const size_t row_arr = padded(1u << (kLogCols + kRowTileLog)) * sizeof(u64);
    const size_t tw_bytes = kRowTwWords * sizeof(u64);
    cudaError_t e;
    if (B == 0xffffffffu) {  // per-device set-up (fused_mul_init_device)
        if ((e = dispatch_cols(k, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, -1)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_rows_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_rows_tensor, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024)) != cudaSuccess) return e;
        if ((e = relin_attr<1>(80 * 1024)) != cudaSuccess) return e;
        if ((e = relin_attr<2>(80 * 1024)) != cudaSuccess) return e;
        if ((e = relin_attr<3>(80 * 1024)) != cudaSuccess) return e;
        if ((e = relin_attr<4>(80 * 1024)) != cudaSuccess) return e;
        if ((e = relin_attr<5>(100 * 1024)) != cudaSuccess) return e;
        if ((e = relin_attr<6>(100 * 1024)) != cudaSuccess) return e;
        return relin_attr<7>(100 * 1024);
    }
    if (B == 0) return cudaSuccess;

    // ops: bit 0 = the part that only needs operand a (row-inverse + exact Q -> P extension of a), bit 1 = everything else
    k_rows_inv<<<dim3(row_tiles, L, B), 4 * kGroup, 4 * row_arr + tw_bytes, k.s>>>(k.tab, k.logN, a, b, ha, hb, ops);
    if ((e = dispatch_cols(k, B, ha, hb, e1p, e2h, nullptr, nullptr, nullptr, (int)ops - 1)) != cudaSuccess) return e;
    if (!(ops & 2u)) return cudaGetLastError();
    k_rows_tensor<<<dim3(row_tiles, LT, B), 4 * kGroup, 4 * row_arr + 2 * tw_bytes, k.s>>>(k.tab, k.logN, a, e1p, e2h, th);
    if ((e = dispatch_cols(k, B, nullptr, nullptr, nullptr, nullptr, th, rh, dh, 3)) != cudaSuccess) return e;
    const dim3 rg(row_tiles, L, B);
    const size_t rs = (2 + L) * row_arr + tw_bytes;
    const bool four = relin_groups(L, B) == 4 && L > 2;
#define PSI_RELIN_CASE(l)                                                                                              \
    case l:                                                                                                            \
        if (four)                                                                                                      \
            k_rows_relin<l, 4><<<rg, 4 * kGroup, rs, k.s>>>(k.tab, k.logN, rh, dh, evk_b, evk_a, mask, out);           \
        else                                                                                                           \
            k_rows_relin<l, 2 + l><<<rg, (2 + l) * kGroup, rs, k.s>>>(k.tab, k.logN, rh, dh, evk_b, evk_a, mask, out); \
        break;
    switch (L) {
        PSI_RELIN_CASE(1)
        PSI_RELIN_CASE(2)
        PSI_RELIN_CASE(3)
        PSI_RELIN_CASE(4)
        PSI_RELIN_CASE(5)
        PSI_RELIN_CASE(6)
        default:
            if (four)
                k_rows_relin<7, 4><<<rg, 4 * kGroup, rs, k.s>>>(k.tab, k.logN, rh, dh, evk_b, evk_a, mask, out);
            else
                k_rows_relin<7, 9><<<rg, 9 * kGroup, rs, k.s>>>(k.tab, k.logN, rh, dh, evk_b, evk_a, mask, out);
            break;
    }
#undef PSI_RELIN_CASE
    return cudaGetLastError();
}

}  // namespace psi
