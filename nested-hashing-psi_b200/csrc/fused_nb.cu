// Fused automorphism key switch of the non-batched FHEHIPPIE path (psi_nonbatched.cu; reference FHEHIPPIE.cpp:61-77:
// every EvalSum step and every EvalMerge rotation is EvalAutomorphism = KeySwitchBV of c1 + AutomorphismTransform).
// Built from the row / column machinery of fused_mul.cuh, three kernels per step instead of INTT, digit lift, NTT and
// inner product as separate whole-polynomial launches:
//
//   k_nb_rows_inv       c1 of every item (EVALUATION)  -> row-inverse halves                  (L limb-polys / item)
//   k_nb_cols_digits    column-inverse, canonical coefficients (N^-1 folded in), BV digit lift with the centred
//                       SwitchModulus, column-forward of the L x L digit limbs
//   k_nb_rows_ks        row-forward of the digits, sum_i digit_i * key_i (+ c0), the EVALUATION-format permutation as a
//                       scatter inside a 128-coefficient block, and the EvalAdd of EvalSum
//
// A digit limb crosses HBM/L2 twice instead of four times, the digit-lift and accumulate kernels disappear, and the
// transforms run as the compiled radix-16 / radix-8 register passes.  Keys come in Montgomery form (times 2^64 mod q_k):
// one 128-bit multiply-accumulate per term and one Montgomery reduction per output, as in k_rows_relin.
// Results are canonical residues, bit-identical to the unfused sequence and to the oracle (tests/test_gpu_nonbatched.py).
#include "fused_mul.cuh"

namespace psi {

// PrecomputeAutoMap (recalled): output position p of the bit-reversed EVALUATION vector reads input position automap(p, g)
__device__ __forceinline__ uint32_t nb_automap(uint32_t p, uint32_t g, uint32_t logN) {
    const uint32_t j = __brev(p) >> (32 - logN);
    const uint32_t idx = (((2 * j + 1) * g) & ((2u << logN) - 1)) >> 1;
    return __brev(idx) >> (32 - logN);
}

// ---- (1) rows, inverse: c1 of four items per CTA ---------------------------------------------------
// grid (R/8, L, ceil(B/4)); cur: [B][2][L][N] EVALUATION; h: [B][L][N] row-inverse halves of c1
__global__ void __launch_bounds__(4 * kGroup, 3) k_nb_rows_inv(const DevTables* __restrict__ tab, uint32_t logN, uint32_t B,
                                                               const u64* __restrict__ cur, u64* __restrict__ h) {
    extern __shared__ __align__(16) u64 smem[];
    const uint32_t N = 1u << logN, L = tab->L;
    const uint32_t m = kLogCols + kRowTileLog, M = 1u << m, P = padded(M);
    const uint32_t g = threadIdx.x / kGroup, tid = threadIdx.x % kGroup;
    const uint32_t l = blockIdx.y, tile_base = blockIdx.x * M;
    const size_t item = (size_t)blockIdx.z * 4 + g;
    const bool active = item < B;
    __shared__ __align__(8) uint64_t tw_bar;
    ulonglong2* tws = reinterpret_cast<ulonglong2*>(smem);
    u64* arr = smem + kRowTwWords;
    if (threadIdx.x == 0) {
        mbar_init(&tw_bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&tw_bar, kRowTwEntries * 16u);
        stage_row_twiddles(tws, tab->mods[l].itw_rows, blockIdx.x, &tw_bar);
    }
    if (active) load_rows(arr + g * P, cur + ((item * 2 + 1) * L + l) * N + tile_base, tid);
    loads_wait();
    __syncthreads();
    mbar_wait(&tw_bar, 0);
    transform_rows<true>(tab, arr, P, 4, [l](uint32_t) { return l; }, active ? g : 4, 4, logN, tile_base, tid, tws);
    if (active) store_rows(arr + g * P, h + (item * L + l) * N + tile_base, tid);
}

// ---- (2) columns: inverse, canonical coefficient, digit lift, forward ------------------------------
// grid (128/8, B), kColGroups groups.  h: [B][L][N]; dh: [B][L(i)][L(k)][N] column-forward halves of the BV digits
template <int L, int LOGN_CT>
__global__ void __launch_bounds__(kColGroups* kGroup, 3)
    k_nb_cols_digits(const DevTables* __restrict__ tab, uint32_t logN, const u64* __restrict__ h, u64* __restrict__ dh) {
    extern __shared__ __align__(16) u64 smem[];
    const uint32_t N = 1u << logN;
    const uint32_t logR = logN - kLogCols, m = logR + kColTileLog, M = 1u << m, P = padded(M);
    const uint32_t g = threadIdx.x / kGroup, tid = threadIdx.x % kGroup;
    const uint32_t c0 = blockIdx.x << kColTileLog;
    const size_t item = blockIdx.y;
    for (uint32_t a = g; a < L; a += kColGroups) load_cols(smem + a * P, h + (item * L + a) * (size_t)N, logR, c0, tid);
    loads_wait();
    __syncthreads();
    transform_cols<true, LOGN_CT>(tab, smem, P, L, [](uint32_t a) { return a; }, g, kColGroups, logN, tid);
    // SetFormat(COEFFICIENT) is complete once N^-1 is applied: canonical residues, what CRTDecompose switches
    for (uint32_t a = g; a < L; a += kColGroups) {
        const ModDev& md = tab->mods[a];
        for (uint32_t j = tid; j < M; j += kGroup) smem[a * P + sl(j)] = shoup_canon(smem[a * P + sl(j)], md.ninv, md.ninv_s, md.q);
    }
    __syncthreads();  // every group reads every limb in the lift below
    // DCRTPoly::CRTDecompose (BV, digit size 0): digit i = limb i switched to every q_k with the centred lift of
    // NativeVector::SwitchModulus; arrays L..2L-1 hold the L limbs of the current digit
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
        group_sync();  // lift, transform and store of digit limb kk all belong to group kk % kColGroups
        transform_cols<false, LOGN_CT>(tab, smem + L * P, P, L, [](uint32_t a) { return a; }, g, kColGroups, logN, tid);
        for (uint32_t kk = g; kk < L; kk += kColGroups)
            store_cols(smem + (L + kk) * P, dh + ((item * L + i) * L + kk) * (size_t)N, logR, c0, tid);
        group_sync();
    }
}

// ---- (3) rows: forward, key-switch inner product, permutation, add ---------------------------------
// grid (R/8, L, B), L groups (digit 0..L-1 of limb blockIdx.y).  key_bR / key_aR: [n_keys][L][L][N] in Montgomery form.
// The item's key slot and inverse automorphism index come from sel_key / sel_ginv[sel_mod ? item % sel_mod : 0];
// slot -1 = the identity (bin 0 of EvalMerge).  ADD: out = cur + sigma_g(keyswitched cur), else out = sigma_g(...).
// A 128-aligned block of positions maps onto a 128-aligned block (the low bits of a bit-reversed position only feed
// the low bits of its image), so the scattered stores of a row stay inside one row.
template <int L, bool ADD>
__global__ void __launch_bounds__(L* kGroup, L <= 4 ? 3 : 2)
    k_nb_rows_ks(const DevTables* __restrict__ tab, uint32_t logN, const u64* __restrict__ cur, const u64* __restrict__ dh,
                 const u64* __restrict__ key_bR, const u64* __restrict__ key_aR, const int* __restrict__ sel_key,
                 const uint32_t* __restrict__ sel_ginv, uint32_t sel_mod, u64* __restrict__ out) {
    extern __shared__ __align__(16) u64 smem[];
    const uint32_t N = 1u << logN;
    const uint32_t m = kLogCols + kRowTileLog, M = 1u << m, P = padded(M);
    const uint32_t g = threadIdx.x / kGroup, tid = threadIdx.x % kGroup;
    const uint32_t kk = blockIdx.y, tile_base = blockIdx.x * M;
    const size_t item = blockIdx.z;
    const size_t LN = (size_t)L * N;
    const uint32_t sel = sel_mod ? (uint32_t)(item % sel_mod) : 0;
    const int slot = sel_key[sel];
    const u64* ct = cur + item * 2 * LN;
    u64* o = out + item * 2 * LN;
    if (slot < 0) {  // uniform over the CTA
        for (uint32_t j = threadIdx.x; j < M; j += blockDim.x) {
            const size_t n = (size_t)kk * N + tile_base + j;
            o[n] = ct[n];
            o[LN + n] = ct[LN + n];
        }
        return;
    }
    const ModDev& md = tab->mods[kk];
    __shared__ __align__(8) uint64_t tw_bar;
    ulonglong2* tws = reinterpret_cast<ulonglong2*>(smem);  // forward twiddles of this row tile
    u64* arr = smem + kRowTwWords;
    if (threadIdx.x == 0) {
        mbar_init(&tw_bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&tw_bar, kRowTwEntries * 16u);
        stage_row_twiddles(tws, md.ftw_rows, blockIdx.x, &tw_bar);
    }
    load_rows(arr + g * P, dh + ((item * L + g) * L + kk) * (size_t)N + tile_base, tid);
    loads_wait();
    __syncthreads();
    mbar_wait(&tw_bar, 0);
    transform_rows<false>(tab, arr, P, L, [kk](uint32_t) { return kk; }, g, L, logN, tile_base, tid, tws);
    __syncthreads();  // the inner product reads all L arrays

    const u64 q = md.q, qinv = md.qinv;
    const u64* kb = key_bR + (size_t)slot * L * LN;
    const u64* ka = key_aR + (size_t)slot * L * LN;
    const uint32_t ginv = sel_ginv[sel];
    for (uint32_t j = threadIdx.x; j < M; j += blockDim.x) {
        const size_t n = (size_t)kk * N + tile_base + j;
        u64 h0 = 0, l0 = 0, h1 = 0, l1 = 0;
#pragma unroll
        for (int i = 0; i < L; i++) {
            const u64 d = lazy_below_2q(arr[i * P + sl(j)], q);  // < 2q + 2^32: L <= 7 terms stay below q * 2^64
            mac128(h0, l0, d, kb[(size_t)i * LN + n]);
            mac128(h1, l1, d, ka[(size_t)i * LN + n]);
        }
        u64 r0 = mont_redc_lazy(h0, l0, q, qinv), r1 = mont_redc_lazy(h1, l1, q, qinv);  // < 2q
        r0 = r0 >= q ? r0 - q : r0;
        r1 = r1 >= q ? r1 - q : r1;
        r0 = addmod(r0, ct[n], q);  // KeySwitchInPlace: c0 += d0, c1 = d1
        const size_t dst = (size_t)kk * N + nb_automap(tile_base + j, ginv, logN);
        if (ADD) {
            r0 = addmod(r0, ct[dst], q);
            r1 = addmod(r1, ct[LN + dst], q);
        }
        o[dst] = r0;
        o[LN + dst] = r1;
    }
}

// ---- launcher ---------------------------------------------------------------------------------------
template <int L>
static cudaError_t nb_launch_L(const KCtx& k, uint32_t B, const u64* cur, u64* h, u64* dh, const u64* key_bR, const u64* key_aR,
                               const int* sel_key, const uint32_t* sel_ginv, uint32_t sel_mod, bool add, u64* out, bool init) {
    const uint32_t logR = k.logN - kLogCols, col_tiles = (1u << kLogCols) >> kColTileLog;
    const uint32_t row_tiles = (1u << logR) >> kRowTileLog;
    const size_t row_arr = padded(1u << (kLogCols + kRowTileLog)) * sizeof(u64), tw_bytes = kRowTwWords * sizeof(u64);
    const size_t col_smem = (size_t)2 * L * padded(1u << (logR + kColTileLog)) * sizeof(u64);
    const size_t ks_smem = (size_t)L * row_arr + tw_bytes;
    cudaError_t e;
    if (init) {  // per-device function attributes (dynamic shared memory above 48 KiB)
        if ((e = cudaFuncSetAttribute(k_nb_rows_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_nb_cols_digits<L, 14>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_nb_cols_digits<L, 13>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_nb_cols_digits<L, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_nb_rows_ks<L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)) != cudaSuccess) return e;
        return cudaFuncSetAttribute(k_nb_rows_ks<L, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    }
    k_nb_rows_inv<<<dim3(row_tiles, L, (B + 3) / 4), 4 * kGroup, 4 * row_arr + tw_bytes, k.s>>>(k.tab, k.logN, B, cur, h);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const dim3 cg(col_tiles, B);
    if (k.logN == 14)
        k_nb_cols_digits<L, 14><<<cg, kColGroups * kGroup, col_smem, k.s>>>(k.tab, k.logN, h, dh);
    else if (k.logN == 13)
        k_nb_cols_digits<L, 13><<<cg, kColGroups * kGroup, col_smem, k.s>>>(k.tab, k.logN, h, dh);
    else
        k_nb_cols_digits<L, 0><<<cg, kColGroups * kGroup, col_smem, k.s>>>(k.tab, k.logN, h, dh);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const dim3 rg(row_tiles, L, B);
    if (add)
        k_nb_rows_ks<L, true><<<rg, L * kGroup, ks_smem, k.s>>>(k.tab, k.logN, cur, dh, key_bR, key_aR, sel_key, sel_ginv, sel_mod, out);
    else
        k_nb_rows_ks<L, false><<<rg, L * kGroup, ks_smem, k.s>>>(k.tab, k.logN, cur, dh, key_bR, key_aR, sel_key, sel_ginv, sel_mod, out);
    return cudaGetLastError();
}

// the fused key switch covers the ring dimensions and limb counts of the fused ct x ct kernels (N >= 1024, L <= 7; BV)
bool nb_fused_supported(const KCtx& k) {
    if (k.logN < kLogCols + kRowTileLog || k.L < 1 || k.L > 7 || k.Lk != 0) return false;
    // the column kernel keeps 2 L column tiles in shared memory
    return (size_t)2 * k.L * padded(1u << (k.logN - kLogCols + kColTileLog)) * sizeof(u64) <= 200 * 1024;
}

static cudaError_t nb_dispatch(const KCtx& k, uint32_t B, const u64* cur, u64* h, u64* dh, const u64* key_bR, const u64* key_aR,
                               const int* sel_key, const uint32_t* sel_ginv, uint32_t sel_mod, bool add, u64* out, bool init) {
    switch (k.L) {
        case 1: return nb_launch_L<1>(k, B, cur, h, dh, key_bR, key_aR, sel_key, sel_ginv, sel_mod, add, out, init);
        case 2: return nb_launch_L<2>(k, B, cur, h, dh, key_bR, key_aR, sel_key, sel_ginv, sel_mod, add, out, init);
        case 3: return nb_launch_L<3>(k, B, cur, h, dh, key_bR, key_aR, sel_key, sel_ginv, sel_mod, add, out, init);
        case 4: return nb_launch_L<4>(k, B, cur, h, dh, key_bR, key_aR, sel_key, sel_ginv, sel_mod, add, out, init);
        case 5: return nb_launch_L<5>(k, B, cur, h, dh, key_bR, key_aR, sel_key, sel_ginv, sel_mod, add, out, init);
        case 6: return nb_launch_L<6>(k, B, cur, h, dh, key_bR, key_aR, sel_key, sel_ginv, sel_mod, add, out, init);
        case 7: return nb_launch_L<7>(k, B, cur, h, dh, key_bR, key_aR, sel_key, sel_ginv, sel_mod, add, out, init);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t nb_fused_init_device(const KCtx& k) {
    if (!nb_fused_supported(k)) return cudaSuccess;
    return nb_dispatch(k, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, false, nullptr, true);
}

cudaError_t launch_nb_keyswitch(const KCtx& k, uint32_t B, const u64* cur, u64* h, u64* dh, const u64* key_bR, const u64* key_aR,
                                const int* sel_key, const uint32_t* sel_ginv, uint32_t sel_mod, bool add, u64* out) {
    if (B == 0) return cudaSuccess;
    return nb_dispatch(k, B, cur, h, dh, key_bR, key_aR, sel_key, sel_ginv, sel_mod, add, out, false);
}

}  // namespace psi
