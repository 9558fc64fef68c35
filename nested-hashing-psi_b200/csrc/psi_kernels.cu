// sm_100a kernels of the BatchedFHEPIE server evaluation (everything except the NTT, which
// lives in ntt.cu).  Each kernel names the OpenFHE operation at the reference call site it
// replaces; the CPU restatement it is checked against is oracle/psi_oracle.c.
#include "psi_kernels.cuh"

namespace psi {

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

__device__ __forceinline__ u64 ld_stream(const u64* p) {
    // read-once data (plaintext DB): keep it out of L1
    u64 v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// ------------------------------------------------------------------------------------------
// Phase 1 — the encrypted one-hot inner product.
// Replaces the EvalMult(ct,pt) / EvalAdd loop and the EvalAdd of minusCompareElement
// (/root/reference/.../BatchedFHEHIPPIE.cpp:101-116).  HBM-bound: every plaintext word is read
// exactly once; BT bins share the two index-ciphertext words held in registers; the sum over pos
// is accumulated lazily in 128 bits and reduced once (canonical residue == term-by-term ModMul/ModAdd).
// ------------------------------------------------------------------------------------------
template <int BT>
__global__ void __launch_bounds__(256) k_mac(const DevTables* __restrict__ tab, uint32_t N, uint32_t L, uint32_t b,
                                             uint32_t E, const u64* __restrict__ pt, const u64* __restrict__ idx,
                                             const u64* __restrict__ minus, u64* __restrict__ acc) {
    const size_t LN = (size_t)L * N;
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= LN) return;
    const uint32_t nbg = (b + BT - 1) / BT;
    const uint32_t hf = blockIdx.y / nbg;
    const uint32_t bin0 = (blockIdx.y % nbg) * BT;
    const ModDev& md = tab->mods[c / N];
    const u64 q = md.q, mu_hi = md.mu_hi, mu_lo = md.mu_lo;

    u64 lo0[BT], hi0[BT], lo1[BT], hi1[BT];
#pragma unroll
    for (int j = 0; j < BT; j++) lo0[j] = hi0[j] = lo1[j] = hi1[j] = 0;

    const u64* ip = idx + (size_t)hf * E * 2 * LN + c;
    const size_t bin_stride = (size_t)E * LN;
    const u64* pp = pt + ((size_t)hf * b + bin0) * bin_stride + c;
    for (uint32_t pos = 0; pos < E; pos++) {
        const u64 i0 = ip[0], i1 = ip[LN];
        ip += 2 * LN;
#pragma unroll
        for (int j = 0; j < BT; j++) {
            if (bin0 + j < b) {
                const u64 pv = ld_stream(pp + (size_t)j * bin_stride);
                mac128(hi0[j], lo0[j], i0, pv);
                mac128(hi1[j], lo1[j], i1, pv);
            }
        }
        pp += LN;
        if ((pos & 127u) == 127u) {  // keep the lazy sum below 2^128 for any E
#pragma unroll
            for (int j = 0; j < BT; j++) {
                lo0[j] = barrett128(hi0[j], lo0[j], q, mu_hi, mu_lo);
                lo1[j] = barrett128(hi1[j], lo1[j], q, mu_hi, mu_lo);
                hi0[j] = hi1[j] = 0;
            }
        }
    }
    const u64 m0 = minus[c], m1 = minus[LN + c];
#pragma unroll
    for (int j = 0; j < BT; j++) {
        if (bin0 + j < b) {
            u64* o = acc + (((size_t)hf * b + bin0 + j) * 2) * LN + c;
            o[0] = addmod(barrett128(hi0[j], lo0[j], q, mu_hi, mu_lo), m0, q);
            o[LN] = addmod(barrett128(hi1[j], lo1[j], q, mu_hi, mu_lo), m1, q);
        }
    }
}

cudaError_t launch_mac(const KCtx& k, uint32_t K, uint32_t b, uint32_t E, const u64* pt, const u64* idx,
                       const u64* minus, u64* acc) {
    const size_t LN = (size_t)k.L * k.N;
    constexpr int BT = 4;
    dim3 grid(cdiv(LN, 256), K * ((b + BT - 1) / BT));
    k_mac<BT><<<grid, 256, 0, k.s>>>(k.tab, k.N, k.L, b, E, pt, idx, minus, acc);
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
            nu = __dadd_rn(nu, __dmul_rn(__ull2double_rn(y[i]), inv));
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
            nu = __dadd_rn(nu, __dmul_rn(tab->tQSfrac[i], __ull2double_rn(xp[i])));
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
cudaError_t launch_slots_to_crt(const KCtx& k, uint32_t n_pt, uint32_t nslots, const long long* slots,
                                const uint32_t* to_crt, u64* out) {
    LAUNCH_1D(k_slots_to_crt, (size_t)n_pt * k.N, n_pt, nslots, slots, to_crt, out)
}

// ------------------------------------------------------------------------------------------
// Integer-pipe peak: independent 32x32+64 multiply-add chains (IMAD.WIDE), the instruction the
// NTT butterflies and the lazy MACs are made of.  Denominator of the integer roofline.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_imad_peak(u64* out, uint32_t iters, uint32_t seed) {
    uint32_t a = threadIdx.x * 2654435761u + seed, b = blockIdx.x * 40503u + 12345u;
    u64 c0 = a, c1 = b, c2 = a ^ b, c3 = a + b, c4 = 1, c5 = 2, c6 = 3, c7 = 4;
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c0) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c1) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c2) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c3) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c4) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c5) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c6) : "r"(a), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c7) : "r"(a), "r"(b));
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}

cudaError_t imad_peak(int device, double* mads_per_second) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return e;
    const unsigned blocks = prop.multiProcessorCount * 8, threads = 256;
    const uint32_t iters = 4096;
    u64* out = nullptr;
    if ((e = cudaMalloc(&out, (size_t)blocks * threads * sizeof(u64))) != cudaSuccess) return e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(t0);
        k_imad_peak<<<blocks, threads>>>(out, iters, rep);
        cudaEventRecord(t1);
        if ((e = cudaEventSynchronize(t1)) != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, t0, t1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    cudaFree(out);
    if (e != cudaSuccess) return e;
    *mads_per_second = (double)blocks * threads * iters * 64.0 / (best * 1e-3);
    return cudaSuccess;
}

}  // namespace psi
