// C ABI (include/psi_b200.h) and kernel orchestration of the B200 BatchedFHEPIE server path.
//
// Replaces, behind plain pointers and sizes, what the reference does through OpenFHE objects:
//   context / relin key        BatchedFHEPSIServer.cpp:21-54  (deserialised CryptoContext, EvalMultKey)
//   plaintext DB + masks       BatchedFHEHIPPIE.cpp:37-82     (vectorizedHCT, preCalcRandomMask)
//   query                      BatchedFHEHIPPIE.hpp:40-48     (setIndex, setMinusCompareElement)
//   run()                      BatchedFHEHIPPIE.cpp:88-129
//   getResultList()            BatchedFHEHIPPIE.hpp:35-38
// There is no CPU path: every entry point that computes needs a CUDA device and says so.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "psi_b200.h"
#include "psi_kernels.cuh"
#include "psi_ctx.cuh"
#include "host_copy.hpp"
#include "../host/hashing.hpp"
#include "../host/psi_host_internal.hpp"

namespace psi {

static thread_local std::string g_last_error;

int set_error(int status, const std::string& msg) {
    g_last_error = msg;
    return status;
}

int cuda_fail(cudaError_t e, const char* what) {
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInitializationError ||
        e == cudaErrorSystemDriverMismatch || e == cudaErrorNotSupported)
        return set_error(PSI_ERR_NO_DEVICE, std::string(what) + ": no usable CUDA device (" + cudaGetErrorString(e) +
                                                "); this library has no CPU path");
    return set_error(PSI_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

typedef unsigned __int128 u128h;

static u64 h_mulmod(u64 a, u64 b, u64 q) { return (u64)(((u128h)a * b) % q); }
static u64 h_powmod(u64 a, u64 e, u64 q) {
    u64 r = 1;
    a %= q;
    for (; e; e >>= 1) {
        if (e & 1) r = h_mulmod(r, a, q);
        a = h_mulmod(a, a, q);
    }
    return r;
}
static u64 h_shoup(u64 w, u64 q) { return (u64)(((u128h)w << 64) / q); }
// c * 2^64 mod q, in split-30 storage (hi30 : lo30)
static u64 h_mont_split(u64 c, u64 q) {
    const u64 m = (u64)(((u128h)c << 64) % q);
    return ((m >> 30) << 32) | (m & 0x3fffffffull);
}
static uint32_t h_bitrev(uint32_t x, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

}  // namespace psi

using namespace psi;

namespace psi {

int ensure_device(psi_ctx* c) {
    CK(cudaSetDevice(c->device));
    return PSI_OK;
}

// PackedEncoding slot permutation (OpenFHE PackedEncoding::SetParams_2n as restated in oracle/psi_oracle.c):
// slot i <-> exponent 5^i, slot i + N/2 <-> cofactor * 5^i, bit-reversed transform order.  The co-factor is 3 in
// the OpenFHE 1.0.x line as recalled and 2N - 1 (the conjugation) in the other reading; the server and the client
// must agree on it whenever the server encodes the database itself (psi_set_packing_cofactor).
static int build_pack_perm(psi_ctx* c, uint32_t mode) {
    const uint32_t N = c->N;
    std::vector<uint32_t> perm(N);
    const u64 m2 = 2ull * N, cofactor = mode == PSI_PACK_COFACTOR_CONJ ? m2 - 1 : 3;
    u64 cur = 1;
    for (uint32_t i = 0; i < N / 2; i++) {
        perm[h_bitrev((uint32_t)((cur - 1) / 2), (int)c->logN)] = i;
        const u64 cof = (cur * cofactor) % m2;
        perm[h_bitrev((uint32_t)((cof - 1) / 2), (int)c->logN)] = i + N / 2;
        cur = (cur * 5) % m2;
    }
    CK(c->to_crt.alloc(N));
    CK(cudaMemcpy(c->to_crt.p, perm.data(), N * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return PSI_OK;
}

static int build_tables(psi_ctx* c) {
    const psi_params& P = c->P;
    const uint32_t N = c->N, L = c->L, Lp = c->Lp, nm = L + Lp + 1 + c->Lk;
    std::vector<u64> tw((size_t)nm * 4 * N);
    DevTables T;
    std::memset(&T, 0, sizeof(T));
    T.N = N;
    T.logN = c->logN;
    T.L = L;
    T.Lp = Lp;
    T.t = P.t;
    CK(c->twiddles.alloc(tw.size()));
    CK(c->twiddles2.alloc(tw.size()));
    std::vector<u64> tw2(tw.size());
    for (uint32_t m = 0; m < nm; m++) {
        const u64 q = m < L ? P.q[m] : (m < L + Lp ? P.p[m - L] : (m == L + Lp ? P.t : P.pk[m - L - Lp - 1]));
        const u64 psi_root = m < L ? P.psi_q[m] : (m < L + Lp ? P.psi_p[m - L] : (m == L + Lp ? P.psi_t : P.psi_pk[m - L - Lp - 1]));
        if (q < 2 || q >= (1ull << 60)) return set_error(PSI_ERR_INVALID, "moduli must be below 2^60 (OpenFHE's maximum for BFVrns)");
        if (h_powmod(psi_root, N, q) != q - 1)
            return set_error(PSI_ERR_INVALID, "psi is not a primitive 2N-th root of unity for one of the moduli");
        u64* w = &tw[((size_t)m * 4 + 0) * N];
        u64* ws = &tw[((size_t)m * 4 + 1) * N];
        u64* iw = &tw[((size_t)m * 4 + 2) * N];
        u64* iws = &tw[((size_t)m * 4 + 3) * N];
        const u64 ipsi = h_powmod(psi_root, q - 2, q);
        u64 pw = 1, ipw = 1;
        for (uint32_t i = 0; i < N; i++) {
            const uint32_t r = h_bitrev(i, (int)c->logN);
            w[r] = pw;
            iw[r] = ipw;
            pw = h_mulmod(pw, psi_root, q);
            ipw = h_mulmod(ipw, ipsi, q);
        }
        for (uint32_t i = 0; i < N; i++) {
            ws[i] = h_shoup(w[i], q);
            iws[i] = h_shoup(iw[i], q);
        }
        ModDev& md = T.mods[m];
        md.q = q;
        const u128h mu = (~(u128h)0) / q;
        md.mu_hi = (u64)(mu >> 64);
        md.mu_lo = (u64)mu;
        md.ninv = h_powmod(N, q - 2, q);
        md.ninv_s = h_shoup(md.ninv, q);
        {
            u64 inv = q;  // Newton iteration for q^-1 mod 2^64 (q odd): 5 steps from 3 correct bits
            for (int it = 0; it < 6; it++) inv *= 2 - q * inv;
            md.qinv = inv;
            md.Rmodq = (u64)((((u128h)1) << 64) % q);
            md.Rmodq_s = h_shoup(md.Rmodq, q);
            md.ninvR = h_mulmod(md.ninv, md.Rmodq, q);
            md.ninvR_s = h_shoup(md.ninvR, q);
        }
        md.w = c->twiddles.p + ((size_t)m * 4 + 0) * N;
        md.ws = c->twiddles.p + ((size_t)m * 4 + 1) * N;
        md.iw = c->twiddles.p + ((size_t)m * 4 + 2) * N;
        md.iws = c->twiddles.p + ((size_t)m * 4 + 3) * N;
        u64* f2 = &tw2[((size_t)m * 2 + 0) * 2 * N];
        u64* i2 = &tw2[((size_t)m * 2 + 1) * 2 * N];
        for (uint32_t i = 0; i < N; i++) {
            f2[2 * i] = w[i];
            f2[2 * i + 1] = ws[i];
            i2[2 * i] = iw[i];
            i2[2 * i + 1] = iws[i];
        }
        md.ftw = reinterpret_cast<const ulonglong2*>(c->twiddles2.p + ((size_t)m * 2 + 0) * 2 * N);
        md.itw = reinterpret_cast<const ulonglong2*>(c->twiddles2.p + ((size_t)m * 2 + 1) * 2 * N);
    }
    CK(cudaMemcpy(c->twiddles2.p, tw2.data(), tw2.size() * sizeof(u64), cudaMemcpyHostToDevice));
    if (c->logN >= 10) {
        // Row-stage twiddles per row tile, in the read order of fused_mul.cu (row tile = 2^10 coefficients,
        // forward plan: radix-16 over row stages u = 0..3 then radix-8 over u = 4..6; inverse: radix-16 over
        // u = 6..3 then radix-8 over u = 2..0):
        //   forward: [0,120)   stage-major, shared by the 8 threads of a row: 8*(2^u - 1) + i
        //            [120,1016) per radix-8 block g (128 of them): 120 + 7 g + (2^r - 1) + j,  u = 4 + r
        //   inverse: [0,960)   per radix-16 block g (64 of them): 15 g + (2^r - 1) + j,        u = 3 + r
        //            [960,1016) stage-major, u = 0..2: 960 + 8*(2^u - 1) + i
        const uint32_t tiles = N >> 10, per_tile = 1016, nmq = L + Lp;
        std::vector<u64> tr((size_t)nmq * 2 * tiles * per_tile * 2);
        CK(c->twiddles_rows.alloc(tr.size()));
        for (uint32_t m = 0; m < nmq; m++) {
            const u64* f2 = &tw2[((size_t)m * 2 + 0) * 2 * N];
            const u64* i2 = &tw2[((size_t)m * 2 + 1) * 2 * N];
            for (uint32_t t = 0; t < tiles; t++) {
                u64* fo = &tr[(((size_t)m * 2 + 0) * tiles + t) * per_tile * 2];
                u64* io = &tr[(((size_t)m * 2 + 1) * tiles + t) * per_tile * 2];
                const uint32_t r0 = t * 8;
                auto gidx = [&](uint32_t u, uint32_t i) { return (1u << (c->logN - 7 + u)) + (r0 << u) + i; };
                auto put = [](u64* dst, uint32_t e, const u64* src, uint32_t g) {
                    dst[2 * e] = src[2 * g];
                    dst[2 * e + 1] = src[2 * g + 1];
                };
                for (uint32_t u = 0; u < 4; u++)
                    for (uint32_t i = 0; i < (8u << u); i++) put(fo, ((8u << u) - 8u) + i, f2, gidx(u, i));
                for (uint32_t g = 0; g < 128; g++)
                    for (uint32_t r = 0; r < 3; r++)
                        for (uint32_t j = 0; j < (1u << r); j++) put(fo, 120 + 7 * g + ((1u << r) - 1) + j, f2, gidx(4 + r, (g << r) + j));
                for (uint32_t g = 0; g < 64; g++)
                    for (uint32_t r = 0; r < 4; r++)
                        for (uint32_t j = 0; j < (1u << r); j++) put(io, 15 * g + ((1u << r) - 1) + j, i2, gidx(3 + r, (g << r) + j));
                for (uint32_t u = 0; u < 3; u++)
                    for (uint32_t i = 0; i < (8u << u); i++) put(io, 960 + ((8u << u) - 8u) + i, i2, gidx(u, i));
            }
            T.mods[m].ftw_rows = reinterpret_cast<const ulonglong2*>(c->twiddles_rows.p + (((size_t)m * 2 + 0) * tiles) * per_tile * 2);
            T.mods[m].itw_rows = reinterpret_cast<const ulonglong2*>(c->twiddles_rows.p + (((size_t)m * 2 + 1) * tiles) * per_tile * 2);
        }
        CK(cudaMemcpy(c->twiddles_rows.p, tr.data(), tr.size() * sizeof(u64), cudaMemcpyHostToDevice));
    }
    CK(cudaMemcpy(c->twiddles.p, tw.data(), tw.size() * sizeof(u64), cudaMemcpyHostToDevice));
    for (uint32_t i = 0; i < L; i++) {
        T.QHatInvModq[i] = P.QHatInvModq[i];
        T.QHatInvModq_s[i] = h_shoup(P.QHatInvModq[i], P.q[i]);
        T.qInv[i] = P.qInv[i];
        T.negPQHatInvModq[i] = P.negPQHatInvModq[i];
        T.negPQHatInvModq_s[i] = h_shoup(P.negPQHatInvModq[i], P.q[i]);
        for (uint32_t j = 0; j < Lp; j++) {
            T.qInvModp[i][j] = P.qInvModp[i][j];
            T.qInvModp_s[i][j] = h_shoup(P.qInvModp[i][j], P.p[j]);
            T.qInvModp_m[i][j] = h_mont_split(P.qInvModp[i][j], P.p[j]);
            T.PHatModq_m[i][j] = h_mont_split(P.PHatModq[i][j], P.q[i]);
            T.PHatModq[i][j] = P.PHatModq[i][j];
            T.PHatModq_s[i][j] = h_shoup(P.PHatModq[i][j], P.q[i]);
        }
        for (uint32_t j = 0; j <= Lp; j++) {
            T.tQS[i][j] = P.tQSHatInvModsDivsModq[i][j];
            T.tQS_s[i][j] = h_shoup(P.tQSHatInvModsDivsModq[i][j], P.q[i]);
            T.tQS_m[i][j] = h_mont_split(P.tQSHatInvModsDivsModq[i][j], P.q[i]);
        }
        for (uint32_t k = 0; k < L; k++) T.qModq[i][k] = P.q[i] % P.q[k];
        T.QHatInvNinv[i] = h_mulmod(P.QHatInvModq[i], T.mods[i].ninv, P.q[i]);
        T.QHatInvNinv_s[i] = h_shoup(T.QHatInvNinv[i], P.q[i]);
        T.negPQHatInvNinv[i] = h_mulmod(P.negPQHatInvModq[i], T.mods[i].ninv, P.q[i]);
        T.negPQHatInvNinv_s[i] = h_shoup(T.negPQHatInvNinv[i], P.q[i]);
    }
    for (uint32_t j = 0; j < Lp; j++) {
        T.PHatInvModp[j] = P.PHatInvModp[j];
        T.PHatInvModp_s[j] = h_shoup(P.PHatInvModp[j], P.p[j]);
        T.pInv[j] = P.pInv[j];
        T.tQSfrac[j] = P.tQSHatInvModsDivsFrac[j];
        for (uint32_t i = 0; i < L; i++) {
            T.QHatModp[j][i] = P.QHatModp[j][i];
            T.QHatModp_s[j][i] = h_shoup(P.QHatModp[j][i], P.p[j]);
            T.QHatModp_m[j][i] = h_mont_split(P.QHatModp[j][i], P.p[j]);
        }
    }
    for (uint32_t a = 0; a <= L; a++)
        for (uint32_t j = 0; j < Lp; j++) T.alphaQModp[a][j] = P.alphaQModp[a][j];
    for (uint32_t a = 0; a <= Lp; a++)
        for (uint32_t i = 0; i < L; i++) T.alphaPModq[a][i] = P.alphaPModq[a][i];
    // ---- variants
    T.fp_fma = P.fp_contract == PSI_FP_FMA;
    T.hps = c->hps;
    for (uint32_t j = 0; j < Lp; j++)
        for (uint32_t i = 0; i <= L; i++) T.tPS[j][i] = P.tPSHatInvModsDivsModp[j][i];
    for (uint32_t i = 0; i < L; i++) T.tPSfrac[i] = P.tPSHatInvModsDivsFrac[i];
    if (c->hybrid) {
        // KeySwitchHYBRID tables: all exact integers, derived from the moduli (OpenFHE: PartQlHatInvModq, PartQlHatModp,
        // PInvModq, PHatInvModp, PHatModq)
        const uint32_t Lk = c->Lk, parts = c->ks_parts, alpha = (L + parts - 1) / parts;
        T.ks_parts = parts;
        T.ks_alpha = alpha;
        T.Lk = Lk;
        for (uint32_t i = 0; i < L; i++) {
            const uint32_t lo = (i / alpha) * alpha, hi = lo + alpha < L ? lo + alpha : L;
            for (uint32_t m = 0; m < L + Lk; m++) {
                const u64 mod = m < L ? P.q[m] : P.pk[m - L];
                u64 r = 1 % mod;
                for (uint32_t u = lo; u < hi; u++)
                    if (u != i) r = h_mulmod(r, P.q[u] % mod, mod);
                T.PartQHatModt[i][m] = r;
            }
            T.PartQHatInvModq[i] = h_powmod(T.PartQHatModt[i][i], P.q[i] - 2, P.q[i]);
            T.PartQHatInvModq_s[i] = h_shoup(T.PartQHatInvModq[i], P.q[i]);
            u64 pk = 1;
            for (uint32_t u = 0; u < Lk; u++) pk = h_mulmod(pk, P.pk[u] % P.q[i], P.q[i]);
            T.PkInvModq[i] = h_powmod(pk, P.q[i] - 2, P.q[i]);
            T.PkInvModq_s[i] = h_shoup(T.PkInvModq[i], P.q[i]);
        }
        for (uint32_t u = 0; u < Lk; u++) {
            u64 hat = 1;
            for (uint32_t v = 0; v < Lk; v++)
                if (v != u) hat = h_mulmod(hat, P.pk[v] % P.pk[u], P.pk[u]);
            T.PkHatInvModpk[u] = h_powmod(hat, P.pk[u] - 2, P.pk[u]);
            T.PkHatInvModpk_s[u] = h_shoup(T.PkHatInvModpk[u], P.pk[u]);
            for (uint32_t i = 0; i < L; i++) {
                u64 r = 1;
                for (uint32_t v = 0; v < Lk; v++)
                    if (v != u) r = h_mulmod(r, P.pk[v] % P.q[i], P.q[i]);
                T.PkHatModq[u][i] = r;
            }
        }
    }
    CK(cudaMalloc(&c->d_tab, sizeof(DevTables)));
    CK(cudaMemcpy(c->d_tab, &T, sizeof(T), cudaMemcpyHostToDevice));

    return build_pack_perm(c, PSI_PACK_COFACTOR_3);
}

static int alloc_work(psi_ctx* c) {
    const size_t N = c->N, L = c->L, LT = c->L + c->Lp, b = c->b, K = c->K;
    CK(c->acc.alloc(K * b * 2 * L * N));
    CK(c->out.alloc(b * 2 * L * N));
    CK(c->out2.alloc(b * 2 * L * N));
    if (K > 1) {
        CK(c->coef.alloc(b * 4 * L * N));
        CK(c->e1.alloc(b * 2 * LT * N));
        CK(c->e2.alloc(b * 2 * LT * N));
        CK(c->ten.alloc(b * 3 * LT * N));
        CK(c->res.alloc(b * 3 * L * N));
        {
            const size_t dig_bv = L * L, dig_hy = (size_t)c->ks_parts * (L + c->Lk);
            CK(c->dig.alloc(b * (dig_bv > dig_hy ? dig_bv : dig_hy) * N));
        }
        if (K > 2) CK(c->prod.alloc(b * 2 * L * N));
    }
    return PSI_OK;
}

// One batched EvalMult(ct,ct) + relinearise over B ciphertext pairs (+ optional mask multiply).
// a, bb: [B][2][L][N] EVALUATION (a = multipliedResult, bb = innerProductResult; the operand order of
// BatchedFHEHIPPIE.cpp:123 matters); out: [B][2][L][N].
// bin0: first bin of the batch inside the context's work buffers — bin groups evaluated concurrently on several
// streams (psi_run_phases) work in disjoint slices of the scratch arrays.
// ops (fused kernels only): 3 = everything, 1 = the part that needs operand a alone, 2 = the rest (see launch_fused_mul).
static int mul_ctct_batch(psi_ctx* c, cudaStream_t s, uint32_t B, const u64* a, const u64* bb, const u64* mask,
                          u64* out, uint32_t* launches, uint32_t bin0 = 0, uint32_t ops = 3) {
    const KCtx k = c->k(s);
    const uint32_t L = c->L, Lp = c->Lp, LT = L + Lp;
    const size_t N = c->N;
    uint32_t nl = 0;
    // per-bin sizes of the scratch arrays (alloc_work)
    // row-inverse halves of the two operands: two regions of b bins each, so that a bin's slice does not depend on
    // how the bins are grouped
    u64* const w_coef = c->coef.p + (size_t)bin0 * 2 * L * N;
    u64* const w_coef2 = c->coef.p + ((size_t)c->b + bin0) * 2 * L * N;
    // the fused kernels keep only the P limbs of the extended first operand ([bin][2][Lp][N]); a bin's slice must not
    // depend on the grouping either (ops = 1 prepares all bins at once, ops = 2 runs per group)
    u64* const w_e1 = c->e1.p + (size_t)bin0 * 2 * (fused_mul_supported(c->k(s)) ? Lp : LT) * N;
    u64* const w_e2 = c->e2.p + (size_t)bin0 * 2 * LT * N;
    u64* const w_ten = c->ten.p + (size_t)bin0 * 3 * LT * N;
    u64* const w_res = c->res.p + (size_t)bin0 * 3 * L * N;
    const size_t dig_per_bin = c->hybrid && (size_t)c->ks_parts * (L + c->Lk) > (size_t)L * L ? (size_t)c->ks_parts * (L + c->Lk) : (size_t)L * L;
    u64* const w_dig = c->dig.p + (size_t)bin0 * dig_per_bin * N;
    if (fused_mul_supported(k)) {
        // the fused relinearisation takes the key and the masks in Montgomery form
        const u64* maskR = mask ? c->maskR.p + (mask - c->mask.p) : nullptr;
        CK(launch_fused_mul(k, B, a, bb, w_coef, w_coef2, w_e1, w_e2, w_ten, w_res, w_dig, c->evk_bR.p, c->evk_aR.p, maskR, out,
                            ops));
        if (launches) *launches += ops == 1 ? 2 : 5;
        return PSI_OK;
    }
    if (ops == 1) return PSI_OK;  // the unfused path has no split: everything happens in the ops = 2 call
    u64* coef1 = w_coef;   // [B*2][L][N]
    u64* coef2 = w_coef2;  // [B*2][L][N]
    // (1) both operands to COEFFICIENT
    NttBatch nb{a, coef1, B * 2 * L, L, L * N, N, L * N, 0, L};
    CK(launch_ntt(k, nb, true)); nl++;
    nb = NttBatch{bb, coef2, B * 2 * L, L, L * N, N, L * N, 0, L};
    CK(launch_ntt(k, nb, true)); nl++;
    // (2) first operand: exact Q -> P extension; Q limbs stay as given (EVALUATION)
    CK(launch_expand_q_to_p(k, B * 2, coef1, w_e1)); nl++;
    nb = NttBatch{w_e1 + (size_t)L * N, w_e1 + (size_t)L * N, B * 2 * Lp, Lp, LT * N, N, LT * N, L, Lp};
    CK(launch_ntt(k, nb, false)); nl++;
    CK(cudaMemcpy2DAsync(w_e1, LT * N * sizeof(u64), a, L * N * sizeof(u64), L * N * sizeof(u64), (size_t)B * 2,
                         cudaMemcpyDeviceToDevice, s));
    if (c->hps) {
        // (3, HPS) second operand: the same exact extension (DCRTPoly::ExpandCRTBasis on both ciphertexts)
        CK(launch_expand_q_to_p(k, B * 2, coef2, w_e2)); nl++;
        nb = NttBatch{w_e2 + (size_t)L * N, w_e2 + (size_t)L * N, B * 2 * Lp, Lp, LT * N, N, LT * N, L, Lp};
        CK(launch_ntt(k, nb, false)); nl++;
        CK(cudaMemcpy2DAsync(w_e2, LT * N * sizeof(u64), bb, L * N * sizeof(u64), L * N * sizeof(u64), (size_t)B * 2,
                             cudaMemcpyDeviceToDevice, s));
    } else {
        // (3, HPSPOVERQ) second operand: P-over-Q fast extension, all limbs back to EVALUATION
        CK(launch_fast_expand_poverq(k, B * 2, coef2, w_e2)); nl++;
        nb = NttBatch{w_e2, w_e2, B * 2 * LT, LT, LT * N, N, LT * N, 0, LT};
        CK(launch_ntt(k, nb, false)); nl++;
    }
    // (4) tensor, (5) COEFFICIENT, (6) scale and round into Q (HPSPOVERQ: by t/P; HPS: by t/Q into P, then P -> Q)
    CK(launch_tensor(k, B, w_e1, w_e2, w_ten)); nl++;
    nb = NttBatch{w_ten, w_ten, B * 3 * LT, LT, LT * N, N, LT * N, 0, LT};
    CK(launch_ntt(k, nb, true)); nl++;
    if (c->hps) {
        CK(launch_scale_round_hps(k, B * 3, w_ten, w_res)); nl++;
    } else {
        CK(launch_scale_round(k, B * 3, w_ten, w_res)); nl++;
    }
    // components 0 and 1 to EVALUATION: group = bin, 2L of its 3L limb-polys
    if (c->hybrid) {
        // (7, HYBRID) digits of c2 over the extended basis Q + pk, inner products with the key, ApproxModDown
        const uint32_t Lk = c->Lk, LE = L + Lk, parts = c->ks_parts, pk0 = L + Lp + 1;
        u64* const ext = w_ten;  // [B][2][LE][N]: the tensor buffer is free again (3 (L + Lp) >= 2 (L + Lk) limbs per bin)
        u64* const sw = w_e1;    // [B][2][L][N]
        CK(launch_hybrid_modup(k, B, w_res, w_dig)); nl++;
        nb = NttBatch{w_dig, w_dig, B * parts * L, L, LE * N, N, LE * N, 0, L};
        CK(launch_ntt(k, nb, false)); nl++;
        nb = NttBatch{w_dig + (size_t)L * N, w_dig + (size_t)L * N, B * parts * Lk, Lk, LE * N, N, LE * N, pk0, Lk};
        CK(launch_ntt(k, nb, false)); nl++;
        CK(launch_hybrid_inner(k, B, w_dig, c->evk_b.p, c->evk_a.p, ext)); nl++;
        nb = NttBatch{ext + (size_t)L * N, ext + (size_t)L * N, B * 2 * Lk, Lk, LE * N, N, LE * N, pk0, Lk};
        CK(launch_ntt(k, nb, true)); nl++;
        CK(launch_hybrid_moddown(k, B, ext, sw)); nl++;
        nb = NttBatch{sw, sw, B * 2 * L, L, L * N, N, L * N, 0, L};
        CK(launch_ntt(k, nb, false)); nl++;
        nb = NttBatch{w_res, w_res, B * 2 * L, 2 * L, 3 * L * N, N, 3 * L * N, 0, L};
        CK(launch_ntt(k, nb, false)); nl++;
        CK(launch_hybrid_finish(k, B, w_res, ext, sw, mask, out)); nl++;
    } else {
        // (7, BV) relinearise: digits of c2 (digit size 0), everything to EVALUATION, accumulate
        CK(launch_relin_digits(k, B, w_res, w_dig)); nl++;
        nb = NttBatch{w_dig, w_dig, B * L * L, L, L * N, N, L * N, 0, L};
        CK(launch_ntt(k, nb, false)); nl++;
        nb = NttBatch{w_res, w_res, B * 2 * L, 2 * L, 3 * L * N, N, 3 * L * N, 0, L};
        CK(launch_ntt(k, nb, false)); nl++;
        CK(launch_relin_accum(k, B, w_res, w_dig, c->evk_b.p, c->evk_a.p, mask, out)); nl++;
    }
    if (launches) *launches += nl;
    return PSI_OK;
}

}  // namespace psi

extern "C" {

const char* psi_last_error(void) { return g_last_error.c_str(); }
const char* psi_version(void) { return "psi_b200 0.3 (sm_100a)"; }

int psi_device_count(int* n) {
    if (!n) return set_error(PSI_ERR_INVALID, "null argument");
    *n = 0;
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
    return PSI_OK;
}

int psi_ctx_create(const psi_params* p, int device, psi_ctx** out) {
    if (!p || !out) return set_error(PSI_ERR_INVALID, "null argument");
    *out = nullptr;
    const uint32_t N = p->N;
    if (N < 8 || (N & (N - 1)) || N > 16384) return set_error(PSI_ERR_INVALID, "ring dimension must be a power of two in [8, 16384]");
    if (p->L < 1 || p->L > PSI_MAX_LIMBS || p->Lp < 1 || p->Lp > PSI_MAX_LIMBS)
        return set_error(PSI_ERR_INVALID, "sizeQ / sizeP out of range");
    if (p->mult_technique != PSI_MULT_HPSPOVERQ && p->mult_technique != PSI_MULT_HPS)
        return set_error(PSI_ERR_INVALID, "MultiplicationTechnique must be HPS or HPSPOVERQ (BEHZ and HPSPOVERQLEVELED are not implemented)");
    if (p->ks_technique != PSI_KS_BV && p->ks_technique != PSI_KS_HYBRID)
        return set_error(PSI_ERR_INVALID, "KeySwitchTechnique must be BV (digit size 0) or HYBRID");
    if (p->fp_contract > PSI_FP_FMA) return set_error(PSI_ERR_INVALID, "unknown floating-point contraction mode");
    if (p->ks_technique == PSI_KS_HYBRID) {
        if (p->Lk < 1 || p->Lk > PSI_MAX_LIMBS || p->ks_num_parts < 1 || p->ks_num_parts > p->L)
            return set_error(PSI_ERR_INVALID, "HYBRID key switching: bad number of digits / special primes");
        const uint32_t alpha = (p->L + p->ks_num_parts - 1) / p->ks_num_parts;
        if ((p->L + alpha - 1) / alpha != p->ks_num_parts) return set_error(PSI_ERR_INVALID, "HYBRID key switching: empty digit");
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
    if (ndev == 0) return set_error(PSI_ERR_NO_DEVICE, "no CUDA device; this library has no CPU path");
    if (device < 0 || device >= ndev) return set_error(PSI_ERR_INVALID, "device ordinal out of range");
    psi_ctx* c = new (std::nothrow) psi_ctx();
    if (!c) return set_error(PSI_ERR_INVALID, "out of host memory");
    c->device = device;
    c->P = *p;
    c->N = N;
    c->L = p->L;
    c->Lp = p->Lp;
    c->hps = p->mult_technique == PSI_MULT_HPS;
    c->hybrid = p->ks_technique == PSI_KS_HYBRID;
    c->Lk = c->hybrid ? p->Lk : 0;
    c->ks_parts = c->hybrid ? p->ks_num_parts : 0;
    while ((1u << c->logN) < N) c->logN++;
    int rc = ensure_device(c);
    if (rc == PSI_OK) rc = build_tables(c);
    if (rc == PSI_OK) {
        cudaError_t e2 = ntt_init_device();
        if (e2 == cudaSuccess) e2 = mac_init_device();
        if (e2 == cudaSuccess) e2 = fused_mul_init_device(c->k(0));
        if (e2 != cudaSuccess) rc = cuda_fail(e2, "kernel attribute set-up");
    }
    if (rc != PSI_OK) {
        psi_ctx_destroy(c);
        return rc;
    }
    c->use_graph = std::getenv("PSI_NO_GRAPH") == nullptr;  // tuning / debugging: direct launches
    *out = c;
    return PSI_OK;
}

static void drop_run_graphs(psi_ctx* c);

int psi_ctx_destroy(psi_ctx* c) {
    if (!c) return PSI_OK;
    cudaSetDevice(c->device);
    drop_run_graphs(c);
    nb_release(c);
    if (c->d_tab) cudaFree(c->d_tab);
    DevBuf<u64>* bufs[] = {&c->twiddles, &c->twiddles2, &c->twiddles_rows, &c->evk_bR, &c->evk_aR, &c->maskR, &c->evk_b, &c->evk_a, &c->pt, &c->mask, &c->idx, &c->idx_in, &c->stage, &c->minus, &c->acc,
                           &c->coef,     &c->e1,    &c->e2,    &c->ten, &c->res, &c->dig, &c->prod,  &c->out, &c->out2, &c->minus_in};
    for (auto* b : bufs) b->release();
    c->to_crt.release();
    for (auto& st : c->aux)
        if (st) cudaStreamDestroy(st);
    if (c->sq_in) cudaStreamDestroy(c->sq_in);
    if (c->sq_out) cudaStreamDestroy(c->sq_out);
    for (auto& st : c->sq_grp)
        if (st) cudaStreamDestroy(st);
    if (c->ev_sq_fork) cudaEventDestroy(c->ev_sq_fork);
    for (auto& e : c->ev_dl)
        if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_slice)
        if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_group)
        if (e) cudaEventDestroy(e);
    if (c->ev_sq) cudaEventDestroy(c->ev_sq);
    for (auto& e : c->ev_join)
        if (e) cudaEventDestroy(e);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->pool_in) cudaFreeHost(c->pool_in);
    if (c->pool_out) cudaFreeHost(c->pool_out);
    if (c->ev_pool_in) cudaEventDestroy(c->ev_pool_in);
    delete c;
    return PSI_OK;
}

int psi_set_packing_cofactor(psi_ctx* c, uint32_t mode) {
    if (!c) return set_error(PSI_ERR_INVALID, "null argument");
    if (mode != PSI_PACK_COFACTOR_3 && mode != PSI_PACK_COFACTOR_CONJ) return set_error(PSI_ERR_INVALID, "unknown packing co-factor");
    int rc = ensure_device(c);
    if (rc) return rc;
    return build_pack_perm(c, mode);
}

int psi_set_encode_lift(psi_ctx* c, uint32_t mode) {
    if (!c) return set_error(PSI_ERR_INVALID, "null argument");
    if (mode != PSI_ENCODE_LIFT_PLAIN && mode != PSI_ENCODE_LIFT_CENTRED) return set_error(PSI_ERR_INVALID, "unknown encode lift");
    c->encode_lift = mode;
    return PSI_OK;
}

int psi_set_relin_key(psi_ctx* c, const uint64_t* evk_b, const uint64_t* evk_a) {
    if (!c || !evk_b || !evk_a) return set_error(PSI_ERR_INVALID, "null argument");
    int rc = ensure_device(c);
    if (rc) return rc;
    // BV: [L][L][N]; HYBRID: [parts][L + Lk][N]
    const size_t n = c->hybrid ? (size_t)c->ks_parts * (c->L + c->Lk) * c->N : (size_t)c->L * c->L * c->N;
    CK(c->evk_b.alloc(n));
    CK(c->evk_a.alloc(n));
    CK(cudaMemcpy(c->evk_b.p, evk_b, n * sizeof(u64), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->evk_a.p, evk_a, n * sizeof(u64), cudaMemcpyHostToDevice));
    if (!c->hybrid) {
        CK(c->evk_bR.alloc(n));
        CK(c->evk_aR.alloc(n));
        const KCtx k = c->k(0);
        CK(launch_to_montgomery(k, c->L, c->evk_b.p, c->evk_bR.p));
        CK(launch_to_montgomery(k, c->L, c->evk_a.p, c->evk_aR.p));
        CK(cudaStreamSynchronize(0));
    }
    c->have_evk = true;
    return PSI_OK;
}

static int db_dims(psi_ctx* c, uint32_t K, uint32_t b, uint32_t E) {
    if (K < 1 || b < 1 || E < 1) return set_error(PSI_ERR_INVALID, "K, b, E must be positive");
    c->K = K;
    c->b = b;
    c->E = E;
    c->have_db = false;
    c->have_query = false;
    c->ran = false;
    const size_t LN = (size_t)c->L * c->N;
    CK(c->pt.alloc((size_t)K * b * E * LN));
    CK(c->mask.alloc((size_t)b * LN));
    CK(c->maskR.alloc((size_t)b * LN));
    if (LN % 128) return set_error(PSI_ERR_INVALID, "L * N must be a multiple of 128");
    CK(c->idx.alloc((size_t)K * E * 2 * LN));
    CK(c->idx_in.alloc((size_t)2 * K * E * 2 * LN));
    CK(c->minus.alloc(2 * LN));
    CK(c->minus_in.alloc(2 * 2 * LN));
    c->n_uploaded = c->n_committed = 0;
    return alloc_work(c);
}

// One database shard: bins [bin_begin, bin_end) of a b_total-bin database become the resident database of this
// context (b_local = bin_end - bin_begin).  The full-array layouts are those of the unsharded calls; for every
// hash function the shard's plaintexts / slot vectors are one contiguous block of them.
static int check_shard(uint32_t b_total, uint32_t bin_begin, uint32_t bin_end) {
    if (bin_begin >= bin_end || bin_end > b_total) return set_error(PSI_ERR_INVALID, "bin range must be a non-empty part of [0, b)");
    return PSI_OK;
}

int psi_db_load_limbs_shard(psi_ctx* c, uint32_t K, uint32_t b_total, uint32_t bin_begin, uint32_t bin_end, uint32_t E,
                            const uint64_t* pt_limbs, const uint64_t* mask_limbs) {
    if (!c || !pt_limbs || !mask_limbs) return set_error(PSI_ERR_INVALID, "null argument");
    int rc = check_shard(b_total, bin_begin, bin_end);
    if (rc) return rc;
    if ((rc = ensure_device(c))) return rc;
    const uint32_t b = bin_end - bin_begin;
    if ((rc = db_dims(c, K, b, E))) return rc;
    const size_t LN = (size_t)c->L * c->N;
    {
        const size_t n_hf = (size_t)b * E, chunk = n_hf < kDbChunk ? n_hf : kDbChunk;
        CK(c->stage.alloc(chunk * LN));
        DevBuf<int> d_bad;  // raised by the re-tiling kernel when a limb is not a canonical residue
        CK(d_bad.alloc(1));
        CK(cudaMemset(d_bad.p, 0, sizeof(int)));
        for (uint32_t hf = 0; hf < K; hf++) {
            const uint64_t* src = pt_limbs + (((size_t)hf * b_total + bin_begin) * E) * LN;
            for (size_t p0 = 0; p0 < n_hf; p0 += chunk) {
                const size_t n = (n_hf - p0) < chunk ? (n_hf - p0) : chunk;
                CK(cudaMemcpy(c->stage.p, src + p0 * LN, n * LN * sizeof(u64), cudaMemcpyHostToDevice));
                CK(launch_retile_pt(0, c->stage.p, c->pt.p, LN, E, (size_t)hf * n_hf + p0, n, true, c->d_tab, d_bad.p));
                CK(cudaStreamSynchronize(0));
            }
        }
        int bad = 0;
        CK(cudaMemcpy(&bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (bad) return set_error(PSI_ERR_INVALID, "plaintext limbs must be canonical residues (value >= q_l found)");
    }
    CK(cudaMemcpy(c->mask.p, mask_limbs + (size_t)bin_begin * LN, (size_t)b * LN * sizeof(u64), cudaMemcpyHostToDevice));
    CK(launch_to_montgomery(c->k(0), b, c->mask.p, c->maskR.p));
    CK(cudaStreamSynchronize(0));
    c->have_db = true;
    return PSI_OK;
}

int psi_db_load_limbs(psi_ctx* c, uint32_t K, uint32_t b, uint32_t E, const uint64_t* pt_limbs,
                      const uint64_t* mask_limbs) {
    return psi_db_load_limbs_shard(c, K, b, 0, b, E, pt_limbs, mask_limbs);
}

// Packed coefficients mod t ([n][N], in [0, t)) -> EVALUATION limbs [n][L][N] at ntt_dst.
//   PLAIN:   every limb is the same vector (residues in [0, t) are below every q_l); OpenFHE >= 1.0 as recalled:
//            PackedEncoding::Encode copies the coefficients into the first tower and SwitchModulus (centred about
//            q_0, so values < t stay) spreads them.
//   CENTRED: c > t/2 becomes q_l - (t - c), the signed representative lifted to q_l (SURVEY.md Appendix A's
//            reading).  Both decrypt identically; only the limbs differ, hence the switch (risk register).
static cudaError_t lift_and_ntt(psi_ctx* c, const KCtx& k, uint32_t n, const u64* d_crt, u64* ntt_dst) {
    const size_t N = c->N, L = c->L;
    if (c->encode_lift == PSI_ENCODE_LIFT_CENTRED) {
        cudaError_t e = launch_centre_lift(k, n, d_crt, ntt_dst);
        if (e != cudaSuccess) return e;
        NttBatch nb{ntt_dst, ntt_dst, n * (uint32_t)L, (uint32_t)L, L * N, N, L * N, 0, (uint32_t)L};
        return launch_ntt(k, nb, false);
    }
    NttBatch nb{d_crt, ntt_dst, n * (uint32_t)L, (uint32_t)L, N, 0, L * N, 0, (uint32_t)L};
    return launch_ntt(k, nb, false);
}

// MakePackedPlaintext + SetFormat(EVALUATION) for n_pt plaintexts, chunked so that the staging
// buffers stay small next to the DB itself.
// tiled_E != 0: dst is the tiled plaintext DB (positions per bin = tiled_E) and the plaintexts are numbers
// p_base .. p_base + n_pt - 1 of it; else dst is flat [n_pt][L][N].
static int encode_into_impl(psi_ctx* c, size_t n_pt, uint32_t nslots, const int64_t* slots, u64* dst, uint32_t tiled_E,
                            size_t p_base = 0) {
    const size_t N = c->N, L = c->L;
    const size_t chunk = n_pt < kDbChunk ? n_pt : kDbChunk;
    if (tiled_E) CK(c->stage.alloc(chunk * L * N));
    DevBuf<long long> d_slots;
    DevBuf<u64> d_crt;
    CK(d_slots.alloc(chunk * nslots));
    CK(d_crt.alloc(chunk * N));
    const KCtx k = c->k(0);
    for (size_t p0 = 0; p0 < n_pt; p0 += chunk) {
        const uint32_t n = (uint32_t)((n_pt - p0) < chunk ? (n_pt - p0) : chunk);
        CK(cudaMemcpy(d_slots.p, slots + p0 * nslots, (size_t)n * nslots * sizeof(int64_t), cudaMemcpyHostToDevice));
        CK(launch_slots_to_crt(k, n, nslots, d_slots.p, c->to_crt.p, d_crt.p));
        {
            NttBatch nb{d_crt.p, d_crt.p, n, 1, N, 0, N, c->L + c->Lp, 1};
            CK(launch_ntt(k, nb, true));  // slots -> coefficients mod t
        }
        // lift of the coefficients from [0, t) to every q_l (see psi_set_encode_lift), then NTT mod q_l
        CK(lift_and_ntt(c, k, n, d_crt.p, tiled_E ? c->stage.p : dst + p0 * L * N));
        if (tiled_E) CK(launch_retile_pt(0, c->stage.p, dst, L * N, tiled_E, p_base + p0, n, true));
        CK(cudaStreamSynchronize(0));
    }
    return PSI_OK;
}

int psi_db_encode_slots_shard(psi_ctx* c, uint32_t K, uint32_t b_total, uint32_t bin_begin, uint32_t bin_end, uint32_t E,
                              uint32_t nslots, const int64_t* slots, const int64_t* mask_slots) {
    if (!c || !slots || !mask_slots) return set_error(PSI_ERR_INVALID, "null argument");
    if (nslots < 1 || nslots > c->N) return set_error(PSI_ERR_INVALID, "batch size must be in [1, N]");
    int rc = check_shard(b_total, bin_begin, bin_end);
    if (rc) return rc;
    const uint32_t b = bin_end - bin_begin;
    // PackedEncoding::Encode rejects |v| >= t
    const uint64_t t = c->P.t;
    const size_t n_hf = (size_t)b * E * nslots, n_mask = (size_t)b * nslots;
    for (uint32_t hf = 0; hf < K; hf++) {
        const int64_t* sv = slots + (((size_t)hf * b_total + bin_begin) * E) * nslots;
        for (size_t i = 0; i < n_hf; i++) {
            const int64_t v = sv[i];
            if ((uint64_t)(v < 0 ? -v : v) >= t) return set_error(PSI_ERR_INVALID, "slot value out of range of the plaintext modulus");
        }
    }
    const int64_t* mv = mask_slots + (size_t)bin_begin * nslots;
    for (size_t i = 0; i < n_mask; i++) {
        const int64_t v = mv[i];
        if ((uint64_t)(v < 0 ? -v : v) >= t) return set_error(PSI_ERR_INVALID, "mask value out of range of the plaintext modulus");
    }
    if ((rc = ensure_device(c))) return rc;
    if ((rc = db_dims(c, K, b, E))) return rc;
    for (uint32_t hf = 0; hf < K; hf++)
        if ((rc = encode_into_impl(c, (size_t)b * E, nslots, slots + (((size_t)hf * b_total + bin_begin) * E) * nslots, c->pt.p, E,
                              (size_t)hf * b * E)))
            return rc;
    if ((rc = encode_into_impl(c, b, nslots, mv, c->mask.p, 0))) return rc;
    CK(launch_to_montgomery(c->k(0), b, c->mask.p, c->maskR.p));
    CK(cudaStreamSynchronize(0));
    c->have_db = true;
    return PSI_OK;
}

int psi_db_encode_slots(psi_ctx* c, uint32_t K, uint32_t b, uint32_t E, uint32_t nslots, const int64_t* slots,
                        const int64_t* mask_slots) {
    return psi_db_encode_slots_shard(c, K, b, 0, b, E, nslots, slots, mask_slots);
}

// Device-resident constructor path: hash the server set, build the nested cuckoo tables, apply the bin
// shuffle, transpose and encode without the table ever visiting the host.
static int hct_on_device(psi_ctx* c, uint64_t hash_seed, uint32_t k, uint32_t e, uint32_t K, uint32_t E, uint32_t b,
                         uint64_t eviction_seed, const uint64_t* items, size_t n, DevBuf<u64>& d_cells) {
    if (k < 1 || e < 1 || K < 2 || E < 1 || b < 1) return set_error(PSI_ERR_INVALID, "table sizes must be positive, K >= 2");
    if (n > 0x7fffffffull) return set_error(PSI_ERR_INVALID, "server set too large for one device build");
    TabulationHashing hashf(hash_seed, k + K);
    DevBuf<u64> d_T, d_items;  // freed on every exit path
    const std::vector<uint64_t>& T = hashf.tables();
    CK(d_T.alloc(T.size()));
    CK(d_items.alloc(n ? n : 1));
    CK(d_cells.alloc((size_t)k * e * K * b * E));
    CK(cudaMemcpy(d_T.p, T.data(), T.size() * sizeof(u64), cudaMemcpyHostToDevice));
    if (n) CK(cudaMemcpy(d_items.p, items, n * sizeof(u64), cudaMemcpyHostToDevice));
    int failed = 0;
    cudaError_t err = hct_build_device(0, d_T.p, k, e, K, b, E, eviction_seed, d_items.p, n, d_cells.p, &failed);
    if (err != cudaSuccess) return cuda_fail(err, "device table build");
    if (failed) return set_error(PSI_ERR_STATE, "(Blocked) Cuckoo hashing error");  // CuckooHashTable.cpp:113
    return PSI_OK;
}

int psi_hct_build_device(psi_ctx* c, uint64_t hash_seed, uint32_t k, uint64_t e, uint32_t K, uint64_t E, uint64_t b,
                         uint64_t eviction_seed, const uint64_t* items, size_t n, uint64_t* cells) {
    if (!c || (!items && n) || !cells) return set_error(PSI_ERR_INVALID, "null argument");
    int rc = ensure_device(c);
    if (rc) return rc;
    DevBuf<u64> d_cells;
    if ((rc = hct_on_device(c, hash_seed, k, (uint32_t)e, K, (uint32_t)E, (uint32_t)b, eviction_seed, items, n, d_cells))) return rc;
    CK(cudaMemcpy(cells, d_cells.p, (size_t)k * e * K * b * E * sizeof(u64), cudaMemcpyDeviceToHost));
    return PSI_OK;
}

int psi_db_build_from_items_shard(psi_ctx* c, uint64_t hash_seed, uint32_t k, uint64_t e, uint32_t K, uint64_t E, uint64_t b,
                                  uint64_t eviction_seed, const uint64_t* items, size_t n, uint64_t shuffle_seed,
                                  uint64_t mask_seed, uint32_t bin_begin, uint32_t bin_end) {
    if (!c || (!items && n)) return set_error(PSI_ERR_INVALID, "null argument");
    const size_t nslots = (size_t)k * e;
    if (nslots > c->N) return set_error(PSI_ERR_INVALID, "batch size exceeds the ring dimension");
    if (b > 65535) return set_error(PSI_ERR_INVALID, "bin size too large");
    int rc = check_shard((uint32_t)b, bin_begin, bin_end);
    if (rc) return rc;
    if (shuffle_seed == PSI_SEED_RANDOM && !(bin_begin == 0 && bin_end == b))
        return set_error(PSI_ERR_INVALID, "a sharded build needs an explicit shuffle seed shared by all shards (draw it once)");
    const uint64_t t = c->P.t;
    for (size_t i = 0; i < n; i++)
        if (items[i] >= t) return set_error(PSI_ERR_INVALID, "slot value out of range of the plaintext modulus");
    if ((rc = ensure_device(c))) return rc;
    DevBuf<u64> d_cells;
    if ((rc = hct_on_device(c, hash_seed, k, (uint32_t)e, K, (uint32_t)E, (uint32_t)b, eviction_seed, items, n, d_cells))) return rc;
    // bin shuffle (BatchedFHEHIPPIE.cpp:25-35) as permutations, masks (:73-82): same generators as the host ctor,
    // drawn for ALL b bins so that every shard of one database sees the same shuffle and the same masks
    std::mt19937 mt = seeded_mt19937(resolve_seed(shuffle_seed));
    const std::vector<uint16_t> perm = makeBinShuffle(k, e, K, b, mt);
    std::vector<int64_t> mask_slots((size_t)b * nslots);
    std::mt19937_64 mm(resolve_seed(mask_seed));
    for (auto& v : mask_slots) v = (int64_t)(mm() % (t - 1) + 1);
    const uint32_t bl = bin_end - bin_begin;
    DevBuf<uint16_t> d_perm;
    DevBuf<u64> d_crt;
    CK(d_perm.alloc(perm.size()));
    CK(cudaMemcpy(d_perm.p, perm.data(), perm.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    if ((rc = db_dims(c, K, bl, (uint32_t)E))) return rc;
    {
        const size_t N = c->N, L = c->L, n_pt = (size_t)K * bl * E, chunk = n_pt < kDbChunk ? n_pt : kDbChunk;
        const KCtx kc = c->k(0);
        CK(d_crt.alloc(chunk * N));
        CK(c->stage.alloc(chunk * L * N));
        for (size_t p0 = 0; p0 < n_pt; p0 += chunk) {
            const uint32_t np = (uint32_t)((n_pt - p0) < chunk ? (n_pt - p0) : chunk);
            CK(launch_cells_to_crt(kc, np, (uint32_t)p0, (uint32_t)nslots, K, (uint32_t)b, (uint32_t)E, bin_begin, bl, d_cells.p,
                                   d_perm.p, c->to_crt.p, d_crt.p));
            {
                NttBatch nb{d_crt.p, d_crt.p, np, 1, N, 0, N, c->L + c->Lp, 1};
                CK(launch_ntt(kc, nb, true));
            }
            CK(lift_and_ntt(c, kc, np, d_crt.p, c->stage.p));
            CK(launch_retile_pt(0, c->stage.p, c->pt.p, L * N, (uint32_t)E, p0, np, true));
            CK(cudaStreamSynchronize(0));
        }
    }
    if ((rc = encode_into_impl(c, bl, (uint32_t)nslots, mask_slots.data() + (size_t)bin_begin * nslots, c->mask.p, 0))) return rc;
    CK(launch_to_montgomery(c->k(0), bl, c->mask.p, c->maskR.p));
    CK(cudaStreamSynchronize(0));
    c->have_db = true;
    return PSI_OK;
}

int psi_db_build_from_items(psi_ctx* c, uint64_t hash_seed, uint32_t k, uint64_t e, uint32_t K, uint64_t E, uint64_t b,
                            uint64_t eviction_seed, const uint64_t* items, size_t n, uint64_t shuffle_seed,
                            uint64_t mask_seed) {
    return psi_db_build_from_items_shard(c, hash_seed, k, e, K, E, b, eviction_seed, items, n, shuffle_seed, mask_seed, 0,
                                         (uint32_t)b);
}

// One bin of the resident database back to the host (tests, in-run parity checks of bench.py):
// pt [K][E][L][N], mask [L][N], canonical residues.
int psi_db_get_bin_limbs(psi_ctx* c, uint32_t bin, uint64_t* pt_limbs, uint64_t* mask_limbs) {
    if (!c || !pt_limbs || !mask_limbs) return set_error(PSI_ERR_INVALID, "null argument");
    if (!c->have_db) return set_error(PSI_ERR_STATE, "no database loaded");
    if (bin >= c->b) return set_error(PSI_ERR_INVALID, "bin out of range");
    int rc = ensure_device(c);
    if (rc) return rc;
    const size_t LN = (size_t)c->L * c->N;
    DevBuf<u64> tmp;
    CK(tmp.alloc((size_t)c->E * LN));
    for (uint32_t hf = 0; hf < c->K; hf++) {
        CK(cudaMemset(tmp.p, 0, (size_t)c->E * LN * sizeof(u64)));
        // retile works on (plaintext index - p0) inside the flat buffer
        CK(launch_retile_pt(0, tmp.p, c->pt.p, LN, c->E, ((size_t)hf * c->b + bin) * c->E, c->E, false));
        CK(cudaMemcpy(pt_limbs + (size_t)hf * c->E * LN, tmp.p, (size_t)c->E * LN * sizeof(u64), cudaMemcpyDeviceToHost));
    }
    CK(cudaMemcpy(mask_limbs, c->mask.p + (size_t)bin * LN, LN * sizeof(u64), cudaMemcpyDeviceToHost));
    return PSI_OK;
}

int psi_db_get_limbs(psi_ctx* c, uint64_t* pt_limbs, uint64_t* mask_limbs) {
    if (!c) return set_error(PSI_ERR_INVALID, "null argument");
    if (!c->have_db) return set_error(PSI_ERR_STATE, "no database loaded");
    int rc = ensure_device(c);
    if (rc) return rc;
    const size_t LN = (size_t)c->L * c->N;
    if (pt_limbs) {
        const size_t n_pt = (size_t)c->K * c->b * c->E, chunk = n_pt < kDbChunk ? n_pt : kDbChunk;
        CK(c->stage.alloc(chunk * LN));
        for (size_t p0 = 0; p0 < n_pt; p0 += chunk) {  // tiled split-30 storage -> flat canonical residues
            const size_t n = (n_pt - p0) < chunk ? (n_pt - p0) : chunk;
            CK(launch_retile_pt(0, c->stage.p, c->pt.p, LN, c->E, p0, n, false));
            CK(cudaMemcpy(pt_limbs + p0 * LN, c->stage.p, n * LN * sizeof(u64), cudaMemcpyDeviceToHost));
        }
    }
    if (mask_limbs) CK(cudaMemcpy(mask_limbs, c->mask.p, (size_t)c->b * LN * sizeof(u64), cudaMemcpyDeviceToHost));
    return PSI_OK;
}

int psi_query_set(psi_ctx* c, const uint64_t* idx, const uint64_t* minus, void* stream) {
    int rc = psi_query_upload(c, idx, minus, stream);
    if (rc) return rc;
    return psi_query_commit(c, stream);
}

static int landing_slot(psi_ctx* c, uint32_t* which) {
    if (!c->have_db) return set_error(PSI_ERR_STATE, "load the database before the query (it fixes K and E)");
    if (c->n_uploaded - c->n_committed >= 2)
        return set_error(PSI_ERR_STATE, "two uploaded queries are already waiting for psi_query_commit");
    *which = c->n_uploaded & 1u;
    return PSI_OK;
}

int psi_query_upload(psi_ctx* c, const uint64_t* idx, const uint64_t* minus, void* stream) {
    if (!c || !idx || !minus) return set_error(PSI_ERR_INVALID, "null argument");
    uint32_t w = 0;
    int rc = landing_slot(c, &w);
    if (rc) return rc;
    if ((rc = ensure_device(c))) return rc;
    const size_t LN = (size_t)c->L * c->N;
    cudaStream_t s = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(c->idx_in.p + w * c->idx_words(), idx, c->idx_words() * sizeof(u64), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->minus_in.p + w * 2 * LN, minus, 2 * LN * sizeof(u64), cudaMemcpyHostToDevice, s));
    c->n_uploaded++;
    return PSI_OK;
}

static int ensure_pool(u64** p, size_t* have, size_t words) {
    if (*p && *have >= words) return PSI_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *have = 0;
    CK(cudaHostAlloc((void**)p, words * sizeof(u64), cudaHostAllocPortable));
    *have = words;
    return PSI_OK;
}

constexpr uint32_t kLimbPieces = 8;  // staging granularity: host copy of piece i+1 overlaps the DMA of piece i

// setIndex + setMinusCompareElement from SEPARATE limb vectors (what a deserialised OpenFHE query holds,
// BatchedFHEPSIServer.cpp:114-141): K*E*2*L pointers in [hf][pos][comp][limb] order + 2*L for minus, N words each,
// pageable memory.  A few host threads copy the vectors of a piece into the pinned pool, the copy engine uploads
// the piece while the next one is being staged.  The vectors may be freed when the call returns.
int psi_query_upload_limbs(psi_ctx* c, const uint64_t* const* idx_limbs, const uint64_t* const* minus_limbs, void* stream) {
    if (!c || !idx_limbs || !minus_limbs) return set_error(PSI_ERR_INVALID, "null argument");
    uint32_t w = 0;
    int rc = landing_slot(c, &w);
    if (rc) return rc;
    if ((rc = ensure_device(c))) return rc;
    const size_t N = c->N, LN = (size_t)c->L * N, W = c->idx_words(), nvec = W / N;
    if ((rc = ensure_pool(&c->pool_in, &c->pool_in_words, W + 2 * LN))) return rc;
    if (!c->ev_pool_in) CK(cudaEventCreateWithFlags(&c->ev_pool_in, cudaEventDisableTiming));
    CK(cudaEventSynchronize(c->ev_pool_in));  // the previous upload has read the pool
    cudaStream_t s = (cudaStream_t)stream;
    u64* pool = c->pool_in;
    for (size_t v = 0; v < 2 * (size_t)c->L; v++) std::memcpy(pool + W + v * N, minus_limbs[v], N * sizeof(u64));
    CK(cudaMemcpyAsync(c->minus_in.p + w * 2 * LN, pool + W, 2 * LN * sizeof(u64), cudaMemcpyHostToDevice, s));
    const int nt = c->host_threads;
    (void)nt;
    for (uint32_t piece = 0; piece < kLimbPieces; piece++) {
        const long v0 = (long)(nvec * piece / kLimbPieces), v1 = (long)(nvec * (piece + 1) / kLimbPieces);
        if (v1 == v0) continue;
#pragma omp parallel for schedule(static) num_threads(nt) if (v1 - v0 >= 16)
        for (long v = v0; v < v1; v++) copy_limb_vector(pool + (size_t)v * N, idx_limbs[v], N);
        CK(cudaMemcpyAsync(c->idx_in.p + w * W + (size_t)v0 * N, pool + (size_t)v0 * N, (size_t)(v1 - v0) * N * sizeof(u64),
                           cudaMemcpyHostToDevice, s));
    }
    CK(cudaEventRecord(c->ev_pool_in, s));
    c->n_uploaded++;
    return PSI_OK;
}

// getResultList into b*2*L separate limb vectors ([bin][comp][limb] order): the download is split into pieces,
// each scattered by host threads as soon as it has arrived.  Synchronous: the vectors are filled on return.
int psi_result_get_limbs(psi_ctx* c, uint64_t* const* out_limbs, void* stream) {
    if (!c || !out_limbs) return set_error(PSI_ERR_INVALID, "null argument");
    if (!c->ran) return set_error(PSI_ERR_STATE, "getResultList() before run()");
    int rc = ensure_device(c);
    if (rc) return rc;
    const size_t N = c->N, W = (size_t)c->b * 2 * c->L * N, nvec = W / N;
    if ((rc = ensure_pool(&c->pool_out, &c->pool_out_words, W))) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    cudaEvent_t ev[kLimbPieces] = {};
    struct EvGuard {
        cudaEvent_t* e;
        ~EvGuard() {
            for (uint32_t i = 0; i < kLimbPieces; i++)
                if (e[i]) cudaEventDestroy(e[i]);
        }
    } guard{ev};
    const u64* src = c->out_buf(c->out_cur);
    for (uint32_t piece = 0; piece < kLimbPieces; piece++) {
        const size_t v0 = nvec * piece / kLimbPieces, v1 = nvec * (piece + 1) / kLimbPieces;
        CK(cudaEventCreateWithFlags(&ev[piece], cudaEventDisableTiming));
        if (v1 > v0) CK(cudaMemcpyAsync(c->pool_out + v0 * N, src + v0 * N, (v1 - v0) * N * sizeof(u64), cudaMemcpyDeviceToHost, s));
        CK(cudaEventRecord(ev[piece], s));
    }
    const int nt = c->host_threads;
    (void)nt;
    for (uint32_t piece = 0; piece < kLimbPieces; piece++) {
        const long v0 = (long)(nvec * piece / kLimbPieces), v1 = (long)(nvec * (piece + 1) / kLimbPieces);
        CK(cudaEventSynchronize(ev[piece]));
#pragma omp parallel for schedule(static) num_threads(nt) if (v1 - v0 >= 16)
        for (long v = v0; v < v1; v++) copy_limb_vector(out_limbs[v], c->pool_out + (size_t)v * N, N);
    }
    return PSI_OK;
}

int psi_debug_set_tuning(psi_ctx* c, int mac_variant, int phase2_groups) {
    if (!c || mac_variant < -1 || mac_variant > 6 || phase2_groups < -1 || phase2_groups > 8) return set_error(PSI_ERR_INVALID, "bad tuning value");
    if (mac_variant >= 0) mac_force_variant(mac_variant);
    if (phase2_groups >= 0) c->p2_groups = (uint32_t)phase2_groups;
    return PSI_OK;
}

int psi_set_host_threads(psi_ctx* c, int n) {
    if (!c || n < 1 || n > 256) return set_error(PSI_ERR_INVALID, "host thread count must be in [1, 256]");
    c->host_threads = n;
    return PSI_OK;
}

int psi_query_landing_ptr(psi_ctx* c, uint32_t which, void** idx, size_t* idx_bytes, void** minus, size_t* minus_bytes) {
    if (!c || !idx || !idx_bytes || !minus || !minus_bytes || which > 1) return set_error(PSI_ERR_INVALID, "bad argument");
    if (!c->have_db) return set_error(PSI_ERR_STATE, "load the database before the query (it fixes K and E)");
    int rc = ensure_device(c);
    if (rc) return rc;
    const size_t LN = (size_t)c->L * c->N;
    *idx = c->idx_in.p + which * c->idx_words();
    *idx_bytes = c->idx_words() * sizeof(u64);
    *minus = c->minus_in.p + which * 2 * LN;
    *minus_bytes = 2 * LN * sizeof(u64);
    return PSI_OK;
}

int psi_query_next_landing(psi_ctx* c, uint32_t* which) {
    if (!c || !which) return set_error(PSI_ERR_INVALID, "null argument");
    return landing_slot(c, which);
}

int psi_query_uploaded(psi_ctx* c, uint32_t which) {
    if (!c) return set_error(PSI_ERR_INVALID, "null argument");
    uint32_t w = 0;
    int rc = landing_slot(c, &w);
    if (rc) return rc;
    if (which != w) return set_error(PSI_ERR_STATE, "landing buffers are filled in turn: see psi_query_next_landing");
    c->n_uploaded++;
    return PSI_OK;
}

int psi_query_commit(psi_ctx* c, void* stream) {
    if (!c) return set_error(PSI_ERR_INVALID, "null argument");
    if (c->n_committed == c->n_uploaded) return set_error(PSI_ERR_STATE, "psi_query_commit before psi_query_upload");
    int rc = ensure_device(c);
    if (rc) return rc;
    const size_t LN = (size_t)c->L * c->N;
    const uint32_t w = c->n_committed & 1u;
    cudaStream_t s = (cudaStream_t)stream;
    // index cts -> tiled split-30, Montgomery form
    CK(launch_retile_idx(c->k(s), c->idx_in.p + w * c->idx_words(), c->idx.p, LN, c->K, c->E));
    CK(cudaMemcpyAsync(c->minus.p, c->minus_in.p + w * 2 * LN, 2 * LN * sizeof(u64), cudaMemcpyDeviceToDevice, s));
    c->n_committed++;
    c->have_query = true;
    return PSI_OK;
}

int psi_run(psi_ctx* c, void* stream) { return psi_run_phases(c, PSI_PHASE_ALL, stream); }

static int phase2_streams(psi_ctx* c) {
    if (c->ev_fork) return PSI_OK;
    cudaError_t e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    for (int i = 0; i < 7 && e == cudaSuccess; i++) {
        e = cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) return cuda_fail(e, "bin-group streams");
    return PSI_OK;
}

// The ct x ct chain + mask of bins [g0, g1) into result (the context's [b][2][L][N] buffer), split into G bin
// groups on concurrent streams forked from / joined into s.  The bins are independent, so one group's kernels fill
// the SMs another group's tails leave idle (measured: 2 groups -3 % at 47 bins, -13 % at 12).
static int phase2_bins(psi_ctx* c, cudaStream_t s, uint32_t g0, uint32_t g1, uint32_t G, u64* result, uint32_t* nl,
                       uint32_t first_ops = 3) {
    const size_t ct = (size_t)2 * c->L * c->N;
    const uint32_t nb = g1 - g0;
    if (G > nb) G = nb;
    if (G > 8) G = 8;
    if (G < 1) G = 1;
    if (G > 1) {
        int rc0 = phase2_streams(c);
        if (rc0) return rc0;
        CK(cudaEventRecord(c->ev_fork, s));
    }
    for (uint32_t g = 0; g < G; g++) {
        const uint32_t b0 = g0 + (uint32_t)((uint64_t)nb * g / G), b1 = g0 + (uint32_t)((uint64_t)nb * (g + 1) / G);
        cudaStream_t sg = g == 0 ? s : c->aux[g - 1];
        if (g > 0) CK(cudaStreamWaitEvent(sg, c->ev_fork, 0));
        const u64* prod = c->acc.p + (size_t)b0 * ct;  // hf = 0
        for (uint32_t hf = 1; hf < c->K; hf++) {
            const bool last = hf + 1 == c->K;
            u64* dst = (last ? result : c->prod.p) + (size_t)b0 * ct;
            int rc = mul_ctct_batch(c, sg, b1 - b0, prod, c->acc.p + ((size_t)hf * c->b + b0) * ct,
                                    last ? c->mask.p + (size_t)b0 * c->L * c->N : nullptr, dst, nl, b0, hf == 1 ? first_ops : 3);
            if (rc) return rc;
            prod = dst;
        }
        if (g > 0) {
            CK(cudaEventRecord(c->ev_join[g - 1], sg));
            CK(cudaStreamWaitEvent(s, c->ev_join[g - 1], 0));
        }
    }
    return PSI_OK;
}

static uint32_t default_groups(const psi_ctx* c, uint32_t nbins) {
    return c->p2_groups ? c->p2_groups : (nbins >= 4 ? 2u : 1u);
}

// Enqueues one evaluation on s (and the auxiliary streams forked from it); result = the buffer this run writes.
static int enqueue_run(psi_ctx* c, uint32_t phases, cudaStream_t s, u64* result, uint32_t* nl_out) {
    const KCtx k = c->k(s);
    uint32_t nl = 0;
    int rc;
    if (phases & PSI_PHASE_INNER_PRODUCT) {
        CK(launch_mac(k, c->K, c->b, c->E, c->pt.p, c->idx.p, c->minus.p, c->acc.p)); nl++;
    }
    if (phases & PSI_PHASE_MULTIPLY_MASK) {
        if (c->K == 1) {
            cudaError_t e1 = launch_mul_ctpt(k, c->b, c->acc.p, c->mask.p, result);
            if (e1 != cudaSuccess) return cuda_fail(e1, "launch_mul_ctpt");
            nl++;
        } else if ((rc = phase2_bins(c, s, 0, c->b, default_groups(c, c->b), result, &nl))) {
            return rc;
        }
    }
    *nl_out = nl;
    return PSI_OK;
}

// everything a captured launch set depends on: dimensions, tuning, and the address of every buffer a kernel touches
static uint64_t run_graph_key(const psi_ctx* c, uint32_t phases, uint32_t out_which) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&h](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
    const void* ptrs[] = {c->pt.p,  c->mask.p, c->maskR.p, c->idx.p, c->minus.p, c->acc.p,   c->coef.p,  c->e1.p,   c->e2.p, c->ten.p,
                          c->res.p, c->dig.p,  c->prod.p,  c->out.p, c->out2.p,  c->evk_b.p, c->evk_a.p, c->evk_bR.p, c->evk_aR.p, c->d_tab};
    for (const void* p : ptrs) mix((uint64_t)(uintptr_t)p);
    for (uint64_t v : {(uint64_t)c->K, (uint64_t)c->b, (uint64_t)c->E, (uint64_t)phases, (uint64_t)out_which, (uint64_t)c->p2_groups,
                       (uint64_t)mac_forced_variant()})
        mix(v);
    return h;
}

static void drop_run_graphs(psi_ctx* c) {
    for (auto& g : c->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    c->graphs.clear();
}

int psi_run_phases(psi_ctx* c, uint32_t phases, void* stream) {
    if (!c) return set_error(PSI_ERR_INVALID, "null argument");
    if (!(phases & PSI_PHASE_ALL)) return set_error(PSI_ERR_INVALID, "no phase selected");
    if (!c->have_db || !c->have_query) return set_error(PSI_ERR_STATE, "run() needs a database and a query");
    if (c->K > 1 && !c->have_evk) return set_error(PSI_ERR_STATE, "EvalMult(ct,ct) needs the relinearisation key");
    int rc = ensure_device(c);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const bool writes_result = (phases & PSI_PHASE_MULTIPLY_MASK) != 0;
    // this run writes the OTHER result buffer; out_cur / ran only change once every launch has been enqueued, so a
    // failure part-way never exposes a half-written buffer through psi_result_get
    const uint32_t next_out = c->out_cur ^ 1u;
    u64* const result = c->out_buf(next_out);
    uint32_t nl = 0;
    // the legacy default stream cannot be captured: it takes the direct launches
    if (c->use_graph && s != nullptr && s != cudaStreamLegacy) {
        const uint64_t key = run_graph_key(c, phases, writes_result ? next_out : 2u);
        psi_ctx::RunGraph* g = nullptr;
        for (auto& e : c->graphs)
            if (e.key == key) g = &e;
        if (!g) {
            if (c->graphs.size() >= 16) drop_run_graphs(c);  // buffers were re-allocated many times: start over
            // the bin-group streams and events must exist before the capture starts
            if ((rc = phase2_streams(c))) return rc;
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            rc = enqueue_run(c, phases, s, result, &nl);
            const cudaError_t ee = cudaStreamEndCapture(s, &graph);
            if (rc != PSI_OK || ee != cudaSuccess) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                if (writes_result) c->ran = false;
                return rc != PSI_OK ? rc : cuda_fail(ee, "cudaStreamEndCapture");
            }
            cudaGraphExec_t exec = nullptr;
            const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ei != cudaSuccess) return cuda_fail(ei, "cudaGraphInstantiate");
            c->graphs.push_back(psi_ctx::RunGraph{key, exec, nl});
            g = &c->graphs.back();
        }
        const cudaError_t el = cudaGraphLaunch(g->exec, s);
        if (el != cudaSuccess) {
            if (writes_result) c->ran = false;
            return cuda_fail(el, "cudaGraphLaunch");
        }
        nl = g->launches;
    } else if ((rc = enqueue_run(c, phases, s, result, &nl))) {
        if (writes_result) c->ran = false;
        return rc;
    }
    c->launches_per_run = nl;
    if (writes_result) {
        c->out_cur = next_out;
        c->ran = true;
    }
    return PSI_OK;
}

// One query, host memory to host memory, with the transfers overlapped INSIDE the query: what the reference's server
// does once per session (BatchedFHEPSIServer.cpp:99-108: setMinusCompareElement, setIndex, run, getResultList).
//   upload     the K*E index ciphertexts cross PCIe in slices (hash function, position range); as soon as a slice
//              has landed it is re-tiled and its part of the inner products is accumulated (launch_mac_range), so
//              when the last slice arrives only that slice's share of phase 1 is still to do;
//   download   the ct x ct chain runs over the bins in kStreamOutGroups groups; the result ciphertexts of a group
//              go back to the host while the next group is evaluated.
// idx, minus, out: host buffers (pinned for full speed) laid out as for psi_query_set / psi_result_get.  Everything
// is enqueued; the caller synchronises `stream` (the last download is ordered into it) before reading out.
// download groups shrink towards the end: only the LAST group's download is exposed, and a bin downloads faster
// (18.7 us at 56 GB/s) than it evaluates (22 us), so the copy engine keeps up with the earlier, larger groups
constexpr uint32_t kStreamSlicesPerHf = 4, kStreamOutGroups = 4;
static const double kStreamOutCut[kStreamOutGroups + 1] = {0.0, 0.38, 0.72, 0.91, 1.0};
// idx / minus / out: contiguous (pinned) host buffers, or null with the limb-vector forms given instead: then every
// upload slice is first gathered into the pinned pool by the host threads (while the copy engine moves the previous
// slice) and every download group is scattered into the result vectors as soon as it has arrived (the call then
// returns with the vectors filled).
static int query_run_streamed_impl(psi_ctx* c, const uint64_t* idx, const uint64_t* minus, uint64_t* out,
                                   const uint64_t* const* idx_limbs, const uint64_t* const* minus_limbs,
                                   uint64_t* const* out_limbs, void* stream) {
    const bool limbs = idx_limbs != nullptr;
    if (!c->have_db) return set_error(PSI_ERR_STATE, "run() needs a database");
    if (c->K > 1 && !c->have_evk) return set_error(PSI_ERR_STATE, "EvalMult(ct,ct) needs the relinearisation key");
    if (c->n_uploaded != c->n_committed) return set_error(PSI_ERR_STATE, "an uploaded query is waiting for psi_query_commit");
    int rc = ensure_device(c);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (!c->sq_in) {
        CK(cudaStreamCreateWithFlags(&c->sq_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c->sq_out, cudaStreamNonBlocking));
        for (auto& e : c->ev_slice) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : c->ev_group) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_sq, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_sq_fork, cudaEventDisableTiming));
        for (auto& e : c->ev_dl) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        // one stream per download group, the earlier group at the higher priority: all groups are enqueued at once,
        // the block scheduler serves the first group first and fills its partial waves with the next one's CTAs
        int least = 0, greatest = 0;
        CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        for (int g = 0; g < 4; g++) {
            int prio = greatest + g;
            if (prio > least) prio = least;
            CK(cudaStreamCreateWithPriority(&c->sq_grp[g], cudaStreamNonBlocking, prio));
        }
    }
    const KCtx k = c->k(s);
    const size_t N = c->N, LN = (size_t)c->L * N, ct = 2 * LN;
    const uint32_t K = c->K, E = c->E, w = c->n_uploaded & 1u;
    const bool split_a = fused_mul_supported(k);
    const int nt = c->host_threads;
    (void)nt;
    if (limbs) {
        if ((rc = ensure_pool(&c->pool_in, &c->pool_in_words, c->idx_words() + 2 * LN))) return rc;
        if ((rc = ensure_pool(&c->pool_out, &c->pool_out_words, (size_t)c->b * ct))) return rc;
        if (!c->ev_pool_in) CK(cudaEventCreateWithFlags(&c->ev_pool_in, cudaEventDisableTiming));
        CK(cudaEventSynchronize(c->ev_pool_in));  // the previous upload has read the pool
        for (size_t v = 0; v < 2 * (size_t)c->L; v++) std::memcpy(c->pool_in + c->idx_words() + v * N, minus_limbs[v], N * sizeof(u64));
        minus = reinterpret_cast<const uint64_t*>(c->pool_in + c->idx_words());
        idx = reinterpret_cast<const uint64_t*>(c->pool_in);
        out = reinterpret_cast<uint64_t*>(c->pool_out);
    }
    // tuning switches (tools/tune_streamed.py): PSI_STREAM_SLICES=n, PSI_STREAM_CUTS="a,b,c"
    uint32_t slices_per_hf = kStreamSlicesPerHf;
    double out_cut[kStreamOutGroups + 1];
    for (uint32_t g = 0; g <= kStreamOutGroups; g++) out_cut[g] = kStreamOutCut[g];
    if (const char* e = std::getenv("PSI_STREAM_SLICES")) {
        const int v = std::atoi(e);
        if (v >= 1 && (uint32_t)v * K <= kMaxStreamSlices) slices_per_hf = (uint32_t)v;
    }
    if (const char* e = std::getenv("PSI_STREAM_CUTS")) {
        double a = 0, b2 = 0, c3 = 0;
        if (std::sscanf(e, "%lf,%lf,%lf", &a, &b2, &c3) == 3 && 0 <= a && a <= b2 && b2 <= c3 && c3 <= 1) {
            out_cut[1] = a;
            out_cut[2] = b2;
            out_cut[3] = c3;
        }
    }
    u64* const land = c->idx_in.p + w * c->idx_words();
    u64* const land_minus = c->minus_in.p + w * 2 * LN;
    uint32_t nl = 0;
    // PSI_STREAM_TIMELINE: debugging aid -- timing events at the milestones of the query, printed (after a
    // synchronisation) as one JSON line on stderr
    const bool timeline = std::getenv("PSI_STREAM_TIMELINE") != nullptr;
    cudaEvent_t tl[8] = {};
    if (timeline)
        for (auto& e : tl) CK(cudaEventCreate(&e));
    if (timeline) CK(cudaEventRecord(tl[0], s));
    // the copy stream starts where the caller's stream is (previous evaluation, previous use of the landing buffer)
    CK(cudaEventRecord(c->ev_sq, s));
    CK(cudaStreamWaitEvent(c->sq_in, c->ev_sq, 0));
    CK(cudaStreamWaitEvent(c->sq_out, c->ev_sq, 0));
    CK(cudaMemcpyAsync(land_minus, minus, ct * sizeof(u64), cudaMemcpyHostToDevice, c->sq_in));
    struct Slice {
        uint32_t hf, p0, p1;
        bool last_of_hf;
    };
    std::vector<Slice> slices;
    for (uint32_t hf = 0; hf < K; hf++) {
        const uint32_t S = E < slices_per_hf ? E : slices_per_hf;
        for (uint32_t sl = 0; sl < S; sl++)
            slices.push_back(Slice{hf, (uint32_t)((uint64_t)E * sl / S), (uint32_t)((uint64_t)E * (sl + 1) / S), sl + 1 == S});
    }
    if (slices.size() > kMaxStreamSlices) return set_error(PSI_ERR_INVALID, "too many hash functions for the streamed path");
    // H2D of slice j, then (ordered by an event) its re-tiling and its share of the inner products on the caller's stream
    auto enqueue_slice = [&](uint32_t j) -> int {
        const Slice& sl = slices[j];
        const size_t off = ((size_t)sl.hf * E + sl.p0) * ct;
        CK(cudaMemcpyAsync(land + off, idx + off, (size_t)(sl.p1 - sl.p0) * ct * sizeof(u64), cudaMemcpyHostToDevice, c->sq_in));
        CK(cudaEventRecord(c->ev_slice[j], c->sq_in));
        CK(cudaStreamWaitEvent(s, c->ev_slice[j], 0));
        if (j == 0) CK(cudaMemcpyAsync(c->minus.p, land_minus, ct * sizeof(u64), cudaMemcpyDeviceToDevice, s));
        CK(launch_retile_idx_range(k, land, c->idx.p, LN, E, sl.hf, sl.p0, sl.p1)); nl++;
        CK(launch_mac_range(k, sl.hf, 1, c->b, E, sl.p0, sl.p1, (sl.p0 > 0 ? 1u : 0u) | (sl.p1 == E ? 2u : 0u), c->pt.p, c->idx.p,
                            c->minus.p, c->acc.p)); nl++;
        // the first operand of the first multiplication (the inner products of hash function 0) is complete while the
        // ciphertexts of hash function 1 are still crossing PCIe: its row-inverse and exact basis extension run now
        if (sl.hf == 0 && sl.last_of_hf && K > 1 && split_a)
            return mul_ctct_batch(c, s, c->b, c->acc.p, c->acc.p + (size_t)c->b * ct, nullptr, nullptr, &nl, 0, 1);
        return PSI_OK;
    };
    auto slice_vectors = [&](uint32_t j, long* v0, long* v1) {
        const size_t off = ((size_t)slices[j].hf * E + slices[j].p0) * ct;
        *v0 = (long)(off / N);
        *v1 = (long)((off + (size_t)(slices[j].p1 - slices[j].p0) * ct) / N);
    };
    if (!limbs) {
        for (uint32_t j = 0; j < slices.size(); j++)
            if ((rc = enqueue_slice(j))) return rc;
    } else {
        // limb vectors: the host threads gather slice j + 1 into the pinned pool while the calling thread enqueues slice j
        // and the copy engine moves it
        long v0, v1;
        slice_vectors(0, &v0, &v1);
#pragma omp parallel for schedule(static) num_threads(nt) if (v1 - v0 >= 16)
        for (long v = v0; v < v1; v++) copy_limb_vector(c->pool_in + (size_t)v * N, idx_limbs[v], N);
        for (uint32_t j = 0; j < slices.size(); j++) {
            if (j + 1 == slices.size()) {
                if ((rc = enqueue_slice(j))) return rc;
                break;
            }
            slice_vectors(j + 1, &v0, &v1);
            int rc_master = PSI_OK;
#pragma omp parallel num_threads(nt)
            {
#pragma omp master
                rc_master = enqueue_slice(j);
#pragma omp for schedule(dynamic, 2) nowait
                for (long v = v0; v < v1; v++) copy_limb_vector(c->pool_in + (size_t)v * N, idx_limbs[v], N);
            }
            if (rc_master) return rc_master;
        }
    }
    if (limbs) CK(cudaEventRecord(c->ev_pool_in, c->sq_in));
    if (timeline) {
        CK(cudaEventRecord(tl[1], c->sq_in));  // upload complete
        CK(cudaEventRecord(tl[2], s));         // inner products complete
    }
    c->n_uploaded++;
    c->n_committed++;
    c->have_query = true;
    const uint32_t next_out = c->out_cur ^ 1u;
    u64* const result = c->out_buf(next_out);
    uint32_t cuts[kStreamOutGroups + 1];
    for (uint32_t g = 0; g <= kStreamOutGroups; g++) cuts[g] = (uint32_t)(out_cut[g] * c->b + 0.5);
    cuts[kStreamOutGroups] = c->b;
    // The groups run one after another on the caller's stream.  PSI_STREAM_CONCURRENT=1 enqueues them all at once on
    // prioritised streams instead -- measured slower (3.21-3.35 ms against 3.19-3.23): the priorities do not stagger the
    // completions, three of four groups finish together and the downloads bunch up (`profiles/r02_streamed_timeline.md`)
    const bool concurrent = std::getenv("PSI_STREAM_CONCURRENT") != nullptr;
    if (concurrent) CK(cudaEventRecord(c->ev_sq_fork, s));
    for (uint32_t g = 0; g < kStreamOutGroups; g++) {
        const uint32_t g0 = cuts[g], g1 = cuts[g + 1];
        if (g1 <= g0) continue;
        cudaStream_t sg = concurrent ? c->sq_grp[g] : s;
        if (concurrent) CK(cudaStreamWaitEvent(sg, c->ev_sq_fork, 0));
        if (K == 1) {
            const KCtx kg = c->k(sg);
            cudaError_t e1 = launch_mul_ctpt(kg, g1 - g0, c->acc.p + (size_t)g0 * ct, c->mask.p + (size_t)g0 * LN, result + (size_t)g0 * ct);
            if (e1 != cudaSuccess) {
                c->ran = false;
                return cuda_fail(e1, "launch_mul_ctpt");
            }
            nl++;
        } else if ((rc = phase2_bins(c, sg, g0, g1, concurrent ? 1u : default_groups(c, g1 - g0), result, &nl, split_a ? 2 : 3))) {
            c->ran = false;
            return rc;
        }
        CK(cudaEventRecord(c->ev_group[g], sg));
        if (timeline) CK(cudaEventRecord(tl[3 + g], sg));  // bin group g evaluated
        CK(cudaStreamWaitEvent(c->sq_out, c->ev_group[g], 0));
        CK(cudaMemcpyAsync(out + (size_t)g0 * ct, result + (size_t)g0 * ct, (size_t)(g1 - g0) * ct * sizeof(u64), cudaMemcpyDeviceToHost,
                           c->sq_out));
        if (limbs) CK(cudaEventRecord(c->ev_dl[g], c->sq_out));
    }
    // join: the caller's stream is done when the last download is
    CK(cudaEventRecord(c->ev_sq, c->sq_out));
    CK(cudaStreamWaitEvent(s, c->ev_sq, 0));
    if (timeline) {
        CK(cudaEventRecord(tl[7], c->sq_out));
        CK(cudaEventSynchronize(tl[7]));
        float ms[8] = {};
        for (int i = 1; i < 8; i++) cudaEventElapsedTime(&ms[i], tl[0], tl[i]);
        std::fprintf(stderr, "{\"streamed_timeline_ms\": {\"upload_done\": %.3f, \"inner_products_done\": %.3f, \"groups_done\": [%.3f, %.3f, %.3f, %.3f], \"last_download_done\": %.3f}}\n",
                     ms[1], ms[2], ms[3], ms[4], ms[5], ms[6], ms[7]);
        for (auto& e : tl) cudaEventDestroy(e);
    }
    c->out_cur = next_out;
    c->launches_per_run = nl;
    c->ran = true;
    if (limbs) {
        // scatter every group into the caller's vectors as soon as it has arrived
        for (uint32_t g = 0; g < kStreamOutGroups; g++) {
            const long v0 = (long)cuts[g] * 2 * c->L, v1 = (long)cuts[g + 1] * 2 * c->L;
            if (v1 <= v0) continue;
            CK(cudaEventSynchronize(c->ev_dl[g]));
#pragma omp parallel for schedule(static) num_threads(nt) if (v1 - v0 >= 16)
            for (long v = v0; v < v1; v++) copy_limb_vector(out_limbs[v], c->pool_out + (size_t)v * N, N);
        }
    }
    return PSI_OK;
}

int psi_query_run_streamed(psi_ctx* c, const uint64_t* idx, const uint64_t* minus, uint64_t* out, void* stream) {
    if (!c || !idx || !minus || !out) return set_error(PSI_ERR_INVALID, "null argument");
    return query_run_streamed_impl(c, idx, minus, out, nullptr, nullptr, nullptr, stream);
}

// The same from / into separately allocated limb vectors (psi_query_upload_limbs / psi_result_get_limbs layouts).
// Synchronous: the result vectors are filled on return.
int psi_query_run_streamed_limbs(psi_ctx* c, const uint64_t* const* idx_limbs, const uint64_t* const* minus_limbs,
                                 uint64_t* const* out_limbs, void* stream) {
    if (!c || !idx_limbs || !minus_limbs || !out_limbs) return set_error(PSI_ERR_INVALID, "null argument");
    return query_run_streamed_impl(c, nullptr, nullptr, nullptr, idx_limbs, minus_limbs, out_limbs, stream);
}

int psi_result_get(psi_ctx* c, uint64_t* out, void* stream) {
    if (!c || !out) return set_error(PSI_ERR_INVALID, "null argument");
    if (!c->ran) return set_error(PSI_ERR_STATE, "getResultList() before run()");
    int rc = ensure_device(c);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, c->out_buf(c->out_cur), (size_t)c->b * 2 * c->L * c->N * sizeof(u64), cudaMemcpyDeviceToHost,
                       (cudaStream_t)stream));
    return PSI_OK;
}

int psi_stream_sync(void* stream) {
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return PSI_OK;
}

int psi_run_launch_count(psi_ctx* c, uint32_t* out) {
    if (!c || !out) return set_error(PSI_ERR_INVALID, "null argument");
    if (!c->ran) return set_error(PSI_ERR_STATE, "launch count is known after the first run()");
    *out = c->launches_per_run;
    return PSI_OK;
}

int psi_result_device_ptr(psi_ctx* c, void** out, size_t* bytes) {
    if (!c || !out || !bytes) return set_error(PSI_ERR_INVALID, "null argument");
    if (!c->have_db) return set_error(PSI_ERR_STATE, "no database loaded");
    *out = c->out_buf(c->out_cur);
    *bytes = (size_t)c->b * 2 * c->L * c->N * sizeof(u64);
    return PSI_OK;
}

int psi_debug_ntt(psi_ctx* c, uint64_t* data, const uint32_t* moduli, uint32_t n_polys, int inverse) {
    if (!c || !data || !moduli) return set_error(PSI_ERR_INVALID, "null argument");
    int rc = ensure_device(c);
    if (rc) return rc;
    const size_t N = c->N;
    DevBuf<u64> d;
    CK(d.alloc((size_t)n_polys * N));
    CK(cudaMemcpy(d.p, data, (size_t)n_polys * N * sizeof(u64), cudaMemcpyHostToDevice));
    const KCtx k = c->k(0);
    // group consecutive polys with the same modulus into one launch
    uint32_t i = 0;
    while (i < n_polys) {
        if (moduli[i] > c->L + c->Lp) {
            d.release();
            return set_error(PSI_ERR_INVALID, "modulus index out of range");
        }
        uint32_t j = i;
        while (j < n_polys && moduli[j] == moduli[i]) j++;
        NttBatch nb{d.p + i * N, d.p + i * N, j - i, 1, N, 0, N, moduli[i], 1};
        cudaError_t e = launch_ntt(k, nb, inverse != 0);
        if (e != cudaSuccess) {
            d.release();
            return cuda_fail(e, "launch_ntt");
        }
        i = j;
    }
    cudaError_t e = cudaMemcpy(data, d.p, (size_t)n_polys * N * sizeof(u64), cudaMemcpyDeviceToHost);
    d.release();
    if (e != cudaSuccess) return cuda_fail(e, "psi_debug_ntt");
    return PSI_OK;
}

int psi_debug_mul_ctct(psi_ctx* c, const uint64_t* ct1, const uint64_t* ct2, uint64_t* out) {
    if (!c || !ct1 || !ct2 || !out) return set_error(PSI_ERR_INVALID, "null argument");
    if (!c->have_evk) return set_error(PSI_ERR_STATE, "EvalMult(ct,ct) needs the relinearisation key");
    int rc = ensure_device(c);
    if (rc) return rc;
    if (c->have_db && c->K < 2)
        return set_error(PSI_ERR_STATE, "debug ct x ct needs work buffers of a K >= 2 database (or no database)");
    // borrow the work buffers with b = 1, K = 2; the guard restores the context and frees the staging buffer on
    // every exit path
    struct Guard {
        psi_ctx* c;
        uint32_t K, b, E;
        DevBuf<u64> in;
        ~Guard() {
            in.release();
            c->K = K;
            c->b = b;
            c->E = E;
        }
    } g{c, c->K, c->b, c->E, {}};
    if (!c->have_db) {
        c->K = 2;
        c->b = 1;
        c->E = 1;
        if ((rc = alloc_work(c))) return rc;
    }
    const size_t ct = (size_t)2 * c->L * c->N;
    CK(g.in.alloc(3 * ct));
    CK(cudaMemcpy(g.in.p, ct1, ct * sizeof(u64), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(g.in.p + ct, ct2, ct * sizeof(u64), cudaMemcpyHostToDevice));
    if ((rc = mul_ctct_batch(c, 0, 1, g.in.p, g.in.p + ct, nullptr, g.in.p + 2 * ct, nullptr))) return rc;
    CK(cudaMemcpy(out, g.in.p + 2 * ct, ct * sizeof(u64), cudaMemcpyDeviceToHost));
    return PSI_OK;
}

int psi_bench_imad_peak(int device, double* mads_per_second) { return psi_bench_pipe_peak(device, 0, mads_per_second); }

int psi_bench_pipe_peak(int device, int kind, double* per_second) {
    if (!per_second) return set_error(PSI_ERR_INVALID, "null argument");
    if (kind < 0 || (kind & 15) > 7 || kind > 255) return set_error(PSI_ERR_INVALID, "unknown micro-benchmark kind");
    cudaError_t e = pipe_peak(device, kind, per_second);
    if (e != cudaSuccess) return cuda_fail(e, "psi_bench_pipe_peak");
    return PSI_OK;
}

}  // extern "C"

namespace psi {
int encode_into(psi_ctx* c, size_t n_pt, uint32_t nslots, const int64_t* slots, u64* dst, uint32_t tiled_E, size_t p_base) {
    return encode_into_impl(c, n_pt, nslots, slots, dst, tiled_E, p_base);
}
}  // namespace psi
