// Kernel launch interface between psi_api.cu (C ABI, orchestration) and the kernel TUs.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "modarith.cuh"
#include "psi_b200.h"

namespace psi {

constexpr int kMaxMods = 3 * PSI_MAX_LIMBS + 1;  // q, p, t, then the HYBRID special primes

// Device-resident copy of psi_params plus derived constants (Shoup companions, relin lifts).
struct DevTables {
    uint32_t N, logN, L, Lp;
    u64 t;
    u64 QHatInvModq[PSI_MAX_LIMBS], QHatInvModq_s[PSI_MAX_LIMBS];
    u64 QHatModp[PSI_MAX_LIMBS][PSI_MAX_LIMBS];
    u64 alphaQModp[PSI_MAX_LIMBS + 1][PSI_MAX_LIMBS];
    double qInv[PSI_MAX_LIMBS];
    u64 negPQHatInvModq[PSI_MAX_LIMBS], negPQHatInvModq_s[PSI_MAX_LIMBS];
    u64 qInvModp[PSI_MAX_LIMBS][PSI_MAX_LIMBS];
    u64 PHatInvModp[PSI_MAX_LIMBS], PHatInvModp_s[PSI_MAX_LIMBS];
    u64 PHatModq[PSI_MAX_LIMBS][PSI_MAX_LIMBS];
    u64 alphaPModq[PSI_MAX_LIMBS + 1][PSI_MAX_LIMBS];
    double pInv[PSI_MAX_LIMBS];
    u64 tQS[PSI_MAX_LIMBS][PSI_MAX_LIMBS + 1];
    double tQSfrac[PSI_MAX_LIMBS];
    u64 qModq[PSI_MAX_LIMBS][PSI_MAX_LIMBS];  // [i][k] = q_i mod q_k (BV digit lift)
    // N^-1 of the inverse NTT folded into the first constant of the two basis extensions
    u64 QHatInvNinv[PSI_MAX_LIMBS], QHatInvNinv_s[PSI_MAX_LIMBS];
    u64 negPQHatInvNinv[PSI_MAX_LIMBS], negPQHatInvNinv_s[PSI_MAX_LIMBS];
    // Shoup companions floor(c * 2^64 / modulus) of the conversion matrices (fused kernels)
    u64 QHatModp_s[PSI_MAX_LIMBS][PSI_MAX_LIMBS];
    u64 qInvModp_s[PSI_MAX_LIMBS][PSI_MAX_LIMBS];
    u64 PHatModq_s[PSI_MAX_LIMBS][PSI_MAX_LIMBS];
    u64 tQS_s[PSI_MAX_LIMBS][PSI_MAX_LIMBS + 1];
    // the same matrices in Montgomery form (times 2^64 mod the row's modulus), stored split-30: the fused
    // kernels form each RNS sum as carry-free IMAD.WIDE partial sums + ONE Montgomery reduction
    u64 QHatModp_m[PSI_MAX_LIMBS][PSI_MAX_LIMBS];
    u64 qInvModp_m[PSI_MAX_LIMBS][PSI_MAX_LIMBS];
    u64 PHatModq_m[PSI_MAX_LIMBS][PSI_MAX_LIMBS];
    u64 tQS_m[PSI_MAX_LIMBS][PSI_MAX_LIMBS + 1];
    // ---- variants of the context (risk register)
    uint32_t fp_fma;       // PSI_FP_FMA: the double sums nu += x * inv are evaluated with fused multiply-adds
    uint32_t hps;          // PSI_MULT_HPS
    // HPS: ScaleAndRound by t/Q with output basis P
    u64 tPS[PSI_MAX_LIMBS][PSI_MAX_LIMBS + 1];
    double tPSfrac[PSI_MAX_LIMBS];
    // HYBRID key switching: Q cut into ks_parts digits of ks_alpha consecutive limbs, extended basis Q + Lk special primes
    uint32_t ks_parts, ks_alpha, Lk, pad_;
    u64 PartQHatInvModq[PSI_MAX_LIMBS], PartQHatInvModq_s[PSI_MAX_LIMBS];
    u64 PartQHatModt[PSI_MAX_LIMBS][2 * PSI_MAX_LIMBS];  // [i][m]: (digit modulus / q_i) mod (m < L ? q_m : pk_{m-L})
    u64 PkInvModq[PSI_MAX_LIMBS], PkInvModq_s[PSI_MAX_LIMBS];
    u64 PkHatInvModpk[PSI_MAX_LIMBS], PkHatInvModpk_s[PSI_MAX_LIMBS];
    u64 PkHatModq[PSI_MAX_LIMBS][PSI_MAX_LIMBS];         // [k][i]
    ModDev mods[kMaxMods];                    // 0..L-1: q, L..L+Lp-1: p, L+Lp: t, L+Lp+1..: pk (HYBRID)
};

// nu += x * y the way the host library's compiler evaluates it (PSI_FP_SEPARATE / PSI_FP_FMA): explicit
// round-to-nearest intrinsics, so that nvcc never contracts or splits on its own
__device__ __forceinline__ double nu_step(double nu, double x, double y, uint32_t fma_mode) {
    return fma_mode ? __fma_rn(x, y, nu) : __dadd_rn(nu, __dmul_rn(x, y));
}

// What every launcher needs: device tables + the host copy of the dimensions + the stream.
struct KCtx {
    const DevTables* tab;
    uint32_t N, logN, L, Lp;
    cudaStream_t s;
    uint32_t Lk = 0;         // HYBRID: size of the key-switching basis
    bool fused_ok = true;    // the fused ct x ct kernels implement HPSPOVERQ + BV only
};

// Batched negacyclic NTT over `n_polys` limb-polynomials of N coefficients.
//   poly i: group g = i / G, limb l = i % G
//   src = src_base + g*src_gs + l*src_ls,  dst = dst_base + g*dst_gs + l*N   (element strides)
//   modulus index = mod_base + (l % mod_period)
struct NttBatch {
    const u64* src;
    u64* dst;
    uint32_t n_polys, G;
    size_t src_gs, src_ls, dst_gs;
    uint32_t mod_base, mod_period;
};
cudaError_t launch_ntt(const KCtx& k, const NttBatch& b, bool inverse);

// Phase 1: acc[hf][bin] = sum_pos idx[hf][pos] (.) pt[hf][bin][pos] + minus
// pt and idx are in tiled split-30 storage (launch_retile_*), minus and acc canonical [..][L][N].
cudaError_t launch_retile_pt(cudaStream_t s, u64* flat, u64* tiled, size_t LN, uint32_t E, size_t p0, size_t n,
                             bool to_tiled, const DevTables* range_tab = nullptr, int* bad = nullptr);
cudaError_t launch_retile_idx(const KCtx& k, const u64* flat, u64* tiled, size_t LN, uint32_t K, uint32_t E);
cudaError_t launch_retile_idx_range(const KCtx& k, const u64* flat, u64* tiled, size_t LN, uint32_t E, uint32_t hf, uint32_t pos0,
                                    uint32_t pos1);
cudaError_t launch_mac(const KCtx& k, uint32_t K, uint32_t b, uint32_t E, const u64* pt, const u64* idx,
                       const u64* minus, u64* acc);
// A slice of the same inner products: hash functions [hf0, hf0 + nhf), positions [pos0, pos1); flags bit 0 = add the
// previous contents of acc, bit 1 = add minusCompareElement
cudaError_t launch_mac_range(const KCtx& k, uint32_t hf0, uint32_t nhf, uint32_t b, uint32_t E, uint32_t pos0, uint32_t pos1,
                             uint32_t flags, const u64* pt, const u64* idx, const u64* minus, u64* acc);
int mac_forced_variant();
void mac_force_variant(int v);  // 0 = choose the bin-block width by shape, 1..3 = force 2 * v bins per CTA (tuning, tests)

// EvalMult(ct,ct) building blocks, all batched over B ciphertexts (see psi_api.cu for the sequence)
cudaError_t launch_expand_q_to_p(const KCtx& k, uint32_t groups, const u64* coef, u64* ext);
cudaError_t launch_fast_expand_poverq(const KCtx& k, uint32_t groups, const u64* coef, u64* ext);
cudaError_t launch_tensor(const KCtx& k, uint32_t B, const u64* e1, const u64* e2, u64* ten);
cudaError_t launch_scale_round(const KCtx& k, uint32_t groups, const u64* ten, u64* res);
// HPS: ScaleAndRound by t/Q into P, then the exact SwitchCRTBasis P -> Q; same buffers as launch_scale_round
cudaError_t launch_scale_round_hps(const KCtx& k, uint32_t groups, const u64* ten, u64* res);
// HYBRID key switching (KeySwitchHYBRID, unfused): see psi_kernels.cu
cudaError_t launch_hybrid_modup(const KCtx& k, uint32_t B, const u64* res, u64* dig);
cudaError_t launch_hybrid_inner(const KCtx& k, uint32_t B, const u64* dig, const u64* evk_b, const u64* evk_a, u64* ext);
cudaError_t launch_hybrid_moddown(const KCtx& k, uint32_t B, const u64* ext, u64* sw);
cudaError_t launch_hybrid_finish(const KCtx& k, uint32_t B, const u64* res_eval, const u64* ext, const u64* sw, const u64* mask,
                                 u64* out);
cudaError_t launch_relin_digits(const KCtx& k, uint32_t B, const u64* res, u64* dig);
cudaError_t launch_relin_accum(const KCtx& k, uint32_t B, const u64* res_eval, const u64* dig, const u64* evk_b,
                               const u64* evk_a, const u64* mask /*nullable*/, u64* out);
// Fused EvalMult(ct,ct) + relinearise (+ mask), fused_mul.cu.  a, b: [B][2][L][N] EVALUATION; scratch:
// ha, hb [B][2][L][N], e1p [B][2][Lp][N], e2h [B][2][LT][N], th [B][3][LT][N], rh [B][2][L][N], dh [B][L][L][N].
bool fused_mul_supported(const KCtx& k);
// per-device function attributes (dynamic shared memory above 48 KiB); called by psi_ctx_create
cudaError_t fused_mul_init_device(const KCtx& k);
cudaError_t ntt_init_device();
cudaError_t mac_init_device();
// dst[g][l][n] = src[g][l][n] * 2^64 mod q_l  (groups of L limb-polys): Montgomery form of constants
cudaError_t launch_to_montgomery(const KCtx& k, uint32_t groups, const u64* src, u64* dst);
// ops: 3 = the whole multiplication; 1 = only the part that needs operand a alone (row-inverse + exact Q -> P
// extension of a: scratch ha, e1p); 2 = the rest, after an ops = 1 launch on the same scratch
cudaError_t launch_fused_mul(const KCtx& k, uint32_t B, const u64* a, const u64* b, u64* ha, u64* hb, u64* e1p,
                             u64* e2h, u64* th, u64* rh, u64* dh, const u64* evk_b, const u64* evk_a, const u64* mask,
                             u64* out, uint32_t ops = 3);
cudaError_t launch_mul_ctpt(const KCtx& k, uint32_t B, const u64* ct, const u64* pt, u64* out);
// Fused automorphism key switch of the non-batched path (fused_nb.cu): cur [B][2][L][N] -> out, scratch h [B][L][N],
// dh [B][L][L][N]; keys in Montgomery form [n_keys][L][L][N]; per-item key slot / inverse index through sel_*
bool nb_fused_supported(const KCtx& k);
cudaError_t nb_fused_init_device(const KCtx& k);
cudaError_t launch_nb_keyswitch(const KCtx& k, uint32_t B, const u64* cur, u64* h, u64* dh, const u64* key_bR, const u64* key_aR,
                                const int* sel_key, const uint32_t* sel_ginv, uint32_t sel_mod, bool add, u64* out);

// Packed encoding, centred lift: crt [n][N] in [0, t) -> out [n][L][N]
cudaError_t launch_centre_lift(const KCtx& k, uint32_t n, const u64* crt, u64* out);
// Packed encoding front end: slot values -> CRT-ordered residues mod t
cudaError_t launch_slots_to_crt(const KCtx& k, uint32_t n_pt, uint32_t nslots, const long long* slots,
                                const uint32_t* to_crt, u64* out);

// Device-side nested cuckoo table build and constructor transposition (hashing_dev.cu)
cudaError_t hct_build_device(cudaStream_t s, const u64* T, uint32_t k, uint32_t e, uint32_t K, uint32_t b, uint32_t E,
                             u64 eviction_seed, const u64* items, size_t n, u64* cells, int* failed);
// plaintexts p0 .. p0+n_pt-1 of the shard [bin_begin, bin_begin + b_local) of a b-bin table (p counts inside the shard)
cudaError_t launch_cells_to_crt(const KCtx& k, uint32_t n_pt, uint32_t p0, uint32_t nslots, uint32_t K, uint32_t b, uint32_t E,
                                uint32_t bin_begin, uint32_t b_local, const u64* cells, const uint16_t* perm,
                                const uint32_t* to_crt, u64* out);

cudaError_t pipe_peak(int device, int kind, double* per_second);

}  // namespace psi
