// Non-batched FHEHIPPIE on the device: one private indexed equality check per outer cell of the
// server's nested cuckoo table, reference FHEHIPPIE.cpp:9-77 (ctor :9-59, run :61-77), FHEHIPPIE.hpp:18-50, run by
// SimpleFHEPSIServer.cpp:126-160 through FHEHIPPIECollection (PIECollection.hpp).
//
// What run() does per PIE, in OpenFHE calls:   for hf, bin:  EvalInnerProduct(indexMatrix[hf], vectorizedCT[hf][bin], b)
//                                               EvalMerge over the bins, EvalMult by preCalcRandomMask[hf]
// What the device does instead: every (pie, hf, bin) triple is one ITEM and the whole collection advances in lock step,
//   (1) item = idx[pie][hf] (.) pt[item]                                        k_nb_mul_ctpt
//   (2) ceil(log2 b) EvalSum steps, all items with the same automorphism key:   the fused key switch of fused_nb.cu
//       (k_nb_rows_inv -> k_nb_cols_digits -> k_nb_rows_ks<ADD>); contexts it does not cover (N < 1024) take the generic
//       kernels below: INTT(c1) -> k_nb_digits -> NTT -> k_nb_ks_apply<ADD>
//   (3) item (.) packed {1, 0, ...}; rotation by -bin with the key of that bin:  k_nb_mul_ctpt, the same key switch <SET>
//   (4) sum over the bins of a (pie, hf) and (.) mask:                           k_nb_sum_mask
// so a collection of P PIEs costs the same 3 (steps + 1) + 3 launches as one PIE.  The key switch follows OpenFHE's
// EvalAutomorphism order (recalled, see oracle/psi_oracle.c): KeySwitchInPlace first, the EVALUATION-format
// permutation after it - fused here as a scatter: the thread that forms the key-switched value at position j stores
// it at the position the inverse index maps j to.  No floating point on this path: all residues are canonical.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "psi_ctx.cuh"

namespace psi {

struct NbState {
    // automorphism keys (EvalSumKeyGen + EvalRotateKeyGen, SimpleFHEPSIClient.cpp:79-90)
    std::vector<uint64_t> key_index;
    DevBuf<u64> key_b, key_a;  // [n_keys][L][L][N]
    DevBuf<u64> key_bR, key_aR;  // the same times 2^64 mod q_k (Montgomery form for the fused key switch)
    bool fused = false;          // fused_nb.cu covers this context (N >= 1024, L <= 7); PSI_NB_UNFUSED=1 forces the generic kernels
    // database: vectorizedCT / preCalcRandomMask of every PIE, EvalMerge's plaintext
    uint32_t n_pie = 0, K = 0, b = 0;
    DevBuf<u64> pt, mask, merge_pt;  // [n_pie][K][b][L][N], [n_pie][K][L][N], [L][N]
    bool have_db = false;
    // work buffers for one chunk of PIEs
    uint32_t chunk_pies = 0;
    DevBuf<u64> idx, cur, nxt, coef, dig, out;
    DevBuf<u64> ext, sw;         // HYBRID key switching: inner products over Q + pk, ApproxModDown scratch
    size_t key_words = 0;        // words of one key component: L*L*N (BV) or numPartQ*(L+Lk)*N (HYBRID)
    DevBuf<int> sel_key;        // [steps + b]: key slot per EvalSum step, then per bin (-1 = identity, bin 0)
    DevBuf<uint32_t> sel_ginv;  // inverse automorphism index of the same entries
    uint32_t n_sum = 0;
    uint32_t launches = 0;
    // copy streams and the events that order uploads, evaluations and downloads of consecutive chunks
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {}, ev_eval[2] = {}, ev_out[2] = {};
    ~NbState() {
        if (s_in) cudaStreamDestroy(s_in);
        if (s_out) cudaStreamDestroy(s_out);
        for (int i = 0; i < 2; i++) {
            if (ev_in[i]) cudaEventDestroy(ev_in[i]);
            if (ev_eval[i]) cudaEventDestroy(ev_eval[i]);
            if (ev_out[i]) cudaEventDestroy(ev_out[i]);
        }
    }
};

void nb_release(psi_ctx* c) {
    delete c->nb;
    c->nb = nullptr;
}

static uint64_t pow_mod_2n(uint64_t g, uint64_t e, uint64_t m) {
    uint64_t r = 1;
    g %= m;
    for (; e; e >>= 1) {
        if (e & 1) r = (r * g) % m;
        g = (g * g) % m;
    }
    return r;
}
// EvalSum_2n / GenerateIndices_2n (OpenFHE advancedshe, recalled): squares of 5, the last one 2N - 1 for a full row
static uint32_t eval_sum_indices(uint32_t N, uint32_t batch, uint64_t* out) {
    if (batch <= 1) return 0;
    const uint64_t m = 2ull * N;
    uint32_t steps = 0, n = 0;
    while ((1u << steps) < batch) steps++;
    uint64_t g = 5;
    for (uint32_t i = 0; i + 1 < steps; i++) {
        out[n++] = g;
        g = (g * g) % m;
    }
    out[n++] = (2ull * batch < m) ? g : m - 1;
    return n;
}
// FindAutomorphismIndex2n: 5^i mod 2N, with 5^-1 for negative i (the unit group of Z_2N has exponent N/2)
static uint64_t rotation_index(uint32_t N, int64_t i) {
    const uint64_t m = 2ull * N;
    const uint64_t g0 = i < 0 ? pow_mod_2n(5, N - 1, m) : 5;
    return pow_mod_2n(g0, (uint64_t)(i < 0 ? -i : i), m);
}

// PrecomputeAutoMap (recalled): output position p of the bit-reversed EVALUATION vector reads input position automap(p, g)
__device__ __forceinline__ uint32_t automap(uint32_t p, uint32_t g, uint32_t logN) {
    const uint32_t j = __brev(p) >> (32 - logN);
    const uint32_t idx = (((2 * j + 1) * g) & ((2u << logN) - 1)) >> 1;
    return __brev(idx) >> (32 - logN);
}

// EvalMult(ct, pt) over items: out[item] = ct[item / ct_div] (.) pt[item * pt_stride]   (pt_stride 0: one plaintext for all)
__global__ void __launch_bounds__(256) k_nb_mul_ctpt(const DevTables* __restrict__ tab, uint32_t N, uint32_t B,
                                                     const u64* __restrict__ ct, uint32_t ct_div, const u64* __restrict__ pt,
                                                     size_t pt_stride, u64* __restrict__ out) {
    const size_t LN = (size_t)tab->L * N;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)B * LN) return;
    const size_t item = tid / LN, c = tid % LN;
    const ModDev& m = tab->mods[c / N];
    const u64 pv = pt[item * pt_stride + c];
    const u64* src = ct + (item / ct_div) * 2 * LN;
    out[(item * 2) * LN + c] = mulmod(src[c], pv, m);
    out[(item * 2 + 1) * LN + c] = mulmod(src[LN + c], pv, m);
}

// DCRTPoly::CRTDecompose (BV, digit size 0) of the c1 component: digit i = limb i (COEFFICIENT), switched to every q_k
// with the centred lift of NativeVector::SwitchModulus.  coef: [B][L][N], dig: [B][L(i)][L(k)][N]
__global__ void __launch_bounds__(256) k_nb_digits(const DevTables* __restrict__ tab, uint32_t N, uint32_t B,
                                                   const u64* __restrict__ coef, u64* __restrict__ dig) {
    const int L = tab->L;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)B * L * N) return;
    const uint32_t n = tid % N, i = (tid / N) % L;
    const size_t item = tid / ((size_t)N * L);
    const u64 v = coef[tid];
    const u64 qi = tab->mods[i].q;
    const bool neg = v > ((qi - 1) >> 1);
#pragma unroll
    for (int k = 0; k < PSI_MAX_LIMBS; k++) {
        if (k < L) {
            const u64 qk = tab->mods[k].q;
            u64 r = v;
            if (k != (int)i) {
                r = (v < qk) ? v : ((v - qk < qk) ? v - qk : v % qk);
                if (neg) r = submod(r, tab->qModq[i][k], qk);
            }
            dig[((item * L + i) * L + k) * N + n] = r;
        }
    }
}

// KeySwitchBV core + KeySwitchInPlace + AutomorphismTransform, and the EvalAdd of EvalSum when ADD.
//   cur: [B][2][L][N] EVAL, dig: [B][L][L][N] EVAL, keys: [n_keys][L][L][N]
//   the item's key slot and inverse index come from sel_key / sel_ginv[sel_mod ? item % sel_mod : 0]; slot -1 = the
//   identity (bin 0 of EvalMerge is not rotated)
//   ADD: out[item] = cur[item] + sigma_g(keyswitched cur[item]);  else out[item] = sigma_g(keyswitched cur[item])
template <bool ADD>
__global__ void __launch_bounds__(256) k_nb_ks_apply(const DevTables* __restrict__ tab, uint32_t N, uint32_t logN, uint32_t B,
                                                     const u64* __restrict__ cur, const u64* __restrict__ dig,
                                                     const u64* __restrict__ key_b, const u64* __restrict__ key_a,
                                                     const int* __restrict__ sel_key, const uint32_t* __restrict__ sel_ginv,
                                                     uint32_t sel_mod, u64* __restrict__ out) {
    const int L = tab->L;
    const size_t LN = (size_t)L * N;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)B * LN) return;
    const size_t item = tid / LN, c = tid % LN;  // c = k*N + j
    const uint32_t sel = sel_mod ? (uint32_t)(item % sel_mod) : 0;
    const int slot = sel_key[sel];
    const u64* ct = cur + item * 2 * LN;
    u64* o = out + item * 2 * LN;
    if (slot < 0) {
        o[c] = ct[c];
        o[LN + c] = ct[LN + c];
        return;
    }
    const ModDev& m = tab->mods[c / N];
    const u64* kb = key_b + (size_t)slot * L * LN;
    const u64* ka = key_a + (size_t)slot * L * LN;
    u64 h0 = 0, l0 = 0, h1 = 0, l1 = 0;
#pragma unroll
    for (int i = 0; i < PSI_MAX_LIMBS; i++) {
        if (i < L) {
            const u64 d = dig[(item * L + i) * LN + c];
            mac128(h0, l0, d, kb[(size_t)i * LN + c]);
            mac128(h1, l1, d, ka[(size_t)i * LN + c]);
        }
    }
    u64 v0 = addmod(barrett128(h0, l0, m.q, m.mu_hi, m.mu_lo), ct[c], m.q);
    u64 v1 = barrett128(h1, l1, m.q, m.mu_hi, m.mu_lo);
    const uint32_t j = (uint32_t)(c % N);
    const size_t dst = (c - j) + automap(j, sel_ginv[sel], logN);
    if (ADD) {
        v0 = addmod(v0, ct[dst], m.q);
        v1 = addmod(v1, ct[LN + dst], m.q);
    }
    o[dst] = v0;
    o[LN + dst] = v1;
}

// ---- HYBRID key switching (KeySwitchHYBRID, recalled; the relinearisation form lives in psi_kernels.cu) ------------
// c1 (COEFFICIENT, [B][L][N]) is cut into ks_parts digits of ks_alpha consecutive limbs; digit j is lifted to every other
// limb of Q + pk by ApproxSwitchCRTBasis, its own limbs keep the coefficients.  dig: [B][parts][L+Lk][N] COEFFICIENT.
__global__ void __launch_bounds__(256) k_nb_hybrid_modup(const DevTables* __restrict__ tab, uint32_t N, uint32_t B,
                                                         const u64* __restrict__ coef, u64* __restrict__ dig) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)B * N) return;
    const size_t item = tid / N;
    const uint32_t n = tid % N;
    const int L = tab->L, Lk = tab->Lk, LE = L + Lk, parts = tab->ks_parts, alpha = tab->ks_alpha;
    u64 x[PSI_MAX_LIMBS], y[PSI_MAX_LIMBS];
#pragma unroll
    for (int i = 0; i < PSI_MAX_LIMBS; i++)
        if (i < L) {
            x[i] = coef[(item * L + i) * N + n];
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
            dig[((item * parts + j) * LE + m) * N + n] = v;
        }
    }
}

// ext[item][comp][m][n] = sum_j dig[item][j][m][n] * key_comp[slot(item)][j][m][n] over the extended basis, EVALUATION
__global__ void __launch_bounds__(256) k_nb_hybrid_inner(const DevTables* __restrict__ tab, uint32_t N, uint32_t B,
                                                         const u64* __restrict__ dig, const u64* __restrict__ key_b,
                                                         const u64* __restrict__ key_a, const int* __restrict__ sel_key,
                                                         uint32_t sel_mod, u64* __restrict__ ext) {
    const int L = tab->L, LE = L + tab->Lk, parts = tab->ks_parts;
    const size_t LEN = (size_t)LE * N;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)B * LEN) return;
    const size_t item = tid / LEN, c = tid % LEN;
    const int slot = sel_key[sel_mod ? (uint32_t)(item % sel_mod) : 0];
    if (slot < 0) return;  // identity item: k_nb_hybrid_finish copies it
    const int m = (int)(c / N);
    const ModDev& md = tab->mods[m < L ? m : tab->L + tab->Lp + 1 + (m - L)];
    const u64* kb = key_b + (size_t)slot * parts * LEN;
    const u64* ka = key_a + (size_t)slot * parts * LEN;
    u64 h0 = 0, l0 = 0, h1 = 0, l1 = 0;
    for (int j = 0; j < parts; j++) {
        const u64 d = dig[(item * parts + j) * LEN + c];
        mac128(h0, l0, d, kb[(size_t)j * LEN + c]);
        mac128(h1, l1, d, ka[(size_t)j * LEN + c]);
    }
    ext[(item * 2) * LEN + c] = barrett128(h0, l0, md.q, md.mu_hi, md.mu_lo);
    ext[(item * 2 + 1) * LEN + c] = barrett128(h1, l1, md.q, md.mu_hi, md.mu_lo);
}

// ApproxModDown, second half, + KeySwitchInPlace (c0 += d0, c1 = d1) + AutomorphismTransform (+ the EvalAdd of EvalSum)
//   cur: [B][2][L][N] EVAL, ext: [B][2][L+Lk][N] (Q limbs EVAL), sw: [B][2][L][N] EVAL
template <bool ADD>
__global__ void __launch_bounds__(256) k_nb_hybrid_finish(const DevTables* __restrict__ tab, uint32_t N, uint32_t logN, uint32_t B,
                                                          const u64* __restrict__ cur, const u64* __restrict__ ext,
                                                          const u64* __restrict__ sw, const int* __restrict__ sel_key,
                                                          const uint32_t* __restrict__ sel_ginv, uint32_t sel_mod,
                                                          u64* __restrict__ out) {
    const int L = tab->L, LE = L + tab->Lk;
    const size_t LN = (size_t)L * N;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)B * LN) return;
    const size_t item = tid / LN, c = tid % LN;
    const uint32_t sel = sel_mod ? (uint32_t)(item % sel_mod) : 0;
    const u64* ct = cur + item * 2 * LN;
    u64* o = out + item * 2 * LN;
    if (sel_key[sel] < 0) {
        o[c] = ct[c];
        o[LN + c] = ct[LN + c];
        return;
    }
    const int i = (int)(c / N);
    const ModDev& m = tab->mods[i];
    const uint32_t j = (uint32_t)(c % N);
    const size_t dst = (c - j) + automap(j, sel_ginv[sel], logN);
#pragma unroll
    for (int comp = 0; comp < 2; comp++) {
        const u64 e = ext[((item * 2 + comp) * LE) * N + c];
        const u64 d = submod(e, sw[(item * 2 + comp) * LN + c], m.q);
        u64 r = mul_shoup(d, tab->PkInvModq[i], tab->PkInvModq_s[i], m.q);
        if (comp == 0) r = addmod(r, ct[c], m.q);
        if (ADD) r = addmod(r, ct[comp * LN + dst], m.q);
        o[comp * LN + dst] = r;
    }
}

// EvalMerge's EvalAdd chain over the b bins of a (pie, hf) group + EvalMult by the random mask (FHEHIPPIE.cpp:73)
__global__ void __launch_bounds__(256) k_nb_sum_mask(const DevTables* __restrict__ tab, uint32_t N, uint32_t G, uint32_t b,
                                                     const u64* __restrict__ items, const u64* __restrict__ mask,
                                                     u64* __restrict__ out) {
    const size_t LN = (size_t)tab->L * N;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (size_t)G * LN) return;
    const size_t grp = tid / LN, c = tid % LN;
    const ModDev& m = tab->mods[c / N];
    u64 s0 = 0, s1 = 0;
    for (uint32_t bin = 0; bin < b; bin++) {
        const u64* ct = items + (grp * b + bin) * 2 * LN;
        s0 = addmod(s0, ct[c], m.q);
        s1 = addmod(s1, ct[LN + c], m.q);
    }
    const u64 mv = mask[grp * LN + c];
    out[(grp * 2) * LN + c] = mulmod(s0, mv, m);
    out[(grp * 2 + 1) * LN + c] = mulmod(s1, mv, m);
}

static inline unsigned blocks_for(size_t total) { return (unsigned)((total + 255) / 256); }

static NbState* nb_state(psi_ctx* c) {
    if (!c->nb) c->nb = new (std::nothrow) NbState();
    return c->nb;
}

static int nb_check_ctx(psi_ctx* c) {
    if (!c) return set_error(PSI_ERR_INVALID, "null argument");
    return ensure_device(c);
}

static int nb_db_dims(psi_ctx* c, NbState* s, uint32_t n_pie, uint32_t K, uint32_t b) {
    if (n_pie < 1 || K < 1 || b < 1) return set_error(PSI_ERR_INVALID, "n_pie, K and b must be positive");
    if ((uint64_t)b + 1 > c->N / 2) return set_error(PSI_ERR_INVALID, "b + 1 slots must fit one row of the packing (N/2)");
    const size_t LN = (size_t)c->L * c->N;
    s->have_db = false;
    s->n_pie = n_pie;
    s->K = K;
    s->b = b;
    CK(s->pt.alloc((size_t)n_pie * K * b * LN));
    CK(s->mask.alloc((size_t)n_pie * K * LN));
    CK(s->merge_pt.alloc(LN));
    s->chunk_pies = 0;
    return PSI_OK;
}

// work buffers + selection tables for chunks of up to `pies` PIEs
static int nb_prepare(psi_ctx* c, NbState* s, uint32_t pies) {
    if (s->chunk_pies >= pies) return PSI_OK;
    const uint32_t L = c->L, N = c->N, K = s->K, b = s->b;
    const size_t LN = (size_t)L * N, items = (size_t)pies * K * b;
    // key slots: EvalSum steps with batch size b (EvalInnerProduct(.., vectorizedCT[hfInd].size()), FHEHIPPIE.cpp:70),
    // then the rotation by -bin of EvalMerge
    uint64_t sum_idx[32];
    s->n_sum = eval_sum_indices(N, b, sum_idx);
    std::vector<int> sel_key(s->n_sum + b);
    std::vector<uint32_t> sel_ginv(s->n_sum + b);
    auto find = [&](uint64_t g) -> int {
        for (size_t i = 0; i < s->key_index.size(); i++)
            if (s->key_index[i] == g) return (int)i;
        return -1;
    };
    for (uint32_t i = 0; i < s->n_sum + b; i++) {
        if (i == s->n_sum) {  // bin 0 stays where it is
            sel_key[i] = -1;
            sel_ginv[i] = 1;
            continue;
        }
        const uint64_t g = i < s->n_sum ? sum_idx[i] : rotation_index(N, -(int64_t)(i - s->n_sum));
        sel_key[i] = find(g);
        if (sel_key[i] < 0)  // OpenFHE: "Could not find an EvalKey for index ..."
            return set_error(PSI_ERR_STATE, "automorphism key for index " + std::to_string(g) + " has not been set");
        sel_ginv[i] = (uint32_t)pow_mod_2n(g, N - 1, 2ull * N);
    }
    CK(s->sel_key.alloc(sel_key.size()));
    CK(s->sel_ginv.alloc(sel_ginv.size()));
    CK(cudaMemcpy(s->sel_key.p, sel_key.data(), sel_key.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->sel_ginv.p, sel_ginv.data(), sel_ginv.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CK(s->idx.alloc((size_t)2 * pies * K * 2 * LN));  // double-buffered: chunk i + 1 uploads while chunk i is evaluated
    CK(s->out.alloc((size_t)2 * pies * K * 2 * LN));
    if (!s->s_in) {
        CK(cudaStreamCreateWithFlags(&s->s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            CK(cudaEventCreateWithFlags(&s->ev_in[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&s->ev_eval[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&s->ev_out[i], cudaEventDisableTiming));
        }
    }
    CK(s->cur.alloc(items * 2 * LN));
    CK(s->nxt.alloc(items * 2 * LN));
    CK(s->coef.alloc(items * LN));
    const size_t dig_polys = c->hybrid ? std::max<size_t>((size_t)L * L, (size_t)c->ks_parts * (L + c->Lk)) : (size_t)L * L;
    CK(s->dig.alloc(items * dig_polys * N));
    if (c->hybrid) {
        CK(s->ext.alloc(items * 2 * (L + c->Lk) * N));
        CK(s->sw.alloc(items * 2 * LN));
    }
    s->chunk_pies = pies;
    return PSI_OK;
}

// the key switch + automorphism of every item: cur -> nxt
template <bool ADD>
static int nb_ks_step(psi_ctx* c, NbState* s, const KCtx& k, uint32_t B, uint32_t sel0, uint32_t sel_mod) {
    const uint32_t L = c->L, N = c->N;
    const size_t LN = (size_t)L * N;
    if (s->fused) {  // three fused kernels (fused_nb.cu)
        CK(launch_nb_keyswitch(k, B, s->cur.p, s->coef.p, s->dig.p, s->key_bR.p, s->key_aR.p, s->sel_key.p + sel0, s->sel_ginv.p + sel0,
                               sel_mod, ADD, s->nxt.p));
        std::swap(s->cur.p, s->nxt.p);
        std::swap(s->cur.n, s->nxt.n);
        s->launches += 3;
        return PSI_OK;
    }
    if (c->hybrid) {  // KeySwitchHYBRID, one kernel per operation
        const uint32_t Lk = c->Lk, LE = L + Lk, parts = c->ks_parts, pk0 = L + c->Lp + 1;
        NttBatch hb{s->cur.p + LN, s->coef.p, B * L, L, 2 * LN, N, LN, 0, L};
        CK(launch_ntt(k, hb, true));
        k_nb_hybrid_modup<<<blocks_for((size_t)B * N), 256, 0, k.s>>>(k.tab, N, B, s->coef.p, s->dig.p);
        CK(cudaGetLastError());
        hb = NttBatch{s->dig.p, s->dig.p, B * parts * L, L, (size_t)LE * N, N, (size_t)LE * N, 0, L};
        CK(launch_ntt(k, hb, false));
        hb = NttBatch{s->dig.p + LN, s->dig.p + LN, B * parts * Lk, Lk, (size_t)LE * N, N, (size_t)LE * N, pk0, Lk};
        CK(launch_ntt(k, hb, false));
        k_nb_hybrid_inner<<<blocks_for((size_t)B * LE * N), 256, 0, k.s>>>(k.tab, N, B, s->dig.p, s->key_b.p, s->key_a.p, s->sel_key.p + sel0,
                                                                        sel_mod, s->ext.p);
        CK(cudaGetLastError());
        hb = NttBatch{s->ext.p + LN, s->ext.p + LN, B * 2 * Lk, Lk, (size_t)LE * N, N, (size_t)LE * N, pk0, Lk};
        CK(launch_ntt(k, hb, true));
        CK(launch_hybrid_moddown(k, B, s->ext.p, s->sw.p));
        hb = NttBatch{s->sw.p, s->sw.p, B * 2 * L, L, LN, N, LN, 0, L};
        CK(launch_ntt(k, hb, false));
        k_nb_hybrid_finish<ADD><<<blocks_for((size_t)B * LN), 256, 0, k.s>>>(k.tab, N, c->logN, B, s->cur.p, s->ext.p, s->sw.p,
                                                                          s->sel_key.p + sel0, s->sel_ginv.p + sel0, sel_mod, s->nxt.p);
        CK(cudaGetLastError());
        std::swap(s->cur.p, s->nxt.p);
        std::swap(s->cur.n, s->nxt.n);
        s->launches += 9;
        return PSI_OK;
    }
    NttBatch nb{s->cur.p + LN, s->coef.p, B * L, L, 2 * LN, N, LN, 0, L};  // c1 of every item to COEFFICIENT
    CK(launch_ntt(k, nb, true));
    k_nb_digits<<<blocks_for((size_t)B * LN), 256, 0, k.s>>>(k.tab, N, B, s->coef.p, s->dig.p);
    CK(cudaGetLastError());
    nb = NttBatch{s->dig.p, s->dig.p, B * L * L, L, LN, N, LN, 0, L};
    CK(launch_ntt(k, nb, false));
    k_nb_ks_apply<ADD><<<blocks_for((size_t)B * LN), 256, 0, k.s>>>(k.tab, N, c->logN, B, s->cur.p, s->dig.p, s->key_b.p, s->key_a.p,
                                                                   s->sel_key.p + sel0, s->sel_ginv.p + sel0, sel_mod, s->nxt.p);
    CK(cudaGetLastError());
    std::swap(s->cur.p, s->nxt.p);
    std::swap(s->cur.n, s->nxt.n);
    s->launches += 4;
    return PSI_OK;
}

}  // namespace psi

using namespace psi;

extern "C" {

int psi_nb_eval_sum_indices(uint32_t N, uint32_t batch_size, uint64_t* out, uint32_t* n) {
    if (!out || !n || N < 4 || (N & (N - 1))) return set_error(PSI_ERR_INVALID, "bad argument");
    *n = eval_sum_indices(N, batch_size, out);
    return PSI_OK;
}

int psi_nb_rotation_index(uint32_t N, int64_t i, uint64_t* out) {
    if (!out || N < 4 || (N & (N - 1))) return set_error(PSI_ERR_INVALID, "bad argument");
    *out = rotation_index(N, i);
    return PSI_OK;
}

int psi_nb_set_automorphism_keys(psi_ctx* c, uint32_t n_keys, const uint64_t* auto_index, const uint64_t* key_b,
                                 const uint64_t* key_a) {
    int rc = nb_check_ctx(c);
    if (rc) return rc;
    if (!n_keys || !auto_index || !key_b || !key_a) return set_error(PSI_ERR_INVALID, "null argument");
    NbState* s = nb_state(c);
    if (!s) return set_error(PSI_ERR_CUDA, "out of host memory");
    const uint64_t m = 2ull * c->N;
    for (uint32_t i = 0; i < n_keys; i++)
        if (!(auto_index[i] & 1) || auto_index[i] >= m)
            return set_error(PSI_ERR_INVALID, "an automorphism index must be odd and below 2N");
    s->key_words = c->hybrid ? (size_t)c->ks_parts * (c->L + c->Lk) * c->N : (size_t)c->L * c->L * c->N;
    const size_t words = (size_t)n_keys * s->key_words;
    CK(s->key_b.alloc(words));
    CK(s->key_a.alloc(words));
    CK(cudaMemcpy(s->key_b.p, key_b, words * sizeof(u64), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->key_a.p, key_a, words * sizeof(u64), cudaMemcpyHostToDevice));
    s->key_index.assign(auto_index, auto_index + n_keys);
    const KCtx k = c->k(0);
    s->fused = nb_fused_supported(k) && std::getenv("PSI_NB_UNFUSED") == nullptr;
    if (s->fused) {
        CK(nb_fused_init_device(k));
        CK(s->key_bR.alloc(words));
        CK(s->key_aR.alloc(words));
        CK(launch_to_montgomery(k, n_keys * c->L, s->key_b.p, s->key_bR.p));
        CK(launch_to_montgomery(k, n_keys * c->L, s->key_a.p, s->key_aR.p));
        CK(cudaStreamSynchronize(0));
    }
    s->chunk_pies = 0;  // the key slots of the selection tables are re-derived
    return PSI_OK;
}

int psi_nb_db_load_limbs(psi_ctx* c, uint32_t n_pie, uint32_t K, uint32_t b, const uint64_t* pt_limbs,
                         const uint64_t* mask_limbs, const uint64_t* merge_limbs) {
    int rc = nb_check_ctx(c);
    if (rc) return rc;
    if (!pt_limbs || !mask_limbs || !merge_limbs) return set_error(PSI_ERR_INVALID, "null argument");
    NbState* s = nb_state(c);
    if (!s) return set_error(PSI_ERR_CUDA, "out of host memory");
    if ((rc = nb_db_dims(c, s, n_pie, K, b))) return rc;
    const size_t LN = (size_t)c->L * c->N;
    CK(cudaMemcpy(s->pt.p, pt_limbs, (size_t)n_pie * K * b * LN * sizeof(u64), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->mask.p, mask_limbs, (size_t)n_pie * K * LN * sizeof(u64), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->merge_pt.p, merge_limbs, LN * sizeof(u64), cudaMemcpyHostToDevice));
    s->have_db = true;
    return PSI_OK;
}

int psi_nb_db_encode_slots(psi_ctx* c, uint32_t n_pie, uint32_t K, uint32_t b, uint32_t nslots, const int64_t* slots,
                           const int64_t* mask_slots) {
    int rc = nb_check_ctx(c);
    if (rc) return rc;
    if (!slots || !mask_slots) return set_error(PSI_ERR_INVALID, "null argument");
    if (nslots < 1 || nslots > c->N) return set_error(PSI_ERR_INVALID, "batch size must be in [1, N]");
    const uint64_t t = c->P.t;
    const size_t n_pt = (size_t)n_pie * K * b, n_mask = (size_t)n_pie * K;
    for (size_t i = 0; i < n_pt * nslots; i++)
        if ((uint64_t)(slots[i] < 0 ? -slots[i] : slots[i]) >= t)
            return set_error(PSI_ERR_INVALID, "slot value out of range of the plaintext modulus");
    for (size_t i = 0; i < n_mask * b; i++)
        if ((uint64_t)(mask_slots[i] < 0 ? -mask_slots[i] : mask_slots[i]) >= t)
            return set_error(PSI_ERR_INVALID, "mask value out of range of the plaintext modulus");
    NbState* s = nb_state(c);
    if (!s) return set_error(PSI_ERR_CUDA, "out of host memory");
    if ((rc = nb_db_dims(c, s, n_pie, K, b))) return rc;
    if ((rc = encode_into(c, n_pt, nslots, slots, s->pt.p, 0))) return rc;
    if ((rc = encode_into(c, n_mask, b, mask_slots, s->mask.p, 0))) return rc;
    const int64_t one = 1;  // EvalMerge: MakePackedPlaintext({1, 0, 0, ...})
    if ((rc = encode_into(c, 1, 1, &one, s->merge_pt.p, 0))) return rc;
    s->have_db = true;
    return PSI_OK;
}

int psi_nb_db_get_limbs(psi_ctx* c, uint64_t* pt_limbs, uint64_t* mask_limbs, uint64_t* merge_limbs) {
    int rc = nb_check_ctx(c);
    if (rc) return rc;
    NbState* s = c->nb;
    if (!s || !s->have_db) return set_error(PSI_ERR_STATE, "no non-batched database loaded");
    const size_t LN = (size_t)c->L * c->N;
    if (pt_limbs) CK(cudaMemcpy(pt_limbs, s->pt.p, (size_t)s->n_pie * s->K * s->b * LN * sizeof(u64), cudaMemcpyDeviceToHost));
    if (mask_limbs) CK(cudaMemcpy(mask_limbs, s->mask.p, (size_t)s->n_pie * s->K * LN * sizeof(u64), cudaMemcpyDeviceToHost));
    if (merge_limbs) CK(cudaMemcpy(merge_limbs, s->merge_pt.p, LN * sizeof(u64), cudaMemcpyDeviceToHost));
    return PSI_OK;
}

int psi_nb_run(psi_ctx* c, uint32_t pie_begin, uint32_t pie_end, const uint64_t* idx, uint64_t* out, void* stream) {
    int rc = nb_check_ctx(c);
    if (rc) return rc;
    NbState* s = c->nb;
    if (!s || !s->have_db) return set_error(PSI_ERR_STATE, "no non-batched database loaded");
    if (s->key_index.empty()) return set_error(PSI_ERR_STATE, "no automorphism keys set");
    if (!idx || !out) return set_error(PSI_ERR_INVALID, "null argument");
    if (pie_begin >= pie_end || pie_end > s->n_pie) return set_error(PSI_ERR_INVALID, "bad PIE range");
    const uint32_t L = c->L, N = c->N, K = s->K, b = s->b;
    const size_t LN = (size_t)L * N;
    // Chunks of PIEs: the index ciphertexts of chunk i + 1 go up and the results of chunk i - 1 come down while chunk i
    // is evaluated (copy streams beside the caller's; index and result buffers are double-buffered, the work buffers
    // are shared because the evaluations are ordered on one stream).  A chunk is at least ~128 items (enough CTAs for
    // every launch), at most a quarter of the range, and its work buffers - (2 + 2 + 1 + L) polynomials per item with BV keys
    // - stay below ~6 GB.
    const uint32_t n_range = pie_end - pie_begin;
    const size_t limbs_per_item = 5 * (size_t)L + (c->hybrid ? std::max<size_t>((size_t)L * L, (size_t)c->ks_parts * (L + c->Lk)) + 2 * (L + c->Lk) + 2 * L
                                                             : (size_t)L * L);
    const size_t per_pie = (size_t)K * b * limbs_per_item * N * sizeof(u64);
    const uint32_t min_pies = (uint32_t)((128 + (size_t)K * b - 1) / ((size_t)K * b));
    uint32_t chunk = std::max<uint32_t>(min_pies, (n_range + 3) / 4);
    chunk = (uint32_t)std::max<size_t>(1, std::min<size_t>(std::min<uint32_t>(chunk, n_range), (6ull << 30) / per_pie));
    if ((rc = nb_prepare(c, s, chunk))) return rc;
    chunk = std::min(chunk, s->chunk_pies);
    cudaStream_t st = (cudaStream_t)stream;
    const KCtx k = c->k(st);
    s->launches = 0;
    const uint32_t n_chunks = (n_range + chunk - 1) / chunk;
    const size_t ct_chunk = (size_t)s->chunk_pies * K * 2 * LN;  // words of one index / result buffer
    auto upload = [&](uint32_t ci) -> int {
        const uint32_t p0 = pie_begin + ci * chunk, np = std::min(chunk, pie_end - p0), w = ci & 1;
        if (ci >= 2) CK(cudaStreamWaitEvent(s->s_in, s->ev_eval[w], 0));  // the evaluation of chunk ci - 2 has read this buffer
        CK(cudaMemcpyAsync(s->idx.p + w * ct_chunk, idx + (size_t)(p0 - pie_begin) * K * 2 * LN, (size_t)np * K * 2 * LN * sizeof(u64),
                           cudaMemcpyHostToDevice, s->s_in));
        CK(cudaEventRecord(s->ev_in[w], s->s_in));
        return PSI_OK;
    };
    auto evaluate = [&](uint32_t ci) -> int {
        const uint32_t p0 = pie_begin + ci * chunk, np = std::min(chunk, pie_end - p0), G = np * K, B = G * b, w = ci & 1;
        const u64* d_idx = s->idx.p + w * ct_chunk;
        u64* d_out = s->out.p + w * ct_chunk;
        CK(cudaStreamWaitEvent(st, s->ev_in[w], 0));
        if (ci >= 2) CK(cudaStreamWaitEvent(st, s->ev_out[w], 0));  // the results of chunk ci - 2 have left this buffer
        // (1) EvalMult(indexMatrix[hf], vectorizedCT[hf][bin])
        k_nb_mul_ctpt<<<blocks_for((size_t)B * LN), 256, 0, st>>>(k.tab, N, B, d_idx, b, s->pt.p + (size_t)p0 * K * b * LN, LN, s->cur.p);
        CK(cudaGetLastError());
        s->launches++;
        // (2) EvalSum
        int r;
        for (uint32_t step = 0; step < s->n_sum; step++)
            if ((r = nb_ks_step<true>(c, s, k, B, step, 0))) return r;
        // (3) EvalMerge: keep slot 0, move it to slot `bin`
        k_nb_mul_ctpt<<<blocks_for((size_t)B * LN), 256, 0, st>>>(k.tab, N, B, s->cur.p, 1, s->merge_pt.p, 0, s->nxt.p);
        CK(cudaGetLastError());
        s->launches++;
        std::swap(s->cur.p, s->nxt.p);
        std::swap(s->cur.n, s->nxt.n);
        if (b > 1 && (r = nb_ks_step<false>(c, s, k, B, s->n_sum, b))) return r;
        // (4) sum over the bins, EvalMult by preCalcRandomMask[hf]
        k_nb_sum_mask<<<blocks_for((size_t)G * LN), 256, 0, st>>>(k.tab, N, G, b, s->cur.p, s->mask.p + (size_t)p0 * K * LN, d_out);
        CK(cudaGetLastError());
        s->launches++;
        CK(cudaEventRecord(s->ev_eval[w], st));
        return PSI_OK;
    };
    auto download = [&](uint32_t ci) -> int {
        const uint32_t p0 = pie_begin + ci * chunk, np = std::min(chunk, pie_end - p0), w = ci & 1;
        CK(cudaStreamWaitEvent(s->s_out, s->ev_eval[w], 0));
        CK(cudaMemcpyAsync(out + (size_t)(p0 - pie_begin) * K * 2 * LN, s->out.p + w * ct_chunk, (size_t)np * K * 2 * LN * sizeof(u64),
                           cudaMemcpyDeviceToHost, s->s_out));
        CK(cudaEventRecord(s->ev_out[w], s->s_out));
        return PSI_OK;
    };
    if ((rc = upload(0)) || (rc = evaluate(0))) return rc;
    for (uint32_t ci = 0; ci < n_chunks; ci++) {
        if (ci + 1 < n_chunks && ((rc = upload(ci + 1)) || (rc = evaluate(ci + 1)))) return rc;  // queued behind chunk ci
        if ((rc = download(ci))) return rc;
    }
    CK(cudaStreamSynchronize(s->s_out));
    CK(cudaStreamSynchronize(st));
    s->launches /= n_chunks;  // per chunk: what one collection-sized launch set costs
    return PSI_OK;
}

int psi_nb_launch_count(psi_ctx* c, uint32_t* out) {
    if (!c || !out) return set_error(PSI_ERR_INVALID, "null argument");
    *out = c->nb ? c->nb->launches : 0;
    return PSI_OK;
}

}  // extern "C"
