// psi_params_generate — stand-alone BFV-RNS context tables for the B200 BatchedFHEPIE path.
//
// In the reference the context is created by the CLIENT through OpenFHE
// (/root/reference/src/Client/FHE/BatchedFHEPSIClient.cpp:22-78: plaintext modulus by bit size,
// depth by eachCuckooTableSize, ring dimension 16384, HEStd_128_classic, library defaults for
// everything else) and reaches the server serialised (BatchedFHEPSIServer.cpp:21-54).  When this
// library is linked next to OpenFHE the adapter copies the tables out of CryptoParametersBFVRNS
// (INTEGRATION.md); this file builds the same tables from (N, t, depth) so the path also runs
// without OpenFHE.  OpenFHE 1.0.x defaults as recalled: 60-bit moduli, largest primes below 2^60
// congruent to 1 mod 2N in descending order, minimal primitive roots, HPSPOVERQ with
// |P| = |Q| limbs continuing below Q, BV key switching with digit size 0.
//
// All tables are residues of products/inverses modulo one word-sized prime, so no multi-precision
// arithmetic is needed (the exact big-integer definitions live in oracle/params_ref.py and
// tests/test_params.py compares the two generators table by table).
#include <cmath>
#include <cstring>
#include <vector>

#include "psi_b200.h"
#include "psi_host_internal.hpp"

namespace psi {

typedef unsigned __int128 u128;

static inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t m) { return (uint64_t)(((u128)a * b) % m); }

static uint64_t powmod(uint64_t a, uint64_t e, uint64_t m) {
    uint64_t r = 1 % m;
    a %= m;
    for (; e; e >>= 1) {
        if (e & 1) r = mulmod(r, a, m);
        a = mulmod(a, a, m);
    }
    return r;
}

static inline uint64_t invmod_prime(uint64_t a, uint64_t p) { return powmod(a % p, p - 2, p); }

bool is_prime_u64(uint64_t n) {
    static const uint64_t bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return false;
    for (uint64_t b : bases) {
        if (n % b == 0) return n == b;
    }
    uint64_t d = n - 1;
    int s = 0;
    while ((d & 1) == 0) {
        d >>= 1;
        ++s;
    }
    for (uint64_t a : bases) {
        uint64_t x = powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool composite = true;
        for (int r = 1; r < s; ++r) {
            x = mulmod(x, x, n);
            if (x == n - 1) {
                composite = false;
                break;
            }
        }
        if (composite) return false;
    }
    return true;
}

// first prime >= 2^bits that is 1 mod m, then step down: the largest such prime below it
static uint64_t first_prime(unsigned bits, uint64_t m) {
    uint64_t x = 1ull << bits;
    uint64_t r = x % m;
    uint64_t q = r ? x + (m - r) + 1 : x + 1;
    while (!is_prime_u64(q)) q += m;
    return q;
}
static uint64_t previous_prime(uint64_t q, uint64_t m) {
    do {
        q -= m;
    } while (!is_prime_u64(q));
    return q;
}

// minimum primitive m-th root of unity modulo prime q (m a power of two dividing q-1)
uint64_t min_primitive_root(uint64_t m, uint64_t q) {
    uint64_t r = 0;
    for (uint64_t x = 2;; ++x) {
        r = powmod(x, (q - 1) / m, q);
        if (powmod(r, m / 2, q) == q - 1) break;
    }
    uint64_t r2 = mulmod(r, r, q), best = r, cur = r;
    for (uint64_t i = 1; i < m / 2; ++i) {
        cur = mulmod(cur, r2, q);
        if (cur < best) best = cur;
    }
    return best;
}

// sizeQ from the BFVrns noise estimate (EvalMult-only branch, BV digit size 0), 60-bit moduli
static uint32_t derive_size_q(uint32_t N, uint64_t t, uint32_t depth) {
    const double dcrt_bits = 60.0, sigma = 3.19, alpha = 36.0;
    const double p = (double)t, Berr = sigma * std::sqrt(alpha), Bkey = 1.0;
    const double delta = 2.0 * std::sqrt((double)N);
    const double Vnorm = Berr * (1.0 + 2.0 * delta * Bkey);
    const double w = std::pow(2.0, dcrt_bits);
    const double C1 = delta * delta * p * Bkey;
    auto logq_bfv = [&](double logq_prev) {
        double noise_ks = delta * (std::floor(logq_prev / (std::log(2.0) * dcrt_bits)) + 1) * w * Berr;
        double C2 = delta * delta * Bkey * Bkey / 2.0 + noise_ks;
        return std::log(4 * p) + (depth - 1) * std::log(C1) + std::log(C1 * Vnorm + depth * C2);
    };
    double logq = logq_bfv(6.0 * std::log(10.0));
    logq = logq_bfv(logq);
    return (uint32_t)std::ceil((std::ceil(logq / std::log(2.0)) + 1.0) / dcrt_bits);
}

// Every table of psi_params that is a function of the moduli alone (all of them except the moduli and roots).
// Requires N, L, Lp, t, q[], p[] set.  The double tables follow OpenFHE's construction as recalled:
// qInv = 1.0 / (double) q_i, frac = (double) remainder / (double) p_i.
static void fill_tables(psi_params* out) {
    const uint32_t L = out->L, Lp = out->Lp;
    const uint64_t t = out->t;
    const uint64_t* q = out->q;
    const uint64_t* p = out->p;
    auto prod_mod = [](const uint64_t* v, uint32_t n, int skip, uint64_t mod) {
        uint64_t r = 1 % mod;
        for (uint32_t i = 0; i < n; ++i)
            if ((int)i != skip) r = mulmod(r, v[i] % mod, mod);
        return r;
    };
    for (uint32_t i = 0; i < L; ++i) {
        uint64_t qhat = prod_mod(q, L, (int)i, q[i]);            // (Q/q_i) mod q_i
        out->QHatInvModq[i] = invmod_prime(qhat, q[i]);
        out->qInv[i] = 1.0 / (double)q[i];
        uint64_t Pmodqi = prod_mod(p, Lp, -1, q[i]);
        out->negPQHatInvModq[i] = (q[i] - mulmod(Pmodqi, out->QHatInvModq[i], q[i])) % q[i];
        for (uint32_t j = 0; j < Lp; ++j) {
            out->qInvModp[i][j] = invmod_prime(q[i], p[j]);
            out->PHatModq[i][j] = prod_mod(p, Lp, (int)j, q[i]); // (P/p_j) mod q_i
        }
    }
    for (uint32_t j = 0; j < Lp; ++j) {
        uint64_t phat = prod_mod(p, Lp, (int)j, p[j]);
        out->PHatInvModp[j] = invmod_prime(phat, p[j]);
        out->pInv[j] = 1.0 / (double)p[j];
        for (uint32_t i = 0; i < L; ++i) out->QHatModp[j][i] = prod_mod(q, L, (int)i, p[j]);
        uint64_t Qmodpj = prod_mod(q, L, -1, p[j]);
        for (uint32_t a = 0; a <= L; ++a) out->alphaQModp[a][j] = mulmod(a, Qmodpj, p[j]);
    }
    for (uint32_t i = 0; i < L; ++i) {
        uint64_t Pmodqi = prod_mod(p, Lp, -1, q[i]);
        for (uint32_t a = 0; a <= Lp; ++a) out->alphaPModq[a][i] = mulmod(a, Pmodqi, q[i]);
    }
    // ScaleAndRound by t/P, output Q.  With S = Q*P and c_i = t*Q*[(S/p_i)^-1]_{p_i}:
    //   frac_i = (c_i mod p_i)/p_i,   floor(c_i/p_i) = -(c_i mod p_i) * p_i^-1  (mod q_j)   [c_i = 0 mod q_j]
    for (uint32_t i = 0; i < Lp; ++i) {
        uint64_t shat = mulmod(prod_mod(q, L, -1, p[i]), prod_mod(p, Lp, (int)i, p[i]), p[i]); // (S/p_i) mod p_i
        uint64_t inv = invmod_prime(shat, p[i]);
        uint64_t rem = mulmod(mulmod(t % p[i], prod_mod(q, L, -1, p[i]), p[i]), inv, p[i]);
        out->tQSHatInvModsDivsFrac[i] = (double)rem / (double)p[i];
        for (uint32_t j = 0; j < L; ++j) {
            uint64_t v = mulmod(rem % q[j], invmod_prime(p[i], q[j]), q[j]);
            out->tQSHatInvModsDivsModq[j][i] = (q[j] - v) % q[j];
        }
    }
    for (uint32_t j = 0; j < L; ++j) {
        uint64_t qhat = prod_mod(q, L, (int)j, q[j]);                       // (Q/q_j) mod q_j
        uint64_t shat = mulmod(qhat, prod_mod(p, Lp, -1, q[j]), q[j]);      // (S/q_j) mod q_j
        out->tQSHatInvModsDivsModq[j][Lp] = mulmod(mulmod(t % q[j], qhat, q[j]), invmod_prime(shat, q[j]), q[j]);
    }
    // HPS: ScaleAndRound by t/Q, output P.  c_i = t*P*[(S/q_i)^-1]_{q_i}: frac_i = (c_i mod q_i)/q_i and
    // floor(c_i/q_i) = -(c_i mod q_i) * q_i^-1 (mod p_j) because c_i = 0 mod p_j; own P limb: t*(P/p_j)*[(S/p_j)^-1]_{p_j}
    for (uint32_t i = 0; i < L; ++i) {
        uint64_t shat = mulmod(prod_mod(q, L, (int)i, q[i]), prod_mod(p, Lp, -1, q[i]), q[i]);  // (S/q_i) mod q_i
        uint64_t rem = mulmod(mulmod(t % q[i], prod_mod(p, Lp, -1, q[i]), q[i]), invmod_prime(shat, q[i]), q[i]);
        out->tPSHatInvModsDivsFrac[i] = (double)rem / (double)q[i];
        for (uint32_t j = 0; j < Lp; ++j) {
            uint64_t v = mulmod(rem % p[j], invmod_prime(q[i], p[j]), p[j]);
            out->tPSHatInvModsDivsModp[j][i] = (p[j] - v) % p[j];
        }
    }
    for (uint32_t j = 0; j < Lp; ++j) {
        uint64_t phat = prod_mod(p, Lp, (int)j, p[j]);                      // (P/p_j) mod p_j
        uint64_t shat = mulmod(phat, prod_mod(q, L, -1, p[j]), p[j]);       // (S/p_j) mod p_j
        out->tPSHatInvModsDivsModp[j][L] = mulmod(mulmod(t % p[j], phat, p[j]), invmod_prime(shat, p[j]), p[j]);
    }
}


int generate_params(uint32_t N, uint64_t t, uint32_t depth, uint32_t L_override, psi_params* out, uint32_t mult_technique,
                    uint32_t ks_technique, uint32_t fp_contract) {
    if (mult_technique > PSI_MULT_HPSPOVERQ || ks_technique > PSI_KS_HYBRID || fp_contract > PSI_FP_FMA)
        return set_error(PSI_ERR_INVALID, "unknown multiplication / key-switching technique or fp mode");
    if (!out) return PSI_ERR_INVALID;
    if (N < 8 || (N & (N - 1)) != 0 || N > 65536) return set_error(PSI_ERR_INVALID, "ring dimension must be a power of two");
    const uint64_t m = 2ull * N;
    if (t < 2 || (t - 1) % m != 0 || !is_prime_u64(t))
        return set_error(PSI_ERR_INVALID, "plaintext modulus must be a prime congruent to 1 mod 2N");
    uint32_t L = L_override ? L_override : derive_size_q(N, t, depth ? depth : 1);
    if (L < 1 || L > PSI_MAX_LIMBS) return set_error(PSI_ERR_INVALID, "sizeQ out of range (1..PSI_MAX_LIMBS)");
    std::memset(out, 0, sizeof(*out));
    // HPSPOVERQ: sizeP = sizeQ; HPS: one limb more (the tensor of two centred lifts needs Q*P > N * Q^2 / 2)
    const uint32_t Lp = mult_technique == PSI_MULT_HPS ? L + 1 : L;
    if (Lp > PSI_MAX_LIMBS) return set_error(PSI_ERR_INVALID, "sizeP out of range (HPS needs sizeQ + 1 <= PSI_MAX_LIMBS)");
    out->N = N;
    out->L = L;
    out->Lp = Lp;
    out->mult_technique = mult_technique;
    out->ks_technique = ks_technique;
    out->fp_contract = fp_contract;
    out->t = t;
    uint64_t* q = out->q;
    uint64_t* p = out->p;
    q[0] = previous_prime(first_prime(60, m), m);
    for (uint32_t i = 1; i < L; ++i) q[i] = previous_prime(q[i - 1], m);
    p[0] = previous_prime(q[L - 1], m);
    for (uint32_t j = 1; j < Lp; ++j) p[j] = previous_prime(p[j - 1], m);
    for (uint32_t i = 0; i < L; ++i) out->psi_q[i] = min_primitive_root(m, q[i]);
    for (uint32_t j = 0; j < Lp; ++j) out->psi_p[j] = min_primitive_root(m, p[j]);
    out->psi_t = min_primitive_root(m, t);
    if (ks_technique == PSI_KS_HYBRID) {
        // OpenFHE ComputeNumLargeDigits (recalled): 3 digits for depth > 3, 2 for depth > 0, else 1; never more than
        // sizeQ; sizeP of the key-switching basis = limbs of the largest digit; special primes continue below
        uint32_t parts = depth > 3 ? 3 : (depth > 0 ? 2 : 1);
        if (parts > L) parts = L;
        const uint32_t alpha = (L + parts - 1) / parts;
        out->ks_num_parts = (L + alpha - 1) / alpha;
        out->Lk = alpha;
        uint64_t prev = p[Lp - 1];
        for (uint32_t i = 0; i < alpha; ++i) {
            prev = previous_prime(prev, m);
            out->pk[i] = prev;
            out->psi_pk[i] = min_primitive_root(m, prev);
        }
    }
    fill_tables(out);
    return PSI_OK;
}


}  // namespace psi

// psi_params for a context whose moduli and roots come from the host library (the adapter reads them from
// OpenFHE's ILDCRTParams): same tables as psi_params_generate, no choice of primes involved.
extern "C" int psi_params_from_moduli(uint32_t N, uint64_t t, uint32_t L, const uint64_t* q, const uint64_t* psi_q, uint32_t Lp,
                                      const uint64_t* p, const uint64_t* psi_p, uint64_t psi_t, psi_params* out) {
    using namespace psi;
    if (!out || !q || !psi_q || !p || !psi_p) return set_error(PSI_ERR_INVALID, "null argument");
    if (N < 8 || (N & (N - 1)) != 0 || N > 65536) return set_error(PSI_ERR_INVALID, "ring dimension must be a power of two");
    if (L < 1 || L > PSI_MAX_LIMBS || Lp < 1 || Lp > PSI_MAX_LIMBS) return set_error(PSI_ERR_INVALID, "sizeQ / sizeP out of range");
    const uint64_t m = 2ull * N;
    std::memset(out, 0, sizeof(*out));
    out->N = N;
    out->L = L;
    out->Lp = Lp;
    out->mult_technique = PSI_MULT_HPSPOVERQ;
    out->ks_technique = PSI_KS_BV;
    out->t = t;
    out->psi_t = psi_t;
    for (uint32_t i = 0; i < L + Lp; i++) {
        const uint64_t mod = i < L ? q[i] : p[i - L], root = i < L ? psi_q[i] : psi_p[i - L];
        if (mod >= (1ull << 60) || (mod - 1) % m != 0 || !is_prime_u64(mod))
            return set_error(PSI_ERR_INVALID, "moduli must be primes below 2^60 congruent to 1 mod 2N");
        for (uint32_t j = 0; j < i; j++)
            if ((j < L ? q[j] : p[j - L]) == mod) return set_error(PSI_ERR_INVALID, "moduli must be pairwise distinct");
        if (powmod(root, N, mod) != mod - 1) return set_error(PSI_ERR_INVALID, "root is not a primitive 2N-th root of unity");
        if (i < L) {
            out->q[i] = mod;
            out->psi_q[i] = root;
        } else {
            out->p[i - L] = mod;
            out->psi_p[i - L] = root;
        }
    }
    if (t < 2 || (t - 1) % m != 0 || powmod(psi_t, N, t) != t - 1)
        return set_error(PSI_ERR_INVALID, "plaintext modulus / root: t = 1 mod 2N and psi_t a primitive 2N-th root needed");
    fill_tables(out);
    return PSI_OK;
}

extern "C" int psi_params_generate(uint32_t N, uint64_t t, uint32_t mult_depth, uint32_t L_override, psi_params* out) {
    return psi::generate_params(N, t, mult_depth ? mult_depth : 1, L_override, out, PSI_MULT_HPSPOVERQ, PSI_KS_BV, PSI_FP_SEPARATE);
}
extern "C" int psi_params_generate_ex(uint32_t N, uint64_t t, uint32_t mult_depth, uint32_t L_override, uint32_t mult_technique,
                                      uint32_t ks_technique, uint32_t fp_contract, psi_params* out) {
    return psi::generate_params(N, t, mult_depth ? mult_depth : 1, L_override, out, mult_technique, ks_technique, fp_contract);
}
