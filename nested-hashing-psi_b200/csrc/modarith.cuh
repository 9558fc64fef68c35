// 64-bit modular arithmetic on the sm_100a integer pipes (IMAD.WIDE / IADD3); no tensor cores:
// the BFV-RNS path is residue arithmetic modulo ~60-bit primes, not a floating-point contraction.
#pragma once
#include <cstdint>

namespace psi {

typedef unsigned long long u64;
typedef unsigned __int128 u128;

// Per-modulus constants + twiddle tables, resident in global memory (L2-hot).
struct ModDev {
    u64 q;
    u64 mu_hi, mu_lo;  // floor(2^128 / q)
    u64 ninv, ninv_s;  // N^-1 mod q and its Shoup companion
    const u64* w;      // psi^bitrev(i)            (forward, Cooley-Tukey)
    const u64* ws;     // floor(w * 2^64 / q)
    const u64* iw;     // psi^-bitrev(i)           (inverse, Gentleman-Sande)
    const u64* iws;
};

__device__ __forceinline__ u64 mulhi64(u64 a, u64 b) { return __umul64hi(a, b); }

// x * w mod q in [0, 2q) for ANY x < 2^64; ws = floor(w * 2^64 / q)  (Shoup / Harvey)
__device__ __forceinline__ u64 mul_shoup_lazy(u64 x, u64 w, u64 ws, u64 q) {
    u64 h = mulhi64(x, ws);
    return x * w - h * q;
}
__device__ __forceinline__ u64 mul_shoup(u64 x, u64 w, u64 ws, u64 q) {
    u64 r = mul_shoup_lazy(x, w, ws, q);
    return r >= q ? r - q : r;
}

// Canonical residue of a 128-bit value (hi:lo) < 2^128 modulo q: the quotient estimate is
// floor(x * mu / 2^128) evaluated modulo 2^64 (sufficient because the true remainder is < 3q).
__device__ __forceinline__ u64 barrett128(u64 hi, u64 lo, u64 q, u64 mu_hi, u64 mu_lo) {
    u64 left_hi = mulhi64(lo, mu_lo);
    u64 m1_lo = lo * mu_hi, m1_hi = mulhi64(lo, mu_hi);
    u64 m2_lo = hi * mu_lo, m2_hi = mulhi64(hi, mu_lo);
    u64 s = m1_lo + left_hi;
    u64 c1 = s < m1_lo;
    u64 s2 = s + m2_lo;
    u64 c2 = s2 < s;
    u64 qhat = hi * mu_hi + m1_hi + m2_hi + c1 + c2;
    u64 r = lo - qhat * q;
    if (r >= q) r -= q;
    if (r >= q) r -= q;
    if (r >= q) r -= q;
    return r;
}
__device__ __forceinline__ u64 barrett128(u128 x, const ModDev& m) {
    return barrett128((u64)(x >> 64), (u64)x, m.q, m.mu_hi, m.mu_lo);
}
__device__ __forceinline__ u64 mulmod(u64 a, u64 b, const ModDev& m) {
    return barrett128(mulhi64(a, b), a * b, m.q, m.mu_hi, m.mu_lo);
}
__device__ __forceinline__ u64 addmod(u64 a, u64 b, u64 q) {
    u64 r = a + b;
    return r >= q ? r - q : r;
}
__device__ __forceinline__ u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }

// 128-bit multiply-accumulate: (hi:lo) += a * b
__device__ __forceinline__ void mac128(u64& hi, u64& lo, u64 a, u64 b) {
    u64 pl = a * b, ph = mulhi64(a, b);
    lo += pl;
    hi += ph + (lo < pl);
}

}  // namespace psi
