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
    u64 qinv;          // q^-1 mod 2^64 (Montgomery reduction, radix R = 2^64)
    u64 ninvR, ninvR_s;  // N^-1 * R mod q: undoes the R^-1 left by a Montgomery-reduced tensor product
    u64 Rmodq, Rmodq_s;  // R mod q
    const ulonglong2* ftw;  // {w, ws} interleaved: one 16-byte load per twiddle (fused kernels)
    const ulonglong2* itw;  // {iw, iws}
    // per row tile (8 rows of 128 coefficients): the 1016 twiddles of the 7 row stages packed in the order
    // the fused row kernels read them (one 16 KiB bulk copy per tile, bank-conflict-free reads), or null
    const ulonglong2* ftw_rows;
    const ulonglong2* itw_rows;
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

// ---- hand-scheduled variants for the NTT inner loop -------------------------------------------
// The multiplier pipe (IMAD on fmaheavy; IMAD.WIDE at a quarter of the FP32 rate, measured) is the NTT's roofline, so the
// butterfly is written to spend exactly 10 multiply issues and as few ALU issues as possible.
__device__ __forceinline__ u64 madw32(uint32_t a, uint32_t b, u64 c) { return c + (u64)a * b; }  // IMAD.WIDE.U32
__device__ __forceinline__ uint32_t lo32(u64 x) { return (uint32_t)x; }
__device__ __forceinline__ uint32_t hi32(u64 x) { return (uint32_t)(x >> 32); }

// floor(x * y / 2^64) with four IMAD.WIDE and one 64-bit add
__device__ __forceinline__ u64 mulhi64_4(u64 x, u64 y) {
    const uint32_t x0 = lo32(x), x1 = hi32(x), y0 = lo32(y), y1 = hi32(y);
    const u64 p00 = madw32(x0, y0, 0);
    const u64 t = madw32(x0, y1, (u64)hi32(p00));
    const u64 t2 = madw32(x1, y0, (u64)lo32(t));
    return madw32(x1, y1, (u64)hi32(t)) + (u64)hi32(t2);
}
// x * w - floor(x * ws / 2^64) * q  (mod 2^64), in [0, 2q) for any x < 2^64; nq = 2^64 - q.
// Low 64 bits of x*w + h*nq: two IMAD.WIDE for the low words, four 32-bit IMADs into the high word.
__device__ __forceinline__ u64 mul_shoup_lazy_nq(u64 x, u64 w, u64 ws, u64 nq) {
    const u64 h = mulhi64_4(x, ws);
    const uint32_t x0 = lo32(x), x1 = hi32(x), w0 = lo32(w), w1 = hi32(w);
    const uint32_t h0 = lo32(h), h1 = hi32(h), n0 = lo32(nq), n1 = hi32(nq);
    u64 r = madw32(x0, w0, 0);
    r = madw32(h0, n0, r);
    uint32_t top = hi32(r);
    top += x0 * w1;
    top += x1 * w0;
    top += h0 * n1;
    top += h1 * n0;
    return ((u64)top << 32) | lo32(r);
}
// Approximate lazy correction: subtracts q2 only when the HIGH WORD already proves x >= q2, so it can
// never underflow; afterwards x < q2 + 2^32.  One 32-bit compare + a predicated 64-bit subtract.
__device__ __forceinline__ u64 lazy_sub_hi(u64 x, u64 q2) {
    uint32_t xl = lo32(x), xh = hi32(x);
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "setp.gt.u32 p, %1, %3;\n\t"
        "@p sub.cc.u32 %0, %0, %2;\n\t"
        "@p subc.u32 %1, %1, %3;\n\t"
        "}"
        : "+r"(xl), "+r"(xh)
        : "r"(lo32(q2)), "r"(hi32(q2)));
    return ((u64)xh << 32) | xl;
}

// Whole forward (Cooley-Tukey, Harvey lazy) butterfly in one PTX block: x' = x~ + v, y' = x~ - v + 2q with
// x~ = lazy_sub_hi(x), v = y * w mod q in [0, 2q).  Everything on 32-bit halves so that ptxas sees the
// carry chains explicitly.
__device__ __forceinline__ void ct_butterfly_asm(u64& x, u64& y, u64 w, u64 ws, u64 q2, u64 nq) {
    asm("{\n\t"
        ".reg .u32 x0,x1,y0,y1,w0,w1,s0,s1,a0,a1,n0,n1,h0,h1,t0,t1,b0,b1,v0,v1,u0,u1,z;\n\t"
        ".reg .u64 p,t,t2,h,r;\n\t"
        ".reg .pred pg;\n\t"
        "mov.b64 {x0,x1}, %0;\n\t"
        "mov.b64 {y0,y1}, %1;\n\t"
        "mov.b64 {w0,w1}, %2;\n\t"
        "mov.b64 {s0,s1}, %3;\n\t"
        "mov.b64 {a0,a1}, %4;\n\t"
        "mov.b64 {n0,n1}, %5;\n\t"
        "mov.u32 z, 0;\n\t"
        "setp.gt.u32 pg, x1, a1;\n\t"
        "@pg sub.cc.u32 x0, x0, a0;\n\t"
        "@pg subc.u32 x1, x1, a1;\n\t"
        "mul.hi.u32 t1, y0, s0;\n\t"
        "mov.b64 t, {t1, z};\n\t"
        "mad.wide.u32 t, y0, s1, t;\n\t"
        "mov.b64 {t0,t1}, t;\n\t"
        "mov.b64 t2, {t0, z};\n\t"
        "mad.wide.u32 t2, y1, s0, t2;\n\t"
        "mov.b64 {b0,b1}, t2;\n\t"
        "mov.b64 h, {t1, z};\n\t"
        "mad.wide.u32 h, y1, s1, h;\n\t"
        "mov.b64 {h0,h1}, h;\n\t"
        "add.cc.u32 h0, h0, b1;\n\t"
        "addc.u32 h1, h1, 0;\n\t"
        "mul.wide.u32 r, y0, w0;\n\t"
        "mad.wide.u32 r, h0, n0, r;\n\t"
        "mov.b64 {v0,v1}, r;\n\t"
        "mad.lo.u32 v1, y0, w1, v1;\n\t"
        "mad.lo.u32 v1, y1, w0, v1;\n\t"
        "mad.lo.u32 v1, h0, n1, v1;\n\t"
        "mad.lo.u32 v1, h1, n0, v1;\n\t"
        "add.cc.u32 u0, x0, v0;\n\t"
        "addc.u32 u1, x1, v1;\n\t"
        "sub.cc.u32 t0, x0, v0;\n\t"
        "subc.u32 t1, x1, v1;\n\t"
        "add.cc.u32 t0, t0, a0;\n\t"
        "addc.u32 t1, t1, a1;\n\t"
        "mov.b64 %0, {u0,u1};\n\t"
        "mov.b64 %1, {t0,t1};\n\t"
        "}"
        : "+l"(x), "+l"(y)
        : "l"(w), "l"(ws), "l"(q2), "l"(nq));
}

// Montgomery reduction (radix 2^64, subtractive form): for T = (hi:lo) < q * 2^64 returns a value in
// [0, 2q) congruent to T * 2^-64 mod q.  m = lo * q^-1 makes T - m*q divisible by 2^64 exactly, so the
// quotient is hi - floor(m*q / 2^64) in (-q, q); adding q makes it non-negative.
__device__ __forceinline__ u64 mont_redc_lazy(u64 hi, u64 lo, u64 q, u64 qinv) {
    const u64 m = lo * qinv;
    return hi - mulhi64(m, q) + q;
}

// 128-bit multiply-accumulate: (hi:lo) += a * b
__device__ __forceinline__ void mac128(u64& hi, u64& lo, u64 a, u64 b) {
    // written on unsigned __int128: the compiler emits four IMAD.WIDE with the carries chained through
    // their addends (about 8 issue slots) instead of separate mul.lo / mul.hi / compare / add (about 13)
    u128 t = ((u128)hi << 64) | lo;
    t += (u128)a * b;
    lo = (u64)t;
    hi = (u64)(t >> 64);
}
// full 64 x 64 -> 128 product
__device__ __forceinline__ void mul128(u64& hi, u64& lo, u64 a, u64 b) {
    const u128 t = (u128)a * b;
    lo = (u64)t;
    hi = (u64)(t >> 64);
}

}  // namespace psi
