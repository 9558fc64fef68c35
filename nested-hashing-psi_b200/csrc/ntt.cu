// Batched negacyclic NTT / INTT over 64-bit prime moduli for sm_100a.
//
// Replaces OpenFHE's ChineseRemainderTransformFTT (ForwardTransformToBitReverse /
// InverseTransformFromBitReverse) that DCRTPoly::SetFormat runs inside EvalMult(ct,ct)
// (/root/reference/.../BatchedFHEHIPPIE.cpp:123) and inside the lazy SetFormat(EVALUATION) of every
// plaintext (BatchedFHEHIPPIE.cpp:108,113,126).
//
// Forward: Cooley-Tukey, natural order in -> bit-reversed order out, twiddles psi^bitrev(i).
// Inverse: Gentleman-Sande, bit-reversed in -> natural out, scaled by N^-1.
// Harvey lazy butterflies with Shoup twiddles: values live in [0,4q) (forward) / [0,2q) (inverse)
// between stages, outputs are canonical.
//
// One CTA owns one limb-polynomial, which stays in shared memory (N*8 bytes <= 128 KiB) for all
// log2(N) stages, so each coefficient crosses HBM/L2 exactly once per direction.  Stages are
// executed in register passes of up to RADIX_LOG stages: a thread gathers 2^r coefficients whose
// butterflies are closed under those r stages, runs them out of registers, and scatters them back.
#include "psi_kernels.cuh"

namespace psi {

constexpr int kNttThreads = 512;
constexpr int kRadixLog = 3;  // 8 coefficients per thread per pass

// physical shared-memory slot of logical coefficient i: one pad word per 32 coefficients keeps the
// stride-2^k gathers of the late (forward) / early (inverse) passes off a single bank group
__device__ __forceinline__ uint32_t sidx(uint32_t i) { return i + (i >> 5); }

template <int R>
__device__ __forceinline__ void fwd_pass(u64* __restrict__ sm, const ModDev& md, uint32_t logN, uint32_t s0,
                                         uint32_t tid, uint32_t nthreads) {
    // stages s0 .. s0+R-1; stage s has m = 2^s groups and butterfly stride t = N >> (s+1).
    // A block of 2^R coefficients {base + k * tl}, tl = N >> (s0+R), is closed under these stages.
    const uint32_t N = 1u << logN;
    const uint32_t tl = N >> (s0 + R);
    const u64 q = md.q, q2 = 2 * q;
    for (uint32_t blk = tid; blk < (N >> R); blk += nthreads) {
        // blk -> (group index at stage s0, offset inside the stride)
        const uint32_t off = blk & (tl - 1);
        const uint32_t grp = blk / tl;  // in [0, 2^s0)
        const uint32_t base = (grp << (logN - s0)) + off;
        u64 v[1 << R];
#pragma unroll
        for (int k = 0; k < (1 << R); k++) v[k] = sm[sidx(base + k * tl)];
#pragma unroll
        for (int r = 0; r < R; r++) {
            // inside the block, stage s0+r pairs k with k + (2^(R-1-r)); sub-group j = k >> (R-r)
            const int half = 1 << (R - 1 - r);
#pragma unroll
            for (int j = 0; j < (1 << r); j++) {
                const uint32_t widx = (1u << (s0 + r)) + (grp << r) + j;
                const u64 w = md.w[widx], ws = md.ws[widx];
#pragma unroll
                for (int i = 0; i < half; i++) {
                    const int a = j * 2 * half + i, b = a + half;
                    u64 u = v[a];
                    if (u >= q2) u -= q2;
                    const u64 x = mul_shoup_lazy(v[b], w, ws, q);
                    v[a] = u + x;
                    v[b] = u - x + q2;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < (1 << R); k++) sm[sidx(base + k * tl)] = v[k];
    }
}

template <int R>
__device__ __forceinline__ void inv_pass(u64* __restrict__ sm, const ModDev& md, uint32_t logN, uint32_t s0,
                                         uint32_t tid, uint32_t nthreads) {
    // Gentleman-Sande stages in forward-stage numbering, executed from s0+R-1 down to s0: stage s
    // has h = 2^s twiddles iw[h + i] and stride t = N >> (s+1).
    const uint32_t N = 1u << logN;
    const uint32_t tl = N >> (s0 + R);
    const u64 q = md.q, q2 = 2 * q;
    for (uint32_t blk = tid; blk < (N >> R); blk += nthreads) {
        const uint32_t off = blk & (tl - 1);
        const uint32_t grp = blk / tl;
        const uint32_t base = (grp << (logN - s0)) + off;
        u64 v[1 << R];
#pragma unroll
        for (int k = 0; k < (1 << R); k++) v[k] = sm[sidx(base + k * tl)];
#pragma unroll
        for (int r = R - 1; r >= 0; r--) {
            const int half = 1 << (R - 1 - r);
#pragma unroll
            for (int j = 0; j < (1 << r); j++) {
                const uint32_t widx = (1u << (s0 + r)) + (grp << r) + j;
                const u64 w = md.iw[widx], ws = md.iws[widx];
#pragma unroll
                for (int i = 0; i < half; i++) {
                    const int a = j * 2 * half + i, b = a + half;
                    const u64 u = v[a], x = v[b];
                    u64 s = u + x;
                    if (s >= q2) s -= q2;
                    v[a] = s;
                    v[b] = mul_shoup_lazy(u - x + q2, w, ws, q);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < (1 << R); k++) sm[sidx(base + k * tl)] = v[k];
    }
}

template <bool INV>
__global__ void __launch_bounds__(kNttThreads) k_ntt(const DevTables* __restrict__ tab, uint32_t logN, NttBatch b) {
    extern __shared__ u64 sm[];
    const uint32_t N = 1u << logN;
    const uint32_t i = blockIdx.x;
    const uint32_t g = i / b.G, l = i % b.G;
    const u64* src = b.src + (size_t)g * b.src_gs + (size_t)l * b.src_ls;
    u64* dst = b.dst + (size_t)g * b.dst_gs + (size_t)l * N;
    const ModDev& md = tab->mods[b.mod_base + (l % b.mod_period)];
    const uint32_t tid = threadIdx.x, nt = blockDim.x;

    for (uint32_t j = tid; j < N; j += nt) sm[sidx(j)] = src[j];
    __syncthreads();

    if (!INV) {
        uint32_t s = 0;
        while (s + kRadixLog <= logN) {
            fwd_pass<kRadixLog>(sm, md, logN, s, tid, nt);
            __syncthreads();
            s += kRadixLog;
        }
        if (logN - s == 2) {
            fwd_pass<2>(sm, md, logN, s, tid, nt);
            __syncthreads();
        } else if (logN - s == 1) {
            fwd_pass<1>(sm, md, logN, s, tid, nt);
            __syncthreads();
        }
        const u64 q = md.q, q2 = 2 * q;
        for (uint32_t j = tid; j < N; j += nt) {
            u64 v = sm[sidx(j)];
            if (v >= q2) v -= q2;
            if (v >= q) v -= q;
            dst[j] = v;
        }
    } else {
        // stages logN-1 .. 0; the remainder (logN mod R) is done first so that all passes align
        uint32_t s = logN;
        const uint32_t rem = logN % kRadixLog;
        if (rem == 2) {
            s -= 2;
            inv_pass<2>(sm, md, logN, s, tid, nt);
            __syncthreads();
        } else if (rem == 1) {
            s -= 1;
            inv_pass<1>(sm, md, logN, s, tid, nt);
            __syncthreads();
        }
        while (s >= (uint32_t)kRadixLog) {
            s -= kRadixLog;
            inv_pass<kRadixLog>(sm, md, logN, s, tid, nt);
            __syncthreads();
        }
        const u64 q = md.q, ninv = md.ninv, ninv_s = md.ninv_s;
        for (uint32_t j = tid; j < N; j += nt) dst[j] = mul_shoup(sm[sidx(j)], ninv, ninv_s, q);
    }
}

cudaError_t ntt_init_device() {
    cudaError_t e = cudaFuncSetAttribute(k_ntt<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_ntt<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

cudaError_t launch_ntt(const KCtx& k, const NttBatch& b, bool inverse) {
    if (b.n_polys == 0) return cudaSuccess;
    const size_t smem = ((size_t)k.N + (k.N >> 5) + 1) * sizeof(u64);
    const int threads = k.N >= 4096 ? kNttThreads : (k.N >= 1024 ? 128 : 32);
    if (inverse)
        k_ntt<true><<<b.n_polys, threads, smem, k.s>>>(k.tab, k.logN, b);
    else
        k_ntt<false><<<b.n_polys, threads, smem, k.s>>>(k.tab, k.logN, b);
    return cudaGetLastError();
}

}  // namespace psi
