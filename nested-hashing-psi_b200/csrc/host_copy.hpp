// Host-side copy of limb vectors between pageable memory and the pinned staging pools (psi_*_limbs entry points).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace psi {

// PSI_PLAIN_MEMCPY=1 (environment): A/B switch back to memcpy
inline bool plain_memcpy_forced() {
    static const bool on = std::getenv("PSI_PLAIN_MEMCPY") != nullptr;
    return on;
}

// One limb vector with non-temporal stores: the destination is written once and read next by the copy engine (gather)
// or much later by the caller (scatter), so write-allocate traffic (a read of every destination line before it is
// overwritten) is pure waste.  MEASURED on the GPU box's host (16 cores, config B, psi_query_run_streamed_limbs): the
// 760 query vectors are in the pool after 2.09-2.19 ms instead of 2.69-2.72, the whole query takes 3.68-3.76 ms
// instead of 4.35-4.48 (profiles/r02_limb_vectors.md).  Falls back to memcpy for a destination that is not 16-byte
// aligned.
inline void copy_limb_vector(void* dst, const void* src, size_t words) {
#if defined(__SSE2__)
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && (words & 7u) == 0 && !plain_memcpy_forced()) {
        const __m128i* s = reinterpret_cast<const __m128i*>(src);
        __m128i* d = reinterpret_cast<__m128i*>(dst);
        for (size_t i = 0; i < words / 2; i += 4) {
            const __m128i a = _mm_loadu_si128(s + i), b = _mm_loadu_si128(s + i + 1), c = _mm_loadu_si128(s + i + 2),
                          e = _mm_loadu_si128(s + i + 3);
            _mm_stream_si128(d + i, a);
            _mm_stream_si128(d + i + 1, b);
            _mm_stream_si128(d + i + 2, c);
            _mm_stream_si128(d + i + 3, e);
        }
        _mm_sfence();
        return;
    }
#endif
    std::memcpy(dst, src, words * sizeof(uint64_t));
}

}  // namespace psi
