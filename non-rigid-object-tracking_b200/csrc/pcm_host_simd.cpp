// pcm_host_simd.cpp -- host-side byte shuffles of the host-buffer entry points: the reference's
// mask is an interleaved H x W x 3 image of which only channel 2 is used
// (maskers/pixel_classification.py:246, main.py:286,343), so every pcm_update scatters a dense
// plane into every third byte and every pcm_iou gathers it back.  SSSE3 where the CPU has it
// (checked at run time), plain loops otherwise.  No part of the masker arithmetic lives here.
#include <cstdint>
#include <cstring>
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define PCM_X86 1
#endif

namespace pcm {

#ifdef PCM_X86
__attribute__((target("ssse3"))) static int gather3_ssse3(const uint8_t* s, uint8_t* d, int n) {
    const __m128i m0 = _mm_setr_epi8(0, 3, 6, 9, 12, 15, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m128i m1 = _mm_setr_epi8(-1, -1, -1, -1, -1, -1, 2, 5, 8, 11, 14, -1, -1, -1, -1, -1);
    const __m128i m2 = _mm_setr_epi8(-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 1, 4, 7, 10, 13);
    int i = 0;
    // a group reads s[3i .. 3i+47]; the last byte that exists is s[3(n-1)]
    for (; i + 17 <= n; i += 16) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 3 * i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 3 * i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 3 * i + 32));
        const __m128i r = _mm_or_si128(_mm_or_si128(_mm_shuffle_epi8(a, m0), _mm_shuffle_epi8(b, m1)), _mm_shuffle_epi8(c, m2));
        _mm_storeu_si128(reinterpret_cast<__m128i*>(d + i), r);
    }
    return i;
}

__attribute__((target("ssse3"))) static int scatter3_ssse3(const uint8_t* s, uint8_t* d, int n) {
    // byte j of the 16 source bytes goes to byte 3j of the 48 destination bytes
    const __m128i p0 = _mm_setr_epi8(0, -1, -1, 1, -1, -1, 2, -1, -1, 3, -1, -1, 4, -1, -1, 5);
    const __m128i p1 = _mm_setr_epi8(-1, -1, 6, -1, -1, 7, -1, -1, 8, -1, -1, 9, -1, -1, 10, -1);
    const __m128i p2 = _mm_setr_epi8(-1, 11, -1, -1, 12, -1, -1, 13, -1, -1, 14, -1, -1, 15, -1, -1);
    const __m128i k0 = _mm_setr_epi8(-1, 0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0, -1);
    const __m128i k1 = _mm_setr_epi8(0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0, -1, 0);
    const __m128i k2 = _mm_setr_epi8(0, -1, 0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0, -1, 0, 0);
    int i = 0;
    // a group rewrites d[3i .. 3i+47] (other channels are written back unchanged)
    for (; i + 17 <= n; i += 16) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
        __m128i* q = reinterpret_cast<__m128i*>(d + 3 * i);
        const __m128i a = _mm_loadu_si128(q), b = _mm_loadu_si128(q + 1), c = _mm_loadu_si128(q + 2);
        const __m128i na = _mm_or_si128(_mm_andnot_si128(k0, a), _mm_shuffle_epi8(v, p0));
        const __m128i nb = _mm_or_si128(_mm_andnot_si128(k1, b), _mm_shuffle_epi8(v, p1));
        const __m128i nc = _mm_or_si128(_mm_andnot_si128(k2, c), _mm_shuffle_epi8(v, p2));
        // a mask changes little from frame to frame: where the caller's bytes already hold the new values nothing is
        // stored, so the cache lines stay clean and the scatter costs their read only
        const __m128i same = _mm_and_si128(_mm_and_si128(_mm_cmpeq_epi8(na, a), _mm_cmpeq_epi8(nb, b)), _mm_cmpeq_epi8(nc, c));
        if (_mm_movemask_epi8(same) == 0xffff) continue;
        _mm_storeu_si128(q, na);
        _mm_storeu_si128(q + 1, nb);
        _mm_storeu_si128(q + 2, nc);
    }
    return i;
}

static bool have_ssse3() {
    static const bool v = __builtin_cpu_supports("ssse3");
    return v;
}
#endif

// dst[i] = src[i * stride], i < n
void gather_strided(const uint8_t* src, int64_t stride, uint8_t* dst, int n) {
    if (stride == 1) { memcpy(dst, src, (size_t)n); return; }
    int i = 0;
#ifdef PCM_X86
    if (stride == 3 && have_ssse3()) i = gather3_ssse3(src, dst, n);
#endif
    for (; i < n; ++i) dst[i] = src[(size_t)i * stride];
}

// dst[i * stride] = src[i], i < n; bytes between the written ones keep their values
void scatter_strided(const uint8_t* src, uint8_t* dst, int64_t stride, int n) {
    if (stride == 1) { memcpy(dst, src, (size_t)n); return; }
    int i = 0;
#ifdef PCM_X86
    if (stride == 3 && have_ssse3()) i = scatter3_ssse3(src, dst, n);
#endif
    for (; i < n; ++i) dst[(size_t)i * stride] = src[i];
}

}  // namespace pcm
