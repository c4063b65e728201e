// Streaming (non-temporal) memory copy for the staging copies of the pageable host paths (pipeline.cu).
// The destination of such a copy is read next by the DMA engine (page-locked staging chunk) or much later by the
// caller (the result array), never soon by this CPU, so the stores bypass the cache: no read-for-ownership of the
// destination lines, i.e. two instead of three passes over the memory bus per byte (glibc's memcpy only switches
// to such stores above a threshold far larger than the 2 - 4 MiB parts the copy pool hands out).
// Measured on an 8-core Xeon, 32 MiB pieces in 8 parts: memcpy 21 GB/s, this 32 GB/s.
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>

__attribute__((target("avx2"))) static void copy_avx2(unsigned char* d, const unsigned char* s, size_t n) {
  const size_t blocks = n / 128;
  for (size_t i = 0; i < blocks; ++i) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 64));
    const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(d + 96), e);
    s += 128;
    d += 128;
  }
  _mm_sfence();
  std::memcpy(d, s, n - blocks * 128);
}

static void copy_sse2(unsigned char* d, const unsigned char* s, size_t n) {
  const size_t blocks = n / 64;
  for (size_t i = 0; i < blocks; ++i) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 32));
    const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(d), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 48), e);
    s += 64;
    d += 64;
  }
  _mm_sfence();
  std::memcpy(d, s, n - blocks * 64);
}
#endif

namespace rg {

// dst and src must not overlap.  Small copies and non-x86 hosts use memcpy.
void stream_copy(void* dst, const void* src, size_t n) {
#if defined(__x86_64__)
  static const bool plain = std::getenv("RG_HOST_COPY_PLAIN") != nullptr;  // A/B switch of tools/host_copy_probe.py
  if (n >= (256u << 10) && !plain) {
    unsigned char* d = static_cast<unsigned char*>(dst);
    const unsigned char* s = static_cast<const unsigned char*>(src);
    size_t head = (64 - (reinterpret_cast<uintptr_t>(d) & 63)) & 63;  // the streaming stores want an aligned target
    std::memcpy(d, s, head);
    d += head;
    s += head;
    n -= head;
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) copy_avx2(d, s, n);
    else copy_sse2(d, s, n);
    return;
  }
#endif
  std::memcpy(dst, src, n);
}

}  // namespace rg
