// host widening bandwidth probe: packed u32 (idx16 | data16 << 16) -> two int32 arrays
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>
static void expand(const uint32_t *src, int32_t *idx, int32_t *dat, size_t n, bool nt)
{
    size_t i = 0;
#ifdef __AVX2__
    const __m256i mask = _mm256_set1_epi32(0xffff);
    for (; i + 8 <= n; i += 8)
    {
        __m256i w = _mm256_loadu_si256((const __m256i *)(src + i));
        __m256i a = _mm256_and_si256(w, mask);
        __m256i b = _mm256_srai_epi32(w, 16);
        if (nt) { _mm256_stream_si256((__m256i *)(idx + i), a); _mm256_stream_si256((__m256i *)(dat + i), b); }
        else { _mm256_storeu_si256((__m256i *)(idx + i), a); _mm256_storeu_si256((__m256i *)(dat + i), b); }
    }
#endif
    for (; i < n; ++i) { idx[i] = src[i] & 0xffff; dat[i] = (int32_t)src[i] >> 16; }
}
int main(int argc, char **argv)
{
    size_t n = 448ull << 20;   // entries (1.8 GB packed -> 3.6 GB out)
    uint32_t *src = (uint32_t *)aligned_alloc(64, n * 4);
    int32_t *idx = (int32_t *)aligned_alloc(64, n * 4), *dat = (int32_t *)aligned_alloc(64, n * 4);
    memset(src, 1, n * 4); memset(idx, 0, n * 4); memset(dat, 0, n * 4);
    printf("hw threads %u\n", std::thread::hardware_concurrency());
    for (int nt = 0; nt < 2; ++nt)
        for (int T : {1, 2, 4, 8, 12, 16, 24, 32})
        {
            double best = 0;
            for (int rep = 0; rep < 3; ++rep)
            {
                auto t0 = std::chrono::steady_clock::now();
                std::vector<std::thread> th;
                for (int t = 0; t < T; ++t)
                {
                    size_t lo = n * t / T / 8 * 8, hi = (t + 1 == T) ? n : n * (t + 1) / T / 8 * 8;
                    th.emplace_back(expand, src + lo, idx + lo, dat + lo, hi - lo, nt != 0);
                }
                for (auto &x : th) x.join();
                double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                best = std::max(best, n * 8.0 / s / 1e9);
            }
            printf("nt=%d threads=%2d  output %.1f GB/s (3.59 GB in %.1f ms)\n", nt, T, best, 3.59 / best * 1e3);
        }
    return 0;
}
