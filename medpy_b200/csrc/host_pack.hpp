// host_pack.hpp -- host-side bit packing of marker volumes for the end-to-end path (used by gc_pybind.cpp; compiled on
// its own by tests/test_host_packer.py).  No CUDA here.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

// ---- marker volumes -> bit planes on the host (bit v & 31 of word v >> 5 = marker[v] != 0) --------------------
// graph_from_voxels receives the markers as bool arrays (generate.py:125-126); crossing PCIe as bits instead of bytes
// saves 1.75 of the 10 bytes per voxel an end-to-end step has to upload.  Worker threads pack block after block in
// order and publish their progress, so the native call can start uploading the image while the tail is still packed.
#if defined(__x86_64__)
__attribute__((target("avx2"))) void pack_words_avx2(const uint8_t* src, uint32_t* dst, size_t nwords)
{
    const __m256i zero = _mm256_setzero_si256();
    for (size_t w = 0; w < nwords; ++w) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + 32 * w));
        dst[w] = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, zero));
    }
}
#endif

void pack_words(const uint8_t* src, uint32_t* dst, size_t nwords)
{
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) { pack_words_avx2(src, dst, nwords); return; }
#endif
    for (size_t w = 0; w < nwords; ++w) {
        uint32_t m = 0;
        for (int b = 0; b < 32; ++b) m |= (uint32_t)(src[32 * w + b] != 0) << b;
        dst[w] = m;
    }
}

struct MarkerPacker {
    const uint8_t* src[2] = {nullptr, nullptr};
    uint32_t* dst[2] = {nullptr, nullptr};
    size_t n = 0, words = 0;
    std::atomic<int64_t> ready{0};            // leading words of BOTH planes written so far
    std::vector<std::thread> workers;
    std::vector<std::atomic<int>> done;       // per block
    size_t block_words = 1 << 16;             // 2 Mi voxels per block
    size_t nblocks = 0;
    std::atomic<size_t> next{0};

    MarkerPacker(const uint8_t* fg, const uint8_t* bg, uint32_t* fgb, uint32_t* bgb, size_t n_)
        : n(n_), words((n_ + 31) / 32), done(((n_ + 31) / 32 + (1 << 16) - 1) / (1 << 16))
    {
        src[0] = fg; src[1] = bg; dst[0] = fgb; dst[1] = bgb;
        nblocks = done.size();
        for (auto& d : done) d.store(0);
        // worker count: at most 8, and no more than the container's CPU quota allows (hardware_concurrency() reports the
        // machine's logical CPUs; oversubscribing a cgroup quota gets every thread of the process throttled)
        unsigned nt = std::thread::hardware_concurrency();
        if (const char* e = std::getenv("MEDPY_GC_PACK_THREADS")) { if (std::atoi(e) > 0) nt = (unsigned)std::atoi(e); }
        else {
            if (FILE* f = std::fopen("/sys/fs/cgroup/cpu.max", "r")) {
                long long q = 0, per = 0;
                if (std::fscanf(f, "%lld %lld", &q, &per) == 2 && q > 0 && per > 0) { const unsigned lim = (unsigned)(q / per / 2); if (lim >= 1 && lim < nt) nt = lim; }
                std::fclose(f);
            }
        }
        nt = nt < 2 ? 1 : (nt > 8 ? 8 : nt);
        if (nblocks < nt) nt = (unsigned)nblocks;
        for (unsigned i = 0; i < nt; ++i) workers.emplace_back([this] { run(); });
    }
    void run()
    {
        for (;;) {
            const size_t b = next.fetch_add(1);
            if (b >= nblocks) break;
            const size_t w0 = b * block_words, w1 = std::min(words, w0 + block_words);
            for (int p = 0; p < 2; ++p) {
                if (!src[p]) continue;
                const size_t full = std::min(w1, n / 32);          // words whose 32 voxels all exist
                if (full > w0) pack_words(src[p] + 32 * w0, dst[p] + w0, full - w0);
                if (w1 > full) {                                   // the last, partial word
                    uint32_t m = 0;
                    for (size_t v = 32 * full; v < n; ++v) m |= (uint32_t)(src[p][v] != 0) << (v - 32 * full);
                    dst[p][full] = m;
                }
            }
            done[b].store(1, std::memory_order_release);
            // advance the contiguous frontier
            int64_t r = ready.load();
            for (;;) {
                if ((size_t)r >= words) break;
                const size_t fb = (size_t)r / block_words;
                if (fb >= nblocks || !done[fb].load(std::memory_order_acquire)) break;
                const int64_t nr = (int64_t)std::min(words, (fb + 1) * block_words);
                if (ready.compare_exchange_weak(r, nr)) r = nr;
            }
        }
    }
    ~MarkerPacker() { for (auto& t : workers) t.join(); }
};

