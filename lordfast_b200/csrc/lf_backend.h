/*
 * lf_backend.h -- the handful of runtime calls the host pipeline (lf_pipeline.inl) needs.
 *
 * Product build (nvcc, LF_EMU undefined): thin wrappers over the CUDA runtime + CUB; this is the
 * only backend liblfgpu.so contains -- there is no CPU path in the shipped library.
 * Test build (g++, LF_EMU defined by tests/emu/cuda_emu.h): the same pipeline source runs its
 * kernels through the fiber emulator so that it can be debugged in a container without a GPU.
 */
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>

#ifndef LF_EMU
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>

typedef cudaStream_t lfb_stream;
typedef cudaEvent_t lfb_event;
#define LFB_CHECK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return lfb_fail(cudaGetErrorString(e_), #expr); } while (0)
/* Preferred shared-memory carve-out (percent of the SM's 256 KB L1/shared array) asked for by every kernel of the
 * library.  The size-class kernels run concurrently and the split is per-SM state; measured on B200 with the config-2
 * step (profiles/r02c_carveout_bandreg.txt): driver default 1.97 ms, 25 % 1.98, 50 % 1.84, 75 % 1.95, 100 % 2.34 --
 * the checkpoints and op planes live on L1 hits, so all-shared is the worst choice.  LF_CARVEOUT overrides (-1 = leave
 * the driver default). */
static inline int lfb_carveout() { static int v = -2; if (v == -2) { const char *e = getenv("LF_CARVEOUT"); v = e ? atoi(e) : 50; } return v; }
/* the attribute is per device: once per (kernel, device the launching thread made current with set_dev) */
static thread_local int lfb_cur_dev = 0;
#define LFB_LAUNCH(kern, grid, block, smem, stream, ...) do { \
        static std::atomic<unsigned long long> lfb_done_{0}; \
        const unsigned long long lfb_bit_ = 1ull << (lfb_cur_dev & 63); \
        if (!(lfb_done_.load(std::memory_order_relaxed) & lfb_bit_)) { if (lfb_carveout() >= 0) cudaFuncSetAttribute((const void *)(kern), cudaFuncAttributePreferredSharedMemoryCarveout, lfb_carveout()); lfb_done_.fetch_or(lfb_bit_); } \
        kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); lfb_launches++; } while (0)
#else
#include <algorithm>
#include <numeric>
#include <vector>
typedef int lfb_stream;
typedef int lfb_event;
#define LFB_CHECK(expr) do { int e_ = (expr); if (e_ != 0) return lfb_fail("emu", #expr); } while (0)
#define LFB_LAUNCH(kern, grid, block, smem, stream, ...) do { emu::launch(emu_dim3(grid), emu_dim3(block), (smem), [&] { kern(__VA_ARGS__); }); lfb_launches++; } while (0)
#endif

static thread_local char lfb_errbuf[512];
static std::atomic<uint64_t> lfb_launches{0}; /* kernels launched by this library (the early emit launches from its own thread) */
static int lfb_fail(const char *what, const char *where)
{
    snprintf(lfb_errbuf, sizeof lfb_errbuf, "%s at %s", what, where);
    return -2; /* LF_ERR_CUDA */
}

#ifndef LF_EMU
static inline int lfb_malloc(void **p, size_t n) { return cudaMalloc(p, n ? n : 1) == cudaSuccess ? 0 : lfb_fail(cudaGetErrorString(cudaGetLastError()), "cudaMalloc"); }
static inline void lfb_free(void *p) { if (p) cudaFree(p); }
static inline int lfb_h2d(void *d, const void *h, size_t n, lfb_stream s) { LFB_CHECK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s)); return 0; }
static inline int lfb_d2h(void *h, const void *d, size_t n, lfb_stream s) { LFB_CHECK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s)); return 0; }
static inline int lfb_memset(void *d, int v, size_t n, lfb_stream s) { LFB_CHECK(cudaMemsetAsync(d, v, n, s)); return 0; }
static inline int lfb_sync(lfb_stream s) { LFB_CHECK(cudaStreamSynchronize(s)); return 0; }
static inline int lfb_last_error() { LFB_CHECK(cudaGetLastError()); return 0; }
static inline void *lfb_host_alloc(size_t n) { void *p = NULL; return cudaMallocHost(&p, n ? n : 1) == cudaSuccess ? p : NULL; }
/* is p pinned host memory (cudaMallocHost / lf_gpu_host_alloc) that a kernel may read directly? */
static inline bool lfb_is_pinned(const void *p) { cudaPointerAttributes a; if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; } return a.type == cudaMemoryTypeHost; }
static inline void lfb_host_free(void *p) { if (p) cudaFreeHost(p); }

struct LfbTemp { void *p = nullptr; size_t cap = 0; };
static inline int lfb_temp(LfbTemp &t, size_t need) { if (need > t.cap) { lfb_free(t.p); t.p = nullptr; t.cap = 0; if (lfb_malloc(&t.p, need + need / 4)) return -2; t.cap = need + need / 4; } return 0; }

static inline int lfb_sort_pairs(LfbTemp &tmp, const uint32_t *kin, uint32_t *kout, const uint32_t *vin, uint32_t *vout, size_t n, lfb_stream s)
{
    size_t need = 0;
    LFB_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, need, kin, kout, vin, vout, (int)n, LF_KEY_SHIFT, 32, s));
    if (lfb_temp(tmp, need)) return -2;
    LFB_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, need, kin, kout, vin, vout, (int)n, LF_KEY_SHIFT, 32, s));
    return 0;
}
struct LfbU32ToU64 { __host__ __device__ unsigned long long operator()(uint32_t v) const { return v; } };
/* out[i] = sum_{j<=i} in[j] (inclusive) or sum_{j<i} (exclusive, n+1 entries with the total last) */
static inline int lfb_scan_incl(LfbTemp &tmp, const uint32_t *in, unsigned long long *out, size_t n, lfb_stream s)
{
    thrust::transform_iterator<LfbU32ToU64, const uint32_t *, unsigned long long> it(in, LfbU32ToU64());
    size_t need = 0;
    LFB_CHECK(cub::DeviceScan::InclusiveSum(nullptr, need, it, out, (int)n, s));
    if (lfb_temp(tmp, need)) return -2;
    LFB_CHECK(cub::DeviceScan::InclusiveSum(tmp.p, need, it, out, (int)n, s));
    return 0;
}
static inline int lfb_scan_excl_total(LfbTemp &tmp, const uint32_t *in, unsigned long long *out, size_t n, lfb_stream s)
{ /* out has n+1 entries: out[0] = 0, out[i+1] = inclusive sum */
    LFB_CHECK(cudaMemsetAsync(out, 0, sizeof(unsigned long long), s));
    return lfb_scan_incl(tmp, in, out + 1, n, s);
}
#else
static inline int lfb_malloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : -2; }
static inline void lfb_free(void *p) { free(p); }
static inline int lfb_h2d(void *d, const void *h, size_t n, lfb_stream) { memcpy(d, h, n); return 0; }
static inline int lfb_d2h(void *h, const void *d, size_t n, lfb_stream) { memcpy(h, d, n); return 0; }
static inline int lfb_memset(void *d, int v, size_t n, lfb_stream) { memset(d, v, n); return 0; }
static inline int lfb_sync(lfb_stream) { return 0; }
static inline int lfb_last_error() { return 0; }
static inline void *lfb_host_alloc(size_t n) { return malloc(n ? n : 1); }
static inline bool lfb_is_pinned(const void *) { return false; }
static inline void lfb_host_free(void *p) { free(p); }
struct LfbTemp { void *p = nullptr; size_t cap = 0; };
static inline int lfb_sort_pairs(LfbTemp &, const uint32_t *kin, uint32_t *kout, const uint32_t *vin, uint32_t *vout, size_t n, lfb_stream)
{
    std::vector<uint32_t> perm(n);
    std::iota(perm.begin(), perm.end(), 0u);
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return kin[a] < kin[b]; });
    for (size_t i = 0; i < n; i++) { kout[i] = kin[perm[i]]; vout[i] = vin[perm[i]]; }
    return 0;
}
static inline int lfb_scan_incl(LfbTemp &, const uint32_t *in, unsigned long long *out, size_t n, lfb_stream)
{ unsigned long long s = 0; for (size_t i = 0; i < n; i++) { s += in[i]; out[i] = s; } return 0; }
static inline int lfb_scan_excl_total(LfbTemp &t, const uint32_t *in, unsigned long long *out, size_t n, lfb_stream s)
{ out[0] = 0; return lfb_scan_incl(t, in, out + 1, n, s); }
#endif

/* grow-only device buffer */
struct LfbBuf {
    void *p = nullptr; size_t cap = 0;
    int reserve(size_t need) { if (need <= cap) return 0; lfb_free(p); p = nullptr; cap = 0; size_t want = need + need / 8 + 256; if (lfb_malloc(&p, want)) return -5; cap = want; return 0; }
    void release() { lfb_free(p); p = nullptr; cap = 0; }
    template <typename T> T *as() const { return (T *)p; }
};
