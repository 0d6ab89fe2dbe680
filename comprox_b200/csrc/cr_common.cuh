// cr_common.cuh -- shared definitions for the comprox-b200 CUDA library.
//
// Two ways to compile the kernels in this directory:
//   * nvcc, sm_100a (the product: libcrgpu.so).  This is the only thing the C ABI ever runs.
//   * g++ with -DCRGPU_SIM (tests/sim only).  "Kernel-logic simulation": every kernel that is written as
//     independent threads (no __syncthreads / shared memory / warp intrinsics) is executed thread by thread
//     on the host so its integer logic can be checked against the oracle in a container without a GPU.
//     The simulation is a TEST HARNESS.  It is never linked into libcrgpu.so, and the product library has no
//     CPU code path: crgpu_create() fails when no CUDA device is present.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef CRGPU_SIM
// ------------------------------------------------------------------ host simulation shim (tests only)
#include <algorithm>
#include <vector>
struct crsim_dim3 { unsigned x, y, z; crsim_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
typedef crsim_dim3 dim3;
extern thread_local crsim_dim3 threadIdx, blockIdx, blockDim, gridDim;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
typedef void* cudaStream_t;
typedef int cudaError_t;
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#define cudaSuccess 0
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
#define CR_LAUNCH(kernel, grid, block, stream, ...)                                            \
    do {                                                                                       \
        crsim_dim3 g_ = (grid), b_ = (block);                                                  \
        gridDim = g_; blockDim = b_;                                                           \
        for (unsigned bz = 0; bz < g_.z; bz++) for (unsigned by = 0; by < g_.y; by++) for (unsigned bx = 0; bx < g_.x; bx++) \
        for (unsigned tz = 0; tz < b_.z; tz++) for (unsigned ty = 0; ty < b_.y; ty++) for (unsigned tx = 0; tx < b_.x; tx++) { \
            blockIdx = crsim_dim3(bx, by, bz); threadIdx = crsim_dim3(tx, ty, tz);             \
            kernel(__VA_ARGS__);                                                               \
        }                                                                                      \
    } while (0)
static inline int cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline int cudaFree(void* p) { free(p); return 0; }
static inline int cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline int cudaFreeHost(void* p) { free(p); return 0; }
enum { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
static inline int cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { if (n) memmove(d, s, n); return 0; }
static inline int cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return 0; }
static inline int cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline int cudaGetLastError() { return 0; }
static inline const char* cudaGetErrorString(int) { return "sim"; }
#else
// ------------------------------------------------------------------ real CUDA
#include <cuda_runtime.h>
extern unsigned long long g_cr_launches;   // kernels launched by this library (crgpu_launch_count)
#define CR_LAUNCH(kernel, grid, block, stream, ...)                                             \
    do {                                                                                        \
        __atomic_fetch_add(&g_cr_launches, 1ull, __ATOMIC_RELAXED);   /* handles work on several host threads */ \
        kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);                                  \
        cudaError_t le_ = cudaGetLastError();                                                   \
        if (le_ != cudaSuccess) {                                                               \
            fprintf(stderr, "crgpu: launch of %s failed: %s (%s:%d)\n", #kernel, cudaGetErrorString(le_), __FILE__, __LINE__); \
            return CRGPU_ERR_CUDA;                                                              \
        }                                                                                       \
    } while (0)
#endif

#define CR_HD __host__ __device__ __forceinline__
#define CR_D  __device__ __forceinline__

// error codes of the C ABI
#include "../../include/crgpu.h"

#define CR_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (cudaError_t)(expr);                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "crgpu: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return CRGPU_ERR_CUDA;                                                             \
        }                                                                                      \
    } while (0)
#define CR_TRY(expr) do { int r_ = (expr); if (r_ != CRGPU_OK) return r_; } while (0)

static inline unsigned cr_div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// ------------------------------------------------------------------ device memory arena-ish helper
// A growable device buffer: keeps its allocation between calls so steady-state runs do no cudaMalloc.
// Growth goes through the stream-ordered allocator (cudaMallocAsync / cudaFreeAsync on the stream the calling thread's handle works on,
// bound by CR_SET_DEVICE): cudaFree synchronises the whole device, which stalled every other handle of a crgpu_compress_batch call each
// time one of them met a larger container (mixed corpus, 8 handles: 9.7 s -> see profiles/round2_summary.md section 3).
#ifndef CRGPU_SIM
extern thread_local cudaStream_t g_cr_alloc_stream;
extern thread_local bool g_cr_alloc_async;
#endif
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return CRGPU_OK;
        const size_t want = n + n / 4 + 256;
#ifndef CRGPU_SIM
        if (g_cr_alloc_async) {
            if (p) cudaFreeAsync(p, g_cr_alloc_stream);
            p = nullptr; cap = 0;
            if (cudaMallocAsync(&p, want, g_cr_alloc_stream) != cudaSuccess) { p = nullptr; cudaGetLastError(); return CRGPU_ERR_OOM; }
            cap = want;
            return CRGPU_OK;
        }
#endif
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; return CRGPU_ERR_OOM; }
        cap = want;
        return CRGPU_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

// ------------------------------------------------------------------ stage timer (CUDA events on the handle's stream)
#include <string>
#include <vector>
struct StageTimer {
    bool enabled = false;
    cudaStream_t stream = 0;
#ifndef CRGPU_SIM
    std::vector<cudaEvent_t> ev;
#endif
    std::vector<std::string> names;
    std::vector<std::pair<std::string, double>> result;
    void begin(cudaStream_t s) { stream = s; names.clear(); if (enabled) mark("start"); }
    void mark(const char* name) {
        if (!enabled) return;
#ifndef CRGPU_SIM
        if (ev.size() <= names.size()) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); }
        cudaEventRecord(ev[names.size()], stream);
#endif
        names.push_back(name);
    }
    void count(const char* name, double v) {   // counters ride along in the report (names start with '#')
        if (!enabled) return;
        for (auto& r : result) if (r.first == name) { r.second += v; return; }
        result.push_back(std::make_pair(std::string(name), v));
    }
    void finish() {   // call after a stream synchronize
        if (!enabled) return;
#ifndef CRGPU_SIM
        for (size_t i = 1; i < names.size(); i++) {
            float ms = 0; cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
            bool found = false;
            for (auto& r : result) if (r.first == names[i]) { r.second += ms; found = true; }
            if (!found) result.push_back(std::make_pair(names[i], (double)ms));
        }
#endif
        names.clear();
    }
};

// ------------------------------------------------------------------ small device helpers
CR_HD uint32_t cr_ld32(const uint8_t* p) { return p[0] | p[1] << 8 | p[2] << 16 | (uint32_t)p[3] << 24; }
CR_HD bool cr_is_lower(uint32_t c) { return c - 'a' < 26u; }
CR_HD bool cr_is_upper(uint32_t c) { return c - 'A' < 26u; }
CR_HD bool cr_is_alpha(uint32_t c) { return ((c | 32) - 'a') < 26u; }

// common prefix length of a[0..) and b[0..), capped at `cap` (<= 255)
CR_HD uint32_t cr_cpl(const uint8_t* a, const uint8_t* b, uint32_t cap) {
    uint32_t j = 0;
    while (j < cap && a[j] == b[j]) j++;
    return j;
}
