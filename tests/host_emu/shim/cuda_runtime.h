// TEST INFRASTRUCTURE: stands in for <cuda_runtime.h> when the kernel headers of qbold_vi_b200/csrc are compiled for the
// HOST (g++ -DQB_HOST_EMU -I tests/host_emu/shim).  A small SIMT emulator: every CUDA thread of a CTA is a host thread;
// warp primitives (__shfl_sync, __ballot_sync, __all_sync, __syncwarp) meet at a per-warp barrier and exchange through a
// per-warp scratch line; __syncthreads is a per-CTA barrier; __shared__ objects are function-local statics (one CTA runs
// at a time); "shared addresses" are 32-bit offsets from an anchor in this module.  Only what the qBOLD kernels use.
#pragma once
#include <pthread.h>
#include <stdint.h>

#include <cmath>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#ifndef QB_HOST_EMU
#error "the cuda_runtime.h shim is for -DQB_HOST_EMU builds only"
#endif

// Entry points of a harness library.  Everything else is built with hidden visibility and without GNU unique symbols
// (-fvisibility=hidden -fno-gnu-unique): several harness libraries live in one test process, and each must keep its OWN
// emulator state, kernel statics (__shared__ objects) and shared-window anchor -- a process-wide unified inline variable
// would put the anchor of one library gigabytes away from the statics of another.
#define QB_EMU_API extern "C" __attribute__((visibility("default")))

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __grid_constant__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct float2 {
    float x, y;
};
struct alignas(16) float4 {
    float x, y, z, w;
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct dim3 {
    unsigned x = 1, y = 1, z = 1;
};
typedef void* cudaStream_t;
typedef int cudaError_t;

namespace qb_emu {

constexpr int kMaxWarps = 32;

struct Warp {
    pthread_barrier_t bar;
    uint32_t word[32];     // exchange line of the warp primitives
    unsigned vote;
};

struct Cta {
    pthread_barrier_t bar;
    Warp warp[kMaxWarps];
    int n_threads = 0;
};

inline Cta& cta() {
    static Cta c;
    return c;
}

inline char smem_anchor;     // "shared window" origin: shared addresses are offsets from here (same module, < 2 GB away)
inline unsigned to_shared(const void* p) { return (unsigned)(int32_t)(static_cast<const char*>(p) - &smem_anchor); }
inline void* from_shared(unsigned a) { return &smem_anchor + (int32_t)a; }

}  // namespace qb_emu

inline thread_local dim3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

namespace qb_emu {

inline Warp& my_warp() { return cta().warp[threadIdx.x >> 5]; }
inline int my_lane() { return (int)(threadIdx.x & 31); }

// all 32 lanes publish a word, then read the word of lane `src`
inline uint32_t exchange(uint32_t mine, int src) {
    Warp& w = my_warp();
    w.word[my_lane()] = mine;
    pthread_barrier_wait(&w.bar);
    const uint32_t got = w.word[src & 31];
    pthread_barrier_wait(&w.bar);       // nobody overwrites the line before everybody has read it
    return got;
}

inline unsigned ballot(bool pred) {
    Warp& w = my_warp();
    w.word[my_lane()] = pred ? 1u : 0u;
    pthread_barrier_wait(&w.bar);
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= (w.word[i] & 1u) << i;
    pthread_barrier_wait(&w.bar);
    return m;
}

// Runs kernel(args...) as a grid of `grid` x `grid_y` CTAs of `block` threads (block a multiple of 32, <= 1024), one
// CTA at a time; smem_bytes = the dynamic shared memory of the launch.
inline std::vector<float>& dynamic_window() {
    static std::vector<float> w;
    return w;
}
inline float* dynamic_smem() { return dynamic_window().data(); }       // `extern __shared__` of the running launch

inline void launch(int grid, int block, const std::function<void()>& kernel, int grid_y = 1, size_t smem_bytes = 0) {
    Cta& c = cta();
    dynamic_window().assign(smem_bytes / sizeof(float) + 4, 0.f);
    blockDim.x = (unsigned)block;
    gridDim.x = (unsigned)grid;
    gridDim.y = (unsigned)grid_y;
    for (int by = 0; by < grid_y; ++by)
    for (int b = 0; b < grid; ++b) {
        c.n_threads = block;
        pthread_barrier_init(&c.bar, nullptr, (unsigned)block);
        for (int w = 0; w < block / 32; ++w) pthread_barrier_init(&c.warp[w].bar, nullptr, 32);
        std::vector<std::thread> threads;
        threads.reserve((size_t)block);
        for (int t = 0; t < block; ++t)
            threads.emplace_back([&kernel, b, by, t]() {
                threadIdx.x = (unsigned)t;
                blockIdx.x = (unsigned)b;
                blockIdx.y = (unsigned)by;
                kernel();
            });
        for (auto& th : threads) th.join();
        pthread_barrier_destroy(&c.bar);
        for (int w = 0; w < block / 32; ++w) pthread_barrier_destroy(&c.warp[w].bar);
    }
}

}  // namespace qb_emu

static inline void __syncthreads() { pthread_barrier_wait(&qb_emu::cta().bar); }
static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&qb_emu::my_warp().bar); }

static inline unsigned __float_as_uint(float f) {
    unsigned u;
    std::memcpy(&u, &f, 4);
    return u;
}
static inline float __uint_as_float(unsigned u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

static inline float __shfl_sync(unsigned, float v, int src) {
    return __uint_as_float(qb_emu::exchange(__float_as_uint(v), src));
}
static inline int __shfl_sync(unsigned, int v, int src) { return (int)qb_emu::exchange((uint32_t)v, src); }
static inline unsigned __shfl_sync(unsigned, unsigned v, int src) { return qb_emu::exchange(v, src); }
static inline unsigned long long __shfl_sync(unsigned m, unsigned long long v, int src) {
    const unsigned long long lo = qb_emu::exchange((uint32_t)v, src);
    const unsigned long long hi = qb_emu::exchange((uint32_t)(v >> 32), src);
    (void)m;
    return lo | (hi << 32);
}
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
    return __uint_as_float(qb_emu::exchange(__float_as_uint(v), qb_emu::my_lane() ^ lane_mask));
}
static inline unsigned __ballot_sync(unsigned, bool pred) { return qb_emu::ballot(pred); }
static inline bool __all_sync(unsigned, bool pred) { return qb_emu::ballot(pred) == 0xffffffffu; }
static inline bool __any_sync(unsigned, bool pred) { return qb_emu::ballot(pred) != 0u; }

static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
template <class T>
static inline T __ldg(const T* p) {
    return *p;
}

// IEEE single operations that the device code asks for by name (no contraction): volatile keeps the host compiler
// from fusing them either
static inline float __fmul_rn(float a, float b) {
    volatile float r = a * b;
    return r;
}
static inline float __fadd_rn(float a, float b) {
    volatile float r = a + b;
    return r;
}
static inline float __fsub_rn(float a, float b) {
    volatile float r = a - b;
    return r;
}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline void sincospif(float x, float* s, float* c) {
    *s = (float)std::sin(M_PI * (double)x);
    *c = (float)std::cos(M_PI * (double)x);
}
static inline size_t __cvta_generic_to_shared(const void* p) { return qb_emu::to_shared(p); }

// host-API names the launch helpers mention (never called in the emulator)
constexpr cudaError_t cudaSuccess = 0;

// ---- additions for the ELBO kernels
using std::isfinite;
using std::isnan;
using std::isinf;
#define __noinline__ __attribute__((noinline))
struct float3 {
    float x, y, z;
};
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
static inline double atomicAdd(double* p, double v) {
    uint64_t o = __atomic_load_n(reinterpret_cast<uint64_t*>(p), __ATOMIC_RELAXED), w;
    double old;
    do {
        std::memcpy(&old, &o, 8);
        const double want = old + v;
        std::memcpy(&w, &want, 8);
    } while (!__atomic_compare_exchange_n(reinterpret_cast<uint64_t*>(p), &o, w, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return old;
}
static inline float atomicAdd(float* p, float v) {
    float old;
    uint32_t o = __atomic_load_n(reinterpret_cast<uint32_t*>(p), __ATOMIC_RELAXED), w;
    do {
        std::memcpy(&old, &o, 4);
        const float want = old + v;
        std::memcpy(&w, &want, 4);
    } while (!__atomic_compare_exchange_n(reinterpret_cast<uint32_t*>(p), &o, w, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return old;
}
static inline int __shfl_xor_sync(unsigned, int v, int lane_mask) {
    return (int)qb_emu::exchange((uint32_t)v, qb_emu::my_lane() ^ lane_mask);
}
static inline double __shfl_xor_sync(unsigned, double v, int lane_mask) {
    uint64_t u;
    std::memcpy(&u, &v, 8);
    const uint64_t lo = qb_emu::exchange((uint32_t)u, qb_emu::my_lane() ^ lane_mask);
    const uint64_t hi = qb_emu::exchange((uint32_t)(u >> 32), qb_emu::my_lane() ^ lane_mask);
    u = lo | (hi << 32);
    std::memcpy(&v, &u, 8);
    return v;
}
static inline double __shfl_sync(unsigned, double v, int src) {
    uint64_t u;
    std::memcpy(&u, &v, 8);
    const uint64_t lo = qb_emu::exchange((uint32_t)u, src);
    const uint64_t hi = qb_emu::exchange((uint32_t)(u >> 32), src);
    u = lo | (hi << 32);
    std::memcpy(&v, &u, 8);
    return v;
}
static inline float __shfl_down_sync(unsigned, float v, unsigned delta) {
    const int src = qb_emu::my_lane() + (int)delta;
    const float got = __uint_as_float(qb_emu::exchange(__float_as_uint(v), src & 31));
    return src < 32 ? got : v;
}
