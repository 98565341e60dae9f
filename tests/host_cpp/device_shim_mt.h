// device_shim_mt.h -- a second way to compile the product's CUDA sources for the HOST: here a CUDA thread is a real host thread
// and a warp has its 32 lanes.  __syncwarp / __syncthreads are pthread barriers, the warp collectives (ballot, shuffle, match,
// reduce, any) exchange their operands through a per-warp slot array between two barriers, atomics are real atomics.  This
// is what the warp-cooperative code of the product runs under on the CPU -- k_sampler_warp: 128-word twist steps across the
// lanes, Lemire rejections ending a batch at a lane, shuffles applied in waves of independent exchanges, the CTA-wide
// copy-out -- which the one-lane emulation of device_shim.h cannot exercise.  Because every barrier is a pthread barrier, the
// build also runs under ThreadSanitizer: a missing __syncwarp between a shared-memory store and another lane's load is a
// data race it reports (tools/tsan_lanes_on_host.sh), and under AddressSanitizer for out-of-bounds shared / global accesses.
// Test infrastructure only (tests/host_cpp/sampler_mt.cpp, tests/test_sampler_lanes_on_host.py).
#pragma once
#include <cuda_runtime.h>
#include <pthread.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#undef __device__
#undef __host__
#undef __global__
#undef __forceinline__
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#undef __shared__
#undef __launch_bounds__
#define __shared__ static
#define __launch_bounds__(...)
#undef __noinline__
#define __noinline__
#define RGK_WARP_LANES 32
#define __CUDA_ARCH_EMULATED_LANES__ 32

struct doh_idx { unsigned x, y, z; };
// a warp: the full-mask barrier is a pthread barrier; a collective over a partial mask (the lanes that took a branch: e.g. the
// shuffle of push_queue among the lanes that queue a ray) meets at a per-mask counter under the warp's mutex
struct DohGroup { unsigned arrived = 0, generation = 0; };
struct DohWarp {
    pthread_barrier_t bar; unsigned long long slot[32];
    std::mutex mu; std::condition_variable cv; std::map<unsigned, DohGroup> groups;
    void meet(unsigned mask) {
        if (mask == 0xffffffffu) { pthread_barrier_wait(&bar); return; }
        std::unique_lock<std::mutex> lock(mu);
        DohGroup& g = groups[mask];
        const unsigned mine = g.generation;
        if (++g.arrived == (unsigned)__builtin_popcount(mask)) { g.arrived = 0; g.generation++; cv.notify_all(); }
        else cv.wait(lock, [&] { return g.generation != mine; });
    }
};
struct DohBlock { pthread_barrier_t bar; std::unique_ptr<DohWarp[]> warps; };
inline thread_local doh_idx doh_threadIdx{0, 0, 0}, doh_blockIdx{0, 0, 0};
inline doh_idx doh_blockDim{1, 1, 1}, doh_gridDim{1, 1, 1};
inline thread_local DohWarp* doh_warp = nullptr;
inline thread_local DohBlock* doh_block = nullptr;
inline thread_local unsigned doh_lane = 0;
#define threadIdx doh_threadIdx
#define blockIdx doh_blockIdx
#define blockDim doh_blockDim
#define gridDim doh_gridDim

static inline void __syncwarp(unsigned mask = 0xffffffffu) { doh_warp->meet(mask); }
static inline void __syncthreads() { pthread_barrier_wait(&doh_block->bar); }
// every collective: publish, barrier, read, barrier (the second barrier keeps a fast lane's next publish from overtaking a slow reader)
template <class F> static inline auto doh_exchange(unsigned mask, unsigned long long mine, F read) {
    doh_warp->slot[doh_lane] = mine;
    doh_warp->meet(mask);
    auto r = read(doh_warp->slot);
    doh_warp->meet(mask);
    return r;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    return doh_exchange(mask, pred ? 1ull : 0ull, [mask](const unsigned long long* s) { unsigned m = 0; for (int l = 0; l < 32; l++) if (((mask >> l) & 1u) && s[l]) m |= 1u << l; return m; });
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src) {
    unsigned long long bits = 0; std::memcpy(&bits, &v, sizeof(T));
    const unsigned long long r = doh_exchange(mask, bits, [src](const unsigned long long* s) { return s[src & 31]; });
    T out; std::memcpy(&out, &r, sizeof(T)); return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int mask) { return __shfl_sync(m, v, (int)(doh_lane ^ (unsigned)mask)); }
template <class T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d) { return __shfl_sync(m, v, doh_lane >= d ? (int)(doh_lane - d) : (int)doh_lane); }
static inline unsigned __match_any_sync(unsigned mask, unsigned v) {
    return doh_exchange(mask, (unsigned long long)v, [mask, v](const unsigned long long* s) { unsigned m = 0; for (int l = 0; l < 32; l++) if (((mask >> l) & 1u) && (unsigned)s[l] == v) m |= 1u << l; return m; });
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    return doh_exchange(mask, (unsigned long long)v, [mask](const unsigned long long* s) { unsigned m = 0; for (int l = 0; l < 32; l++) if ((mask >> l) & 1u) m = std::max(m, (unsigned)s[l]); return m; });
}

template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((unsigned long long)a * b) >> 32); }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline float atomicAdd(float* p, float v) {
    uint32_t* q = reinterpret_cast<uint32_t*>(p); uint32_t old = __atomic_load_n(q, __ATOMIC_RELAXED);
    for (;;) { const float n = __uint_as_float(old) + v; if (__atomic_compare_exchange_n(q, &old, __float_as_uint(n), false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return __uint_as_float(old); }
}
using std::isnan; using std::min; using std::max;
static inline float __uint2float_rn(uint32_t x) { return (float)x; }

// ---- a launch: the blocks one after the other, the threads of a block as host threads (block sizes are multiples of 32 or below 32)
template <class... KArgs, class... Args>
inline void doh_launch(void (*kernel)(KArgs...), size_t grid, size_t block, size_t, cudaStream_t, Args&&... args) {
    doh_gridDim = {(unsigned)grid, 1, 1}; doh_blockDim = {(unsigned)block, 1, 1};
    static const bool trace = std::getenv("DOH_TRACE") != nullptr;      // one line per launch on stderr (finding a launch that never ends)
    if (trace) std::fprintf(stderr, "doh_launch %p grid %zu block %zu\n", (void*)kernel, grid, block);
    for (unsigned b = 0; b < (unsigned)grid; b++) {
        DohBlock blk;
        const unsigned nwarps = ((unsigned)block + 31u) / 32u;
        blk.warps.reset(new DohWarp[nwarps]);
        pthread_barrier_init(&blk.bar, nullptr, (unsigned)block);
        for (unsigned w = 0; w < nwarps; w++) pthread_barrier_init(&blk.warps[w].bar, nullptr, std::min(32u, (unsigned)block - 32u * w));
        std::vector<std::thread> th;
        for (unsigned t = 0; t < (unsigned)block; t++)
            th.emplace_back([&, t, b]() {
                doh_blockIdx = {b, 0, 0}; doh_threadIdx = {t, 0, 0}; doh_block = &blk; doh_warp = &blk.warps[t / 32u]; doh_lane = t % 32u;
                kernel(args...);
            });
        for (auto& x : th) x.join();
        for (unsigned w = 0; w < nwarps; w++) pthread_barrier_destroy(&blk.warps[w].bar);
        pthread_barrier_destroy(&blk.bar);
    }
    doh_gridDim = {1, 1, 1}; doh_blockDim = {1, 1, 1};
}

// ---- the CUDA runtime calls of the host code
static inline cudaError_t doh_cudaMalloc(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t doh_cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t doh_cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t = nullptr) { std::memset(p, v, n); return cudaSuccess; }
static inline cudaError_t doh_cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t doh_cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t doh_ok() { return cudaSuccess; }
static inline cudaError_t doh_cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t doh_cudaEventOp(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t doh_cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
static inline cudaError_t doh_cudaMemGetInfo(size_t* f, size_t* t) { *f = *t = (size_t)512 << 20; return cudaSuccess; }
static inline cudaError_t doh_cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 1; return cudaSuccess; }
template <class F> static inline cudaError_t doh_cudaOccupancy(int* n, F, int, size_t) { *n = 1; return cudaSuccess; }
template <class F> static inline cudaError_t doh_cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
#define cudaMalloc(p, n) doh_cudaMalloc((void**)(p), (n))
#define cudaMallocHost(p, n) doh_cudaMalloc((void**)(p), (n))
#define cudaFree(p) doh_cudaFree((void*)(p))
#define cudaFreeHost(p) doh_cudaFree((void*)(p))
#define cudaMemsetAsync doh_cudaMemsetAsync
#define cudaMemcpyAsync doh_cudaMemcpyAsync
#define cudaMemcpy doh_cudaMemcpy
#define cudaStreamSynchronize(s) doh_ok()
#define cudaGetLastError() doh_ok()
#define cudaEventCreate doh_cudaEventCreate
#define cudaEventRecord doh_cudaEventOp
#define cudaEventSynchronize doh_cudaEventOp
#define cudaEventDestroy doh_cudaEventOp
#define cudaEventQuery(e) doh_ok()
#define cudaEventElapsedTime doh_cudaEventElapsedTime
#define cudaMemGetInfo doh_cudaMemGetInfo
#define cudaDeviceGetAttribute doh_cudaDeviceGetAttribute
#define cudaOccupancyMaxActiveBlocksPerMultiprocessor doh_cudaOccupancy
#define cudaFuncSetAttribute doh_cudaFuncSetAttribute
#define cudaGetErrorString(e) "host emulation"
