// device_shim.h -- lets g++ compile the traversal headers of the product (rgk_b200/csrc/trace_device.cuh, bvh_device.cuh)
// for the HOST, so that the device source itself -- not a restatement of it -- can be checked against the oracle without a
// GPU (tests/test_device_on_host.py).  A "warp" has one lane here: ballots return bit 0, shuffles return their argument,
// and DevScene::refill_threshold = 1 makes the persistent-warp drivers refill after every ray.  The two inline-PTX helpers
// of trace_device.cuh have host branches (#ifndef __CUDA_ARCH__).  For the wavefront renderer (render.cu, host loop included:
// gen_host_sources.py rewrites its launches into doh_launch) a kernel launch is a loop over blocks and threads, and the CUDA
// runtime calls of the host loop are mapped onto malloc / memcpy / no-ops.  Test infrastructure only.
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstring>

#include <cstdlib>
#include <utility>

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

template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }      // streaming store: cache hint only
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((unsigned long long)a * b) >> 32); }
#define RGK_WARP_LANES 1                                                  // k_sampler_warp: batches of one draw
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline unsigned __ballot_sync(unsigned, int pred) { return pred ? 1u : 0u; }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T, int) { return T(0); }     // another lane: there is none (all uses are sums)
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p += v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p += v; return o; }
#include <algorithm>
using std::isnan; using std::min; using std::max;       // CUDA has these in the global namespace
static inline float __uint2float_rn(uint32_t x) { return (float)x; }
static inline float atomicAdd(float* p, float v) { const float o = *p; *p += v; return o; }
static inline void __syncthreads() {}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned) { return v; }

// ---- launch geometry: mutable stand-ins for the built-in variables
struct doh_idx { unsigned x, y, z; };
inline doh_idx doh_threadIdx{0, 0, 0}, doh_blockIdx{0, 0, 0}, doh_blockDim{1, 1, 1}, doh_gridDim{1, 1, 1};
#define threadIdx doh_threadIdx
#define blockIdx doh_blockIdx
#define blockDim doh_blockDim
#define gridDim doh_gridDim
template <class... KArgs, class... Args>
inline void doh_launch(void (*kernel)(KArgs...), size_t grid, size_t block, size_t, cudaStream_t, Args&&... args) {
    doh_gridDim = {(unsigned)grid, 1, 1}; doh_blockDim = {(unsigned)block, 1, 1};
    for (unsigned b = 0; b < (unsigned)grid; b++)
        for (unsigned t = 0; t < (unsigned)block; t++) { doh_blockIdx.x = b; doh_threadIdx.x = t; kernel(args...); }
    doh_blockIdx.x = 0; doh_threadIdx.x = 0; doh_blockDim = {1, 1, 1}; doh_gridDim = {1, 1, 1};
}

// ---- the CUDA runtime calls of the host loop
static inline cudaError_t doh_cudaMalloc(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t doh_cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t doh_cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t = nullptr) { std::memset(p, v, n); return cudaSuccess; }
static inline cudaError_t doh_cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t doh_cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t doh_ok() { return cudaSuccess; }
static inline cudaError_t doh_cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t doh_cudaEventOp(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }     // arguments are evaluated
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
