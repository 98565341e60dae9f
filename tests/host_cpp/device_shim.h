// device_shim.h -- lets g++ compile the traversal headers of the product (rgk_b200/csrc/trace_device.cuh, bvh_device.cuh)
// for the HOST, so that the device source itself -- not a restatement of it -- can be checked against the oracle without a
// GPU (tests/test_device_on_host.py).  A "warp" has one lane here: ballots return bit 0, shuffles return their argument,
// and DevScene::refill_threshold = 1 makes the persistent-warp drivers refill after every ray.  The two inline-PTX helpers
// of trace_device.cuh have host branches (#ifndef __CUDA_ARCH__).  Test infrastructure only.
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstring>

#undef __device__
#undef __host__
#undef __global__
#undef __forceinline__
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#undef __noinline__
#define __noinline__

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline unsigned __ballot_sync(unsigned, int pred) { return pred ? 1u : 0u; }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p += v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p += v; return o; }
