// trace.cu -- batch entry points for closest-hit and shadow traversal
// (rgk_trace_closest / rgk_trace_shadow and their *_device variants).
//
// Persistent-thread kernels: the grid is sized to the machine (SMs x resident CTAs) and
// every warp pulls work in 32-ray packets from a global counter, so long rays do not hold a
// whole CTA's slot hostage.  Compiled with -fmad=false (see trace_device.cuh).
#include <cstdlib>
#include "trace_device.cuh"
#include "bvh_device.cuh"

namespace {

#ifndef RGK_MIN_BLOCKS
#define RGK_MIN_BLOCKS 1
#endif
#ifndef RGK_TRACE_THREADS_MAX
#define RGK_TRACE_THREADS_MAX 256
#endif
constexpr int TRACE_THREADS_MAX = RGK_TRACE_THREADS_MAX;

template <bool COUNT>
__device__ __forceinline__ void flush_counts(const TravCount& c, uint32_t nrays, rgk_trav_stats* stats) {
    if (!COUNT) return;
    // warp-aggregate, one atomic per counter per warp
    unsigned long long v[8] = {nrays, c.inner, c.leaf, c.refs, c.tests, c.exact, c.prefiltered, c.wrong};
#pragma unroll
    for (int k = 0; k < 8; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd((unsigned long long*)&stats->rays, v[0]);
        atomicAdd((unsigned long long*)&stats->inner, v[1]);
        atomicAdd((unsigned long long*)&stats->leaf, v[2]);
        atomicAdd((unsigned long long*)&stats->refs, v[3]);
        atomicAdd((unsigned long long*)&stats->tests, v[4]); atomicAdd((unsigned long long*)&stats->exact, v[5]);
        atomicAdd((unsigned long long*)&stats->prefiltered, v[6]); atomicAdd((unsigned long long*)&stats->prefilter_wrong, v[7]);
    }
}

template <bool COUNT, int VARIANT>
__global__ void __launch_bounds__(TRACE_THREADS_MAX, RGK_MIN_BLOCKS)
k_trace_closest(DevScene S, const rgk_ray* __restrict__ rays, const uint32_t* __restrict__ ignore, uint64_t n,
                rgk_hit* __restrict__ hits, rgk_trav_stats* stats, unsigned long long* next) {
    TravCount cnt{0, 0, 0, 0, 0, 0, 0};
    uint32_t mine = 0;
    trace_rays<VARIANT, false, COUNT>(S, (uint32_t)n, next, cnt, mine,
        [&](uint32_t i, Traverser<false, COUNT>& T) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(rays + i));
            const float4 b = __ldg(reinterpret_cast<const float4*>(rays + i) + 1);
            return T.init(S, a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, ignore ? __ldg(ignore + i) : RGK_NO_TRIANGLE);
        },
        [&](uint32_t i, bool found, const HitRec& h) {
            rgk_hit out;
            out.triangle = found ? h.tri : RGK_NO_TRIANGLE; out.t = found ? h.t : __int_as_float(0x7f800000);
            if (found) { out.a = 1.0f - h.alpha - h.beta; out.b = h.alpha; out.c = h.beta; }
            else { out.a = 0.0f; out.b = 0.0f; out.c = 0.0f; }
            hits[i] = out;
        });
    flush_counts<COUNT>(cnt, mine, stats);
}

template <bool COUNT, int VARIANT>
__global__ void __launch_bounds__(TRACE_THREADS_MAX, RGK_MIN_BLOCKS)
k_trace_shadow(DevScene S, const float* __restrict__ pa, const float* __restrict__ pb, uint64_t n,
               uint8_t* __restrict__ visible, rgk_trav_stats* stats, unsigned long long* next) {
    TravCount cnt{0, 0, 0, 0, 0, 0, 0};
    uint32_t mine = 0;
    trace_rays<VARIANT, true, COUNT>(S, (uint32_t)n, next, cnt, mine,
        [&](uint32_t i, Traverser<true, COUNT>& T) {
            // Ray(from, to, eps) (src/ray.hpp:15-22) + Scene::Visibility (src/scene.cpp:670-673)
            const float ax = pa[3 * (size_t)i], ay = pa[3 * (size_t)i + 1], az = pa[3 * (size_t)i + 2];
            const float ex = pb[3 * (size_t)i] - ax, ey = pb[3 * (size_t)i + 1] - ay, ez = pb[3 * (size_t)i + 2] - az;
            const float d2 = ex * ex + ey * ey + ez * ez;
            const float inv = 1.0f / sqrtf(d2), len = sqrtf(d2);
            const float e20 = S.epsilon * 20.0f;
            return T.init(S, ax, ay, az, ex * inv, ey * inv, ez * inv, 0.0f + e20, len - e20, RGK_NO_TRIANGLE);
        },
        [&](uint32_t i, bool found, const HitRec&) { visible[i] = found ? 0 : 1; });
    flush_counts<COUNT>(cnt, mine, stats);
}

// ---- wide BVH (RGK_TRAVERSAL_BVH, the default): BVH pass over all rays, kd pass over the deferred (ambiguous) ones ----------

template <bool COUNT>
__global__ void __launch_bounds__(TRACE_THREADS_MAX, RGK_MIN_BLOCKS)
k_bvh_closest(DevScene S, const rgk_ray* __restrict__ rays, const uint32_t* __restrict__ ignore, uint64_t n,
              rgk_hit* __restrict__ hits, BvhStats* stats, unsigned long long* next, uint32_t* __restrict__ list, uint32_t* list_count) {
    BvhCount cnt{0, 0, 0};
    uint32_t mine = 0, deferred = 0;
    trace_bvh<false, COUNT>(S, (uint32_t)n, next, cnt, mine, deferred,
        [&](uint32_t i, BvhTraverser<false, COUNT>& T) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(rays + i));
            const float4 b = __ldg(reinterpret_cast<const float4*>(rays + i) + 1);
            return T.init(S, a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, ignore ? __ldg(ignore + i) : RGK_NO_TRIANGLE);
        },
        [&](uint32_t i, bool found, const HitRec& h) {
            rgk_hit out;
            out.triangle = found ? h.tri : RGK_NO_TRIANGLE; out.t = found ? h.t : __int_as_float(0x7f800000);
            if (found) { out.a = 1.0f - h.alpha - h.beta; out.b = h.alpha; out.c = h.beta; }
            else { out.a = 0.0f; out.b = 0.0f; out.c = 0.0f; }
            hits[i] = out;
        },
        [&](uint32_t i) { list[atomicAdd(list_count, 1u)] = i; });
    flush_bvh_counts<COUNT>(cnt, mine, deferred, stats);
}

template <bool COUNT>
__global__ void __launch_bounds__(TRACE_THREADS_MAX, RGK_MIN_BLOCKS)
k_bvh_shadow(DevScene S, const float* __restrict__ pa, const float* __restrict__ pb, uint64_t n, uint8_t* __restrict__ visible,
             BvhStats* stats, unsigned long long* next, uint32_t* __restrict__ list, uint32_t* list_count) {
    BvhCount cnt{0, 0, 0};
    uint32_t mine = 0, deferred = 0;
    trace_bvh<true, COUNT>(S, (uint32_t)n, next, cnt, mine, deferred,
        [&](uint32_t i, BvhTraverser<true, COUNT>& T) {
            const float ax = pa[3 * (size_t)i], ay = pa[3 * (size_t)i + 1], az = pa[3 * (size_t)i + 2];
            const float ex = pb[3 * (size_t)i] - ax, ey = pb[3 * (size_t)i + 1] - ay, ez = pb[3 * (size_t)i + 2] - az;
            const float d2 = ex * ex + ey * ey + ez * ez;
            const float inv = 1.0f / sqrtf(d2), len = sqrtf(d2);
            const float e20 = S.epsilon * 20.0f;
            return T.init(S, ax, ay, az, ex * inv, ey * inv, ez * inv, 0.0f + e20, len - e20, RGK_NO_TRIANGLE);
        },
        [&](uint32_t i, bool found, const HitRec&) { visible[i] = found ? 0 : 1; },
        [&](uint32_t i) { list[atomicAdd(list_count, 1u)] = i; });
    flush_bvh_counts<COUNT>(cnt, mine, deferred, stats);
}

// the kd kernels over the deferred list (count read on the device: no host round trip between the two passes)
template <int VARIANT>
__global__ void __launch_bounds__(TRACE_THREADS_MAX, RGK_MIN_BLOCKS)
k_trace_closest_list(DevScene S, const rgk_ray* __restrict__ rays, const uint32_t* __restrict__ ignore, const uint32_t* __restrict__ list,
                     const uint32_t* __restrict__ list_count, rgk_hit* __restrict__ hits, unsigned long long* next) {
    TravCount cnt{0, 0, 0, 0, 0, 0, 0};
    uint32_t mine = 0;
    trace_rays<VARIANT, false, false>(S, *list_count, next, cnt, mine,
        [&](uint32_t k, Traverser<false, false>& T) {
            const uint32_t i = list[k];
            const float4 a = __ldg(reinterpret_cast<const float4*>(rays + i));
            const float4 b = __ldg(reinterpret_cast<const float4*>(rays + i) + 1);
            return T.init(S, a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, ignore ? __ldg(ignore + i) : RGK_NO_TRIANGLE);
        },
        [&](uint32_t k, bool found, const HitRec& h) {
            const uint32_t i = list[k];
            rgk_hit out;
            out.triangle = found ? h.tri : RGK_NO_TRIANGLE; out.t = found ? h.t : __int_as_float(0x7f800000);
            if (found) { out.a = 1.0f - h.alpha - h.beta; out.b = h.alpha; out.c = h.beta; }
            else { out.a = 0.0f; out.b = 0.0f; out.c = 0.0f; }
            hits[i] = out;
        });
}

template <int VARIANT>
__global__ void __launch_bounds__(TRACE_THREADS_MAX, RGK_MIN_BLOCKS)
k_trace_shadow_list(DevScene S, const float* __restrict__ pa, const float* __restrict__ pb, const uint32_t* __restrict__ list,
                    const uint32_t* __restrict__ list_count, uint8_t* __restrict__ visible, unsigned long long* next) {
    TravCount cnt{0, 0, 0, 0, 0, 0, 0};
    uint32_t mine = 0;
    trace_rays<VARIANT, true, false>(S, *list_count, next, cnt, mine,
        [&](uint32_t k, Traverser<true, false>& T) {
            const uint32_t i = list[k];
            const float ax = pa[3 * (size_t)i], ay = pa[3 * (size_t)i + 1], az = pa[3 * (size_t)i + 2];
            const float ex = pb[3 * (size_t)i] - ax, ey = pb[3 * (size_t)i + 1] - ay, ez = pb[3 * (size_t)i + 2] - az;
            const float d2 = ex * ex + ey * ey + ez * ez;
            const float inv = 1.0f / sqrtf(d2), len = sqrtf(d2);
            const float e20 = S.epsilon * 20.0f;
            return T.init(S, ax, ay, az, ex * inv, ey * inv, ez * inv, 0.0f + e20, len - e20, RGK_NO_TRIANGLE);
        },
        [&](uint32_t k, bool found, const HitRec&) { visible[list[k]] = found ? 0 : 1; });
}

} // namespace

namespace {
// rgk_device_cfg::trace_threads / kd_variant select the CTA size and the kd control structure of the batch entry points
// (A/B knobs; results are identical)
int trace_grid(rgk_context* ctx, int threads) {
    int sms = 148, per = 8;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_trace_closest<false, 2>, threads, 0);
    return sms * (per > 0 ? per : 1);
}

} // namespace

rgk_status launch_trace_closest(rgk_context* ctx, const rgk_ray* d_rays, const uint32_t* d_ignore, uint64_t n,
                                rgk_hit* d_hits, rgk_trav_stats* d_stats) {
    if (n == 0) return RGK_OK;
    if (n > 0xFFFFFFF0ull) return rgk_fail(ctx, RGK_ERR_UNSUPPORTED, "more than 2^32 rays in one batch");
    unsigned long long* next = (unsigned long long*)rgk_scratch(ctx, 3, 256);
    if (!next) return rgk_fail(ctx, RGK_ERR_NOMEM, "scratch allocation failed");
    RGK_CUDA(ctx, cudaMemsetAsync(next, 0, 24, ctx->stream));
    const uint64_t warps = (n + 31) / 32;
    const int threads = (int)ctx->cfg.trace_threads, variant = (int)ctx->cfg.kd_variant;
    const int grid = (int)std::min<uint64_t>(trace_grid(ctx, threads), (warps + threads / 32 - 1) / (threads / 32));
    if (ctx->dev.bvh_nodes && !d_stats) {          // wide BVH for every ray, then the kd-tree for the deferred ones
        uint32_t* list = (uint32_t*)rgk_scratch(ctx, 4, n * 4);
        if (!list) return rgk_fail(ctx, RGK_ERR_NOMEM, "scratch allocation failed");
        uint32_t* list_count = (uint32_t*)(next + 2);
        if (ctx->counting) k_bvh_closest<true><<<grid, threads, 0, ctx->stream>>>(ctx->dev, d_rays, d_ignore, n, d_hits, ctx->d_bvh_stats, next, list, list_count);
        else k_bvh_closest<false><<<grid, threads, 0, ctx->stream>>>(ctx->dev, d_rays, d_ignore, n, d_hits, ctx->d_bvh_stats, next, list, list_count);
        const int g2 = std::min(grid, 148);
        if (variant == 2) k_trace_closest_list<2><<<g2, threads, 0, ctx->stream>>>(ctx->dev, d_rays, d_ignore, list, list_count, d_hits, next + 1);
        else k_trace_closest_list<6><<<g2, threads, 0, ctx->stream>>>(ctx->dev, d_rays, d_ignore, list, list_count, d_hits, next + 1);
        ctx->launches += 2;
        RGK_CUDA(ctx, cudaGetLastError());
        return RGK_OK;
    }
    if (d_stats) k_trace_closest<true, 2><<<grid, threads, 0, ctx->stream>>>(ctx->dev, d_rays, d_ignore, n, d_hits, d_stats, next);
    else if (variant == 6) k_trace_closest<false, 6><<<grid, threads, 0, ctx->stream>>>(ctx->dev, d_rays, d_ignore, n, d_hits, nullptr, next);
    else if (variant == 2) k_trace_closest<false, 2><<<grid, threads, 0, ctx->stream>>>(ctx->dev, d_rays, d_ignore, n, d_hits, nullptr, next);
    ctx->launches++;
    RGK_CUDA(ctx, cudaGetLastError());
    return RGK_OK;
}

rgk_status launch_trace_shadow(rgk_context* ctx, const float* d_a, const float* d_b, uint64_t n,
                               uint8_t* d_visible, rgk_trav_stats* d_stats) {
    if (n == 0) return RGK_OK;
    if (n > 0xFFFFFFF0ull) return rgk_fail(ctx, RGK_ERR_UNSUPPORTED, "more than 2^32 rays in one batch");
    unsigned long long* next = (unsigned long long*)rgk_scratch(ctx, 3, 256);
    if (!next) return rgk_fail(ctx, RGK_ERR_NOMEM, "scratch allocation failed");
    RGK_CUDA(ctx, cudaMemsetAsync(next, 0, 24, ctx->stream));
    const uint64_t warps = (n + 31) / 32;
    const int threads = (int)ctx->cfg.trace_threads, variant = (int)ctx->cfg.kd_variant;
    const int grid = (int)std::min<uint64_t>(trace_grid(ctx, threads), (warps + threads / 32 - 1) / (threads / 32));
    if (ctx->dev.bvh_nodes && !d_stats) {
        uint32_t* list = (uint32_t*)rgk_scratch(ctx, 4, n * 4);
        if (!list) return rgk_fail(ctx, RGK_ERR_NOMEM, "scratch allocation failed");
        uint32_t* list_count = (uint32_t*)(next + 2);
        if (ctx->counting) k_bvh_shadow<true><<<grid, threads, 0, ctx->stream>>>(ctx->dev, d_a, d_b, n, d_visible, ctx->d_bvh_stats + 1, next, list, list_count);
        else k_bvh_shadow<false><<<grid, threads, 0, ctx->stream>>>(ctx->dev, d_a, d_b, n, d_visible, ctx->d_bvh_stats + 1, next, list, list_count);
        const int g2 = std::min(grid, 148);
        if (variant == 2) k_trace_shadow_list<2><<<g2, threads, 0, ctx->stream>>>(ctx->dev, d_a, d_b, list, list_count, d_visible, next + 1);
        else k_trace_shadow_list<6><<<g2, threads, 0, ctx->stream>>>(ctx->dev, d_a, d_b, list, list_count, d_visible, next + 1);
        ctx->launches += 2;
        RGK_CUDA(ctx, cudaGetLastError());
        return RGK_OK;
    }
    if (d_stats) k_trace_shadow<true, 2><<<grid, threads, 0, ctx->stream>>>(ctx->dev, d_a, d_b, n, d_visible, d_stats, next);
    else if (variant == 6) k_trace_shadow<false, 6><<<grid, threads, 0, ctx->stream>>>(ctx->dev, d_a, d_b, n, d_visible, nullptr, next);
    else if (variant == 2) k_trace_shadow<false, 2><<<grid, threads, 0, ctx->stream>>>(ctx->dev, d_a, d_b, n, d_visible, nullptr, next);
    ctx->launches++;
    RGK_CUDA(ctx, cudaGetLastError());
    return RGK_OK;
}
