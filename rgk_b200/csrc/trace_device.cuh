// trace_device.cuh -- kd-tree traversal + triangle test, device side.
//
// Follows Scene::FindIntersectKdOtherThan (src/scene_intersect.cpp:211-327) and
// Triangle::TestIntersection (src/primitives.cpp:75-166): this translation unit is compiled
// with -fmad=false (the reference's x86-64 build has no FMA), IEEE division / sqrt (nvcc
// defaults), and the plane distance is evaluated in fp64 exactly as the reference does, so
// hit triangle, t and barycentrics are bit-identical to the CPU.  What differs is the ORDER
// in which independent accept conditions are evaluated (cheap rejections first) and the
// control structure (persistent warps, descend-to-leaf / process-leaf phases, idle lanes
// refilled from a global counter) -- neither changes any result.
#pragma once
#include "rgk_internal.h"

struct TravCount { uint32_t inner, leaf, refs, tests; };
// Per-ray traversal stack (local memory, lane-interleaved by the hardware); kept outside the Traverser so that the
// scalar ray / interval state stays in registers.
struct TravStack { uint32_t node[RGK_STACK_CAP]; float tmin[RGK_STACK_CAP], tmax[RGK_STACK_CAP]; };
struct HitRec { uint32_t tri; float t, alpha, beta; };  // alpha/beta as returned by TestIntersection

#define RGK_REFILL_THRESHOLD 8   // refill when at least this many lanes of the warp are idle

template <bool ANY, bool COUNT>
struct Traverser {
    float ox, oy, oz, dx, dy, dz, ix, iy, iz, tfar;
    uint32_t ignore;
    uint32_t node; float tmin, tmax;
    int sp;
    HitRec res;

    // Root slab test (src/scene_intersect.cpp:223-232). false: the ray misses the scene box.
    __device__ __forceinline__ bool init(const DevScene& S, float ox_, float oy_, float oz_, float dx_, float dy_, float dz_,
                                         float tnear, float tfar_, uint32_t ignore_) {
        ox = ox_; oy = oy_; oz = oz_; dx = dx_; dy = dy_; dz = dz_; tfar = tfar_; ignore = ignore_;
        res.tri = RGK_NO_TRIANGLE; res.t = __int_as_float(0x7f800000); res.alpha = 0.0f; res.beta = 0.0f;
        ix = 1.f / dx; iy = 1.f / dy; iz = 1.f / dz;
        float t0 = tnear, t1 = tfar_;
        {
            float tn = (S.bb[0] - ox) * ix, tf = (S.bb[1] - ox) * ix;
            if (tn > tf) { const float q = tn; tn = tf; tf = q; }
            t0 = tn > t0 ? tn : t0; t1 = tf < t1 ? tf : t1;
            if (t0 > t1) return false;
        }
        {
            float tn = (S.bb[2] - oy) * iy, tf = (S.bb[3] - oy) * iy;
            if (tn > tf) { const float q = tn; tn = tf; tf = q; }
            t0 = tn > t0 ? tn : t0; t1 = tf < t1 ? tf : t1;
            if (t0 > t1) return false;
        }
        {
            float tn = (S.bb[4] - oz) * iz, tf = (S.bb[5] - oz) * iz;
            if (tn > tf) { const float q = tn; tn = tf; tf = q; }
            t0 = tn > t0 ? tn : t0; t1 = tf < t1 ? tf : t1;
            if (t0 > t1) return false;
        }
        sp = 0; node = 0u; tmin = t0; tmax = t1;
        return !(tfar < tmin);          // "if(r.far < tmin) break" on the root pop (:253)
    }

    // Inner-node steps until `node` is a leaf (src/scene_intersect.cpp:294-321). Pushing one child and popping it
    // straight away, as the reference does, is the same as stepping into it: the re-checked tfar < tmin is unchanged.
    __device__ __forceinline__ uint2 descend(const DevScene& S, TravStack& K, TravCount& cnt) {
        uint2 w = __ldg(S.nodes + node);
        while ((w.y & 3u) != 3u) {
            if (COUNT) cnt.inner++;
            const uint32_t axis = w.y & 3u;
            const float split = __uint_as_float(w.x);
            const float oa = axis == 0u ? ox : (axis == 1u ? oy : oz);
            const float da = axis == 0u ? dx : (axis == 1u ? dy : dz);
            const float ia = axis == 0u ? ix : (axis == 1u ? iy : iz);
            const float tplane = (split - oa) * ia;
            const bool below_first = (oa < split) || (oa == split && da <= 0.0f);
            const uint32_t other = w.y >> 2;
            const uint32_t first = below_first ? node + 1u : other;
            const uint32_t second = below_first ? other : node + 1u;
            if (tplane > tmax || tplane <= 0.0f) node = first;
            else if (tplane < tmin) node = second;
            else {
                K.node[sp] = second; K.tmin[sp] = tplane; K.tmax[sp] = tmax; ++sp;
                node = first; tmax = tplane;
            }
            w = __ldg(S.nodes + node);
        }
        return w;
    }

    // One leaf (src/scene_intersect.cpp:255-292). true: traversal is over (closest: this leaf produced the hit;
    // ANY: some triangle was accepted).  The acceptance conditions of one triangle -- plane not parallel, t inside
    // [tmin-eps, tmax+eps], t < best so far, barycentrics inside -- are a pure conjunction, so they are evaluated
    // cheapest-first; the values they are evaluated ON are computed exactly as Triangle::TestIntersection does.
    __device__ __forceinline__ bool leaf(const DevScene& S, uint2 w, TravCount& cnt) {
        if (COUNT) cnt.leaf++;
        const float eps = S.epsilon;
        const float lo = tmin - eps, hi = tmax + eps;
        bool hit = false;
        const uint32_t n = w.y >> 2, start = w.x;
        for (uint32_t p = 0; p < n; p++) {
            const uint32_t ti = __ldg(S.refs + start + p);
            if (COUNT) cnt.refs++;
            if (ti == ignore) continue;
            if (COUNT) cnt.tests++;
            const float4* rec = S.tri_isect + 3 * (size_t)ti;
            const float4 r0 = __ldg(rec);
            const double dt = (double)(dx * r0.x + dy * r0.y + dz * r0.z);
            if (dt != dt) continue;                                      // std::isnan(dot)
            if (dt < (double)eps && dt > (double)(-eps)) continue;      // parallel to the plane
            const double dot2 = (double)(ox * r0.x + oy * r0.y + oz * r0.z);
            const float t = (float)(-((double)r0.w + dot2) / dt);
            if (t < lo || t > hi) continue;                              // outside this node's interval (:272)
            if (!(t < res.t)) continue;                                  // not closer than the best so far (:275)
            const float4 r1 = __ldg(rec + 1);
            const float4 r2 = __ldg(rec + 2);
            const uint32_t flags = __float_as_uint(r2.w);
            const uint32_t code = flags & 3u;
            const float o1 = (code == 0u) ? oy : ox, d1 = (code == 0u) ? dy : dx;
            const float o2 = (code == 2u) ? oy : oz, d2 = (code == 2u) ? dy : dz;
            const float q0x = (o1 + d1 * t) - r1.x;
            const float q0y = (o2 + d2 * t) - r1.y;
            float alpha, beta;
            if (flags & 4u) {                                            // |q1.x| < eps: uncommon case
                beta = q0x / r2.x;
                if (beta < 0.0f || beta > 1.0f) continue;
                alpha = (q0y - beta * r2.y) / r1.w;
            } else {
                beta = (q0y * r1.z - q0x * r1.w) / r2.z;
                if (beta < 0.0f || beta > 1.0f) continue;
                alpha = (q0x - beta * r2.x) / r1.z;
            }
            if (alpha < 0.0f || (alpha + beta) > 1.0f) continue;
            res.tri = ti; res.t = t; res.alpha = alpha; res.beta = beta;
            if (ANY) return true;
            hit = true;
        }
        return hit;
    }

    // Next stack entry. false: nothing left, or the whole traversal ends because tfar < tmin (:253).
    __device__ __forceinline__ bool pop(const TravStack& K) {
        if (sp == 0) return false;
        --sp;
        node = K.node[sp]; tmin = K.tmin[sp]; tmax = K.tmax[sp];
        return !(tfar < tmin);
    }
};

// Persistent-warp driver.  `fetch(i, T)` loads work item i and calls T.init(...) (returns its result);
// `commit(i, found, res)` stores the result.  Every lane owns one ray at a time; lanes whose ray is finished
// stay idle until at least RGK_REFILL_THRESHOLD lanes of the warp are idle (or all are), then the warp grabs
// that many new items from the global counter with one atomic (warp-ballot work redistribution).
template <bool ANY, bool COUNT, class Fetch, class Commit>
__device__ __forceinline__ void trace_persistent(const DevScene& S, uint32_t count, unsigned long long* work,
                                                 TravCount& cnt, uint32_t& done, Fetch fetch, Commit commit) {
    Traverser<ANY, COUNT> T;
    TravStack K;
    const unsigned lane = threadIdx.x & 31;
    bool active = false, exhausted = false;
    uint32_t item = 0;
    for (;;) {
        __syncwarp();
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (idle != 0u && !exhausted && (__popc(idle) >= RGK_REFILL_THRESHOLD || idle == 0xffffffffu)) {
            const int leader = __ffs(idle) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(work, (unsigned long long)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (base + __popc(idle) >= count) exhausted = true;
            if (!active) {
                const unsigned long long mine = base + __popc(idle & ((1u << lane) - 1u));
                if (mine < count) {
                    item = (uint32_t)mine;
                    done++;
                    if (fetch(item, T)) active = true;
                    else commit(item, false, T.res);
                }
            }
        }
        if (__ballot_sync(0xffffffffu, active) == 0u) { if (exhausted) break; else continue; }
        if (active) {
            const uint2 w = T.descend(S, K, cnt);
            if (T.leaf(S, w, cnt)) { commit(item, true, T.res); active = false; }
            else if (!T.pop(K)) { commit(item, false, T.res); active = false; }
        }
    }
}
