// trace_device.cuh -- kd-tree traversal + triangle test, device side.
//
// Follows Scene::FindIntersectKdOtherThan (src/scene_intersect.cpp:211-327) and
// Triangle::TestIntersection (src/primitives.cpp:75-166) operation for operation: this
// translation unit is compiled with -fmad=false (the reference's x86-64 build has no FMA),
// IEEE division / sqrt (nvcc defaults), and the plane distance is evaluated in fp64 exactly
// as the reference does, so hit triangle, t and barycentrics are bit-identical to the CPU.
#pragma once
#include "rgk_internal.h"

struct TravCount { uint32_t inner, leaf, refs, tests; };

struct HitRec { uint32_t tri; float t, alpha, beta; };  // alpha/beta as returned by TestIntersection

// Triangle::TestIntersection with the ray-independent part read from the 48-byte record.
__device__ __forceinline__ bool tri_test(const float4* __restrict__ rec, float eps,
                                         float ox, float oy, float oz, float dx, float dy, float dz,
                                         float& t, float& alpha, float& beta) {
    const float4 r0 = __ldg(rec);
    const double dt = (double)(dx * r0.x + dy * r0.y + dz * r0.z);
    if (dt != dt) return false;                                   // std::isnan(dot)
    if (dt < (double)eps && dt > (double)(-eps)) return false;   // ray parallel to the plane
    const double dot2 = (double)(ox * r0.x + oy * r0.y + oz * r0.z);
    t = (float)(-((double)r0.w + dot2) / dt);
    const float4 r1 = __ldg(rec + 1);
    const float4 r2 = __ldg(rec + 2);
    const uint32_t flags = __float_as_uint(r2.w);
    const uint32_t code = flags & 3u;
    const float o1 = (code == 0u) ? oy : ox, d1 = (code == 0u) ? dy : dx;
    const float o2 = (code == 2u) ? oy : oz, d2 = (code == 2u) ? dy : dz;
    const float q0x = (o1 + d1 * t) - r1.x;
    const float q0y = (o2 + d2 * t) - r1.y;
    if (flags & 4u) {                                              // |q1.x| < eps: uncommon case
        beta = q0x / r2.x;
        if (beta < 0.0f || beta > 1.0f) return false;
        alpha = (q0y - beta * r2.y) / r1.w;
    } else {
        beta = (q0y * r1.z - q0x * r1.w) / r2.z;
        if (beta < 0.0f || beta > 1.0f) return false;
        alpha = (q0x - beta * r2.x) / r1.z;
    }
    if (alpha < 0.0f || (alpha + beta) > 1.0f) return false;
    return true;
}

// ANY = false: closest hit with the reference's first-hit-leaf early exit.
// ANY = true : Scene::Visibility's boolean -- "some triangle is accepted in some visited
//              leaf" -- which is traversal-order independent, so the first accepted hit ends it (A3).
template <bool ANY, bool COUNT>
__device__ __forceinline__ bool kd_traverse(const DevScene& S, float ox, float oy, float oz,
                                            float dx, float dy, float dz, float tnear, float tfar,
                                            uint32_t ignore, HitRec& res, TravCount& cnt) {
    res.tri = RGK_NO_TRIANGLE; res.t = __int_as_float(0x7f800000); res.alpha = 0.0f; res.beta = 0.0f;
    const float o[3] = {ox, oy, oz}, d[3] = {dx, dy, dz};
    float inv[3];
    float t0 = tnear, t1 = tfar;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        inv[i] = 1.f / d[i];
        float tn = (S.bb[2 * i] - o[i]) * inv[i];
        float tf = (S.bb[2 * i + 1] - o[i]) * inv[i];
        if (tn > tf) { const float tmp = tn; tn = tf; tf = tmp; }
        t0 = tn > t0 ? tn : t0;
        t1 = tf < t1 ? tf : t1;
        if (t0 > t1) return false;
    }
    uint32_t st_node[RGK_STACK_CAP];
    float st_tmin[RGK_STACK_CAP], st_tmax[RGK_STACK_CAP];
    int sp = 1;
    st_node[0] = 0u; st_tmin[0] = t0; st_tmax[0] = t1;
    const float eps = S.epsilon;
    while (sp > 0) {
        --sp;
        uint32_t node = st_node[sp];
        float tmin = st_tmin[sp], tmax = st_tmax[sp];
        if (tfar < tmin) break;
        // descend without touching the stack while only one child is visited
        for (;;) {
            const uint2 w = __ldg(S.nodes + node);
            if ((w.y & 3u) == 3u) {
                if (COUNT) cnt.leaf++;
                bool hit = false;
                const uint32_t n = w.y >> 2, start = w.x;
                for (uint32_t p = 0; p < n; p++) {
                    const uint32_t ti = __ldg(S.refs + start + p);
                    if (COUNT) cnt.refs++;
                    if (ti == ignore) continue;
                    if (COUNT) cnt.tests++;
                    float t, a, b;
                    if (tri_test(S.tri_isect + 3 * (size_t)ti, eps, ox, oy, oz, dx, dy, dz, t, a, b)) {
                        if (t < tmin - eps || t > tmax + eps) continue;
                        if (ANY) { res.tri = ti; res.t = t; res.alpha = a; res.beta = b; return true; }
                        if (t < res.t) { res.tri = ti; res.t = t; res.alpha = a; res.beta = b; hit = true; }
                    }
                }
                if (hit) return true;
                break;
            }
            if (COUNT) cnt.inner++;
            const uint32_t axis = w.y & 3u;
            const float split = __uint_as_float(w.x);
            const float oa = axis == 0u ? ox : (axis == 1u ? oy : oz);
            const float da = axis == 0u ? dx : (axis == 1u ? dy : dz);
            const float ia = axis == 0u ? inv[0] : (axis == 1u ? inv[1] : inv[2]);
            const float tplane = (split - oa) * ia;
            const bool below_first = (oa < split) || (oa == split && da <= 0.0f);
            const uint32_t first = below_first ? node + 1u : (w.y >> 2);
            const uint32_t second = below_first ? (w.y >> 2) : node + 1u;
            if (tplane > tmax || tplane <= 0.0f) {
                node = first;                       // popping it next would re-check tfar < tmin: unchanged tmin
            } else if (tplane < tmin) {
                node = second;
            } else {
                st_node[sp] = second; st_tmin[sp] = tplane; st_tmax[sp] = tmax; ++sp;
                node = first; tmax = tplane;
            }
        }
    }
    return false;
}
