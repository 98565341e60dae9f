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

struct TravCount { uint32_t inner, leaf, refs, tests, exact, prefiltered, wrong; };
// Per-ray traversal stack (local memory, lane-interleaved by the hardware); kept outside the Traverser so that the
// scalar ray / interval state stays in registers.  An entry is (far child, bits of its tmin): the reference's
// NodeToDo{node, tmin, tmax} (src/scene_intersect.cpp:242) without tmax, because a pushed child's interval is
// [tplane, current tmax] and the current interval then becomes [tmin, tplane], so by induction the tmax of entry k is
// the tmin of entry k-1 (the root's tmax for k = 0) -- one 8-byte store per push instead of a 4- and an 8-byte one.
struct TravStack { uint2 e[RGK_STACK_CAP]; };
#ifndef RGK_CAND_CAP
#define RGK_CAND_CAP 4
#endif
// RGK_CAND_CAP:           // deferred exact tests per leaf before a flush
struct HitRec { uint32_t tri; float t, alpha, beta; };  // alpha/beta as returned by TestIntersection

// x, y or z by axis code as two predicated selects (the ternary chain compiles to a branch diamond with its own
// reconvergence barrier otherwise; this sits in the innermost loop of the issue-bound traversal)
__device__ __forceinline__ float sel_axis(uint32_t axis, float x, float y, float z) {
#ifdef __CUDA_ARCH__
    float r;
    asm("{\n\t.reg .pred p0, p1;\n\tsetp.eq.u32 p0, %1, 0;\n\tsetp.eq.u32 p1, %1, 1;\n\t"
        "selp.f32 %0, %3, %4, p1;\n\tselp.f32 %0, %2, %0, p0;\n\t}"
        : "=f"(r) : "r"(axis), "f"(x), "f"(y), "f"(z));
    return r;
#else
    return axis == 0u ? x : (axis == 1u ? y : z);      // host build of this header (tests/host_cpp/device_on_host.cpp)
#endif
}
// approximate reciprocal of the conservative pre-rejections (never part of a result)
__device__ __forceinline__ float rcp_approx(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

// three-input max / min (FMNMX3 on sm_100a); NaN operands are ignored like fmaxf / fminf ignore them
__device__ __forceinline__ float max3f(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
#else
    return fmaxf(fmaxf(a, b), c);
#endif
}
__device__ __forceinline__ float min3f(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
#else
    return fminf(fminf(a, b), c);
#endif
}

template <bool ANY, bool COUNT>
struct Traverser {
    float ox, oy, oz, dx, dy, dz, ix, iy, iz, tfar, troot;
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
        sp = 0; node = 0u; tmin = t0; tmax = t1; troot = t1;
        return !(tfar < tmin);          // "if(r.far < tmin) break" on the root pop (:253)
    }

    // Inner-node steps until `node` is a leaf (src/scene_intersect.cpp:294-321). Pushing one child and popping it
    // straight away, as the reference does, is the same as stepping into it: the re-checked tfar < tmin is unchanged.
    __device__ __forceinline__ uint2 descend(const DevScene& S, TravStack& K, TravCount& cnt) {
        uint2 w = __ldg(S.nodes + node);
        while ((w.y & 3u) != 3u) {
            if (COUNT) cnt.inner++;
            const uint32_t other = w.y >> 2, near = node + 1u;
            const uint32_t axis = w.y & 3u;
            const float split = __uint_as_float(w.x);
            const float oa = sel_axis(axis, ox, oy, oz);
            const float ia = sel_axis(axis, ix, iy, iz);
            const float diff = split - oa;                 // its sign is the exact sign of (split - oa)
            const float tplane = diff * ia;
            bool below_first = diff > 0.0f;                // oa < split
            if (diff == 0.0f) below_first = sel_axis(axis, dx, dy, dz) <= 0.0f;   // oa == split: decided by the direction (rare)
            const uint32_t first = below_first ? near : other;
            const uint32_t second = below_first ? other : near;
            node = first;
            if (tplane > tmax || tplane <= 0.0f) {}
            else if (tplane < tmin) node = second;
            else {
                K.e[sp] = make_uint2(second, __float_as_uint(tplane)); ++sp;
                tmax = tplane;
            }
            w = __ldg(S.nodes + node);
        }
        return w;
    }

    // Conservative fp32 pre-rejection bounds for the references of the current leaf.  The scan computes
    // t32 = -(w + dot2) * rcp.approx(dot): the sum is one fp32 rounding of the exact sum (2^-24), the approximate
    // reciprocal and the product add < 2^-22, and the reference's own t is the exact quotient rounded (2^-24), so
    // |t32 - t| < 2^-21 |t|.  A triangle the exact test accepts has lo <= t <= hi, hence
    // t32 >= lo - 2^-21 |lo| and t32 <= hi + 2^-21 |hi|; the bounds below are widened by 2^-18 (and 1e-30 against
    // flush-to-zero), so "t32 < lo_c or t32 > hi_c" never rejects what the exact test accepts.  Overflow gives +-inf
    // (correctly outside), NaN compares false (kept for the exact test).
    __device__ __forceinline__ void leaf_bounds(float eps, float& lo, float& hi, float& lo_c, float& hi_c) const {
        lo = tmin - eps; hi = tmax + eps;
        lo_c = (lo - fabsf(lo) * 3.814697265625e-6f) - 1e-30f;
        hi_c = (hi + fabsf(hi) * 3.814697265625e-6f) + 1e-30f;
    }
    // one reference of the leaf scan: true = survives the pre-rejection (needs the exact test); t32 = its approximate t
    __device__ __forceinline__ bool prescreen(const float4 r0, float eps, float lo_c, float hi_c, float& t32) const {
        const float dtf = dx * r0.x + dy * r0.y + dz * r0.z;
        const float dot2f = ox * r0.x + oy * r0.y + oz * r0.z;
        const float rcp = rcp_approx(dtf);
        t32 = -(r0.w + dot2f) * rcp;
        return !((fabsf(dtf) < eps) || (t32 < lo_c) || (t32 > hi_c));
    }

    // Second conservative filter, applied to the survivors of the scan before the exact arithmetic: the approximate hit
    // point o + d t32, projected like TestIntersection projects it, against the triangle's widened 2-D bounds
    // (DevScene::ref_bounds, built by tri_prefilter_bounds in host_scene.cpp, which bounds the test's own rounding).
    // The run-time margin covers the difference between this point and the one the exact test computes:
    // |t32 - t| < 2^-21 |t| moves it by < 2^-21 |t| (|d| = 1), and the two evaluations of o + d t round differently by
    // < 2^-22 (|o| + |t|); 2^-17 (|t32| + |ox| + |oy| + |oz|) is 16x that.  NaNs compare false (candidate kept).
    // Most survivors of the scan cross the plane inside the leaf but far from the triangle: this settles them with one
    // 16-byte load instead of the fp64 divide and two more record loads.
    __device__ __forceinline__ bool outside_bounds(const DevScene& S, uint32_t p, float t32) const {
        const float4 b = __ldg(S.ref_bounds + p);
        const uint32_t code = __float_as_uint(b.x) & 3u;
        const float o1 = (code == 0u) ? oy : ox, d1 = (code == 0u) ? dy : dx;
        const float o2 = (code == 2u) ? oy : oz, d2 = (code == 2u) ? dy : dz;
        const float m = (((fabsf(t32) + fabsf(ox)) + fabsf(oy)) + fabsf(oz)) * 7.62939453125e-6f;
        const float p1 = __fmaf_rn(d1, t32, o1), p2 = __fmaf_rn(d2, t32, o2);
        return (p1 + m < b.x) || (p1 - m > b.y) || (p2 + m < b.z) || (p2 - m > b.w);
    }
    // one survivor of the scan: pre-filter, then the exact test
    __device__ __forceinline__ bool candidate(const DevScene& S, uint2 c, float lo, float hi, TravCount& cnt) {
        const bool out = outside_bounds(S, c.x, __uint_as_float(c.y));
        if (COUNT) {                                  // counting instantiation: also proves the filter right
            if (__ldg(S.refs + c.x) == ignore) return false;
            if (out) cnt.prefiltered++; else cnt.exact++;
            const bool acc = exact_test(S, c.x, lo, hi);
            if (acc && out) cnt.wrong++;
            return acc;
        }
        if (out) return false;
        return exact_test(S, c.x, lo, hi);
    }

    // Triangle::TestIntersection proper for the reference at position p of the reference list that survived the
    // pre-rejection, on the reference's operation order (src/primitives.cpp:85-164), followed by the leaf's accept
    // rule (src/scene_intersect.cpp:261,272-283).  The ignored triangle is filtered here, not in the scan, so that
    // the scan reads one 16-byte plane per reference and nothing else.
    __device__ __forceinline__ bool exact_test(const DevScene& S, uint32_t p, float lo, float hi) {
        const uint32_t ti = __ldg(S.refs + p);
        if (ti == ignore) return false;
        const float4 r0 = __ldg(S.ref_planes + p);
        const float dtf = dx * r0.x + dy * r0.y + dz * r0.z;             // glm::dot(direction, planeN), fp32
        if (dtf != dtf) return false;                                    // std::isnan(dot)
        const float dot2f = ox * r0.x + oy * r0.y + oz * r0.z;
        const float t = (float)(-((double)r0.w + (double)dot2f) / (double)dtf);
        if (t < lo || t > hi) return false;                              // outside this node's interval (:272)
        if (!(t < res.t)) return false;                                  // not closer than the best so far (:275)
        const float4* rec = S.tri_isect + 3 * (size_t)ti;
        const float4 r1 = __ldg(rec + 1);
        const float4 r2 = __ldg(rec + 2);
        const uint32_t flags = __float_as_uint(r2.w);
        const uint32_t code = flags & 3u;
        const float o1 = (code == 0u) ? oy : ox, d1 = (code == 0u) ? dy : dx;
        const float o2 = (code == 2u) ? oy : oz, d2 = (code == 2u) ? dy : dz;
        const float q0x = (o1 + d1 * t) - r1.x;
        const float q0y = (o2 + d2 * t) - r1.y;
        float alpha, beta;
        if (flags & 4u) {                                                // |q1.x| < eps: uncommon case
            beta = q0x / r2.x;
            if (beta < 0.0f || beta > 1.0f) return false;
            alpha = (q0y - beta * r2.y) / r1.w;
        } else {
            const float num = q0y * r1.z - q0x * r1.w, den = r2.z;
            // RN(num/den) < 0 or > 1 decided without dividing when num, den are ordinary numbers: the rounded quotient
            // is > 1 exactly when |num| > |den| with equal signs (the next float above |den| already gives a quotient
            // > 1 + 2^-24), and < 0 exactly when the signs differ and the quotient does not round to -0 (|num| > |den| *
            // 2^-100 with |den| in (1e-6, 1e6) keeps every intermediate a normal number).  Everything else divides.
            const float an = fabsf(num), ad = fabsf(den);
            if (ad > 1e-6f && ad < 1e6f && an < 1e30f) {
                const bool same = (num < 0.0f) == (den < 0.0f);
                if (same && an > ad) return false;                       // beta > 1
                if (!same && an > ad * 7.8886090522101181e-31f) return false;   // beta < 0
            }
            beta = num / den;
            if (beta < 0.0f || beta > 1.0f) return false;
            alpha = (q0x - beta * r2.x) / r1.z;
        }
        if (alpha < 0.0f || (alpha + beta) > 1.0f) return false;
        res.tri = ti; res.t = t; res.alpha = alpha; res.beta = beta;
        return true;
    }

    // One leaf (src/scene_intersect.cpp:255-292). true: traversal is over (closest: this leaf produced the hit;
    // ANY: some triangle was accepted).  The acceptance conditions of one triangle -- not the ignored one, plane not
    // parallel, t inside [tmin-eps, tmax+eps], t < best so far, barycentrics inside -- are a pure conjunction, so the
    // scan applies the cheap CONSERVATIVE rejection (leaf_bounds / prescreen) first and defers the survivors: they are
    // evaluated exactly, in leaf order, after the scan.
    __device__ __forceinline__ bool leaf(const DevScene& S, uint2 w, TravCount& cnt) {
        if (COUNT) cnt.leaf++;
        const float eps = S.epsilon;
        float lo, hi, lo_c, hi_c;
        leaf_bounds(eps, lo, hi, lo_c, hi_c);
        const uint32_t pend = w.x + (w.y >> 2);
        uint2 cand[RGK_CAND_CAP];
        int nc = 0;
        bool hit = false;
        for (uint32_t p = w.x; p < pend; p++) {
            if (COUNT) { cnt.refs++; if (__ldg(S.refs + p) != ignore) cnt.tests++; }
            float t32;
            if (prescreen(__ldg(S.ref_planes + p), eps, lo_c, hi_c, t32)) {
                cand[nc++] = make_uint2(p, __float_as_uint(t32));
                if (nc == RGK_CAND_CAP) {          // list full: evaluate what we have, in order
                    for (int k = 0; k < RGK_CAND_CAP; k++) {
                        if (candidate(S, cand[k], lo, hi, cnt)) { if (ANY) return true; hit = true; }
                    }
                    nc = 0;
                }
            }
        }
        for (int k = 0; k < nc; k++) {
            if (candidate(S, cand[k], lo, hi, cnt)) { if (ANY) return true; hit = true; }
        }
        return hit;
    }

    // Next stack entry. false: nothing left, or the whole traversal ends because tfar < tmin (:253).
    __device__ __forceinline__ bool pop(const TravStack& K) {
        if (sp == 0) return false;
        --sp;
        const uint2 e = K.e[sp];
        node = e.x; tmin = __uint_as_float(e.y);
        tmax = sp ? __uint_as_float(K.e[sp - 1].y) : troot;
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
        if (idle != 0u && !exhausted && (__popc(idle) >= (int)S.refill_threshold || idle == 0xffffffffu)) {
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


// ------------------------------------------------------------------------------------------------------------
// Phase-synchronised traversal (VARIANT 6, the default).  Same work per lane as trace_persistent, but the three
// phases of an iteration are separated by explicit warp barriers executed by all 32 lanes: with independent thread
// scheduling the hardware does not wait for the lanes still scanning their leaf before it lets the first finished
// lane run its exact tests, so in trace_persistent the expensive exact path (fp64 divide, two more record loads, two
// divides) executes with 1-2 lanes at a time on incoherent rays (ncu: 28 % of the warp instructions at 1.5 lanes).
// Here every lane first descends, then every lane scans (collecting survivors of the pre-rejection; a full candidate
// list just suspends the scan), then all lanes that have candidates evaluate them together.
template <bool ANY, bool COUNT, class Fetch, class Commit>
__device__ __forceinline__ void trace_phased(const DevScene& S, uint32_t count, unsigned long long* work,
                                             TravCount& cnt, uint32_t& done, Fetch fetch, Commit commit) {
    Traverser<ANY, COUNT> T;
    TravStack K;
    const unsigned lane = threadIdx.x & 31;
    const float eps = S.epsilon;
    bool active = false, exhausted = false, in_leaf = false, hit = false;
    uint32_t item = 0, p = 0, pend = 0;
    uint2 cand[RGK_CAND_CAP];                  // (position in the reference list, bits of its approximate t)
    int nc = 0;
    float lo = 0.0f, hi = 0.0f, lo_c = 0.0f, hi_c = 0.0f;
    for (;;) {
        __syncwarp();
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (idle != 0u && !exhausted && (__popc(idle) >= (int)S.refill_threshold || idle == 0xffffffffu)) {
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
                    if (fetch(item, T)) { active = true; in_leaf = false; }
                    else commit(item, false, T.res);
                }
            }
        }
        if (__ballot_sync(0xffffffffu, active) == 0u) { if (exhausted) break; else continue; }
        // ---- phase 1: descend to the next leaf
        if (active && !in_leaf) {
            const uint2 w = T.descend(S, K, cnt);
            if (COUNT) cnt.leaf++;
            p = w.x; pend = w.x + (w.y >> 2); hit = false; nc = 0; in_leaf = true;
            T.leaf_bounds(eps, lo, hi, lo_c, hi_c);
        }
        __syncwarp();
        // ---- phase 2: scan the leaf's references with the conservative pre-rejection
        if (active) {
            // two references per trip (both planes requested before either is used); the second one is a harmless
            // repeat of the first when the leaf has an odd number left
            while (p < pend && nc <= RGK_CAND_CAP - 2) {
                const bool two = p + 1u < pend;
                const uint32_t p2 = two ? p + 1u : p;
                const float4 ra = __ldg(S.ref_planes + p), rb = __ldg(S.ref_planes + p2);
                if (COUNT) {
                    cnt.refs += two ? 2u : 1u;
                    if (__ldg(S.refs + p) != T.ignore) cnt.tests++;
                    if (two && __ldg(S.refs + p2) != T.ignore) cnt.tests++;
                }
                float ta, tb;
                const bool sa = T.prescreen(ra, eps, lo_c, hi_c, ta), sb = T.prescreen(rb, eps, lo_c, hi_c, tb) && two;
                if (sa) cand[nc++] = make_uint2(p, __float_as_uint(ta));
                if (sb) cand[nc++] = make_uint2(p2, __float_as_uint(tb));
                p = p2 + 1u;
            }
        }
        __syncwarp();
        // ---- phase 3: exact tests of the survivors (leaf order), together
        if (active && nc > 0) {
            for (int k = 0; k < nc; k++)
                if (T.candidate(S, cand[k], lo, hi, cnt)) { hit = true; if (ANY) break; }
            nc = 0;
        }
        if (active && (p == pend || (ANY && hit))) {             // leaf finished (src/scene_intersect.cpp:290-292)
            in_leaf = false;
            if (hit) { commit(item, true, T.res); active = false; }
            else if (!T.pop(K)) { commit(item, false, T.res); active = false; }
        }
    }
}

// VARIANT 2: descend-to-leaf / process-leaf per lane with idle-lane refill; VARIANT 6: the same with the phases
// separated by warp barriers.  (A warp-voted state machine, a bounded-phase state machine and a child-prefetching
// descent were measured in round 1 and removed: all slower than these two -- see DESIGN.md 4.)
template <int VARIANT, bool ANY, bool COUNT, class Fetch, class Commit>
__device__ __forceinline__ void trace_rays(const DevScene& S, uint32_t count, unsigned long long* work,
                                           TravCount& cnt, uint32_t& done, Fetch fetch, Commit commit) {
    if (VARIANT == 6) trace_phased<ANY, COUNT>(S, count, work, cnt, done, fetch, commit);
    else trace_persistent<ANY, COUNT>(S, count, work, cnt, done, fetch, commit);
}
