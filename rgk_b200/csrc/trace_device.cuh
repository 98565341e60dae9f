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
struct TravStack { uint32_t node[RGK_STACK_CAP]; float2 range[RGK_STACK_CAP]; };   // range = (tmin, tmax)
#ifndef RGK_CAND_CAP
#define RGK_CAND_CAP 4
#endif
// RGK_CAND_CAP:           // deferred exact tests per leaf before a flush
struct HitRec { uint32_t tri; float t, alpha, beta; };  // alpha/beta as returned by TestIntersection


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
    // PREFETCH: the words of BOTH children are requested as soon as the node word is known and the chosen one is
    // selected afterwards, which takes the node-load latency out of the per-level dependency chain (the decision
    // arithmetic overlaps with the loads).  Pays one extra load per level; used for incoherent batches, which are
    // latency-bound rather than issue-bound.
    template <bool PREFETCH>
    __device__ __forceinline__ uint2 descend(const DevScene& S, TravStack& K, TravCount& cnt) {
        uint2 w = __ldg(S.nodes + node);
        while ((w.y & 3u) != 3u) {
            if (COUNT) cnt.inner++;
            const uint32_t other = w.y >> 2, near = node + 1u;
            uint2 wn, wo;
            if (PREFETCH) { wn = __ldg(S.nodes + near); wo = __ldg(S.nodes + other); }
            const uint32_t axis = w.y & 3u;
            const float split = __uint_as_float(w.x);
            const float oa = axis == 0u ? ox : (axis == 1u ? oy : oz);
            const float ia = axis == 0u ? ix : (axis == 1u ? iy : iz);
            const float diff = split - oa;                 // its sign is the exact sign of (split - oa)
            const float tplane = diff * ia;
            bool below_first = diff > 0.0f;                // oa < split
            if (diff == 0.0f) {                            // oa == split: decided by the direction (rare)
                const float da = axis == 0u ? dx : (axis == 1u ? dy : dz);
                below_first = da <= 0.0f;
            }
            const uint32_t first = below_first ? near : other;
            const uint32_t second = below_first ? other : near;
            bool take_first = true;
            if (tplane > tmax || tplane <= 0.0f) {}
            else if (tplane < tmin) take_first = false;
            else {
                K.node[sp] = second; K.range[sp] = make_float2(tplane, tmax); ++sp;
                tmax = tplane;
            }
            node = take_first ? first : second;
            if (PREFETCH) w = (node == near) ? wn : wo;
            else w = __ldg(S.nodes + node);
        }
        return w;
    }

    // Triangle::TestIntersection proper for one reference that survived the pre-rejection, on the reference's
    // operation order (src/primitives.cpp:85-164), followed by the leaf's accept rule (src/scene_intersect.cpp:272-283).
    __device__ __forceinline__ bool exact_test(const DevScene& S, uint32_t ti, float lo, float hi) {
        const float4* rec = S.tri_isect + 3 * (size_t)ti;
        const float4 r0 = __ldg(rec);
        const float dtf = dx * r0.x + dy * r0.y + dz * r0.z;             // glm::dot(direction, planeN), fp32
        if (dtf != dtf) return false;                                    // std::isnan(dot)
        const float dot2f = ox * r0.x + oy * r0.y + oz * r0.z;
        const float t = (float)(-((double)r0.w + (double)dot2f) / (double)dtf);
        if (t < lo || t > hi) return false;                              // outside this node's interval (:272)
        if (!(t < res.t)) return false;                                  // not closer than the best so far (:275)
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
    // ANY: some triangle was accepted).  The acceptance conditions of one triangle -- plane not parallel, t inside
    // [tmin-eps, tmax+eps], t < best so far, barycentrics inside -- are a pure conjunction, so the scan applies a cheap
    // CONSERVATIVE rejection first and defers the survivors: they are evaluated exactly, in leaf order, after the scan
    // (lanes of the warp then run the rare fp64 path together instead of one at a time).
    //
    // Pre-rejection: t32 = -(w + dot2) * rcp.approx(dot).  The sum is one fp32 rounding of the exact sum (2^-24), the
    // approximate reciprocal and the product add < 2^-22, and the reference's own t is the exact quotient rounded
    // (2^-24): |t32 - t| < 2^-21 |t|.  A margin of 2^-19 |t32| (+1e-30 against flush-to-zero) therefore never rejects
    // a triangle the exact test accepts; overflow gives +-inf (correctly outside), NaN compares false (kept).
    __device__ __forceinline__ bool leaf(const DevScene& S, uint2 w, TravCount& cnt) {
        if (COUNT) cnt.leaf++;
        const float eps = S.epsilon;
        const float lo = tmin - eps, hi = tmax + eps;
        const uint32_t n = w.y >> 2;
        const uint32_t* __restrict__ rp = S.refs + w.x;
        uint32_t cand[RGK_CAND_CAP];
        int nc = 0;
        bool hit = false;
        const float4* __restrict__ pp = S.ref_planes + w.x;   // plane of reference j, stored next to the reference
        for (uint32_t p = 0; p < n; p++) {
            const uint32_t ti = __ldg(rp + p);
            const float4 r0 = __ldg(pp + p);
            if (COUNT) { cnt.refs++; if (ti != ignore) cnt.tests++; }
            const float dtf = dx * r0.x + dy * r0.y + dz * r0.z;
            const float dot2f = ox * r0.x + oy * r0.y + oz * r0.z;
            float rcp;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(dtf));
            const float t32 = -(r0.w + dot2f) * rcp;
            const float m = __fmaf_rn(fabsf(t32), 1.9073486328125e-6f, 1e-30f);
            const bool reject = (ti == ignore) || (fabsf(dtf) < eps) || (t32 + m < lo) || (t32 - m > hi);
            if (!reject) {
                cand[nc++] = ti;
                if (nc == RGK_CAND_CAP) {          // list full: evaluate what we have, in order
                    for (int k = 0; k < RGK_CAND_CAP; k++) {
                        if (exact_test(S, cand[k], lo, hi)) { if (ANY) return true; hit = true; }
                    }
                    nc = 0;
                }
            }
        }
        for (int k = 0; k < nc; k++) {
            if (exact_test(S, cand[k], lo, hi)) { if (ANY) return true; hit = true; }
        }
        return hit;
    }

    // Next stack entry. false: nothing left, or the whole traversal ends because tfar < tmin (:253).
    __device__ __forceinline__ bool pop(const TravStack& K) {
        if (sp == 0) return false;
        --sp;
        node = K.node[sp]; const float2 r = K.range[sp]; tmin = r.x; tmax = r.y;
        return !(tfar < tmin);
    }
};

// Persistent-warp driver.  `fetch(i, T)` loads work item i and calls T.init(...) (returns its result);
// `commit(i, found, res)` stores the result.  Every lane owns one ray at a time; lanes whose ray is finished
// stay idle until at least RGK_REFILL_THRESHOLD lanes of the warp are idle (or all are), then the warp grabs
// that many new items from the global counter with one atomic (warp-ballot work redistribution).
template <bool ANY, bool COUNT, bool PREFETCH, class Fetch, class Commit>
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
            const uint2 w = T.template descend<PREFETCH>(S, K, cnt);
            if (T.leaf(S, w, cnt)) { commit(item, true, T.res); active = false; }
            else if (!T.pop(K)) { commit(item, false, T.res); active = false; }
        }
    }
}


// ------------------------------------------------------------------------------------------------------------
// Phase-synchronised traversal (VARIANT 6).  Same work per lane as trace_persistent, but the three phases of an
// iteration are separated by explicit warp barriers executed by all 32 lanes: with independent thread scheduling the
// hardware does not wait for the lanes still scanning their leaf before it lets the first finished lane run its
// exact tests, so in trace_persistent the expensive exact path (fp64 divide, two more record loads, two divides)
// executes with 1-2 lanes at a time on incoherent rays (ncu: 28 % of the warp instructions at 1.5 lanes).  Here
// every lane first descends, then every lane scans (collecting survivors of the pre-rejection; a full candidate list
// just suspends the scan), then all lanes that have candidates evaluate them together.
template <bool ANY, bool COUNT, class Fetch, class Commit>
__device__ __forceinline__ void trace_phased(const DevScene& S, uint32_t count, unsigned long long* work,
                                             TravCount& cnt, uint32_t& done, Fetch fetch, Commit commit) {
    Traverser<ANY, COUNT> T;
    TravStack K;
    const unsigned lane = threadIdx.x & 31;
    const float eps = S.epsilon;
    bool active = false, exhausted = false, in_leaf = false, hit = false;
    uint32_t item = 0, p = 0, pend = 0;
    uint32_t cand[RGK_CAND_CAP];
    int nc = 0;
    float lo = 0.0f, hi = 0.0f;
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
            const uint2 w = T.template descend<false>(S, K, cnt);
            if (COUNT) cnt.leaf++;
            p = w.x; pend = w.x + (w.y >> 2); hit = false; nc = 0; in_leaf = true;
            lo = T.tmin - eps; hi = T.tmax + eps;
        }
        __syncwarp();
        // ---- phase 2: scan the leaf's references with the conservative pre-rejection
        if (active) {
            while (p < pend && nc < RGK_CAND_CAP) {
                const uint32_t ti = __ldg(S.refs + p);
                const float4 r0 = __ldg(S.ref_planes + p);
                ++p;
                if (COUNT) { cnt.refs++; if (ti != T.ignore) cnt.tests++; }
                const float dtf = T.dx * r0.x + T.dy * r0.y + T.dz * r0.z;
                const float dot2f = T.ox * r0.x + T.oy * r0.y + T.oz * r0.z;
                float rcp;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(dtf));
                const float t32 = -(r0.w + dot2f) * rcp;
                const float m = __fmaf_rn(fabsf(t32), 1.9073486328125e-6f, 1e-30f);
                const bool reject = (ti == T.ignore) || (fabsf(dtf) < eps) || (t32 + m < lo) || (t32 - m > hi);
                if (!reject) cand[nc++] = ti;
            }
        }
        __syncwarp();
        // ---- phase 3: exact tests of the survivors (leaf order), together
        if (active && nc > 0) {
            for (int k = 0; k < nc; k++)
                if (T.exact_test(S, cand[k], lo, hi)) { hit = true; if (ANY) break; }
            nc = 0;
        }
        if (active && (p == pend || (ANY && hit))) {             // leaf finished (src/scene_intersect.cpp:290-292)
            in_leaf = false;
            if (hit) { commit(item, true, T.res); active = false; }
            else if (!T.pop(K)) { commit(item, false, T.res); active = false; }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Warp-voted traversal (RGK_TRAVERSAL == 3).  Every lane is a small state machine over the same ray state:
//   INNER  -- at an inner node (or about to read its node word)
//   LEAF   -- scanning the references of a leaf with the cheap fp32 pre-rejection
//   EXACT  -- one reference survived the pre-rejection and needs Triangle::TestIntersection proper (fp64 divide,
//             barycentric divides)
//   IDLE   -- ray finished, waiting for a refill
// Each iteration the warp votes and executes only the block most lanes are waiting for (a few steps of it), so a
// lane never idles behind another lane's long leaf or rare exact test: waiting groups only ever gain lanes while the
// total stays 32, hence every group is served eventually.  Which lane evaluates what, in which order along ITS OWN
// ray, is unchanged -- results are identical to the sequential traversal.
enum { TM_IDLE = 0, TM_INNER = 1, TM_LEAF = 2, TM_EXACT = 3 };
#define RGK_STEPS_INNER 3
#define RGK_STEPS_LEAF 3

template <bool ANY, bool COUNT, class Fetch, class Commit>
__device__ __forceinline__ void trace_voted(const DevScene& S, uint32_t count, unsigned long long* work,
                                            TravCount& cnt, uint32_t& done, Fetch fetch, Commit commit) {
    Traverser<ANY, COUNT> T;
    TravStack K;
    const unsigned lane = threadIdx.x & 31;
    const float eps = S.epsilon;
    int mode = TM_IDLE;
    bool exhausted = false, hit = false;
    uint32_t item = 0, p = 0, pend = 0, cand = 0;
    for (;;) {
        const unsigned m_idle = __ballot_sync(0xffffffffu, mode == TM_IDLE);
        if (m_idle != 0u && !exhausted && (__popc(m_idle) >= (int)S.refill_threshold || m_idle == 0xffffffffu)) {
            const int leader = __ffs(m_idle) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(work, (unsigned long long)__popc(m_idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (base + __popc(m_idle) >= count) exhausted = true;
            if (mode == TM_IDLE) {
                const unsigned long long mine = base + __popc(m_idle & ((1u << lane) - 1u));
                if (mine < count) {
                    item = (uint32_t)mine;
                    done++;
                    if (fetch(item, T)) mode = TM_INNER;
                    else commit(item, false, T.res);
                }
            }
        }
        const int n_inner = __popc(__ballot_sync(0xffffffffu, mode == TM_INNER));
        const int n_leaf = __popc(__ballot_sync(0xffffffffu, mode == TM_LEAF));
        const int n_exact = __popc(__ballot_sync(0xffffffffu, mode == TM_EXACT));
        if ((n_inner | n_leaf | n_exact) == 0) { if (exhausted) break; else continue; }
        if (n_exact >= n_inner && n_exact >= n_leaf) {
            if (mode == TM_EXACT) {
                // Triangle::TestIntersection proper, on the reference's operation order (src/primitives.cpp:85-164)
                const float4* rec = S.tri_isect + 3 * (size_t)cand;
                const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1), r2 = __ldg(rec + 2);
                const float dtf = T.dx * r0.x + T.dy * r0.y + T.dz * r0.z;
                const float dot2f = T.ox * r0.x + T.oy * r0.y + T.oz * r0.z;
                const float t = (float)(-((double)r0.w + (double)dot2f) / (double)dtf);
                bool ok = !(t < T.tmin - eps || t > T.tmax + eps) && (t < T.res.t);
                float alpha = 0.0f, beta = 0.0f;
                if (ok) {
                    const uint32_t flags = __float_as_uint(r2.w);
                    const uint32_t code = flags & 3u;
                    const float o1 = (code == 0u) ? T.oy : T.ox, d1 = (code == 0u) ? T.dy : T.dx;
                    const float o2 = (code == 2u) ? T.oy : T.oz, d2 = (code == 2u) ? T.dy : T.dz;
                    const float q0x = (o1 + d1 * t) - r1.x;
                    const float q0y = (o2 + d2 * t) - r1.y;
                    if (flags & 4u) {
                        beta = q0x / r2.x;
                        ok = !(beta < 0.0f || beta > 1.0f);
                        alpha = (q0y - beta * r2.y) / r1.w;
                    } else {
                        beta = (q0y * r1.z - q0x * r1.w) / r2.z;
                        ok = !(beta < 0.0f || beta > 1.0f);
                        alpha = (q0x - beta * r2.x) / r1.z;
                    }
                    if (ok) ok = !(alpha < 0.0f || (alpha + beta) > 1.0f);
                }
                mode = TM_LEAF;
                if (ok) {
                    T.res.tri = cand; T.res.t = t; T.res.alpha = alpha; T.res.beta = beta;
                    hit = true;
                    if (ANY) { commit(item, true, T.res); mode = TM_IDLE; }
                }
            }
        } else if (n_leaf >= n_inner) {
            if (mode == TM_LEAF) {
                const float lo = T.tmin - eps, hi = T.tmax + eps;
#pragma unroll 1
                for (int s = 0; s < RGK_STEPS_LEAF && mode == TM_LEAF; s++) {
                    if (p == pend) {                       // leaf finished (src/scene_intersect.cpp:290-292)
                        if (hit) { commit(item, true, T.res); mode = TM_IDLE; }
                        else if (T.pop(K)) mode = TM_INNER;
                        else { commit(item, false, T.res); mode = TM_IDLE; }
                        break;
                    }
                    const uint32_t ti = __ldg(S.refs + p);
                    const float4 r0 = __ldg(S.ref_planes + p);
                    ++p;
                    if (COUNT) cnt.refs++;
                    if (ti == T.ignore) continue;
                    if (COUNT) cnt.tests++;
                    const float dtf = T.dx * r0.x + T.dy * r0.y + T.dz * r0.z;
                    const float dot2f = T.ox * r0.x + T.oy * r0.y + T.oz * r0.z;
                    // conservative fp32 pre-rejection (see Traverser::leaf): never rejects what the exact test accepts
                    const float n32 = -(r0.w + dot2f);
                    const float hi2 = ANY ? hi : fminf(hi, T.res.t);
                    const float a = lo * dtf, b = hi2 * dtf;
                    const float ma = fabsf(a) * 9.5367431640625e-7f + 1e-30f, mb = fabsf(b) * 9.5367431640625e-7f + 1e-30f;
                    const bool pos = dtf > 0.0f;
                    const bool out_lo = pos ? (n32 < a - ma) : (n32 > a + ma);
                    const bool out_hi = pos ? (n32 > b + mb) : (n32 < b - mb);
                    const bool sane = fabsf(a) < 1e30f && fabsf(b) < 1e30f && fabsf(n32) < 1e30f;
                    const bool reject = (dtf != dtf) || (dtf < eps && dtf > -eps) || ((out_lo || out_hi) && sane);
                    if (!reject) { cand = ti; mode = TM_EXACT; }
                }
            }
        } else {
            if (mode == TM_INNER) {
#pragma unroll 1
                for (int s = 0; s < RGK_STEPS_INNER; s++) {
                    const uint2 w = __ldg(S.nodes + T.node);
                    if ((w.y & 3u) == 3u) {
                        if (COUNT) cnt.leaf++;
                        p = w.x; pend = w.x + (w.y >> 2); hit = false; mode = TM_LEAF;
                        break;
                    }
                    if (COUNT) cnt.inner++;
                    const uint32_t axis = w.y & 3u;
                    const float split = __uint_as_float(w.x);
                    const float oa = axis == 0u ? T.ox : (axis == 1u ? T.oy : T.oz);
                    const float da = axis == 0u ? T.dx : (axis == 1u ? T.dy : T.dz);
                    const float ia = axis == 0u ? T.ix : (axis == 1u ? T.iy : T.iz);
                    const float tplane = (split - oa) * ia;
                    const bool below_first = (oa < split) || (oa == split && da <= 0.0f);
                    const uint32_t other = w.y >> 2;
                    const uint32_t first = below_first ? T.node + 1u : other;
                    const uint32_t second = below_first ? other : T.node + 1u;
                    if (tplane > T.tmax || tplane <= 0.0f) T.node = first;
                    else if (tplane < T.tmin) T.node = second;
                    else {
                        K.node[T.sp] = second; K.range[T.sp] = make_float2(tplane, T.tmax); ++T.sp;
                        T.node = first; T.tmax = tplane;
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Bounded-phase traversal (VARIANT 4).  Same per-lane state machine as the voted variant, but every iteration of
// the warp runs all three blocks with a small step budget each: up to S.steps_inner inner-node steps for lanes that
// are descending, up to S.steps_leaf references of the pre-rejection scan for lanes inside a leaf, then the exact
// tests of lanes whose leaf is finished (or whose candidate list is full).  A lane never waits longer than one
// block budget for a slower neighbour (in the phase structure of variant 2 it waits for the slowest lane of the
// whole phase: mean/max of 5 inner steps and 7 references over 32 lanes is what leaves 5 of 32 lanes busy on
// incoherent rays).  Per-ray evaluation order is unchanged; results are identical.
template <bool ANY, bool COUNT, class Fetch, class Commit>
__device__ __forceinline__ void trace_bounded(const DevScene& S, uint32_t count, unsigned long long* work,
                                              TravCount& cnt, uint32_t& done, Fetch fetch, Commit commit) {
    Traverser<ANY, COUNT> T;
    TravStack K;
    const unsigned lane = threadIdx.x & 31;
    const float eps = S.epsilon;
    const int budget_inner = (int)S.steps_inner, budget_leaf = (int)S.steps_leaf;
    int mode = TM_IDLE;
    bool exhausted = false, hit = false;
    uint32_t item = 0, p = 0, pend = 0;
    uint32_t cand[RGK_CAND_CAP];
    int nc = 0;
    float lo = 0.0f, hi = 0.0f;
    for (;;) {
        const unsigned m_idle = __ballot_sync(0xffffffffu, mode == TM_IDLE);
        if (m_idle != 0u && !exhausted && (__popc(m_idle) >= (int)S.refill_threshold || m_idle == 0xffffffffu)) {
            const int leader = __ffs(m_idle) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(work, (unsigned long long)__popc(m_idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (base + __popc(m_idle) >= count) exhausted = true;
            if (mode == TM_IDLE) {
                const unsigned long long mine = base + __popc(m_idle & ((1u << lane) - 1u));
                if (mine < count) {
                    item = (uint32_t)mine;
                    done++;
                    if (fetch(item, T)) mode = TM_INNER;
                    else commit(item, false, T.res);
                }
            }
        } else if (m_idle == 0xffffffffu) break;          // everything idle and nothing left to fetch
        __syncwarp();
        // ---- block A: inner-node steps
        if (mode == TM_INNER) {
#pragma unroll 1
            for (int s = 0; s < budget_inner; s++) {
                const uint2 w = __ldg(S.nodes + T.node);
                if ((w.y & 3u) == 3u) {
                    if (COUNT) cnt.leaf++;
                    p = w.x; pend = w.x + (w.y >> 2); hit = false; nc = 0; mode = TM_LEAF;
                    lo = T.tmin - eps; hi = T.tmax + eps;
                    break;
                }
                if (COUNT) cnt.inner++;
                const uint32_t axis = w.y & 3u;
                const float split = __uint_as_float(w.x);
                const float oa = axis == 0u ? T.ox : (axis == 1u ? T.oy : T.oz);
                const float ia = axis == 0u ? T.ix : (axis == 1u ? T.iy : T.iz);
                const float diff = split - oa;
                const float tplane = diff * ia;
                bool below_first = diff > 0.0f;
                if (diff == 0.0f) { const float da = axis == 0u ? T.dx : (axis == 1u ? T.dy : T.dz); below_first = da <= 0.0f; }
                const uint32_t other = w.y >> 2, near = T.node + 1u;
                const uint32_t first = below_first ? near : other, second = below_first ? other : near;
                if (tplane > T.tmax || tplane <= 0.0f) T.node = first;
                else if (tplane < T.tmin) T.node = second;
                else { K.node[T.sp] = second; K.range[T.sp] = make_float2(tplane, T.tmax); ++T.sp; T.node = first; T.tmax = tplane; }
            }
        }
        __syncwarp();
        // ---- block B: pre-rejection scan of the leaf's references
        bool run_exact = false;
        if (mode == TM_LEAF) {
#pragma unroll 1
            for (int s = 0; s < budget_leaf && p < pend && nc < RGK_CAND_CAP; s++) {
                const uint32_t ti = __ldg(S.refs + p);
                const float4 r0 = __ldg(S.ref_planes + p);
                ++p;
                if (COUNT) { cnt.refs++; if (ti != T.ignore) cnt.tests++; }
                const float dtf = T.dx * r0.x + T.dy * r0.y + T.dz * r0.z;
                const float dot2f = T.ox * r0.x + T.oy * r0.y + T.oz * r0.z;
                float rcp;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(dtf));
                const float t32 = -(r0.w + dot2f) * rcp;
                const float m = __fmaf_rn(fabsf(t32), 1.9073486328125e-6f, 1e-30f);
                const bool reject = (ti == T.ignore) || (fabsf(dtf) < eps) || (t32 + m < lo) || (t32 - m > hi);
                if (!reject) cand[nc++] = ti;
            }
            run_exact = (nc > 0) && (p == pend || nc == RGK_CAND_CAP);
        }
        __syncwarp();
        // ---- block C: exact tests (leaf order), then leaf completion
        if (run_exact) {
            for (int k = 0; k < nc; k++)
                if (T.exact_test(S, cand[k], lo, hi)) { hit = true; if (ANY) break; }
            nc = 0;
            if (ANY && hit) { commit(item, true, T.res); mode = TM_IDLE; }
        }
        if (mode == TM_LEAF && p == pend && nc == 0) {           // src/scene_intersect.cpp:290-292
            if (hit) { commit(item, true, T.res); mode = TM_IDLE; }
            else if (T.pop(K)) mode = TM_INNER;
            else { commit(item, false, T.res); mode = TM_IDLE; }
        }
    }
}

// VARIANT 2: descend-to-leaf / process-leaf phases with idle-lane refill; VARIANT 3: warp-voted state machine;
// VARIANT 4: bounded-phase state machine; VARIANT 5: variant 2 with child prefetch in the descent.
template <int VARIANT, bool ANY, bool COUNT, class Fetch, class Commit>
__device__ __forceinline__ void trace_rays(const DevScene& S, uint32_t count, unsigned long long* work,
                                           TravCount& cnt, uint32_t& done, Fetch fetch, Commit commit) {
    if (VARIANT == 4) trace_bounded<ANY, COUNT>(S, count, work, cnt, done, fetch, commit);
    else if (VARIANT == 3) trace_voted<ANY, COUNT>(S, count, work, cnt, done, fetch, commit);
    else if (VARIANT == 6) trace_phased<ANY, COUNT>(S, count, work, cnt, done, fetch, commit);
    else if (VARIANT == 5) trace_persistent<ANY, COUNT, true>(S, count, work, cnt, done, fetch, commit);
    else trace_persistent<ANY, COUNT, false>(S, count, work, cnt, done, fetch, commit);
}
