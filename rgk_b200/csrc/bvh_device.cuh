// bvh_device.cuh -- wide-BVH traversal (RGK_TRAVERSAL_BVH, the default), device side.
//
// No reference counterpart: RGKrt traverses its kd-tree only (src/scene_intersect.cpp:211-327).  This finds the
// GLOBALLY closest hit of a ray through the 4-wide BVH of host_bvh.cpp with Triangle::TestIntersection's own arithmetic
// (src/primitives.cpp:75-166, the same operation order as Traverser::exact_test), so whenever the kd-tree's answer is
// the global closest hit the result is bit-identical.  The kd rule can differ from that in two ways only (DESIGN.md 8,
// tests/bvh_study.py): it accepts a triangle inside [leaf tmin - eps, leaf tmax + eps] and stops at the first leaf that
// accepts one, so (1) of two hits less than eps apart it may return the farther one, and (2) hits within eps of the ends
// of the root interval are accepted or not depending on the leaf.  Both cases are detected here -- a second accepted hit
// within 2 eps of the best, or a best hit within eps of the root interval's ends.  A third case is (3) a hit on the very
// edge of a triangle that only touches a kd cell: a ray running along that cell's face may never visit the leaf holding
// the triangle, so the kd-tree reports nothing there; a best hit with a barycentric coordinate within 2^-15 of the
// triangle's boundary (more when the ray's own position uncertainty is larger than that against the triangle's size: far
// origins, tiny triangles) is therefore treated the same way.  Such rays (~3e-4 of them) are NOT committed: they are appended
// to a list that the kd kernels re-trace.  The kd-tree stays the authority.
//
// Boxes bound the region the exact test accepts (the triangle's extents plus the corners of the projected -- for the
// |q1.x| < eps branch: sheared -- triangle lifted onto the stored plane, host_scene.cpp); conservativeness against the
// exact test's rounding (the accepted hit point can lie ~2^-22 (|o| + |t|) outside that region) comes from a per-ray
// margin m = 2^-17 (|ox| + |oy| + |oz| + |t1|) added to every box, and from widening the interval bounds by 2^-20 relative.
//
// Slab test (round 2).  Per ray and axis the sign of 1/d says which of a child's two planes is entered first, so the ray
// carries (a) the position of its near / far plane rows inside a node (a 16-byte offset: the node stores lo and hi as
// separate rows, so "near" is a choice of ADDRESS, not a min/max per child) and (b) two constants c = -(o -+ m) / d, the
// margin shifting the near plane towards the ray and the far plane away from it.  A child's entry and exit distances are
// then 3 + 3 fused multiply-adds and 2 + 2 (three-input) max / min: 10 instructions instead of the 24 of
// (plane - o) * (1/d), min, max per axis.  The FMA rounds plane/d + c once; c itself carries one rounding of (o -+ m)/d,
// i.e. a position error of 2^-24 |o| -- 1/128 of the margin, which was sized 32x above what the exact test needs.
#pragma once
#include "trace_device.cuh"

struct BvhCount { uint32_t nodes, tests, slots; };
constexpr uint32_t BVH_DONE = 0x7fffffffu;      // also the code of an empty child slot

// SORT: 1 = children entered by ascending entry distance (sorting network).  A/B knobs, results identical: 0 = slot order
// (RGK_BVH_SHADOW_NOSORT; the answer of an any-hit query does not depend on the order: a firm hit anywhere blocks, border
// hits alone defer), 2 = nearest child first, the others pushed in slot order (RGK_BVH_CLOSEST_NEAREST; a closest-hit answer
// does not depend on the order either: a node holding a hit inside the final 2-eps window is never culled, because its
// entry distance is below every limit the traversal ever had).
template <bool ANY, bool COUNT, int SORT = 1>
struct BvhTraverser {
    float ox, oy, oz, dx, dy, dz, ix, iy, iz;
    float cnx, cny, cnz, cfx, cfy, cfz;          // near / far slab constants: t = plane * (1/d) + c
    uint32_t rows;                               // byte offsets of the near rows inside a node: x | y << 8 | z << 16 (the far row is the other of the pair)
    float lo_t, hi_t;                            // accept interval of the exact test: root interval -+ eps
    float firm_lo, firm_hi;                      // hits inside are beyond the reach of the kd rule's end effects
    float low_w;                                 // lo_t widened (box test)
    float limit;                                 // boxes entered only up to here: min(hi_t, best + 2 eps), widened
    float second_t;                              // smallest t of an accepted hit other than the best
    float marg;                                  // the per-ray box margin m (also scales the edge test)
    uint32_t ignore;
    bool border;                                 // ANY: an accepted hit outside the firm interval (or on an edge) was seen
    bool best_edge;                              // closest: the best hit lies on the boundary of its triangle
    bool degenerate;                             // set by init: the ray must go to the kd pass untouched
    int sp;
    HitRec res;

    // Same root slab test as the kd traversal (src/scene_intersect.cpp:223-232): a ray the kd-tree rejects at the root is
    // rejected here.  false: no hit.  Rays with a zero (or NaN, or reciprocal-overflowing) direction component are marked
    // `degenerate` and left to the kd pass: with the origin exactly on a split plane the kd traversal's plane distance is
    // 0 * inf = NaN, and a NaN interval bound makes its leaves accept hits anywhere along the line (:272,308-318) --
    // reference behaviour that only the kd traversal itself reproduces.
    __device__ __forceinline__ bool init(const DevScene& S, float ox_, float oy_, float oz_, float dx_, float dy_, float dz_,
                                         float tnear, float tfar, uint32_t ignore_) {
        ox = ox_; oy = oy_; oz = oz_; dx = dx_; dy = dy_; dz = dz_; ignore = ignore_;
        res.tri = RGK_NO_TRIANGLE; res.t = __int_as_float(0x7f800000); res.alpha = 0.0f; res.beta = 0.0f;
        second_t = __int_as_float(0x7f800000); border = false; best_edge = false; sp = 0;
        ix = 1.f / dx; iy = 1.f / dy; iz = 1.f / dz;
        const float inf = __int_as_float(0x7f800000);
        // (1e30 rather than infinity: plane * (1/d) of the slab test must not overflow for coordinates below 3e8)
        degenerate = !(fabsf(ix) < 1e30f) || !(fabsf(iy) < 1e30f) || !(fabsf(iz) < 1e30f);
        (void)inf;
        if (degenerate) return true;
        float t0 = tnear, t1 = tfar;
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
        if (tfar < t0) return false;             // "if(r.far < tmin) break" on the root pop (:253)
        const float eps = S.epsilon;
        lo_t = t0 - eps; hi_t = t1 + eps;
        firm_lo = t0 + eps; firm_hi = t1 - eps;
        const float m = (((fabsf(ox) + fabsf(oy)) + fabsf(oz)) + fabsf(t1)) * 7.62939453125e-6f;
        marg = m;
        {   // d > 0: the near plane is lo, met at (lo - (o + m)) / d; d < 0: the near plane is hi, met at (hi - (o - m)) / d
            const bool sx = ix < 0.0f, sy = iy < 0.0f, sz = iz < 0.0f;
            cnx = -(((sx ? ox - m : ox + m)) * ix); cfx = -(((sx ? ox + m : ox - m)) * ix);
            cny = -(((sy ? oy - m : oy + m)) * iy); cfy = -(((sy ? oy + m : oy - m)) * iy);
            cnz = -(((sz ? oz - m : oz + m)) * iz); cfz = -(((sz ? oz + m : oz - m)) * iz);
            rows = (sx ? 16u : 0u) | ((sy ? 48u : 32u) << 8) | ((sz ? 80u : 64u) << 16);
        }
        low_w = (lo_t - fabsf(lo_t) * 9.5367431640625e-7f) - 1e-30f;
        set_limit(eps);
        return true;
    }
    __device__ __forceinline__ void set_limit(float eps) {
        float l = hi_t;
        if (!ANY) { const float w = res.t + 2.0f * eps; l = w < l ? w : l; }
        limit = (l + fabsf(l) * 9.5367431640625e-7f) + 1e-30f;
    }

    // entry distance of a child from its near / far planes (rows picked by the ray's direction signs), +inf on a miss
    __device__ __forceinline__ float child_entry(float nx, float fx, float ny, float fy, float nz, float fz) const {
        const float tn = max3f(__fmaf_rn(nx, ix, cnx), __fmaf_rn(ny, iy, cny), fmaxf(__fmaf_rn(nz, iz, cnz), low_w));
        const float tf = min3f(__fmaf_rn(fx, ix, cfx), __fmaf_rn(fy, iy, cfy), fminf(__fmaf_rn(fz, iz, cfz), limit));
        return tn <= tf ? tn : __int_as_float(0x7f800000);
    }

    // next stack entry still inside the limit, or BVH_DONE
    __device__ __forceinline__ uint32_t pop(const TravStack& K) {
        while (sp > 0) {
            --sp;
            const uint2 e = K.e[sp];
            if (__uint_as_float(e.y) <= limit) return e.x;
        }
        return BVH_DONE;
    }

    // One inner node: returns the next thing to visit (nearest entered child, else the next stack entry, else BVH_DONE).
    __device__ __forceinline__ uint32_t step(const DevScene& S, uint32_t node, TravStack& K, BvhCount& cnt) {
        if (COUNT) cnt.nodes++;
        // rows of the node: lo.x hi.x lo.y hi.y lo.z hi.z codes (16 bytes each); the near row of an axis is at `rows`, the far one
        // at the other position of the pair (offset ^ 16); 32-bit offsets from the array base (nodes < 2^25)
        const char* base = reinterpret_cast<const char*>(S.bvh_nodes);
        const uint32_t at = node << 7;
        const uint32_t rx = rows & 0xffu, ry = (rows >> 8) & 0xffu, rz = rows >> 16;
        const float4 nx = __ldg(reinterpret_cast<const float4*>(base + (at + rx))), fx = __ldg(reinterpret_cast<const float4*>(base + (at + (rx ^ 16u))));
        const float4 ny = __ldg(reinterpret_cast<const float4*>(base + (at + ry))), fy = __ldg(reinterpret_cast<const float4*>(base + (at + (ry ^ 16u))));
        const float4 nz = __ldg(reinterpret_cast<const float4*>(base + (at + rz))), fz = __ldg(reinterpret_cast<const float4*>(base + (at + (rz ^ 16u))));
        const float4 cf = __ldg(reinterpret_cast<const float4*>(base + (at + 96u)));
        float t0 = child_entry(nx.x, fx.x, ny.x, fy.x, nz.x, fz.x);
        float t1 = child_entry(nx.y, fx.y, ny.y, fy.y, nz.y, fz.y);
        float t2 = child_entry(nx.z, fx.z, ny.z, fy.z, nz.z, fz.z);
        float t3 = child_entry(nx.w, fx.w, ny.w, fy.w, nz.w, fz.w);
        uint32_t c0 = __float_as_uint(cf.x), c1 = __float_as_uint(cf.y), c2 = __float_as_uint(cf.z), c3 = __float_as_uint(cf.w);
        const float inf = __int_as_float(0x7f800000);
        if (SORT == 2) {
            // nearest entered child (first of equals), then the rest in slot order
            float tm = t0; uint32_t cm = c0; int im = 0;
            if (t1 < tm) { tm = t1; cm = c1; im = 1; }
            if (t2 < tm) { tm = t2; cm = c2; im = 2; }
            if (t3 < tm) { tm = t3; cm = c3; im = 3; }
            if (!(tm < inf)) return pop(K);
            if (im != 0 && t0 < inf) { K.e[sp] = make_uint2(c0, __float_as_uint(t0)); ++sp; }
            if (im != 1 && t1 < inf) { K.e[sp] = make_uint2(c1, __float_as_uint(t1)); ++sp; }
            if (im != 2 && t2 < inf) { K.e[sp] = make_uint2(c2, __float_as_uint(t2)); ++sp; }
            if (im != 3 && t3 < inf) { K.e[sp] = make_uint2(c3, __float_as_uint(t3)); ++sp; }
            return cm;
        }
        if (SORT == 0) {
            uint32_t nx = BVH_DONE; float tx = inf;
#define RGK_ENTER(t, c) if (t < inf) { if (nx != BVH_DONE) { K.e[sp] = make_uint2(nx, __float_as_uint(tx)); ++sp; } nx = c; tx = t; }
            RGK_ENTER(t0, c0) RGK_ENTER(t1, c1) RGK_ENTER(t2, c2) RGK_ENTER(t3, c3)
#undef RGK_ENTER
            return nx != BVH_DONE ? nx : pop(K);
        }
        // sorting network (0,1)(2,3)(0,2)(1,3)(1,2): ascending entry distance, misses (+inf) last
#define RGK_CSWAP(ta, ca, tb, cb) { const bool s_ = tb < ta; const float tt_ = s_ ? tb : ta; tb = s_ ? ta : tb; ta = tt_; \
                                    const uint32_t cc_ = s_ ? cb : ca; cb = s_ ? ca : cb; ca = cc_; }
        RGK_CSWAP(t0, c0, t1, c1) RGK_CSWAP(t2, c2, t3, c3) RGK_CSWAP(t0, c0, t2, c2) RGK_CSWAP(t1, c1, t3, c3) RGK_CSWAP(t1, c1, t2, c2)
#undef RGK_CSWAP
        if (!(t0 < inf)) return pop(K);
        if (t3 < inf) { K.e[sp] = make_uint2(c3, __float_as_uint(t3)); ++sp; }
        if (t2 < inf) { K.e[sp] = make_uint2(c2, __float_as_uint(t2)); ++sp; }
        if (t1 < inf) { K.e[sp] = make_uint2(c1, __float_as_uint(t1)); ++sp; }
        return c0;
    }

    // Triangle::TestIntersection for leaf slot p on the reference's operation order (src/primitives.cpp:85-164; the
    // arithmetic of Traverser::exact_test), accepted inside [lo_t, hi_t].  Returns true when the traversal can stop
    // (ANY: a firm hit).
    __device__ __forceinline__ bool test(const DevScene& S, uint32_t p, float eps, BvhCount& cnt) {
        if (COUNT) cnt.slots++;
        const float4 r0 = __ldg(S.bvh_planes + p);
        const float dtf = dx * r0.x + dy * r0.y + dz * r0.z;
        const float dot2f = ox * r0.x + oy * r0.y + oz * r0.z;
        if (dtf != dtf) return false;                                    // std::isnan(dot)
        if (fabsf(dtf) < eps) return false;                              // parallel (:91)
        // conservative fp32 pre-rejection (see Traverser::leaf_bounds): |t32 - t| < 2^-21 |t|, bounds widened by 2^-20
        {
            const float t32 = -(r0.w + dot2f) * rcp_approx(dtf);
            if (t32 < low_w || t32 > limit) return false;
        }
        const uint32_t ti = __ldg(S.bvh_refs + p);
        if (ti == ignore) return false;
        if (COUNT) cnt.tests++;
        const float t = (float)(-((double)r0.w + (double)dot2f) / (double)dtf);
        if (t < lo_t || t > hi_t) return false;
        if (!ANY && t > res.t + 2.0f * eps) return false;                // beyond the window of the best so far
        // Grazing hits: the plane misses the triangle's own extents by up to ~an ulp of the coordinates (fp32 plane, fp32 vertices), and
        // along a ray at cosine |dtf| to it that is ulp / |dtf| of ray parameter -- more than eps for |dtf| below ~1e-2: the hit then
        // lies in a kd cell beyond the triangle's extents, which does not reference it (Cornell box, a ray 0.04 degrees off the
        // ceiling: the leaf holding the ceiling ends 1.9e-4 before the hit).  Bound of the miss: 2^-22 |P|_1 (normalised fp32 normal,
        // fp32 offset -n.v0: ~3 roundings of magnitude |v|_1 2^-24, 4x what the Cornell case shows).
        // The kernel is issue-bound (14 more instructions per accepted hit cost 3 % of the closest-hit launches): a two-instruction
        // bound first, |P|_1 <= |o|_1 + sqrt(3) t <= sqrt(3) 2^17 m (0.0546875 = 1.75 / 32), the exact |P|_1 for the ~1 % that pass it.
        bool grazing = fabsf(dtf) * eps < marg * 0.0546875f;
        if (grazing) {
            const float p1 = (fabsf(ox + dx * t) + fabsf(oy + dy * t)) + fabsf(oz + dz * t);
            grazing = fabsf(dtf) * eps < p1 * 2.384185791015625e-7f;
        }
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
        if (flags & 4u) {
            beta = q0x / r2.x;
            if (beta < 0.0f || beta > 1.0f) return false;
            alpha = (q0y - beta * r2.y) / r1.w;
        } else {
            beta = (q0y * r1.z - q0x * r1.w) / r2.z;
            if (beta < 0.0f || beta > 1.0f) return false;
            alpha = (q0x - beta * r2.x) / r1.z;
        }
        if (alpha < 0.0f || (alpha + beta) > 1.0f) return false;
        // "on the boundary" in units of this ray's position uncertainty: an eighth of the margin m (8x the rounding of o + d t) over the
        // triangle's smallest projected height |denom| / longest projected edge, and never less than 2^-15
        const float longest = fmaxf(fmaxf(fabsf(r1.z), fabsf(r1.w)), fmaxf(fabsf(r2.x), fabsf(r2.y)));
        float delta = fmaxf(3.0517578125e-5f, 0.125f * marg * longest / fabsf(r2.z));
        // A triangle of the |q1.x| < eps branch with q1.x != 0 is tested as its sheared twin (q1.x dropped: host_scene.cpp
        // adds that region to the boxes): alpha and beta are up to s = |q1.x / q2.x| (1 + |q2.y / q1.y|) (times |alpha'| <
        // 4/3 while s < 1/4) away from the true barycentrics, and the kd-tree holds the TRUE triangle -- a hit less than
        // that inside the sheared boundary may lie in a cell that does not reference it.  Wider boundary (everything at s >= 1/4).
        if ((flags & 4u) && r1.z != 0.0f) delta += 2.0f * (fabsf(r1.z / r2.x) * (1.0f + fabsf(r2.y / r1.w)));
        const bool edge = !(alpha >= delta) || !(beta >= delta) || !((alpha + beta) <= 1.0f - delta) || (flags & 8u) || grazing;     // 8: plane off its own vertices (host_scene.cpp)
        if (ANY) {
            if (!edge && t >= firm_lo && t <= firm_hi) return true;
            border = true;
            return false;
        }
        if (t < res.t) {
            second_t = res.t;
            res.tri = ti; res.t = t; res.alpha = alpha; res.beta = beta;
            best_edge = edge;
            set_limit(eps);
        } else if (t < second_t) second_t = t;
        return false;
    }

    // the triangles of a leaf child code; true = stop (ANY with a firm hit)
    __device__ __forceinline__ bool leaf(const DevScene& S, uint32_t code, BvhCount& cnt) {
        const float eps = S.epsilon;
        const uint32_t first = code & 0x1fffffffu, n = ((code >> 29) & 3u) + 1u;
        for (uint32_t j = 0; j < n; j++)
            if (test(S, first + j, eps, cnt)) return true;
        return false;
    }

    // closest: does the answer depend on the kd rule?  (ANY: only border hits were found)
    __device__ __forceinline__ bool ambiguous(float eps) const {
        if (ANY) return border;
        if (res.tri == RGK_NO_TRIANGLE) return false;
        return best_edge || second_t <= res.t + 2.0f * eps || res.t < firm_lo || res.t > firm_hi;
    }
};

// Persistent-warp driver, same work distribution as trace_phased.  `fetch(i, T)` loads item i and calls T.init;
// `commit(i, found, res)` stores a settled result; `defer(i)` hands an ambiguous item to the kd-tree pass.
template <bool ANY, bool COUNT, int SORT = 1, class Fetch, class Commit, class Defer>
__device__ __forceinline__ void trace_bvh(const DevScene& S, uint32_t count, unsigned long long* work, BvhCount& cnt, uint32_t& done,
                                          uint32_t& deferred, Fetch fetch, Commit commit, Defer defer) {
    BvhTraverser<ANY, COUNT, SORT> T;
    TravStack K;
    const unsigned lane = threadIdx.x & 31;
    const float eps = S.epsilon;
    bool active = false, exhausted = false, stop = false;
    uint32_t item = 0, cur = BVH_DONE;
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
                    if (fetch(item, T)) {
                        if (T.degenerate) { defer(item); deferred++; }
                        else { active = true; cur = 0u; stop = false; }
                    } else commit(item, false, T.res);
                }
            }
        }
        if (__ballot_sync(0xffffffffu, active) == 0u) { if (exhausted) break; else continue; }
        // ---- phase 1: inner nodes until the next leaf (or the end)
        if (active) {
            while (cur < BVH_DONE) cur = T.step(S, cur, K, cnt);
        }
        __syncwarp();
        // ---- phase 2: the leaf's triangles, together
        if (active && cur != BVH_DONE) {
            stop = T.leaf(S, cur, cnt);
            cur = stop ? BVH_DONE : T.pop(K);
        }
        if (active && cur == BVH_DONE) {
            if (!stop && T.ambiguous(eps)) { defer(item); deferred++; }
            else commit(item, ANY ? stop : (T.res.tri != RGK_NO_TRIANGLE), T.res);
            active = false;
        }
    }
}

// warp-aggregated flush of a launch's BVH counters (one atomic per counter per warp)
template <bool COUNT>
__device__ __forceinline__ void flush_bvh_counts(const BvhCount& c, uint32_t nrays, uint32_t deferred, BvhStats* stats) {
    unsigned long long v[5] = {nrays, deferred, COUNT ? c.nodes : 0u, COUNT ? c.tests : 0u, COUNT ? c.slots : 0u};
#pragma unroll
    for (int k = 0; k < (COUNT ? 5 : 2); k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&stats->rays, v[0]);
        if (v[1]) atomicAdd(&stats->ambiguous, v[1]);
        if (COUNT) { atomicAdd(&stats->nodes, v[2]); atomicAdd(&stats->tests, v[3]); atomicAdd(&stats->slots, v[4]); }
    }
}
