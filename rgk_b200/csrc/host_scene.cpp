// host_scene.cpp -- host side of rgk_scene_commit: what Scene::Commit does on the CPU
// (src/scene.cpp:294-429), producing the flattened arrays the kernels read.
//
// north_star keeps the SAH kd-tree build on the host.  The build follows the reference's
// procedure step for step (events per node, std::sort with the (pos, BEGIN<END) comparator,
// SAH sweep with ISECT 80 / TRAV 2 / EMPTY_BONUS 0.5, up to two axis retries, max depth
// log2(n)+8) because closest-hit triangle IDs are only bit-exact on the *same* tree with
// the same leaf order.  It emits the preorder node array directly
// (left child = i+1, right child index stored in the parent), i.e. Scene::Compress's
// output (src/scene.cpp:606-657), without building a pointer tree first.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <cstdlib>
#include <stdexcept>
#include <exception>
#include <thread>
#include "rgk_internal.h"

namespace {

struct F3 { float x, y, z; };
inline F3 ld3(const float* p) { return F3{p[0], p[1], p[2]}; }
inline F3 sub(F3 a, F3 b) { return F3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float dot3(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline F3 cross3(F3 x, F3 y) { return F3{x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
inline float comp(F3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

struct Event { float pos; int tri; int kind; };  // kind 0 = BEGIN, 1 = END

struct Builder {
    const std::vector<float>* ev;   // ev[axis][2*tri + {0,1}] = min / max of the triangle on that axis
    std::vector<uint32_t>& nodes;
    std::vector<uint32_t>& refs;
    unsigned max_depth;
    unsigned deepest = 0;
    // Host parallelism: the two children of a node are independent, so above `fork_depth` the right
    // child is built by another thread into its own arrays and stitched in afterwards (child indices and leaf
    // reference offsets shifted).  Every node is still split by the reference's procedure on the same triangle order,
    // so the arrays are byte-identical to the sequential build (tests compare them with the oracle's).
    // A fork happens wherever both children are large and the shared budget of extra threads is not used up (not at a
    // fixed depth: SAH splits are uneven, a fixed-depth fork leaves most of the work with one thread).
    std::atomic<int>* spare_threads = nullptr;     // extra threads that may still be started
    size_t fork_min_tris = 20000;

    void emit_leaf(const std::vector<uint32_t>& tris) {
        nodes.push_back((uint32_t)refs.size());
        nodes.push_back(((uint32_t)tris.size() << 2) | 3u);
        refs.insert(refs.end(), tris.begin(), tris.end());
    }

    // UncompressedKdNode::Subdivide (src/scene.cpp:431-574) fused with CompressRec (:637-657)
    void build(std::vector<uint32_t>& tris, float bb[3][2], unsigned depth) {
        deepest = std::max(deepest, depth);
        const unsigned n = (unsigned)tris.size();
        if (depth >= max_depth || n < 2) { emit_leaf(tris); return; }
        const float size[3] = {bb[0][1] - bb[0][0], bb[1][1] - bb[1][0], bb[2][1] - bb[2][0]};
        unsigned axis = (unsigned)(std::max_element(size, size + 3) - size);
        std::vector<Event> events(2 * (size_t)n);
        int best_offset = -1; float best_pos = 0.0f;
        for (unsigned attempt = 0;; attempt++) {
            const std::vector<float>& a = ev[axis];
            for (unsigned i = 0; i < n; i++) {
                const int t = (int)tris[i];
                events[2 * i] = Event{a[2 * t], t, 0};
                events[2 * i + 1] = Event{a[2 * t + 1], t, 1};
            }
            std::sort(events.begin(), events.end(), [](const Event& l, const Event& r) {
                if (l.pos == r.pos) return l.kind < r.kind;
                return l.pos < r.pos;
            });
            const float lo = bb[axis][0], hi = bb[axis][1];
            const unsigned a2 = (axis + 1) % 3, a3 = (axis + 2) % 3;
            const float inv_total_sa = 1.f / (2.f * (size[0] * size[1] + size[0] * size[2] + size[1] * size[2]));
            const float nosplit_cost = 80.0f * n;
            float best_cost = std::numeric_limits<float>::infinity();
            best_offset = -1;
            int n_before = 0, n_after = (int)n;
            for (unsigned i = 0; i < 2 * n; i++) {
                if (events[i].kind == 1) n_after--;
                const float pos = events[i].pos;
                if (pos > lo && pos < hi) {
                    const float below = 2 * (size[a2] * size[a3] + (pos - lo) * size[a2] + (pos - lo) * size[a3]);
                    const float above = 2 * (size[a2] * size[a3] + (hi - pos) * size[a2] + (hi - pos) * size[a3]);
                    const float p_before = below * inv_total_sa, p_after = above * inv_total_sa;
                    const float bonus = (n_before == 0 || n_after == 0) ? 0.5f : 0.f;
                    const float cost = 2.0f + 80.0f * (1.f - bonus) * (p_before * n_before + p_after * n_after);
                    if (cost < best_cost) { best_cost = cost; best_offset = (int)i; best_pos = pos; }
                }
                if (events[i].kind == 0) n_before++;
            }
            if (best_offset != -1 && !(best_cost > nosplit_cost)) break;
            if (attempt >= 2) { emit_leaf(tris); return; }
            axis = (axis + 1) % 3;
        }
        std::vector<uint32_t> below, above;
        for (int i = 0; i < best_offset; i++) if (events[i].kind == 0) below.push_back((uint32_t)events[i].tri);
        for (unsigned i = (unsigned)best_offset + 1; i < 2 * n; i++) if (events[i].kind == 1) above.push_back((uint32_t)events[i].tri);
        std::vector<Event>().swap(events);
        std::vector<uint32_t>().swap(tris);
        const size_t me = nodes.size();
        uint32_t bits; std::memcpy(&bits, &best_pos, 4);
        nodes.push_back(bits);
        nodes.push_back(axis);
        float cb[3][2]; std::memcpy(cb, bb, sizeof cb);
        cb[axis][1] = best_pos;
        bool fork = false;
        if (spare_threads && above.size() >= fork_min_tris && below.size() >= fork_min_tris) {
            if (spare_threads->fetch_sub(1) > 0) fork = true; else spare_threads->fetch_add(1);
        }
        if (fork) {
            std::vector<uint32_t> rn, rr;
            Builder right{ev, rn, rr, max_depth};
            right.spare_threads = spare_threads; right.fork_min_tris = fork_min_tris;
            float rb[3][2]; std::memcpy(rb, bb, sizeof rb);
            rb[axis][0] = best_pos;
            // a throw on either side (bad_alloc on a multi-million-triangle scene) must not reach std::terminate: the worker's
            // is carried over in an exception_ptr, the parent's unwinds through the joiner, and both surface from
            // rgk_scene_commit as an error status
            std::exception_ptr worker_error;
            std::thread worker([&] { try { right.build(above, rb, depth + 1); } catch (...) { worker_error = std::current_exception(); } });
            {
                struct Joiner { std::thread& t; std::atomic<int>* spare; ~Joiner() { if (t.joinable()) t.join(); spare->fetch_add(1); } } joiner{worker, spare_threads};
                build(below, cb, depth + 1);
            }
            if (worker_error) std::rethrow_exception(worker_error);
            const uint32_t node_off = (uint32_t)(nodes.size() / 2), ref_off = (uint32_t)refs.size();
            nodes[me + 1] = axis | (node_off << 2);
            nodes.reserve(nodes.size() + rn.size());
            for (size_t i = 0; i < rn.size(); i += 2) {
                uint32_t w0 = rn[i], w1 = rn[i + 1];
                if ((w1 & 3u) == 3u) w0 += ref_off;              // leaf: first reference
                else w1 += node_off << 2;                         // inner: index of the far child
                nodes.push_back(w0); nodes.push_back(w1);
            }
            refs.insert(refs.end(), rr.begin(), rr.end());
            deepest = std::max(deepest, right.deepest);
            return;
        }
        build(below, cb, depth + 1);
        nodes[me + 1] = axis | ((uint32_t)(nodes.size() / 2) << 2);
        std::memcpy(cb, bb, sizeof cb);
        cb[axis][0] = best_pos;
        build(above, cb, depth + 1);
    }
};

unsigned measure_depth(const std::vector<uint32_t>& nodes) {
    // iterative preorder walk (the given tree may come from the caller)
    unsigned deepest = 0;
    std::vector<std::pair<uint32_t, unsigned>> st;
    if (!nodes.empty()) st.push_back({0u, 0u});
    const size_t nn = nodes.size() / 2;
    while (!st.empty()) {
        auto [i, d] = st.back(); st.pop_back();
        if (i >= nn) throw std::runtime_error("kd-tree: child index out of range");
        deepest = std::max(deepest, d);
        const uint32_t w1 = nodes[2 * i + 1];
        if ((w1 & 3u) == 3u) continue;
        st.push_back({w1 >> 2, d + 1});
        st.push_back({i + 1, d + 1});
        if (d > 4096) throw std::runtime_error("kd-tree: cycle");
    }
    return deepest;
}

} // namespace

// Conservative bounds of the region in which Triangle::TestIntersection can accept a hit point, in the triangle's
// 2-D projection (the dominant-axis projection the test itself uses, src/primitives.cpp:104-133): the box of the three
// projected vertices widened by W, an upper bound of how far outside the true triangle the test's own fp32 rounding can
// still accept.  The test accepts iff 0 <= beta <= 1, alpha >= 0, alpha + beta <= 1 with
//   beta = (q0y q1x - q0x q1y) / den,  alpha = (q0x - beta q2x) / q1x        (|q1x| >= eps)
//   beta = q0x / q2x,                  alpha = (q0y - beta q2y) / q1y        (|q1x| <  eps)
// For a point within a few triangle sizes L the rounding of the numerators is <= 4u L^2 (u = 2^-24) and of den
// <= 4u L^2, so beta is off by db <= 16u L^2 / |den| and alpha by da <= db |q2x / q1x| + 12u L / |q1x|.  An accepted
// point is (alpha_fl q1 + beta_fl q2), which lies in the vertex box, minus (da q1 + db q2), whose size is at most
// 23u L^3 / |den| (1 + L / |q1x|) + 17u L^2 / |q1x| (|q1x| standing for whatever alpha is divided by).  W takes 128u for
// every term (5x slack) and the filter is disabled (infinite box) when den is within 2^-16 L^2 of cancelling.  The
// device adds the ray-dependent part (error of its fp32 t and of o + d t) at run time.  The axis code rides in the two
// low mantissa bits of the first bound, which is first moved 8 ulps outwards.
static void tri_prefilter_bounds(const float* r, uint32_t code, float* out) {
    const double v0x = r[4], v0y = r[5], q1x = r[6], q1y = r[7], q2x = r[8], q2y = r[9], den = r[10];
    const double L = std::max(std::max(std::fabs(q1x), std::fabs(q1y)), std::max(std::fabs(q2x), std::fabs(q2y)));
    const double K = std::ldexp(1.0, -17);
    const double inf = std::numeric_limits<double>::infinity();
    double W = inf;
    const bool flagged = (code & 4u) != 0;
    const double den_eff = flagged ? std::min(std::fabs(den), std::fabs(q2x * q1y)) : std::fabs(den);
    const double div_a = flagged ? std::min(std::fabs(q1y), std::fabs(q2x)) : std::fabs(q1x);
    if (L > 0 && den_eff > std::ldexp(L * L, -16) && div_a > 0 && std::isfinite(L))
        W = K * (L * L * L / den_eff * (1.0 + L / div_a) + L * L / div_a + L);
    float b[4];
    if (!(W < inf)) {
        b[0] = -std::numeric_limits<float>::max(); b[1] = std::numeric_limits<float>::max(); b[2] = b[0]; b[3] = b[1];
    } else {
        const double lo1 = v0x + std::min(0.0, std::min(q1x, q2x)) - W, hi1 = v0x + std::max(0.0, std::max(q1x, q2x)) + W;
        const double lo2 = v0y + std::min(0.0, std::min(q1y, q2y)) - W, hi2 = v0y + std::max(0.0, std::max(q1y, q2y)) + W;
        const float ninf = -std::numeric_limits<float>::infinity(), pinf = std::numeric_limits<float>::infinity();
        b[0] = std::nextafterf((float)lo1, ninf); b[1] = std::nextafterf((float)hi1, pinf);
        b[2] = std::nextafterf((float)lo2, ninf); b[3] = std::nextafterf((float)hi2, pinf);
        if (!(std::fabs(b[0]) > 1e-30f)) b[0] = -1e-30f;       // keep the bit surgery below away from zero / denormals
        for (int k = 0; k < 8; k++) b[0] = std::nextafterf(b[0], ninf);
        if (!std::isfinite(b[0]) || !std::isfinite(b[1]) || !std::isfinite(b[2]) || !std::isfinite(b[3])) {
            b[0] = -std::numeric_limits<float>::max(); b[1] = std::numeric_limits<float>::max(); b[2] = b[0]; b[3] = b[1];
        }
    }
    uint32_t bits; std::memcpy(&bits, &b[0], 4);
    bits = (bits & ~3u) | (code & 3u);                          // moves the value by < 4 ulps; 8 were reserved
    std::memcpy(&b[0], &bits, 4);
    out[0] = b[0]; out[1] = b[1]; out[2] = b[2]; out[3] = b[3];
}

void host_scene_commit(const rgk_scene_desc* d, const rgk_kdtree* tree, const rgk_device_cfg& cfg, HostScene& hs) {
    hs = HostScene();
    const uint32_t nt = d->n_triangles, nv = d->n_vertices;
    if (nt == 0) throw std::runtime_error("scene has no triangles");
    if (!d->positions || !d->normals || !d->tangents || !d->texcoords || !d->indices || !d->meshes || !d->materials)
        throw std::runtime_error("scene description has null arrays");
    for (uint32_t i = 0; i < 3 * nt; i++)
        if (d->indices[i] >= nv) throw std::runtime_error("triangle index out of range");
    // triangles -> material through the mesh ranges; areal lights per emissive mesh (src/scene.cpp:149-206,215-249)
    std::vector<uint32_t> tri_mat(nt, 0xFFFFFFFFu);
    uint32_t cursor = 0;
    for (uint32_t m = 0; m < d->n_meshes; m++) {
        const rgk_mesh& me = d->meshes[m];
        if (me.first_triangle != cursor || me.first_triangle + me.n_triangles > nt) throw std::runtime_error("mesh ranges must tile [0, n_triangles) in order");
        if (me.material >= d->n_materials) throw std::runtime_error("mesh material index out of range");
        cursor += me.n_triangles;
        for (uint32_t t = 0; t < me.n_triangles; t++) tri_mat[me.first_triangle + t] = me.material;
    }
    if (cursor != nt) throw std::runtime_error("mesh ranges must tile [0, n_triangles) in order");
    for (uint32_t i = 0; i < d->n_materials; i++) {
        const rgk_material& m = d->materials[i];
        if (m.bxdf > RGK_BXDF_LTC_GGX_DIFFUSE) throw std::runtime_error("Unsupported BRDF id in config!");
        if (m.bxdf == RGK_BXDF_MIX && (m.mix_a < 0 || m.mix_b < 0 || (uint32_t)m.mix_a >= i || (uint32_t)m.mix_b >= i))
            throw std::runtime_error("mix material refers to a material that was not (yet) defined");
        const int32_t tx[3] = {m.tex_diffuse, m.tex_color, m.tex_bump};
        for (int32_t t : tx) if (t >= (int32_t)d->n_textures) throw std::runtime_error("texture index out of range");
        if (m.bxdf >= RGK_BXDF_LTC_BECKMANN && !(m.roughness >= 0.0f && m.roughness <= 1.0f))
            throw std::runtime_error("LTC roughness must be in [0,1]");   // assert in src/LTC/ltc.cpp:60
    }

    // planes: Triangle::CalculatePlane (src/primitives.cpp:24-36), GLM cross / normalize formulas
    hs.planes.resize(4 * (size_t)nt);
    hs.tri_shade.resize(4 * (size_t)nt);
    for (uint32_t i = 0; i < nt; i++) {
        const uint32_t va = d->indices[3 * i], vb = d->indices[3 * i + 1], vc = d->indices[3 * i + 2];
        const F3 v0 = ld3(d->positions + 3 * va), v1 = ld3(d->positions + 3 * vb), v2 = ld3(d->positions + 3 * vc);
        const F3 d0 = sub(v1, v0), d1 = sub(v2, v0);
        const F3 c = cross3(d1, d0);
        const float inv = 1.0f / std::sqrt(dot3(c, c));
        const F3 n = F3{c.x * inv, c.y * inv, c.z * inv};
        float* p = &hs.planes[4 * (size_t)i];
        p[0] = n.x; p[1] = n.y; p[2] = n.z; p[3] = -dot3(n, v0);
        uint32_t* s = &hs.tri_shade[4 * (size_t)i];
        s[0] = va; s[1] = vb; s[2] = vc; s[3] = tri_mat[i];
    }

    // areal lights: areas, descending sort, power (src/scene.cpp:323-340)
    float total_areal = 0.0f;
    for (uint32_t m = 0; m < d->n_meshes; m++) {
        const rgk_mesh& me = d->meshes[m];
        const float* e = d->materials[me.material].emission;
        if (!(e[0] > 0 || e[1] > 0 || e[2] > 0) || me.n_triangles == 0) continue;
        std::vector<std::pair<float, unsigned>> ta;
        float total_area = 0.0f;
        for (uint32_t t = 0; t < me.n_triangles; t++) {
            const uint32_t ti = me.first_triangle + t;
            const F3 a = ld3(d->positions + 3 * d->indices[3 * ti]), b = ld3(d->positions + 3 * d->indices[3 * ti + 1]),
                     c = ld3(d->positions + 3 * d->indices[3 * ti + 2]);
            const F3 x = cross3(sub(a, b), sub(c, b));
            const float area = 0.5f * std::sqrt(dot3(x, x));   // Triangle::GetArea, src/primitives.cpp:38-45
            ta.push_back({area, ti});
            total_area += area;
        }
        std::sort(ta.rbegin(), ta.rend());
        DevArealLight al{};
        al.total_area = total_area;
        al.emission[0] = e[0]; al.emission[1] = e[1]; al.emission[2] = e[2];
        al.power = total_area * (e[0] + e[1] + e[2]);
        al.first = (uint32_t)hs.areal_tris.size(); al.count = (uint32_t)ta.size();
        for (auto& p : ta) hs.areal_tris.push_back(DevArealTri{p.first, p.second});
        hs.areal_lights.push_back(al);
        total_areal += al.power;
    }
    float total_point = 0.0f;
    for (uint32_t i = 0; i < d->n_point_lights; i++) total_point += d->point_lights[i].intensity * 4.0f * 3.14159265358979323846264338327950288f;

    // per-axis triangle extents, scene bbox, epsilon (src/scene.cpp:364-395)
    std::vector<float> ev[3];
    float mn[3], mx[3];
    for (int ax = 0; ax < 3; ax++) {
        ev[ax].resize(2 * (size_t)nt);
        for (uint32_t i = 0; i < nt; i++) {
            const float a = d->positions[3 * d->indices[3 * i] + ax], b = d->positions[3 * d->indices[3 * i + 1] + ax],
                        c = d->positions[3 * d->indices[3 * i + 2] + ax];
            const auto mm = std::minmax({a, b, c});
            ev[ax][2 * i] = mm.first; ev[ax][2 * i + 1] = mm.second;
        }
        const auto mm = std::minmax_element(ev[ax].begin(), ev[ax].end());
        mn[ax] = *mm.first; mx[ax] = *mm.second;
    }
    const float xs = mx[0] - mn[0], ys = mx[1] - mn[1], zs = mx[2] - mn[2];
    const float diameter = std::sqrt(xs * xs + ys * ys + zs * zs);
    const float eps = 0.00001f * diameter;
    float bb[3][2];
    for (int ax = 0; ax < 3; ax++) { bb[ax][0] = mn[ax] - eps; bb[ax][1] = mx[ax] + eps; }

    // intersection records: the ray-independent part of Triangle::TestIntersection (src/primitives.cpp:83,104-133,141,149)
    hs.tri_isect.resize(12 * (size_t)nt);
    hs.tri_bounds.resize(4 * (size_t)nt);
    const bool want_bvh = cfg.traversal == RGK_TRAVERSAL_BVH;
    std::vector<float> ev_bvh[3];                              // ev with the corners of the accept regions added (wide BVH only)
    if (want_bvh) for (int ax = 0; ax < 3; ax++) ev_bvh[ax] = ev[ax];
    for (uint32_t i = 0; i < nt; i++) {
        const float* p = &hs.planes[4 * (size_t)i];
        const float px = std::fabs(p[0]), py = std::fabs(p[1]), pz = std::fabs(p[2]);
        int i1, i2; uint32_t code;
        if (px > py && px > pz) { i1 = 1; i2 = 2; code = 0; }
        else if (py > pz) { i1 = 0; i2 = 2; code = 1; }
        else { i1 = 0; i2 = 1; code = 2; }
        const uint32_t* s = &hs.tri_shade[4 * (size_t)i];
        const F3 v0 = ld3(d->positions + 3 * s[0]), v1 = ld3(d->positions + 3 * s[1]), v2 = ld3(d->positions + 3 * s[2]);
        const float q1x = comp(v1, i1) - comp(v0, i1), q1y = comp(v1, i2) - comp(v0, i2);
        const float q2x = comp(v2, i1) - comp(v0, i1), q2y = comp(v2, i2) - comp(v0, i2);
        const float denom = q2y * q1x - q2x * q1y;
        if (q1x > -eps && q1x < eps) code |= 4u;
        float* r = &hs.tri_isect[12 * (size_t)i];
        r[0] = p[0]; r[1] = p[1]; r[2] = p[2]; r[3] = p[3];
        r[4] = comp(v0, i1); r[5] = comp(v0, i2); r[6] = q1x; r[7] = q1y;
        r[8] = q2x; r[9] = q2y; r[10] = denom; std::memcpy(&r[11], &code, 4);
        tri_prefilter_bounds(r, code, &hs.tri_bounds[4 * (size_t)i]);
        // What TestIntersection accepts is a planar region, not the triangle: the PROJECTED triangle -- with v1 moved to v0's
        // first projected coordinate in the |q1.x| < eps branch (src/primitives.cpp:141-147 drops q1.x: a shear of up to eps)
        // -- lifted onto the STORED fp32 plane, whose normal is off by ~2^-24 / sin(angle) for a sliver (:24-36).  The wide
        // BVH's boxes hold the three corners of that region as well (the kd-tree keeps the true extents: ev feeds the
        // reference's build).  A triangle whose plane misses one of its own vertices by more than eps / 4 along the dominant
        // axis is marked (flag 8): the kd-tree references it by its true extents, so its hits always go to the kd pass.
        const int k = 3 - i1 - i2;
        if (p[k] != 0.0f && std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]) && std::isfinite(p[3])) {
            auto lift = [&](double a, double b) { return -((double)p[3] + (double)p[i1] * a + (double)p[i2] * b) / (double)p[k]; };
            const F3 vs[3] = {v0, v1, v2};
            bool off_plane = false;
            for (int c = 0; c < 3; c++) {
                if (!(std::fabs(lift(comp(vs[c], i1), comp(vs[c], i2)) - (double)comp(vs[c], k)) <= 0.25 * (double)eps)) off_plane = true;
                double w[3];
                w[i1] = (c == 1 && (code & 4u)) ? comp(v0, i1) : comp(vs[c], i1); w[i2] = comp(vs[c], i2);
                w[k] = lift(w[i1], w[i2]);
                if (!want_bvh || !std::isfinite(w[k])) continue;
                for (int ax = 0; ax < 3; ax++) {
                    float lo = (float)w[ax], hi = lo;
                    for (int n = 0; n < 2; n++) { lo = std::nextafterf(lo, -std::numeric_limits<float>::infinity()); hi = std::nextafterf(hi, std::numeric_limits<float>::infinity()); }
                    ev_bvh[ax][2 * (size_t)i] = std::min(ev_bvh[ax][2 * (size_t)i], lo);
                    ev_bvh[ax][2 * (size_t)i + 1] = std::max(ev_bvh[ax][2 * (size_t)i + 1], hi);
                }
            }
            if (off_plane) { code |= 8u; std::memcpy(&r[11], &code, 4); }
        }
    }

    // Triangles whose record can produce NaN barycentrics (projected edges exactly collinear: denom == 0, or q2.x == 0 /
    // q1.y == 0 in the |q1.x| < eps branch) while their plane is finite: the reference's TestIntersection ACCEPTS such a hit
    // (every comparison with NaN is false, src/primitives.cpp:141-164) wherever the ray crosses the triangle's plane inside a
    // kd leaf that references it -- not a geometric event, so no bounding box can stand in for it.  Scenes that have one keep
    // the kd-tree for every ray (the wide BVH is not built); exact zero-area triangles are harmless (NaN plane: rejected).
    size_t nan_prone = 0;
    for (uint32_t i = 0; i < nt; i++) {
        const float* r = &hs.tri_isect[12 * (size_t)i];
        uint32_t flags; std::memcpy(&flags, &r[11], 4);
        const bool finite_plane = std::isfinite(r[0]) && std::isfinite(r[1]) && std::isfinite(r[2]) && std::isfinite(r[3]);
        const bool degenerate = (flags & 4u) ? (r[8] == 0.0f || r[7] == 0.0f) : (r[10] == 0.0f);
        bool finite_rec = true;
        for (int k = 4; k < 11; k++) finite_rec = finite_rec && std::isfinite(r[k]);
        if (finite_plane && (degenerate || !finite_rec)) nan_prone++;
    }
    hs.nan_prone_triangles = (uint32_t)std::min<size_t>(nan_prone, 0xFFFFFFFFu);
    // the wide BVH (rgk_device_cfg::traversal == RGK_TRAVERSAL_BVH, the default) is independent of the kd-tree: built on its
    // own thread meanwhile
    hs.bvh_nodes.clear(); hs.bvh_order.clear(); hs.bvh_depth = 0;
    std::thread bvh_thread; std::exception_ptr bvh_error;
    if (want_bvh && nan_prone == 0)
        bvh_thread = std::thread([&] { try { host_bvh_build(ev_bvh, nt, cfg, hs); } catch (...) { bvh_error = std::current_exception(); } });
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } bvh_joiner{bvh_thread};     // also on the throwing paths below

    // kd-tree
    unsigned deepest;
    if (tree) {
        if (!tree->nodes || tree->n_nodes == 0 || (tree->n_refs && !tree->refs)) throw std::runtime_error("given kd-tree is empty");
        hs.nodes.assign(tree->nodes, tree->nodes + 2 * (size_t)tree->n_nodes);
        hs.refs.assign(tree->refs, tree->refs + tree->n_refs);
        for (size_t i = 0; i < hs.nodes.size() / 2; i++) {
            const uint32_t w0 = hs.nodes[2 * i], w1 = hs.nodes[2 * i + 1];
            if ((w1 & 3u) == 3u && (uint64_t)w0 + (w1 >> 2) > hs.refs.size()) throw std::runtime_error("kd-tree: leaf range out of bounds");
        }
        for (uint32_t r : hs.refs) if (r >= nt) throw std::runtime_error("kd-tree: triangle reference out of range");
        deepest = measure_depth(hs.nodes);
    } else {
        Builder b{ev, hs.nodes, hs.refs, (unsigned)(int)(std::log2(nt) + 8)};
        // rgk_device_cfg::build_threads threads in total (0 = all cores, at most 64); 1 forces the sequential build
        unsigned threads = std::min(64u, std::max(1u, std::thread::hardware_concurrency()));
        if (cfg.build_threads) threads = std::min(64u, cfg.build_threads);
        std::atomic<int> spare((int)threads - 1);
        if (threads > 1) b.spare_threads = &spare;
        std::vector<uint32_t> all(nt);
        for (uint32_t i = 0; i < nt; i++) all[i] = i;
        b.build(all, bb, 0);
        deepest = b.deepest;
    }
    if (deepest + 2 > RGK_STACK_CAP) throw std::runtime_error("kd-tree deeper than the traversal stack capacity");
    if (bvh_thread.joinable()) bvh_thread.join();
    if (bvh_error) std::rethrow_exception(bvh_error);

    rgk_scene_info& in = hs.info;
    in.epsilon = eps;
    for (int ax = 0; ax < 3; ax++) { in.bbox[2 * ax] = bb[ax][0]; in.bbox[2 * ax + 1] = bb[ax][1]; }
    in.n_nodes = (uint32_t)(hs.nodes.size() / 2); in.n_refs = (uint32_t)hs.refs.size(); in.n_triangles = nt;
    in.n_areal_lights = (uint32_t)hs.areal_lights.size(); in.max_depth = deepest;
    in.total_point_power = total_point; in.total_areal_power = total_areal;
}
