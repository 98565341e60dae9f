// rgk_oracle.cpp -- CPU restatement of RGKrt's hot path.  TEST INFRASTRUCTURE.
//
// Plain scalar C++ (own vector struct, no GLM, no <random>), one function per
// reference function, each citing the reference file:line it follows (paths are
// relative to the reference tree, Enhex/RGK).  Only tests/,
// __graft_entry__.smoke() and bench.py's CPU-baseline legs may load the library
// built from this file (oracle/_build/librgk_oracle.so); the product
// (rgk_b200/) never includes, links or calls it.
//
// Pinning: the reference ships no tests or golden vectors.  This
// restatement is pinned against the reference ITSELF compiled here
// (oracle/_ref/librgk_ref.so, the unmodified sources + a GLM-formula shim) by
// tests/test_oracle_vs_ref.py and against the fixtures under tests/golden/ that
// were generated from that build (tools/make_golden.py).
//
// Third-party arithmetic restated here because it is not in the reference tree:
//   GLM (version unpinned upstream; formulas of 0.9.7/0.9.8)
//   libstdc++ <random>/<algorithm> as shipped with GCC 13.3: mt19937,
//   generate_canonical<float,24>, uniform_int_distribution (Lemire), std::shuffle
//   (pairwise), SURVEY App. C.  std::sort is used where the reference uses it.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <string>
#include <thread>
#include <vector>
#include "rgk_b200.h"

namespace {

// ------------------------------------------------------------------ vectors
struct V2 { float x, y; };
struct V3 { float x, y, z; float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); } };
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 v3(const float* p) { return V3{p[0], p[1], p[2]}; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
// GLM compute_dot<vec3>: tmp = a*b; tmp.x + tmp.y + tmp.z
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(V3 v) { return std::sqrt(dot(v, v)); }
// GLM normalize: v * inversesqrt(dot(v,v)), inversesqrt(x) = 1/sqrt(x)
inline V3 normalize(V3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }
inline V3 cross(V3 x, V3 y) { return v3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
inline float distance2(V3 a, V3 b) { V3 d = a - b; return dot(d, d); }
inline float gmax(float a, float b) { return (a < b) ? b : a; }   // glm::max
inline float gmin(float a, float b) { return (b < a) ? b : a; }   // glm::min
inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
inline float gangle(V3 x, V3 y) { return std::acos(gclamp(dot(x, y), -1.0f, 1.0f)); } // gtx/vector_angle
const float PI_F = 3.14159265358979323846264338327950288f; // glm::pi<float>()

struct M3 { V3 c[3]; }; // column major
inline V3 mul(const M3& m, V3 v) {
    return v3(m.c[0].x * v.x + m.c[1].x * v.y + m.c[2].x * v.z,
              m.c[0].y * v.x + m.c[1].y * v.y + m.c[2].y * v.z,
              m.c[0].z * v.x + m.c[1].z * v.y + m.c[2].z * v.z);
}
inline float det3(const M3& m) { // GLM determinant(mat3)
    return + m.c[0].x * (m.c[1].y * m.c[2].z - m.c[2].y * m.c[1].z)
           - m.c[1].x * (m.c[0].y * m.c[2].z - m.c[2].y * m.c[0].z)
           + m.c[2].x * (m.c[0].y * m.c[1].z - m.c[1].y * m.c[0].z);
}
inline M3 inv3(const M3& m) { // GLM compute_inverse<mat3>
    const float a00 = m.c[0].x, a01 = m.c[0].y, a02 = m.c[0].z;
    const float a10 = m.c[1].x, a11 = m.c[1].y, a12 = m.c[1].z;
    const float a20 = m.c[2].x, a21 = m.c[2].y, a22 = m.c[2].z;
    const float ood = 1.0f / (+ a00 * (a11 * a22 - a21 * a12) - a10 * (a01 * a22 - a21 * a02) + a20 * (a01 * a12 - a11 * a02));
    M3 I;
    I.c[0].x = + (a11 * a22 - a21 * a12) * ood;
    I.c[1].x = - (a10 * a22 - a20 * a12) * ood;
    I.c[2].x = + (a10 * a21 - a20 * a11) * ood;
    I.c[0].y = - (a01 * a22 - a21 * a02) * ood;
    I.c[1].y = + (a00 * a22 - a20 * a02) * ood;
    I.c[2].y = - (a00 * a21 - a20 * a01) * ood;
    I.c[0].z = + (a01 * a12 - a11 * a02) * ood;
    I.c[1].z = - (a00 * a12 - a10 * a02) * ood;
    I.c[2].z = + (a00 * a11 - a10 * a01) * ood;
    return I;
}

struct Quat { float w, x, y, z; };
inline V3 qrot(const Quat& q, V3 v) { // GLM quat * vec3
    const V3 qv = v3(q.x, q.y, q.z);
    const V3 uv = cross(qv, v);
    const V3 uuv = cross(qv, uv);
    return v + ((uv * q.w) + uuv) * 2.0f;
}
inline Quat qinverse(const Quat& q) { // conjugate(q) / dot(q,q); dot(quat) = (x*x + y*y) + (z*z + w*w)
    const float d = (q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w);
    return Quat{q.w / d, -q.x / d, -q.y / d, -q.z / d};
}
inline Quat angle_axis(float a, V3 ax) { const float s = std::sin(a * 0.5f); V3 vs = ax * s; return Quat{std::cos(a * 0.5f), vs.x, vs.y, vs.z}; }

// src/glm.cpp:3-33 RotationBetweenVectors
Quat rotation_between(V3 start, V3 dest) {
    start = normalize(start); dest = normalize(dest);
    const float cosTheta = dot(start, dest);
    V3 axis;
    if (cosTheta < -1 + 0.001f) {
        axis = cross(v3(0, 1, 0), start);
        if (length(axis) < 0.01) axis = cross(v3(1, 0, 0), start);
        axis = normalize(axis);
        return angle_axis(PI_F, axis);
    }
    axis = cross(start, dest);
    const float s = std::sqrt((1 + cosTheta) * 2);
    const float invs = 1 / s;
    return Quat{s * 0.5f, axis.x * invs, axis.y * invs, axis.z * invs};
}
// src/glm.hpp:18-35 SystemTransform
struct Frame {
    Quat g2l, l2g;
    V3 toLocal(V3 v) const { return qrot(g2l, v); }
    V3 toGlobal(V3 v) const { return qrot(l2g, v); }
};
Frame system_transform(V3 global, V3 local) { Frame f; f.g2l = rotation_between(global, local); f.l2g = qinverse(f.g2l); return f; }

// ------------------------------------------------------------------ colours (src/radiance.hpp)
struct RGB { float r, g, b; };
inline RGB rgb(float r, float g, float b) { return RGB{r, g, b}; }
inline float rgbmax(RGB c) { return gmax(gmax(c.r, c.g), c.b); }
inline void rgbclamp(RGB& c, float v) { if (c.r > v) c.r = v; if (c.g > v) c.g = v; if (c.b > v) c.b = v; }

// ------------------------------------------------------------------ scene
struct Tri { uint32_t va, vb, vc; uint32_t mat; float p[4]; };
struct ArealLight { std::vector<std::pair<float, uint32_t>> tris; float total_area = 0.0f; RGB emission{0, 0, 0}; float power = 0.0f; };
struct Light { int type; V3 pos; RGB color; float intensity; float size; V3 normal; bool valid; }; // type 0 FULL_SPHERE, 1 HEMISPHERE
struct Tex { uint32_t kind, w, h; RGB color; std::vector<float> texels; };

struct KdNodeU { // UncompressedKdNode, src/scene.hpp:185-210
    bool leaf = true; unsigned depth = 0; float bb[3][2];
    std::vector<uint32_t> tris; KdNodeU* ch0 = nullptr; KdNodeU* ch1 = nullptr; int axis = 0; float pos = 0;
    ~KdNodeU() { delete ch0; delete ch1; }
};

struct Scene {
    std::vector<V3> pos, nrm, tan; std::vector<V2> uv;
    std::vector<Tri> tris;
    std::vector<rgk_material> mats;
    std::vector<Tex> texs;
    std::vector<rgk_point_light> plights;
    std::vector<ArealLight> alights;
    float total_areal_power = 0, total_point_power = 0;
    rgk_sky sky{};
    std::vector<float> ltcM[2], ltcA[2]; // 0 GGX, 1 Beckmann
    bool thinglass = false;
    std::vector<float> ev[3]; // xevents/yevents/zevents
    float bb[3][2]; float epsilon = 0.0001f;
    std::vector<uint32_t> nodes, refs; // CompressedKdNode words, compressed_triangles
};

// Triangle::CalculatePlane, src/primitives.cpp:24-36
void calc_plane(const Scene& s, Tri& t) {
    V3 v0 = s.pos[t.va], v1 = s.pos[t.vb], v2 = s.pos[t.vc];
    V3 d0 = v1 - v0, d1 = v2 - v0;
    V3 n = normalize(cross(d1, d0));
    float d = -dot(n, v0);
    t.p[0] = n.x; t.p[1] = n.y; t.p[2] = n.z; t.p[3] = d;
}
// Triangle::GetArea, src/primitives.cpp:38-45
float tri_area(const Scene& s, const Tri& t) {
    V3 a = s.pos[t.va], b = s.pos[t.vb], c = s.pos[t.vc];
    return 0.5f * length(cross(a - b, c - b));
}

// UncompressedKdNode::Subdivide, src/scene.cpp:431-574
const float EMPTY_BONUS = 0.5f, ISECT_COST = 80.0f, TRAV_COST = 2.0f; // src/scene.hpp:181-183
struct BBEvent { float pos; int tri; int type; }; // type 0 BEGIN, 1 END
void subdivide(const Scene& s, KdNodeU* nd, unsigned max_depth) {
    if (nd->depth >= max_depth) return;
    const unsigned n = nd->tris.size();
    if (n < 2) return;
    float sizes[3] = {nd->bb[0][1] - nd->bb[0][0], nd->bb[1][1] - nd->bb[1][0], nd->bb[2][1] - nd->bb[2][0]};
    unsigned axis = std::max_element(sizes, sizes + 3) - sizes;
    unsigned retries = 0;
    std::vector<BBEvent> events;
    int best_offset; float best_cost, best_pos;
    for (;;) {
        const std::vector<float>& all = s.ev[axis];
        events.assign(2 * n, BBEvent{});
        for (unsigned i = 0; i < n; i++) {
            int t = nd->tris[i];
            events[2 * i + 0] = BBEvent{all[2 * t + 0], t, 0};
            events[2 * i + 1] = BBEvent{all[2 * t + 1], t, 1};
        }
        std::sort(events.begin(), events.end(), [](const BBEvent& a, const BBEvent& b) {
            if (a.pos == b.pos) return a.type < b.type;
            return a.pos < b.pos;
        });
        const float lo = nd->bb[axis][0], hi = nd->bb[axis][1];
        const float BBsize[3] = {sizes[0], sizes[1], sizes[2]};
        best_offset = -1; best_cost = std::numeric_limits<float>::infinity(); best_pos = best_cost;
        const float nosplit_cost = ISECT_COST * n;
        const unsigned axis2 = (axis + 1) % 3, axis3 = (axis + 2) % 3;
        const float invTotalSA = 1.f / (2.f * (BBsize[0] * BBsize[1] + BBsize[0] * BBsize[2] + BBsize[1] * BBsize[2]));
        int n_before = 0, n_after = n;
        for (unsigned i = 0; i < 2 * n; i++) {
            if (events[i].type == 1) n_after--;
            const float pos = events[i].pos;
            if (pos > lo && pos < hi) {
                float below = 2 * (BBsize[axis2] * BBsize[axis3] + (pos - lo) * BBsize[axis2] + (pos - lo) * BBsize[axis3]);
                float above = 2 * (BBsize[axis2] * BBsize[axis3] + (hi - pos) * BBsize[axis2] + (hi - pos) * BBsize[axis3]);
                float p_before = below * invTotalSA, p_after = above * invTotalSA;
                float bonus = (n_before == 0 || n_after == 0) ? EMPTY_BONUS : 0.f;
                float cost = TRAV_COST + ISECT_COST * (1.f - bonus) * (p_before * n_before + p_after * n_after);
                if (cost < best_cost) { best_cost = cost; best_offset = i; best_pos = pos; }
            }
            if (events[i].type == 0) n_before++;
        }
        if (best_offset == -1 || best_cost > nosplit_cost) {
            if (retries < 2) { retries++; axis = (axis + 1) % 3; continue; }
            return;
        }
        break;
    }
    nd->leaf = false;
    nd->ch0 = new KdNodeU(); nd->ch1 = new KdNodeU();
    nd->ch0->depth = nd->ch1->depth = nd->depth + 1;
    nd->axis = axis; nd->pos = best_pos;
    for (unsigned i = 0; i < (unsigned)best_offset; ++i) if (events[i].type == 0) nd->ch0->tris.push_back(events[i].tri);
    for (unsigned i = best_offset + 1; i < 2 * n; ++i) if (events[i].type == 1) nd->ch1->tris.push_back(events[i].tri);
    std::memcpy(nd->ch0->bb, nd->bb, sizeof nd->bb); std::memcpy(nd->ch1->bb, nd->bb, sizeof nd->bb);
    nd->ch0->bb[axis][1] = best_pos; nd->ch1->bb[axis][0] = best_pos;
    events.clear(); events.shrink_to_fit();
    subdivide(s, nd->ch0, max_depth);
    subdivide(s, nd->ch1, max_depth);
}
// Scene::CompressRec, src/scene.cpp:637-657 with CompressedKdNode's encoding, src/scene.hpp:212-253
void compress_rec(Scene& s, const KdNodeU* nd) {
    if (nd->leaf) {
        s.nodes.push_back((uint32_t)s.refs.size());
        s.nodes.push_back(((uint32_t)nd->tris.size() << 2) | 0x03);
        for (uint32_t t : nd->tris) s.refs.push_back(t);
    } else {
        const size_t my = s.nodes.size();
        uint32_t bits; std::memcpy(&bits, &nd->pos, 4);
        s.nodes.push_back(bits); s.nodes.push_back((uint32_t)nd->axis);
        compress_rec(s, nd->ch0);
        s.nodes[my + 1] = (s.nodes[my + 1] & 0x03) | ((uint32_t)(s.nodes.size() / 2) << 2);
        compress_rec(s, nd->ch1);
    }
}

// Scene::Commit, src/scene.cpp:294-429
void commit(Scene& s, const rgk_scene_desc* d, const rgk_kdtree* tree) {
    const uint32_t nt = s.tris.size();
    for (auto& t : s.tris) calc_plane(s, t);
    s.total_areal_power = 0.0f;
    for (auto& al : s.alights) {
        for (auto& p : al.tris) { float a = tri_area(s, s.tris[p.second]); p.first = a; al.total_area += a; }
        const float* e = s.mats[s.tris[al.tris[0].second].mat].emission;
        al.emission = rgb(e[0], e[1], e[2]);
        std::sort(al.tris.rbegin(), al.tris.rend());
        al.power = al.total_area * (al.emission.r + al.emission.g + al.emission.b);
        s.total_areal_power += al.power;
    }
    s.total_point_power = 0.0f;
    for (auto& l : s.plights) s.total_point_power += l.intensity * 4.0f * PI_F;
    for (int ax = 0; ax < 3; ax++) {
        s.ev[ax].resize(2 * (size_t)nt);
        for (uint32_t i = 0; i < nt; i++) {
            const Tri& t = s.tris[i];
            float a = s.pos[t.va][ax], b = s.pos[t.vb][ax], c = s.pos[t.vc][ax];
            auto p = std::minmax({a, b, c});
            s.ev[ax][2 * i] = p.first; s.ev[ax][2 * i + 1] = p.second;
        }
    }
    float mn[3], mx[3];
    for (int ax = 0; ax < 3; ax++) {
        auto p = std::minmax_element(s.ev[ax].begin(), s.ev[ax].end());
        mn[ax] = *p.first; mx[ax] = *p.second;
    }
    float xs = mx[0] - mn[0], ys = mx[1] - mn[1], zs = mx[2] - mn[2];
    float diameter = std::sqrt(xs * xs + ys * ys + zs * zs);
    s.epsilon = 0.00001f * diameter;
    for (int ax = 0; ax < 3; ax++) { s.bb[ax][0] = mn[ax] - s.epsilon; s.bb[ax][1] = mx[ax] + s.epsilon; }
    (void)d;
    if (tree) {
        s.nodes.assign(tree->nodes, tree->nodes + 2 * (size_t)tree->n_nodes);
        s.refs.assign(tree->refs, tree->refs + tree->n_refs);
    } else {
        KdNodeU root;
        for (uint32_t i = 0; i < nt; i++) root.tris.push_back(i);
        std::memcpy(root.bb, s.bb, sizeof s.bb);
        int l = std::log2(nt) + 8;
        subdivide(s, &root, l);
        compress_rec(s, &root);
    }
    for (int ax = 0; ax < 3; ax++) { s.ev[ax].clear(); s.ev[ax].shrink_to_fit(); }
}

// ------------------------------------------------------------------ traversal
struct Ray { V3 o, d; float tnear = 0.0f, tfar = 10000.0f; }; // src/ray.hpp
struct Hit { uint32_t tri; float t, a, b, c; };
struct Counters { uint64_t inner = 0, leaf = 0, refs = 0, tests = 0; };

// Triangle::TestIntersection, src/primitives.cpp:75-166
inline bool test_intersection(const Scene& s, const Tri& tr, const Ray& r, float& t, float& a, float& b) {
    const float eps = s.epsilon;
    const V3 planeN = v3(tr.p[0], tr.p[1], tr.p[2]);
    double dt = dot(r.d, planeN);
    if (std::isnan(dt)) return false;
    if (dt < eps && dt > -eps) return false;
    double dot2 = dot(r.o, planeN);
    t = -(tr.p[3] + dot2) / dt;
    int i1, i2;
    const float px = std::fabs(planeN.x), py = std::fabs(planeN.y), pz = std::fabs(planeN.z);
    if (px > py && px > pz) { i1 = 1; i2 = 2; }
    else if (py > pz) { i1 = 0; i2 = 2; }
    else { i1 = 0; i2 = 1; }
    const V3 vert0 = s.pos[tr.va], vert1 = s.pos[tr.vb], vert2 = s.pos[tr.vc];
    const float pt0 = r.o[i1] + r.d[i1] * t, pt1 = r.o[i2] + r.d[i2] * t;
    const float q0x = pt0 - vert0[i1], q0y = pt1 - vert0[i2];
    const float q1x = vert1[i1] - vert0[i1], q1y = vert1[i2] - vert0[i2];
    const float q2x = vert2[i1] - vert0[i1], q2y = vert2[i2] - vert0[i2];
    float alpha, beta;
    if (q1x > -eps && q1x < eps) {
        beta = q0x / q2x;
        if (beta < 0 || beta > 1) return false;
        alpha = (q0y - beta * q2y) / q1y;
    } else {
        beta = (q0y * q1x - q0x * q1y) / (q2y * q1x - q2x * q1y);
        if (beta < 0 || beta > 1) return false;
        alpha = (q0x - beta * q2x) / q1x;
    }
    if (alpha < 0 || (alpha + beta) > 1.0) return false;
    a = alpha; b = beta;
    return true;
}

// Scene::FindIntersectKdOtherThan, src/scene_intersect.cpp:211-327 (ignore == RGK_NO_TRIANGLE:
// Scene::FindIntersectKd, :4-116; the WithThinglass variant :330-455 is result-identical because
// Material::is_thinglass is never set).
Hit find_intersect(const Scene& s, const Ray& r, uint32_t ignore, Counters* cnt) {
    Hit res; res.tri = RGK_NO_TRIANGLE; res.t = std::numeric_limits<float>::infinity(); res.a = res.b = res.c = 0;
    float t0 = r.tnear, t1 = r.tfar;
    for (int i = 0; i < 3; ++i) {
        float invRayDir = 1.f / r.d[i];
        float tNear = (s.bb[i][0] - r.o[i]) * invRayDir;
        float tFar = (s.bb[i][1] - r.o[i]) * invRayDir;
        if (tNear > tFar) std::swap(tNear, tFar);
        t0 = tNear > t0 ? tNear : t0;
        t1 = tFar < t1 ? tFar : t1;
        if (t0 > t1) return res;
    }
    const float invDir[3] = {1.f / r.d.x, 1.f / r.d.y, 1.f / r.d.z};
    struct ToDo { uint32_t node; float tmin, tmax; };
    ToDo todo[200];
    int todo_size = 1;
    todo[0] = ToDo{0, t0, t1};
    const uint32_t* N = s.nodes.data();
    while (todo_size > 0) {
        todo_size--;
        const uint32_t node = todo[todo_size].node;
        const float tmin = todo[todo_size].tmin, tmax = todo[todo_size].tmax;
        if (r.tfar < tmin) break;
        const uint32_t w0 = N[2 * node], w1 = N[2 * node + 1];
        if ((w1 & 3) == 3) {
            if (cnt) cnt->leaf++;
            bool hit = false;
            const uint32_t n = w1 >> 2, start = w0;
            for (uint32_t p = 0; p < n; p++) {
                const uint32_t i = s.refs[start + p];
                if (cnt) cnt->refs++;
                if (i == ignore) continue;
                float t, a, b;
                if (cnt) cnt->tests++;
                if (test_intersection(s, s.tris[i], r, t, a, b)) {
                    if (t < tmin - s.epsilon || t > tmax + s.epsilon) continue;
                    if (t < res.t) {
                        res.tri = i; res.t = t;
                        float c = 1.0f - a - b;
                        res.a = c; res.b = a; res.c = b;
                        hit = true;
                    }
                }
            }
            if (hit) return res;
        } else {
            if (cnt) cnt->inner++;
            const int axis = w1 & 3;
            float split; std::memcpy(&split, &w0, 4);
            const float tplane = (split - r.o[axis]) * invDir[axis];
            const bool belowFirst = (r.o[axis] < split) || (r.o[axis] == split && r.d[axis] <= 0);
            uint32_t first, second;
            if (belowFirst) { first = node + 1; second = w1 >> 2; }
            else { first = w1 >> 2; second = node + 1; }
            if (tplane > tmax || tplane <= 0) todo[todo_size++] = ToDo{first, tmin, tmax};
            else if (tplane < tmin) todo[todo_size++] = ToDo{second, tmin, tmax};
            else { todo[todo_size++] = ToDo{second, tplane, tmax}; todo[todo_size++] = ToDo{first, tmin, tplane}; }
        }
    }
    return res;
}

// Ray(from,to,eps) src/ray.hpp:15-22 + Scene::Visibility src/scene.cpp:670-673
bool visibility(const Scene& s, V3 a, V3 b, Counters* cnt) {
    Ray r; r.o = a;
    V3 diff = b - a;
    r.d = normalize(diff);
    float len = length(diff);
    const float e = s.epsilon * 20.0f;
    r.tnear = 0.0f + e; r.tfar = len - e;
    return find_intersect(s, r, RGK_NO_TRIANGLE, cnt).tri == RGK_NO_TRIANGLE;
}

// ------------------------------------------------------------------ sampler (src/sampler.cpp + libstdc++ 13)
struct MT19937 {
    uint32_t mt[624]; int idx;
    explicit MT19937(uint32_t seed) {
        mt[0] = seed;
        for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    uint32_t next() {
        if (idx >= 624) {
            for (int k = 0; k < 624; k++) {
                uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
        return y;
    }
};
// generate_canonical<float,24> over a 32-bit engine (bits/random.tcc): one draw / 2^32, clamped below 1
inline float canonical(MT19937& g) {
    float ret = (float)g.next() / 4294967296.0f;
    if (ret >= 1.0f) ret = std::nextafterf(1.0f, 0.0f);
    return ret;
}
// uniform_real_distribution<float>(a,b)(g) = canonical*(b-a)+a
inline float uniform_real(MT19937& g, float a, float b) { return canonical(g) * (b - a) + a; }
// uniform_int_distribution<unsigned long>{0, range-1} on a 32-bit-range engine: Lemire (bits/uniform_int_dist.h)
inline uint32_t lemire(MT19937& g, uint32_t range) {
    uint64_t product = (uint64_t)g.next() * (uint64_t)range;
    uint32_t low = (uint32_t)product;
    if (low < range) {
        uint32_t threshold = (0u - range) % range;
        while (low < threshold) { product = (uint64_t)g.next() * (uint64_t)range; low = (uint32_t)product; }
    }
    return (uint32_t)(product >> 32);
}
// std::shuffle, pairwise variant (bits/stl_algo.h); valid while n*n <= 2^32-1
template <class T> void std_shuffle(T* first, uint32_t n, MT19937& g) {
    if (n == 0) return;
    uint32_t i = 1;
    if ((n % 2) == 0) { uint32_t d = lemire(g, 2); std::swap(first[i], first[d]); i++; }
    while (i != n) {
        const uint32_t swap_range = i + 1;
        const uint32_t x = lemire(g, swap_range * (swap_range + 1));
        const uint32_t p1 = x / (swap_range + 1), p2 = x % (swap_range + 1);
        std::swap(first[i], first[p1]); i++;
        std::swap(first[i], first[p2]); i++;
    }
}
// round_up_to_square, src/sampler.cpp:77-83
uint32_t round_up_to_square(uint32_t x) {
    float s = (float)std::sqrt((double)x);  // std::sqrt(unsigned) promotes to double, assigned to float
    float i; float frac = std::modf(s, &i);
    if (frac < 0.0001f) return (uint32_t)(i * i);
    return (uint32_t)((i + 1) * (i + 1));
}
// StratifiedSampler(seed, 64, ms): OfflineSampler ctor + PrepareSamples + Get1D/Get2D, src/sampler.cpp:5-36,85-116
struct Sampler {
    uint32_t set_size, dims;
    std::vector<float> s1; std::vector<V2> s2; // [dim][set]
    MT19937 gen;
    uint32_t cur1 = 0, cur2 = 0, cur_set = 0xFFFFFFFFu;
    Sampler(uint32_t seed, uint32_t dims_, uint32_t ms) : set_size(round_up_to_square(ms)), dims(dims_), gen(seed) {}
    void prepare() {
        s1.resize((size_t)dims * set_size); s2.resize((size_t)dims * set_size);
        for (uint32_t dim = 0; dim < dims; dim++) {
            float* a = &s1[(size_t)dim * set_size]; V2* b = &s2[(size_t)dim * set_size];
            for (uint32_t k = 0; k < set_size; k++) {
                float begin = k / (float)set_size, len = 1.0f / (float)set_size;
                a[k] = begin + uniform_real(gen, 0.0f, len);
            }
            std_shuffle(a, set_size, gen);
            uint32_t sq = (uint32_t)(std::sqrt((double)set_size) + 0.5f);
            for (uint32_t sy = 0; sy < sq; sy++)
                for (uint32_t sx = 0; sx < sq; sx++) {
                    float len = 1.0f / (float)sq, bx = sx / (float)sq, by = sy / (float)sq;
                    float x = bx + uniform_real(gen, 0.0f, len);
                    float y = by + uniform_real(gen, 0.0f, len);
                    b[sy * sq + sx] = V2{x, y};
                }
            std_shuffle(b, set_size, gen);
        }
    }
    void advance() { if (cur_set == 0xFFFFFFFFu) prepare(); cur1 = cur2 = 0; cur_set++; }
    float get1d() { return (cur1 < dims) ? s1[(size_t)(cur1++) * set_size + cur_set] : uniform_real(gen, 0.0f, 1.0f); }
    V2 get2d() {
        if (cur2 < dims) return s2[(size_t)(cur2++) * set_size + cur_set];
        float x = uniform_real(gen, 0.0f, 1.0f); float y = uniform_real(gen, 0.0f, 1.0f); return V2{x, y};
    }
};

// ------------------------------------------------------------------ RandomUtils (src/random_utils.hpp)
inline V2 disc_uniform(V2 s) { // :12-16
    float r = std::sqrt(s.x);
    float a = (float)((double)(s.y * 2.0f) * M_PI);
    return V2{r * std::sin(a), r * std::cos(a)};
}
inline V3 hemi_cos_z(V2 s) { // :39-43
    V2 p = disc_uniform(s);
    float z = std::sqrt(gmax(0.00001f, 1 - p.x * p.x - p.y * p.y));
    return v3(p.x, p.y, z);
}
inline V3 sphere_uniform(V2 s) { // :49-56
    float z = s.x * 2.0f - 1.0f;
    float a = (float)((double)s.y * 6.283185);
    float r = std::sqrt(1 - z * z);
    return v3(r * std::cos(a), r * std::sin(a), z);
}
inline V3 hemi_cos_y(V2 s) { // Sample2DToHemisphereCosine :32-36 (y up)
    V2 p = disc_uniform(s);
    float y = std::sqrt(gmax(0.00001f, 1 - p.x * p.x - p.y * p.y));
    return v3(p.x, y, p.y);
}
// RotationFromY, src/glm.cpp:36-60
inline Quat rotation_from_y(V3 dest) {
    dest = normalize(dest);
    const float cosTheta = dest.y;
    if (cosTheta < -1 + 0.00001f) return angle_axis(PI_F, v3(1.0f, 0.0f, 0.0f));
    const V3 axis = cross(v3(0.0f, 1.0f, 0.0f), dest);
    const float s = std::sqrt((1 + cosTheta) * 2);
    const float invs = 1 / s;
    return Quat{s * 0.5f, axis.x * invs, axis.y * invs, axis.z * invs};
}
inline V3 hemi_cos_directed(V2 s, V3 direction) { return qrot(rotation_from_y(direction), hemi_cos_y(s)); } // :45-47,76-78
inline bool decide_and_rescale(float& sample, float probability) { // :63-73
    if (probability == 0.0f) return false;
    if (probability == 1.0f) return true;
    if (sample < probability) { sample /= probability; return true; }
    sample = (sample - probability) / (1.0f - probability);
    return false;
}

// ------------------------------------------------------------------ textures (src/texture.cpp:35-102, src/texture.hpp:64-80)
inline float repeat(float x) { return x - std::floor(x); }
RGB tex_fetch(const Scene& s, int32_t id, V2 uv) {
    if (id < 0) return rgb(0, 0, 0); // EmptyTexture
    const Tex& t = s.texs[id];
    if (t.kind == 0) return t.color;
    const int W = t.w, H = t.h;
    float x = repeat(uv.x) * W - 0.5f, y = repeat(uv.y) * H - 0.5f;
    float ix0f, iy0f;
    float fx = std::modf(x, &ix0f), fy = std::modf(y, &iy0f);
    int ix0 = ix0f, iy0 = iy0f;
    int ix1 = (ix0 != W - 1) ? ix0 + 1 : ix0;
    int iy1 = (iy0 != H - 1) ? iy0 + 1 : iy0;
    if (ix0 == -1) ix0 = 0;
    if (iy0 == -1) iy0 = 0;
    auto px = [&](int yy, int xx) { const float* p = &t.texels[3 * ((size_t)yy * W + xx)]; return rgb(p[0], p[1], p[2]); };
    RGB c00 = px(iy0, ix0), c01 = px(iy0, ix1), c10 = px(iy1, ix0), c11 = px(iy1, ix1);
    fy = 1.0f - fy; fx = 1.0f - fx;
    auto lerp = [](float w, RGB a, RGB b) { float v = 1.0f - w; return rgb(w * a.r + v * b.r, w * a.g + v * b.g, w * a.b + v * b.b); };
    RGB c0s = lerp(fx, c00, c01), c1s = lerp(fx, c10, c11);
    return lerp(fy, c0s, c1s);
}
void tex_slopes(const Scene& s, int32_t id, V2 uv, float& right, float& bottom) {
    right = bottom = 0.0f;
    if (id < 0) return;
    const Tex& t = s.texs[id];
    if (t.kind == 0) return;
    const int W = t.w, H = t.h;
    int x = repeat(uv.x) * W - 0.5f, y = repeat(uv.y) * H - 0.5f;
    int x2 = (x != W - 1) ? x + 1 : x, y2 = (y != H - 1) ? y + 1 : y;
    if (x == -1) x = 0;
    if (y == -1) y = 0;
    auto mean = [&](int yy, int xx) { const float* p = &t.texels[3 * ((size_t)yy * W + xx)]; return (p[0] + p[1] + p[2]) / 3; };
    float here = mean(y, x);
    right = here - mean(y, x2);
    bottom = here - mean(y2, x);
}

// ------------------------------------------------------------------ LTC (src/LTC/ltc.cpp)
void ltc_bilinear(const Scene& s, int which, float theta, float alpha, M3& M, float& amp) { // :20-57
    float t = gmax(0.0f, gmin(1.0f, theta / (0.5f * 3.14159f)));
    float a = gmax(0.0f, gmin(1.0f, std::sqrt(alpha)));
    if (t >= 1.0f) t = 0.999f;
    if (a >= 1.0f) a = 0.999f;
    const int sz = 63;
    int t1 = std::floor(t * sz), t2 = t1 + 1, a1 = std::floor(a * sz), a2 = a1 + 1;
    float dt1 = t * sz - t1, dt2 = t2 - t * sz, da1 = a * sz - a1, da2 = a2 - a * sz;
    const float* Mt = s.ltcM[which].data(); const float* At = s.ltcA[which].data();
    const float* m11 = Mt + 9 * (a1 + t1 * 64); const float* m12 = Mt + 9 * (a2 + t1 * 64);
    const float* m21 = Mt + 9 * (a1 + t2 * 64); const float* m22 = Mt + 9 * (a2 + t2 * 64);
    float r[9];
    for (int k = 0; k < 9; k++) r[k] = m11[k] * dt2 * da2 + m12[k] * dt2 * da1 + m21[k] * dt1 * da2 + m22[k] * dt1 * da1;
    M.c[0] = v3(r[0], r[1], r[2]); M.c[1] = v3(r[3], r[4], r[5]); M.c[2] = v3(r[6], r[7], r[8]);
    amp = At[a1 + t1 * 64] * dt2 * da2 + At[a2 + t1 * 64] * dt2 * da1 + At[a1 + t2 * 64] * dt1 * da2 + At[a2 + t2 * 64] * dt1 * da1;
}
float ltc_pdf(const Scene& s, int which, V3 N, V3 Vr, V3 Vi, float alpha) { // :59-87
    V3 tangent = cross(N, Vi), Vi_cast = cross(tangent, N);
    M3 rot; rot.c[0] = Vi_cast; rot.c[1] = tangent; rot.c[2] = N;
    M3 unrot = inv3(rot);
    V3 Vr3 = mul(unrot, Vr);
    float theta = gangle(Vi, N);
    M3 M; float amp; ltc_bilinear(s, which, theta, alpha, M, amp);
    M3 invM = inv3(M);
    V3 p = normalize(mul(invM, Vr3));
    V3 L_ = mul(M, p);
    float l = length(L_);
    float detM = det3(M);
    float J = detM / (l * l * l);
    float D = 1.0f / 3.14159f * gmax(0.0f, p.z);
    return amp * D / J;
}
V3 ltc_random(const Scene& s, int which, V3 N, V3 Vi, float roughness, V3 rnd) { // :113-143
    V3 tangent = cross(N, Vi), Vi_cast = cross(tangent, N);
    M3 rot; rot.c[0] = Vi_cast; rot.c[1] = tangent; rot.c[2] = N;
    float theta = gangle(Vi, N);
    M3 M; float amp; ltc_bilinear(s, which, gmax(theta, PI_F / 4.0f), roughness, M, amp);
    V3 q = mul(M, rnd);
    if (q.z < 0.0001f) q.z = 0.0001f;
    q = mul(rot, q);
    return normalize(q);
}

// ------------------------------------------------------------------ BxDFs (src/bxdf/bxdf.hpp:107-159, src/bxdf/bxdf.cpp:192-423)
const V3 UPZ = {0.0f, 0.0f, 1.0f};
void fresnel_dielectric(float eta, float cosTheta, float& R, float& cosT) { // bxdf.cpp:332-354
    if (cosTheta < 0.0f) { eta = 1.0f / eta; cosTheta = -cosTheta; }
    float sinThetaTSq = eta * eta * (1.0f - cosTheta * cosTheta);
    if (sinThetaTSq > 1.0f) { R = 1.0f; cosT = 0.0f; return; }
    float cosThetaTrans = std::sqrt(gmax(1.0f - sinThetaTSq, 0.0f));
    float Rs = (eta * cosTheta - cosThetaTrans) / (eta * cosTheta + cosThetaTrans);
    float Rp = (eta * cosThetaTrans - cosTheta) / (eta * cosThetaTrans + cosTheta);
    R = 0.5f * (Rs * Rs + Rp * Rp); cosT = cosThetaTrans;
}
RGB bxdf_value(const Scene& s, uint32_t mi, V3 Vi, V3 Vr, V2 uv) {
    const rgk_material& m = s.mats[mi];
    switch (m.bxdf) {
    case RGK_BXDF_DIFFUSE: { // bxdf.cpp:192-195
        if (Vi.z <= 0 || Vr.z <= 0) return rgb(0, 0, 0);
        RGB c = tex_fetch(s, m.tex_diffuse, uv); return rgb(c.r / PI_F, c.g / PI_F, c.b / PI_F); }
    case RGK_BXDF_MIX: { // bxdf.cpp:235-239
        RGB a = bxdf_value(s, m.mix_a, Vi, Vr, uv), b = bxdf_value(s, m.mix_b, Vi, Vr, uv);
        float w = m.amount, v = 1.0f - m.amount;
        return rgb(w * a.r + v * b.r, w * a.g + v * b.g, w * a.b + v * b.b); }
    case RGK_BXDF_MIRROR: { // bxdf.cpp:265-270
        V3 refl = v3(-Vi.x, -Vi.y, Vi.z);
        if (std::fabs(dot(refl, Vr) - 1) < 0.0001f) return tex_fetch(s, m.tex_color, uv);
        return rgb(0, 0, 0); }
    case RGK_BXDF_DIELECTRIC: { // bxdf.cpp:356-378
        float eta = (Vi.z < 0) ? m.ior : (float)(1.0 / m.ior);
        float R, cosT; fresnel_dielectric(eta, Vi.z, R, cosT);
        RGB c = tex_fetch(s, m.tex_color, uv);
        if (Vi.z * Vr.z > 0) {
            V3 refl = v3(-Vi.x, -Vi.y, Vi.z);
            if (std::fabs(dot(Vr, refl) - 1) < 0.001f) return rgb(c.r * R, c.g * R, c.b * R);
            return rgb(0, 0, 0);
        } else {
            V3 refr = v3(-Vi.x * eta, -Vi.y * eta, (Vi.z > 0) ? -cosT : cosT);
            float T = 1.0f - R;
            if (std::fabs(dot(Vr, refr) - 1) < 0.001f) return rgb(c.r * T, c.g * T, c.b * T);
            return rgb(0, 0, 0);
        } }
    case RGK_BXDF_TRANSPARENT: { // bxdf.cpp:412-417
        V3 inv = v3(-Vi.x, -Vi.y, -Vi.z);
        if (std::fabs(dot(inv, Vr) - 1) < 0.0001f) return rgb(1, 1, 1);
        return rgb(0, 0, 0); }
    case RGK_BXDF_LTC_BECKMANN: case RGK_BXDF_LTC_GGX: { // bxdf.hpp:110-114
        if (Vi.z <= 0 || Vr.z <= 0) return rgb(0, 0, 0);
        RGB c = tex_fetch(s, m.tex_color, uv);
        float p = ltc_pdf(s, m.bxdf == RGK_BXDF_LTC_GGX ? 0 : 1, UPZ, Vi, Vr, m.roughness);
        return rgb(p * c.r, p * c.g, p * c.b); }
    case RGK_BXDF_LTC_BECKMANN_DIFFUSE: case RGK_BXDF_LTC_GGX_DIFFUSE: { // bxdf.hpp:128-136
        if (Vi.z <= 0 || Vr.z <= 0) return rgb(0, 0, 0);
        RGB diff = tex_fetch(s, m.tex_diffuse, uv), spec = tex_fetch(s, m.tex_color, uv);
        float p = ltc_pdf(s, m.bxdf == RGK_BXDF_LTC_GGX_DIFFUSE ? 0 : 1, UPZ, Vi, Vr, m.roughness);
        return rgb(p * spec.r + diff.r / PI_F, p * spec.g + diff.g / PI_F, p * spec.b + diff.b / PI_F); }
    }
    return rgb(0, 0, 0);
}
void bxdf_sample(const Scene& s, uint32_t mi, V3 Vi, V2 uv, V2 sample, V3& dir, RGB& w, bool& may_leak) {
    const rgk_material& m = s.mats[mi];
    may_leak = false;
    switch (m.bxdf) {
    case RGK_BXDF_DIFFUSE: { // bxdf.cpp:197-204
        if (Vi.z <= 0) { dir = v3(0, 1, 0); w = rgb(0, 0, 0); return; }
        dir = hemi_cos_z(sample); w = tex_fetch(s, m.tex_diffuse, uv); return; }
    case RGK_BXDF_MIX: { // bxdf.cpp:241-249
        if (decide_and_rescale(sample.x, m.amount)) bxdf_sample(s, m.mix_a, Vi, uv, sample, dir, w, may_leak);
        else bxdf_sample(s, m.mix_b, Vi, uv, sample, dir, w, may_leak);
        return; }
    case RGK_BXDF_MIRROR: { // bxdf.cpp:272-276
        dir = v3(-Vi.x, -Vi.y, Vi.z); w = tex_fetch(s, m.tex_color, uv); return; }
    case RGK_BXDF_DIELECTRIC: { // bxdf.cpp:380-408
        float eta = (Vi.z < 0) ? m.ior : (float)(1.0 / m.ior);
        float R, cosT; fresnel_dielectric(eta, std::fabs(Vi.z), R, cosT);
        RGB c = tex_fetch(s, m.tex_color, uv);
        if (decide_and_rescale(sample.x, R)) { dir = v3(-Vi.x, -Vi.y, Vi.z); w = c; return; }
        cosT = std::fabs(cosT);
        dir = v3(-Vi.x * eta, -Vi.y * eta, (Vi.z > 0) ? -cosT : cosT); w = c; may_leak = true; return; }
    case RGK_BXDF_TRANSPARENT: { // bxdf.cpp:419-423
        dir = v3(-Vi.x, -Vi.y, -Vi.z); w = rgb(1, 1, 1); may_leak = true; return; }
    case RGK_BXDF_LTC_BECKMANN: case RGK_BXDF_LTC_GGX: { // bxdf.hpp:115-121
        V3 v = hemi_cos_z(sample);
        v = ltc_random(s, m.bxdf == RGK_BXDF_LTC_GGX ? 0 : 1, UPZ, Vi, m.roughness, v);
        dir = v;
        if (v.z <= 0) { w = rgb(0, 0, 0); return; }
        w = tex_fetch(s, m.tex_color, uv); return; }
    case RGK_BXDF_LTC_BECKMANN_DIFFUSE: case RGK_BXDF_LTC_GGX_DIFFUSE: { // bxdf.hpp:137-158
        RGB diff = tex_fetch(s, m.tex_diffuse, uv), spec = tex_fetch(s, m.tex_color, uv);
        float dp = diff.r + diff.g + diff.b, sp = spec.r + spec.g + spec.b;
        float prob = dp / (dp + sp + 0.0001f);
        if (decide_and_rescale(sample.x, prob)) {
            if (Vi.z <= 0) { dir = v3(0, 1, 0); w = rgb(0, 0, 0); return; }
            dir = hemi_cos_z(sample); w = diff; return;
        }
        V3 v = hemi_cos_z(sample);
        v = ltc_random(s, m.bxdf == RGK_BXDF_LTC_GGX_DIFFUSE ? 0 : 1, UPZ, Vi, m.roughness, v);
        dir = v;
        if (v.z <= 0) { w = rgb(0, 0, 0); return; }
        w = spec; return; }
    }
    dir = v3(0, 1, 0); w = rgb(0, 0, 0);
}

// ------------------------------------------------------------------ lights & sky (src/scene.cpp:686-763, src/primitives.cpp:61-73)
Light random_light(const Scene& s, V2 choice, float light_sample, V2 tri_sample) {
    Light none; none.type = 0; none.pos = v3(0, 0, 0); none.color = rgb(0, 0, 0); none.intensity = 0; none.size = 0; none.normal = v3(0, 0, 0); none.valid = false;
    float total = s.total_point_power + s.total_areal_power;
    if (total <= 0.0f) return none;
    float q = choice.x * total;
    if (q < s.total_point_power) {
        for (size_t i = 0; i < s.plights.size(); i++) {
            q -= s.plights[i].intensity * 4.0f * PI_F;
            if (q <= 0.0f) {
                const rgk_point_light& p = s.plights[i];
                Light l; l.type = 0; l.pos = v3(p.position); l.color = rgb(p.color[0], p.color[1], p.color[2]);
                l.intensity = p.intensity; l.size = p.size; l.normal = v3(0, 0, 0); l.valid = true; return l;
            }
        }
        return none;
    }
    q = choice.y * s.total_areal_power;
    for (size_t i = 0; i < s.alights.size(); i++) {
        q -= s.alights[i].power;
        if (q <= 0.0f) {
            const ArealLight& al = s.alights[i];
            float p = light_sample * al.total_area;
            for (size_t j = 0; j < al.tris.size(); j++) {
                p -= al.tris[j].first;
                if (p <= 0.0f) {
                    const Tri& t = s.tris[al.tris[j].second];
                    // Triangle::GetRandomPoint (note b/c naming swap of the reference)
                    V2 r = tri_sample;
                    V3 a = s.pos[t.va], c = s.pos[t.vb], b = s.pos[t.vc];
                    V3 Va = a - c, Vb = b - c;
                    if (r.x + r.y > 1.0f) { r.x = 1.0f - r.x; r.y = 1.0f - r.y; }
                    Light l; l.type = 1; l.pos = c + r.x * Va + r.y * Vb; l.color = al.emission; l.intensity = 1.0f;
                    l.size = 0; l.normal = s.nrm[t.va]; l.valid = true; return l;
                }
            }
            return none;
        }
    }
    return none;
}
RGB sky_radiance(const Scene& s, V3 dir) { // src/scene.cpp:748-763
    if (s.sky.mode == 0) return rgb(s.sky.color[0] * s.sky.intensity, s.sky.color[1] * s.sky.intensity, s.sky.color[2] * s.sky.intensity);
    float alpha = std::asin(dir.y);
    float beta = -std::atan2(dir.x, dir.z);
    beta += s.sky.rotate * 0.0174533f;
    float x = beta / (2.0f * PI_F) + 0.5f, y = alpha / PI_F + 0.5f;
    RGB c = tex_fetch(s, s.sky.envmap, V2{x, y});
    return rgb(c.r * s.sky.intensity, c.g * s.sky.intensity, c.b * s.sky.intensity);
}

// ------------------------------------------------------------------ camera (src/camera.cpp)
void camera_init(rgk_camera* c, V3 pos, V3 la, V3 up, float yview, float xview, int xres, int yres, float focus, float ls) { // :7-24
    V3 direction = normalize(la - pos);
    V3 left = normalize(cross(up, direction));
    V3 cup = normalize(cross(left, direction));
    V3 vx = -xview * left * focus;     // (-xview * left) * focus
    V3 vy = yview * cup * focus;
    V3 vs = pos + direction * focus - 0.5f * vy - 0.5f * vx;
    auto put = [](float* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; };
    put(c->origin, pos); put(c->lookat, la); put(c->direction, direction); put(c->cameraup, cup); put(c->cameraleft, left);
    put(c->viewscreen, vs); put(c->viewscreen_x, vx); put(c->viewscreen_y, vy);
    c->lens_size = ls; c->xsize = xres; c->ysize = yres;
}
Ray camera_ray(const rgk_camera* c, int x, int y, int xres, int yres, V2 off, V2 lens) { // :26-46
    float fx = (x + off.x) / (float)xres, fy = (y + off.y) / (float)yres;
    V3 p = v3(c->viewscreen) + fx * v3(c->viewscreen_x) + fy * v3(c->viewscreen_y);
    V3 o = v3(c->origin);
    if (c->lens_size != 0.0f) {
        V2 d = disc_uniform(lens);
        V2 lenso = V2{d.x * c->lens_size, d.y * c->lens_size};
        o = o + lenso.x * v3(c->cameraleft) + lenso.y * v3(c->cameraup);
    }
    Ray r; r.o = o; r.d = normalize(p - o); // Ray(from, dir), src/ray.hpp:9-13
    return r;
}

// ------------------------------------------------------------------ path tracer (src/path_tracer.cpp)
struct RenderCounters { uint64_t closest = 0, shadow = 0, samples = 0; Counters trav_closest, trav_shadow; };
struct PathPoint {
    bool infinity = false; V3 pos, lightN, faceN; Frame fr; V3 Vr; uint32_t mat = 0; V2 uv; RGB emission; RGB contribution;
    RGB light_from_source{0, 0, 0};   // light path only (src/path_tracer.hpp:57)
};
struct SideEffect { int x, y; RGB q; };   // PixelRenderResult::side_effects, src/tracer.hpp

// PathTracer::GeneratePath, src/path_tracer.cpp:110-306
void generate_path(const Scene& s, const rgk_render_params& P, Ray r, Sampler& smp, std::vector<PathPoint>& path, RenderCounters& rc,
                   unsigned depth, float russian) {
    path.clear();
    RGB cum = rgb(1, 1, 1);
    Ray cur = r;
    unsigned n = 0;
    uint32_t last = RGK_NO_TRIANGLE;
    while (n < depth) {
        n++;
        rc.closest++;
        Hit i = find_intersect(s, cur, last, &rc.trav_closest);
        PathPoint p;
        p.contribution = cum;
        if (i.tri == RGK_NO_TRIANGLE) { p.infinity = true; p.Vr = -cur.d; path.push_back(p); break; }
        const Tri& tr = s.tris[i.tri];
        p.pos = cur.o + i.t * cur.d; // Ray::t, src/ray.hpp:27
        auto interp = [&](V3 x, V3 y, V3 z) { return i.a * x + i.b * y + i.c * z; };
        p.faceN = interp(s.nrm[tr.va], s.nrm[tr.vb], s.nrm[tr.vc]);
        if (std::isnan(p.faceN.x)) {
            p.faceN = s.nrm[tr.va];
            if (std::isnan(p.faceN.x)) { p.faceN = s.nrm[tr.vb];
                if (std::isnan(p.faceN.x)) { p.faceN = s.nrm[tr.vc];
                    if (std::isnan(p.faceN.x)) return; } }
        }
        if (length(p.faceN) <= 0.0f) return;
        p.faceN = normalize(p.faceN);
        p.Vr = -cur.d;
        const rgk_material& mat = s.mats[tr.mat];
        p.mat = tr.mat;
        V2 ta = s.uv[tr.va], tb = s.uv[tr.vb], tc = s.uv[tr.vc];
        p.uv = V2{i.a * ta.x + i.b * tb.x + i.c * tc.x, i.a * ta.y + i.b * tb.y + i.c * tc.y};
        p.emission = rgb(mat.emission[0], mat.emission[1], mat.emission[2]);
        if (mat.tex_bump >= 0) { // !mat.bumpmap->Empty()
            float right, bottom; tex_slopes(s, mat.tex_bump, p.uv, right, bottom);
            V3 tangent = interp(s.tan[tr.va], s.tan[tr.vb], s.tan[tr.vc]);
            if (tangent.x * tangent.x + tangent.y * tangent.y + tangent.z * tangent.z < 0.001f) p.lightN = p.faceN;
            else {
                tangent = normalize(tangent);
                V3 bitangent = normalize(cross(p.faceN, tangent));
                V3 tangent2 = cross(bitangent, p.faceN);
                p.lightN = normalize(p.faceN + (tangent2 * right + bitangent * bottom) * P.bumpmap_scale);
                if (std::isnan(p.lightN.x)) p.lightN = p.faceN;
            }
        } else p.lightN = p.faceN;
        p.fr = system_transform(p.lightN, UPZ);
        V2 sample = smp.get2d();
        V3 dir; RGB tc_; bool may_leak;
        bxdf_sample(s, tr.mat, p.fr.toLocal(p.Vr), p.uv, sample, dir, tc_, may_leak);
        bool inside = dir.z < 0;
        dir = p.fr.toGlobal(dir);
        if (!(dot(dir, p.faceN) * dot(p.Vr, p.faceN) > 0) && !may_leak) n += 10000;
        float rcoef = (!mat.no_russian && russian > 0.0f && n > 1) ? 1.0f / russian : 1.0f;
        cum = rgb(rcoef * cum.r, rcoef * cum.g, rcoef * cum.b);         // Spectrum *= float : q*r
        cum = rgb(tc_.r * cum.r, tc_.g * cum.g, tc_.b * cum.b);         // Spectrum *= Spectrum : o.r*r
        path.push_back(p);
        if (rgbmax(cum) < 0.001f) break;
        if (!mat.no_russian && russian >= 0.0f && smp.get1d() > russian) break;
        if (n > depth) break;
        Ray nr; nr.o = p.pos + p.faceN * s.epsilon * 10.0f * (inside ? -1.0f : 1.0f);
        nr.d = normalize(normalize(dir)); // glm::normalize(dir) then Ray(from,dir) normalises again
        cur = nr;
        last = i.tri;
    }
}

// Camera::GetCoordsFromDirection, src/camera.cpp:48-83
bool coords_from_direction(const rgk_camera* c, V3 dir, int& x, int& y) {
    const V3 N = v3(c->direction);
    const float q = dot(dir, N);
    if (q < 0.0001) return false;
    const float t = dot(v3(c->viewscreen) - v3(c->origin), N) / q;
    if (t <= 0) return false;
    const V3 p = v3(c->origin) + dir * t;
    const V3 v1 = v3(c->viewscreen_x), v2 = v3(c->viewscreen_y);
    const V3 vp = p - v3(c->viewscreen);
    const float plen = length(vp);
    const float v1_cast_len = plen * dot(normalize(vp), normalize(v1));
    const float v2_cast_len = plen * dot(normalize(vp), normalize(v2));
    const float x_ratio = v1_cast_len / length(v1), y_ratio = v2_cast_len / length(v2);
    if (x_ratio < 0.0f || x_ratio > 1.0f || y_ratio < 0.0f || y_ratio > 1.0f) return false;
    x = (int)(c->xsize * x_ratio); y = (int)(c->ysize * y_ratio);
    return true;
}

// PathTracer::TracePath, src/path_tracer.cpp:308-512.  reverse == 0: the light path is empty (GeneratePath with depth 0
// draws nothing); reverse > 0: a light path of up to `reverse` vertices, its connections to the camera (side effects:
// radiance splatted to other pixels with count 0) and to every vertex of the camera path.
RGB trace_path(const Scene& s, const rgk_camera* cam, const rgk_render_params& P, const Ray& r, Sampler& smp, std::vector<PathPoint>& path,
               std::vector<PathPoint>& light_path, std::vector<SideEffect>& side_effects, RenderCounters& rc) {
    const V3 camerapos = r.o;
    V2 areal_sample = smp.get2d();
    V2 lightdir_sample = smp.get2d();
    V2 choice = smp.get2d(); float ls = smp.get1d();
    Light light = random_light(s, choice, ls, areal_sample);
    generate_path(s, P, r, smp, path, rc, P.depth, P.russian);
    V3 main_light_dir = v3(0, 0, 0);
    if (light.type == 0) {
        V3 dir = sphere_uniform(areal_sample); light.pos = light.pos + light.size * dir;
        if (P.reverse) main_light_dir = hemi_cos_directed(lightdir_sample, normalize(dir));
    } else if (P.reverse) main_light_dir = hemi_cos_directed(lightdir_sample, light.normal);
    light_path.clear();
    if (P.reverse && light.valid) {
        Ray light_ray; light_ray.o = light.pos + s.epsilon * light.normal * 100.0f; light_ray.d = normalize(main_light_dir);
        generate_path(s, P, light_ray, smp, light_path, rc, P.reverse, -1.0f);
        const float dfac = (light.type == 0) ? 1.0f : gmax(0.0f, dot(main_light_dir, light.normal));
        const float k0 = light.intensity * dfac;
        const RGB light_at_path_start = rgb(light.color.r * k0, light.color.g * k0, light.color.b * k0);
        for (PathPoint& p : light_path) {
            const RGB light_here = rgb(p.contribution.r * light_at_path_start.r, p.contribution.g * light_at_path_start.g, p.contribution.b * light_at_path_start.b);
            p.light_from_source = light_here;
            if (p.infinity) continue;
            rc.shadow++;
            if (!visibility(s, p.pos, camerapos, &rc.trav_shadow)) continue;
            const V3 direction = normalize(p.pos - camerapos);
            const RGB f = bxdf_value(s, p.mat, p.fr.toLocal(p.Vr), p.fr.toLocal(-direction), p.uv);
            RGB q = rgb(light_here.r * f.r, light_here.g * f.g, light_here.b * f.b);
            const float G = gmax(0.0f, dot(p.lightN, -direction)) / distance2(camerapos, p.pos);
            if (G >= 0.00001f && !std::isnan(q.r)) {
                q = rgb(q.r * G, q.g * G, q.b * G);
                int x2, y2;
                // (the reference accepts x_ratio == 1, which addresses one pixel past the row: dropped here)
                if (coords_from_direction(cam, direction, x2, y2) && x2 < cam->xsize && y2 < cam->ysize) side_effects.push_back(SideEffect{x2, y2, q});
            }
        }
    }
    RGB total = rgb(0, 0, 0);
    for (size_t n = 0; n < path.size(); n++) {
        const PathPoint& p = path[n];
        if (p.infinity) {
            RGB sky = sky_radiance(s, p.Vr);
            total = rgb(total.r + sky.r * p.contribution.r, total.g + sky.g * p.contribution.g, total.b + sky.b * p.contribution.b);
            continue;
        }
        RGB here = rgb(0, 0, 0);
        if (light.valid) {
            rc.shadow++;
            if (visibility(s, light.pos, p.pos, &rc.trav_shadow)) {
                V3 Vi = normalize(light.pos - p.pos);
                RGB f = bxdf_value(s, p.mat, p.fr.toLocal(Vi), p.fr.toLocal(p.Vr), p.uv);
                float G = std::fabs(dot(p.lightN, Vi)) / distance2(light.pos, p.pos);
                float df = (light.type == 0) ? 1.0f : gmax(0.0f, dot(-Vi, light.normal));
                float k = light.intensity * df;
                RGB inc = rgb(light.color.r * k, light.color.g * k, light.color.b * k);
                RGB fg = rgb(G * f.r, G * f.g, G * f.b);
                here = rgb(here.r + inc.r * fg.r, here.g + inc.g * fg.g, here.b + inc.b * fg.b);
            }
        }
        for (const PathPoint& l : light_path) {                       // "Reverse light", src/path_tracer.cpp:462-480
            if (l.infinity) continue;
            rc.shadow++;
            if (!visibility(s, l.pos, p.pos, &rc.trav_shadow)) continue;
            const V3 light_to_p = normalize(p.pos - l.pos), p_to_light = -light_to_p;
            const RGB f_light = bxdf_value(s, l.mat, l.fr.toLocal(light_to_p), l.fr.toLocal(l.Vr), l.uv);
            const RGB f_point = bxdf_value(s, p.mat, p.fr.toLocal(p.Vr), p.fr.toLocal(p_to_light), p.uv);
            const float G = std::fabs(dot(p.lightN, p_to_light)) / distance2(l.pos, p.pos);
            RGB ff = rgb(f_point.r * f_light.r, f_point.g * f_light.g, f_point.b * f_light.b);   // Spectrum * Spectrum: o.r * r
            ff = rgb(G * ff.r, G * ff.g, G * ff.b);
            here = rgb(here.r + l.light_from_source.r * ff.r, here.g + l.light_from_source.g * ff.g, here.b + l.light_from_source.b * ff.b);
        }
        if (dot(p.faceN, p.Vr) > 0) here = rgb(here.r + p.emission.r, here.g + p.emission.g, here.b + p.emission.b);
        rgbclamp(here, P.clamp);
        total = rgb(total.r + here.r * p.contribution.r, total.g + here.g * p.contribution.g, total.b + here.b * p.contribution.b);
    }
    rgbclamp(total, P.clamp);
    if (std::isnan(total.r) || total.r < 0.0f) total.r = 0.0f;
    if (std::isnan(total.g) || total.g < 0.0f) total.g = 0.0f;
    if (std::isnan(total.b) || total.b < 0.0f) total.b = 0.0f;
    return total;
}

// Tracer::Render + PathTracer::RenderPixel for one task, src/tracer.cpp:6-37, src/path_tracer.cpp:42-78
void render_task(const Scene& s, const rgk_camera* cam, const rgk_render_params& P, const rgk_task& t, uint32_t seed,
                 float* rgb_sum, uint32_t* count, RenderCounters& rc) {
    std::vector<PathPoint> path, light_path;
    std::vector<SideEffect> side_effects;
    for (uint32_t y = t.y1; y < t.y2; y++)
        for (uint32_t x = t.x1; x < t.x2; x++) {
            seed += 0x42424242u;
            Sampler smp(seed, 64, P.multisample);
            RGB tot = rgb(0, 0, 0);
            for (uint32_t i = 0; i < P.multisample; i++) {
                smp.advance();
                V2 coords = smp.get2d();
                V2 lens = V2{0, 0};
                if (cam->lens_size != 0.0f) lens = smp.get2d();
                Ray r = camera_ray(cam, x, y, P.xres, P.yres, coords, lens);
                RGB q = trace_path(s, cam, P, r, smp, path, light_path, side_effects, rc);
                tot = rgb(tot.r + q.r, tot.g + q.g, tot.b + q.b);
                rc.samples++;
            }
            const size_t px = (size_t)y * P.xres + x;
            rgb_sum[3 * px] += tot.r; rgb_sum[3 * px + 1] += tot.g; rgb_sum[3 * px + 2] += tot.b;
            count[px] += P.multisample;
            for (const SideEffect& e : side_effects) {              // AddPixel(x2, y2, r, 0), src/tracer.cpp:20-26
                const size_t q = (size_t)e.y * P.xres + e.x;
                rgb_sum[3 * q] += e.q.r; rgb_sum[3 * q + 1] += e.q.g; rgb_sum[3 * q + 2] += e.q.b;
            }
            side_effects.clear();
        }
}

void run_parallel(uint64_t n, int nthreads, const std::function<void(uint64_t, uint64_t, int)>& f) {
    if (nthreads <= 1 || n < 2) { f(0, n, 0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) {
        uint64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
        th.emplace_back([=, &f] { f(lo, hi, t); });
    }
    for (auto& t : th) t.join();
}

} // namespace

// =================================================================== C ABI
extern "C" {

const char* rgko_describe(void) { return "rgk_oracle: scalar CPU restatement of the RGKrt hot path (test infrastructure)"; }

void* rgko_scene_create(const rgk_scene_desc* d, const rgk_kdtree* tree) {
    Scene* s = new Scene();
    s->pos.resize(d->n_vertices); s->nrm.resize(d->n_vertices); s->tan.resize(d->n_vertices); s->uv.resize(d->n_vertices);
    for (uint32_t i = 0; i < d->n_vertices; i++) {
        s->pos[i] = v3(d->positions + 3 * i); s->nrm[i] = v3(d->normals + 3 * i); s->tan[i] = v3(d->tangents + 3 * i);
        s->uv[i] = V2{d->texcoords[2 * i], d->texcoords[2 * i + 1]};
    }
    s->mats.assign(d->materials, d->materials + d->n_materials);
    for (uint32_t i = 0; i < d->n_textures; i++) {
        Tex t; t.kind = d->textures[i].kind; t.w = d->textures[i].width; t.h = d->textures[i].height;
        t.color = rgb(d->textures[i].color[0], d->textures[i].color[1], d->textures[i].color[2]);
        if (t.kind == 1) t.texels.assign(d->textures[i].texels, d->textures[i].texels + 3 * (size_t)t.w * t.h);
        s->texs.push_back(std::move(t));
    }
    s->tris.resize(d->n_triangles);
    for (uint32_t mi = 0; mi < d->n_meshes; mi++) {
        const rgk_mesh& m = d->meshes[mi];
        const float* e = d->materials[m.material].emission;
        const bool light_source = e[0] > 0 || e[1] > 0 || e[2] > 0; // Radiance::isNonZero
        ArealLight al;
        for (uint32_t t = 0; t < m.n_triangles; t++) {
            const uint32_t ti = m.first_triangle + t;
            Tri& tr = s->tris[ti];
            tr.va = d->indices[3 * ti]; tr.vb = d->indices[3 * ti + 1]; tr.vc = d->indices[3 * ti + 2]; tr.mat = m.material;
            if (light_source) al.tris.push_back(std::make_pair(0.0f, ti));
        }
        if (light_source && !al.tris.empty()) s->alights.push_back(al);
    }
    s->plights.assign(d->point_lights, d->point_lights + d->n_point_lights);
    s->sky = d->sky;
    const rgk_ltc_table* lt[2] = {&d->ltc_ggx, &d->ltc_beckmann};
    for (int k = 0; k < 2; k++)
        if (lt[k]->M) { s->ltcM[k].assign(lt[k]->M, lt[k]->M + 4096 * 9); s->ltcA[k].assign(lt[k]->amplitude, lt[k]->amplitude + 4096); }
    s->thinglass = d->thinglass != 0;
    commit(*s, d, tree);
    return s;
}
void rgko_scene_destroy(void* h) { delete (Scene*)h; }

static void depth_rec(const uint32_t* N, uint32_t i, uint32_t d, uint32_t& mx) {
    mx = std::max(mx, d);
    if ((N[2 * i + 1] & 3) == 3) return;
    depth_rec(N, i + 1, d + 1, mx); depth_rec(N, N[2 * i + 1] >> 2, d + 1, mx);
}
int rgko_scene_get_info(void* h, rgk_scene_info* o) {
    Scene& s = *(Scene*)h;
    o->epsilon = s.epsilon;
    for (int ax = 0; ax < 3; ax++) { o->bbox[2 * ax] = s.bb[ax][0]; o->bbox[2 * ax + 1] = s.bb[ax][1]; }
    o->n_nodes = s.nodes.size() / 2; o->n_refs = s.refs.size(); o->n_triangles = s.tris.size(); o->n_areal_lights = s.alights.size();
    o->max_depth = 0; if (!s.nodes.empty()) depth_rec(s.nodes.data(), 0, 0, o->max_depth);
    o->total_point_power = s.total_point_power; o->total_areal_power = s.total_areal_power;
    return 0;
}
int rgko_scene_get_kdtree(void* h, uint32_t* nodes, uint32_t* refs) {
    Scene& s = *(Scene*)h;
    std::memcpy(nodes, s.nodes.data(), 4 * s.nodes.size()); std::memcpy(refs, s.refs.data(), 4 * s.refs.size());
    return 0;
}
int rgko_scene_get_planes(void* h, float* planes) {
    Scene& s = *(Scene*)h;
    for (size_t i = 0; i < s.tris.size(); i++) std::memcpy(planes + 4 * i, s.tris[i].p, 16);
    return 0;
}

int rgko_trace_closest(void* h, const rgk_ray* rays, const uint32_t* ignore, uint64_t n, rgk_hit* hits, rgk_trav_stats* st, int nthreads) {
    const Scene& s = *(Scene*)h;
    std::vector<Counters> cs(std::max(1, nthreads));
    run_parallel(n, nthreads, [&](uint64_t lo, uint64_t hi, int tid) {
        Counters c;
        for (uint64_t i = lo; i < hi; i++) {
            Ray r; r.o = v3(rays[i].origin); r.d = v3(rays[i].direction); r.tnear = rays[i].tnear; r.tfar = rays[i].tfar;
            Hit ht = find_intersect(s, r, ignore ? ignore[i] : RGK_NO_TRIANGLE, st ? &c : nullptr);
            hits[i].triangle = ht.tri; hits[i].t = ht.t; hits[i].a = ht.a; hits[i].b = ht.b; hits[i].c = ht.c;
        }
        cs[tid] = c;
    });
    if (st) { std::memset(st, 0, sizeof *st); st->rays = n; for (auto& c : cs) { st->inner += c.inner; st->leaf += c.leaf; st->refs += c.refs; st->tests += c.tests; } }
    return 0;
}
int rgko_trace_shadow(void* h, const float* a, const float* b, uint64_t n, uint8_t* visible, rgk_trav_stats* st, int nthreads) {
    const Scene& s = *(Scene*)h;
    std::vector<Counters> cs(std::max(1, nthreads));
    run_parallel(n, nthreads, [&](uint64_t lo, uint64_t hi, int tid) {
        Counters c;
        for (uint64_t i = lo; i < hi; i++) visible[i] = visibility(s, v3(a + 3 * i), v3(b + 3 * i), st ? &c : nullptr) ? 1 : 0;
        cs[tid] = c;
    });
    if (st) { std::memset(st, 0, sizeof *st); st->rays = n; for (auto& c : cs) { st->inner += c.inner; st->leaf += c.leaf; st->refs += c.refs; st->tests += c.tests; } }
    return 0;
}

void rgko_camera_init(rgk_camera* c, const float pos[3], const float la[3], const float up[3], float yview, float xview,
                      int32_t xres, int32_t yres, float focus_plane, float lens_size) {
    camera_init(c, v3(pos), v3(la), v3(up), yview, xview, xres, yres, focus_plane, lens_size);
}
int rgko_camera_rays(const rgk_camera* c, uint32_t xres, uint32_t yres, const int32_t* xy, const float* off, const float* lens, uint64_t n, rgk_ray* rays) {
    for (uint64_t i = 0; i < n; i++) {
        Ray r = camera_ray(c, xy[2 * i], xy[2 * i + 1], xres, yres, V2{off[2 * i], off[2 * i + 1]}, lens ? V2{lens[2 * i], lens[2 * i + 1]} : V2{0, 0});
        rays[i].origin[0] = r.o.x; rays[i].origin[1] = r.o.y; rays[i].origin[2] = r.o.z;
        rays[i].direction[0] = r.d.x; rays[i].direction[1] = r.d.y; rays[i].direction[2] = r.d.z;
        rays[i].tnear = r.tnear; rays[i].tfar = r.tfar;
    }
    return 0;
}

// GenerateTaskList, src/render_driver.cpp:30-46 (+ RenderTask::midpoint, src/tracer.hpp:18)
uint32_t rgko_generate_tasks(uint32_t tile, uint32_t xres, uint32_t yres, rgk_task* out, uint32_t cap) {
    struct T { rgk_task t; float mx, my; };
    std::vector<T> tasks;
    for (uint32_t yp = 0; yp < yres; yp += tile)
        for (uint32_t xp = 0; xp < xres; xp += tile) {
            T t; t.t = rgk_task{xp, std::min(xres, xp + tile), yp, std::min(yres, yp + tile)};
            t.mx = (t.t.x1 + t.t.x2) / 2.0f; t.my = (t.t.y1 + t.t.y2) / 2.0f;
            tasks.push_back(t);
        }
    const float cx = xres / 2.0f, cy = yres / 2.0f;
    auto dist = [&](const T& a) { float dx = cx - a.mx, dy = cy - a.my; return std::sqrt(dx * dx + dy * dy); };
    std::sort(tasks.begin(), tasks.end(), [&](const T& a, const T& b) { return dist(a) < dist(b); });
    for (uint32_t i = 0; i < tasks.size() && i < cap; i++) out[i] = tasks[i].t;
    return (uint32_t)tasks.size();
}

uint32_t rgko_sampler_set_size(uint32_t ms) { return round_up_to_square(ms); }
int rgko_sampler_tables(const uint32_t* seeds, uint32_t n_seeds, uint32_t ms, uint32_t n1d, uint32_t n2d, float* out1d, float* out2d) {
    for (uint32_t si = 0; si < n_seeds; si++) {
        Sampler s(seeds[si], 64, ms);
        s.prepare();
        const uint32_t ss = s.set_size;
        for (uint32_t d = 0; d < n1d; d++) std::memcpy(out1d + ((size_t)si * n1d + d) * ss, &s.s1[(size_t)d * ss], 4 * (size_t)ss);
        for (uint32_t d = 0; d < n2d; d++) std::memcpy(out2d + ((size_t)si * n2d + d) * ss * 2, &s.s2[(size_t)d * ss], 8 * (size_t)ss);
    }
    return 0;
}
// raw mt19937 stream (unit test of the generator)
int rgko_mt19937(uint32_t seed, uint32_t n, uint32_t* out) { MT19937 g(seed); for (uint32_t i = 0; i < n; i++) out[i] = g.next(); return 0; }

// RenderDriver::RenderRound, src/render_driver.cpp:144-190: tasks are independent; worker threads pull
// task indices from an atomic counter (the reference uses a ctpl pool); every pixel belongs to one task,
// so adding straight into the shared framebuffer equals the reference's per-task buffer + Accumulate.
int rgko_render_round(void* h, const rgk_camera* cam, const rgk_render_params* P, const rgk_task* tasks, uint32_t n_tasks,
                      uint32_t seedstart, uint32_t seedcount_base, float* rgb_sum, uint32_t* count, rgk_round_stats* st, int nthreads) {
    const Scene& s = *(Scene*)h;
    nthreads = std::max(1, nthreads);
    std::vector<RenderCounters> rcs(nthreads);
    std::atomic<uint32_t> next(0);
    auto t0 = std::chrono::high_resolution_clock::now();
    // reverse > 0: side effects land on pixels of other tasks, so every task renders into its own full-frame buffer which
    // is then added to the total (EXRTexture::Accumulate, src/render_driver.cpp:176-181); buffers are added in task
    // order, which is what the reference does with one worker thread (with more, its order is a race).
    const size_t npx = (size_t)P->xres * P->yres;
    std::vector<std::vector<float>> tsum; std::vector<std::vector<uint32_t>> tcnt;
    if (P->reverse) { tsum.resize(n_tasks); tcnt.resize(n_tasks); }
    auto worker = [&](int tid) {
        for (;;) {
            uint32_t i = next.fetch_add(1);
            if (i >= n_tasks) break;
            if (P->reverse) {
                tsum[i].assign(3 * npx, 0.0f); tcnt[i].assign(npx, 0u);
                render_task(s, cam, *P, tasks[i], seedstart + seedcount_base + i, tsum[i].data(), tcnt[i].data(), rcs[tid]);
            } else render_task(s, cam, *P, tasks[i], seedstart + seedcount_base + i, rgb_sum, count, rcs[tid]);
        }
    };
    if (nthreads == 1) worker(0);
    else { std::vector<std::thread> th; for (int t = 0; t < nthreads; t++) th.emplace_back(worker, t); for (auto& t : th) t.join(); }
    for (uint32_t i = 0; P->reverse && i < n_tasks; i++) {
        for (size_t k = 0; k < 3 * npx; k++) rgb_sum[k] += tsum[i][k];
        for (size_t k = 0; k < npx; k++) count[k] += tcnt[i][k];
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    if (st) {
        std::memset(st, 0, sizeof *st);
        for (auto& r : rcs) { st->closest_rays += r.closest; st->shadow_rays += r.shadow; st->samples += r.samples; }
        st->gpu_ms = std::chrono::duration<float, std::milli>(t1 - t0).count(); // wall-clock ms of the CPU round
    }
    return 0;
}
// Traversal work counters of a render (for SURVEY 8d's B_sample): closest then shadow.
int rgko_render_round_counters(void* h, const rgk_camera* cam, const rgk_render_params* P, const rgk_task* tasks, uint32_t n_tasks,
                               uint32_t seedstart, uint32_t seedcount_base, rgk_trav_stats* closest, rgk_trav_stats* shadow) {
    const Scene& s = *(Scene*)h;
    std::vector<float> fb(3 * (size_t)P->xres * P->yres, 0.0f); std::vector<uint32_t> cnt((size_t)P->xres * P->yres, 0);
    RenderCounters rc;
    for (uint32_t i = 0; i < n_tasks; i++) render_task(s, cam, *P, tasks[i], seedstart + seedcount_base + i, fb.data(), cnt.data(), rc);
    closest->rays = rc.closest; closest->inner = rc.trav_closest.inner; closest->leaf = rc.trav_closest.leaf; closest->refs = rc.trav_closest.refs; closest->tests = rc.trav_closest.tests;
    shadow->rays = rc.shadow; shadow->inner = rc.trav_shadow.inner; shadow->leaf = rc.trav_shadow.leaf; shadow->refs = rc.trav_shadow.refs; shadow->tests = rc.trav_shadow.tests;
    return 0;
}

#define RGK_ORACLE_BVH_STACK 256
// ---- design study (NOT a reference path): what a wide BVH would cost and how often its hit differs ----------------
// A binned-SAH BVH over the same triangles, collapsed to `width` children per node, traversed front to back for the
// GLOBAL closest hit with the same Triangle::TestIntersection arithmetic.  Reports, per ray, nodes visited, child boxes
// tested and triangles tested, and how many rays end on a different triangle than FindIntersectKdOtherThan (exact ties
// and epsilon cases at kd leaf boundaries).  Used by tests/bvh_study.py for DESIGN.md "Next"; never by a test gate.
namespace {
struct BNode { float lo[3], hi[3]; int left, right, first, count; };   // binary; count > 0 = leaf over order[first..]
struct WNode { int nchild; float lo[8][3], hi[8][3]; int child[8]; int first[8], count[8]; };   // child < 0: leaf slot
struct BvhStudy {
    std::vector<BNode> b; std::vector<uint32_t> order; std::vector<WNode> w;
};
void tri_box(const Scene& s, uint32_t t, float lo[3], float hi[3]) {
    const V3 v[3] = {s.pos[s.tris[t].va], s.pos[s.tris[t].vb], s.pos[s.tris[t].vc]};
    for (int k = 0; k < 3; k++) { lo[k] = std::min(v[0][k], std::min(v[1][k], v[2][k])); hi[k] = std::max(v[0][k], std::max(v[1][k], v[2][k])); }
}
float half_area(const float lo[3], const float hi[3]) { const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2]; return x * y + y * z + z * x; }
int build_binary(const Scene& s, BvhStudy& B, int first, int count, int leaf_size) {
    BNode n; for (int k = 0; k < 3; k++) { n.lo[k] = 1e30f; n.hi[k] = -1e30f; }
    float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = first; i < first + count; i++) {
        float lo[3], hi[3]; tri_box(s, B.order[i], lo, hi);
        for (int k = 0; k < 3; k++) { n.lo[k] = std::min(n.lo[k], lo[k]); n.hi[k] = std::max(n.hi[k], hi[k]); const float c = 0.5f * (lo[k] + hi[k]); clo[k] = std::min(clo[k], c); chi[k] = std::max(chi[k], c); }
    }
    n.left = n.right = -1; n.first = first; n.count = count;
    const int me = (int)B.b.size(); B.b.push_back(n);
    if (count <= leaf_size) return me;
    int axis = 0; for (int k = 1; k < 3; k++) if (chi[k] - clo[k] > chi[axis] - clo[axis]) axis = k;
    if (!(chi[axis] > clo[axis])) return me;
    const int NB = 16; int cnt[NB] = {0}; float blo[NB][3], bhi[NB][3];
    for (int q = 0; q < NB; q++) for (int k = 0; k < 3; k++) { blo[q][k] = 1e30f; bhi[q][k] = -1e30f; }
    const float scale = NB / (chi[axis] - clo[axis]);
    auto bin_of = [&](uint32_t t) { float lo[3], hi[3]; tri_box(s, t, lo, hi); int q = (int)((0.5f * (lo[axis] + hi[axis]) - clo[axis]) * scale); return std::min(NB - 1, std::max(0, q)); };
    for (int i = first; i < first + count; i++) {
        const int q = bin_of(B.order[i]); float lo[3], hi[3]; tri_box(s, B.order[i], lo, hi); cnt[q]++;
        for (int k = 0; k < 3; k++) { blo[q][k] = std::min(blo[q][k], lo[k]); bhi[q][k] = std::max(bhi[q][k], hi[k]); }
    }
    float best = 1e30f; int split = -1;
    for (int sp = 1; sp < NB; sp++) {
        float l0[3] = {1e30f, 1e30f, 1e30f}, h0[3] = {-1e30f, -1e30f, -1e30f}, l1[3] = {1e30f, 1e30f, 1e30f}, h1[3] = {-1e30f, -1e30f, -1e30f}; int c0 = 0, c1 = 0;
        for (int q = 0; q < sp; q++) if (cnt[q]) { c0 += cnt[q]; for (int k = 0; k < 3; k++) { l0[k] = std::min(l0[k], blo[q][k]); h0[k] = std::max(h0[k], bhi[q][k]); } }
        for (int q = sp; q < NB; q++) if (cnt[q]) { c1 += cnt[q]; for (int k = 0; k < 3; k++) { l1[k] = std::min(l1[k], blo[q][k]); h1[k] = std::max(h1[k], bhi[q][k]); } }
        if (!c0 || !c1) continue;
        const float cost = c0 * half_area(l0, h0) + c1 * half_area(l1, h1);
        if (cost < best) { best = cost; split = sp; }
    }
    if (split < 0) return me;
    auto mid = std::partition(B.order.begin() + first, B.order.begin() + first + count, [&](uint32_t t) { return bin_of(t) < split; });
    const int nl = (int)(mid - (B.order.begin() + first));
    if (nl == 0 || nl == count) return me;
    const int l = build_binary(s, B, first, nl, leaf_size), r = build_binary(s, B, first + nl, count - nl, leaf_size);
    B.b[me].left = l; B.b[me].right = r; B.b[me].count = 0;
    return me;
}
int collapse(BvhStudy& B, int bn, int width) {
    std::vector<int> kids = {bn};
    for (;;) {          // open the child with the largest surface until `width` children
        int pick = -1; float area = -1.0f;
        for (size_t i = 0; i < kids.size(); i++) { const BNode& n = B.b[kids[i]]; if (n.count == 0) { const float a = half_area(n.lo, n.hi); if (a > area) { area = a; pick = (int)i; } } }
        if (pick < 0 || (int)kids.size() >= width) break;
        const BNode n = B.b[kids[pick]]; kids[pick] = n.left; kids.push_back(n.right);
    }
    const int me = (int)B.w.size(); B.w.emplace_back();
    WNode wn{}; wn.nchild = (int)kids.size();
    for (size_t i = 0; i < kids.size(); i++) {
        const BNode& n = B.b[kids[i]];
        for (int k = 0; k < 3; k++) { wn.lo[i][k] = n.lo[k]; wn.hi[i][k] = n.hi[k]; }
        if (n.count > 0) { wn.child[i] = -1; wn.first[i] = n.first; wn.count[i] = n.count; }
        else { wn.child[i] = 0; wn.first[i] = kids[i]; wn.count[i] = 0; }
    }
    B.w[me] = wn;
    for (int i = 0; i < wn.nchild; i++) if (wn.child[i] == 0) { const int c = collapse(B, wn.first[i], width); B.w[me].child[i] = c; }
    return me;
}
} // namespace

// out[0..5] = rays, wide nodes visited, child boxes tested, triangles tested, rays whose triangle differs from the kd-tree's,
// of those how many have |dt| <= 2 eps;  out[6] = wide nodes, out[7] = bytes of a 32-byte-per-child wide-node layout,
// out[8] = rays with more than one hit within 2 eps of the closest (the ones an arbiter would hand to the kd traversal),
// out[9] = rays that differ WITHOUT being flagged so (must be 0 for the arbiter plan to be bit-exact)
int rgko_bvh_study(void* h, const rgk_ray* rays, const uint32_t* ignore, uint64_t n, int width, int leaf_size, double* out) {
    const Scene& s = *(Scene*)h;
    BvhStudy B; B.order.resize(s.tris.size());
    for (size_t i = 0; i < B.order.size(); i++) B.order[i] = (uint32_t)i;
    build_binary(s, B, 0, (int)B.order.size(), leaf_size);
    collapse(B, 0, width);
    uint64_t nodes = 0, boxes = 0, tests = 0, differ = 0, near_tie = 0, ambiguous = 0, differ_unflagged = 0;
    const float window = 2.0f * s.epsilon;      // the arbiter's window: every hit within it of the closest one is kept
    std::vector<std::pair<float, uint32_t>> found;
    for (uint64_t i = 0; i < n; i++) {
        found.clear();
        Ray r; r.o = v3(rays[i].origin); r.d = v3(rays[i].direction); r.tnear = rays[i].tnear; r.tfar = rays[i].tfar;
        const uint32_t ign = ignore ? ignore[i] : RGK_NO_TRIANGLE;
        const float inv[3] = {1.f / r.d.x, 1.f / r.d.y, 1.f / r.d.z};
        float best_t = r.tfar; uint32_t best = RGK_NO_TRIANGLE;
        struct E { int node; float t; }; E stack[256]; int sp = 0; stack[sp++] = {0, r.tnear - s.epsilon};
        while (sp) {
            const E e = stack[--sp];
            if (e.t > best_t + window) continue;
            const WNode& wn = B.w[e.node]; nodes++;
            E hitc[9]; int hc = 0; int leafc[8]; int lc = 0;
            for (int c = 0; c < wn.nchild; c++) {
                boxes++;
                float t0 = r.tnear - s.epsilon, t1 = best_t + window;
                for (int k = 0; k < 3; k++) { float a = (wn.lo[c][k] - r.o[k]) * inv[k], b = (wn.hi[c][k] - r.o[k]) * inv[k]; if (a > b) std::swap(a, b); t0 = std::max(t0, a - 1e-4f * std::fabs(a)); t1 = std::min(t1, b + 1e-4f * std::fabs(b)); }
                if (t0 > t1) continue;
                if (wn.child[c] < 0) leafc[lc++] = c; else hitc[hc++] = {wn.child[c], t0};
            }
            for (int q = 0; q < lc; q++) {
                const int c = leafc[q];
                for (int j = wn.first[c]; j < wn.first[c] + wn.count[c]; j++) {
                    const uint32_t ti = B.order[j];
                    if (ti == ign) continue;
                    tests++;
                    float t, a, b;
                    if (test_intersection(s, s.tris[ti], r, t, a, b) && t >= r.tnear - s.epsilon && t <= r.tfar + s.epsilon) {
                        if (t <= best_t + window) found.push_back({t, ti});
                        if (t < best_t || (t == best_t && ti < best)) { best_t = t; best = ti; }
                    }
                }
            }
            for (int a = 1; a < hc; a++) { const E x = hitc[a]; int b = a; while (b > 0 && hitc[b - 1].t < x.t) { hitc[b] = hitc[b - 1]; b--; } hitc[b] = x; }   // far first: nearest popped first
            for (int q = 0; q < hc; q++) stack[sp++] = hitc[q];
        }
        int close = 0;
        for (const auto& f : found) if (f.first <= best_t + window) close++;
        const bool amb = close > 1 || (best != RGK_NO_TRIANGLE && best_t < r.tnear + s.epsilon);   // also: a hit behind / at the offset origin
        if (amb) ambiguous++;
        const Hit k = find_intersect(s, r, ign, nullptr);
        if (k.tri != best && !amb) differ_unflagged++;
        if (k.tri != best) { if (getenv("RGKO_BVH_VERBOSE")) std::fprintf(stderr, "ray %llu: kd tri %u t %.9g | bvh tri %u t %.9g\n", (unsigned long long)i, k.tri, (double)k.t, best, (double)best_t);
            differ++; if (k.tri != RGK_NO_TRIANGLE && best != RGK_NO_TRIANGLE && std::fabs(k.t - best_t) <= 2 * s.epsilon) near_tie++; }
    }
    out[0] = (double)n; out[1] = (double)nodes; out[2] = (double)boxes; out[3] = (double)tests; out[4] = (double)differ; out[5] = (double)near_tie;
    out[6] = (double)B.w.size(); out[7] = (double)B.w.size() * width * 32.0; out[8] = (double)ambiguous; out[9] = (double)differ_unflagged;
    return 0;
}

// ---- CPU mirror of the product's opt-in wide-BVH traversal (rgk_b200/csrc/bvh_device.cuh) -------------------------
// Walks the PRODUCT-built node / order arrays (rgk_host_scene_get_bvh) with the decision logic of BvhTraverser -- per-ray
// box margin, widened interval bounds, sorted child order, accept interval, 2-epsilon ambiguity rule -- and the oracle's
// TestIntersection.  tests/test_bvh_host.py uses it to check, without a GPU, that every ray the BVH pass would COMMIT has
// exactly the kd-tree's answer and that the deferred fraction is small.  Not a reference path.
namespace {
struct Bvh4Out { bool deferred, found; Hit hit; };
Bvh4Out bvh4_trace(const Scene& s, const float* nodes, const uint32_t* order, const Ray& r, uint32_t ignore, bool any, uint64_t* counters) {
    Bvh4Out out{}; out.hit.tri = RGK_NO_TRIANGLE; out.hit.t = std::numeric_limits<float>::infinity(); out.hit.a = out.hit.b = out.hit.c = 0;
    const float INF = std::numeric_limits<float>::infinity();
    // a zero (or NaN / reciprocal-overflowing) direction component: the kd traversal's plane distance can be 0 * inf = NaN (origin exactly on a split plane),
    // and a NaN interval bound accepts hits anywhere along the line (src/scene_intersect.cpp:272,308-318) -- kd rule only
    const float inv[3] = {1.f / r.d.x, 1.f / r.d.y, 1.f / r.d.z};
    if (!(std::fabs(inv[0]) < INF) || !(std::fabs(inv[1]) < INF) || !(std::fabs(inv[2]) < INF)) { out.deferred = true; return out; }
    float t0 = r.tnear, t1 = r.tfar;
    for (int k = 0; k < 3; k++) {
        float tn = (s.bb[k][0] - r.o[k]) * inv[k], tf = (s.bb[k][1] - r.o[k]) * inv[k];
        if (tn > tf) std::swap(tn, tf);
        t0 = tn > t0 ? tn : t0; t1 = tf < t1 ? tf : t1;
        if (t0 > t1) return out;
    }
    if (r.tfar < t0) return out;
    const float eps = s.epsilon;
    const float lo_t = t0 - eps, hi_t = t1 + eps, firm_lo = t0 + eps, firm_hi = t1 - eps;
    const float m = (((std::fabs(r.o.x) + std::fabs(r.o.y)) + std::fabs(r.o.z)) + std::fabs(t1)) * 7.62939453125e-6f;
    const float op[3] = {r.o.x + m, r.o.y + m, r.o.z + m}, om[3] = {r.o.x - m, r.o.y - m, r.o.z - m};
    const float low_w = (lo_t - std::fabs(lo_t) * 9.5367431640625e-7f) - 1e-30f;
    float best_t = INF, second_t = INF, limit;
    auto set_limit = [&]() { float l = hi_t; if (!any) { const float w = best_t + 2.0f * eps; l = w < l ? w : l; } limit = (l + std::fabs(l) * 9.5367431640625e-7f) + 1e-30f; };
    set_limit();
    bool border = false, stop = false, best_edge = false;
    struct E { uint32_t code; float t; }; E stack[RGK_ORACLE_BVH_STACK]; int sp = 0;
    auto pop = [&]() -> uint32_t { while (sp > 0) { --sp; if (stack[sp].t <= limit) return stack[sp].code; } return 0x7fffffffu; };
    uint32_t cur = 0;
    while (cur != 0x7fffffffu) {
        if (cur < 0x7fffffffu) {
            if (counters) counters[0]++;
            const float* n = nodes + 32 * (size_t)cur;
            E e[4];
            for (int c = 0; c < 4; c++) {
                float tn = low_w, tf = limit;
                float lo3[3], hi3[3];
                for (int k = 0; k < 3; k++) { const float a = (n[8 * k + c] - op[k]) * inv[k], b = (n[8 * k + 4 + c] - om[k]) * inv[k]; lo3[k] = std::fmin(a, b); hi3[k] = std::fmax(a, b); }
                tn = std::fmax(std::fmax(lo3[0], lo3[1]), std::fmax(lo3[2], low_w));
                tf = std::fmin(std::fmin(hi3[0], hi3[1]), std::fmin(hi3[2], limit));
                uint32_t code; std::memcpy(&code, n + 24 + c, 4);
                e[c] = {code, tn <= tf ? tn : INF};
            }
            auto cswap = [&](int a, int b) { if (e[b].t < e[a].t) std::swap(e[a], e[b]); };
            cswap(0, 1); cswap(2, 3); cswap(0, 2); cswap(1, 3); cswap(1, 2);
            if (!(e[0].t < INF)) { cur = pop(); continue; }
            for (int c = 3; c >= 1; c--) if (e[c].t < INF) { if (sp >= RGK_ORACLE_BVH_STACK) { out.deferred = true; return out; } stack[sp++] = e[c]; }
            cur = e[0].code;
            continue;
        }
        const uint32_t first = cur & 0x1fffffffu, cnt = ((cur >> 29) & 3u) + 1u;
        for (uint32_t j = 0; j < cnt && !stop; j++) {
            const uint32_t ti = order[first + j];
            if (ti == ignore) continue;
            if (counters) counters[1]++;
            float t, a, b;
            if (!test_intersection(s, s.tris[ti], r, t, a, b)) continue;
            if (t < lo_t || t > hi_t) continue;
            if (!any && t > best_t + 2.0f * eps) continue;
            bool edge;
            {   // as BvhTraverser::test: boundary width from the ray's margin over the triangle's smallest projected height
                const Tri& tr = s.tris[ti];
                const V3 pn = v3(tr.p[0], tr.p[1], tr.p[2]);
                const float px = std::fabs(pn.x), py = std::fabs(pn.y), pz = std::fabs(pn.z);
                int i1, i2;
                if (px > py && px > pz) { i1 = 1; i2 = 2; } else if (py > pz) { i1 = 0; i2 = 2; } else { i1 = 0; i2 = 1; }
                const V3 v0 = s.pos[tr.va], v1 = s.pos[tr.vb], v2 = s.pos[tr.vc];
                const float q1x = v1[i1] - v0[i1], q1y = v1[i2] - v0[i2], q2x = v2[i1] - v0[i1], q2y = v2[i2] - v0[i2];
                const float den = q2y * q1x - q2x * q1y;
                const float longest = std::fmax(std::fmax(std::fabs(q1x), std::fabs(q1y)), std::fmax(std::fabs(q2x), std::fabs(q2y)));
                float delta = std::fmax(3.0517578125e-5f, 0.125f * m * longest / std::fabs(den));
                if (q1x > -eps && q1x < eps && q1x != 0.0f) delta += 2.0f * (std::fabs(q1x / q2x) * (1.0f + std::fabs(q2y / q1y)));    // sheared accept region (src/primitives.cpp:141-147)
                edge = !(a >= delta) || !(b >= delta) || !((a + b) <= 1.0f - delta);
                if (std::fabs(dot(r.d, pn)) * eps < m * 0.0546875f) {      // grazing: the plane's miss of the extents is more than eps of ray parameter (cheap bound, then |P|_1)
                    const float p1 = (std::fabs(r.o.x + r.d.x * t) + std::fabs(r.o.y + r.d.y * t)) + std::fabs(r.o.z + r.d.z * t);
                    if (std::fabs(dot(r.d, pn)) * eps < p1 * 2.384185791015625e-7f) edge = true;
                }
                // flag 8 of the product's record: the stored plane misses one of the triangle's own vertices by more than eps / 4 along the dominant axis
                const int k = 3 - i1 - i2;
                if (tr.p[k] != 0.0f && std::isfinite(tr.p[0]) && std::isfinite(tr.p[1]) && std::isfinite(tr.p[2]) && std::isfinite(tr.p[3])) {
                    const V3 vs[3] = {v0, v1, v2};
                    for (int c = 0; c < 3; c++) {
                        const double lifted = -((double)tr.p[3] + (double)tr.p[i1] * (double)vs[c][i1] + (double)tr.p[i2] * (double)vs[c][i2]) / (double)tr.p[k];
                        if (!(std::fabs(lifted - (double)vs[c][k]) <= 0.25 * (double)eps)) edge = true;
                    }
                }
            }
            if (any) { if (!edge && t >= firm_lo && t <= firm_hi) stop = true; else border = true; continue; }
            if (t < best_t) { second_t = best_t; best_t = t; best_edge = edge; out.hit.tri = ti; out.hit.t = t; out.hit.a = 1.0f - a - b; out.hit.b = a; out.hit.c = b; set_limit(); }
            else if (t < second_t) second_t = t;
        }
        cur = stop ? 0x7fffffffu : pop();
    }
    if (any) { out.found = stop; out.deferred = !stop && border; return out; }
    out.found = out.hit.tri != RGK_NO_TRIANGLE;
    out.deferred = out.found && (best_edge || second_t <= best_t + 2.0f * eps || best_t < firm_lo || best_t > firm_hi);
    return out;
}
} // namespace

// closest: status[i] = 1 when the BVH pass defers ray i to the kd-tree, else hits[i] is what it commits.
// counters: [0] wide nodes visited, [1] exact tests
int rgko_bvh4_closest(void* h, const float* nodes, const uint32_t* order, const rgk_ray* rays, const uint32_t* ignore, uint64_t n,
                      rgk_hit* hits, uint8_t* status, uint64_t* counters) {
    const Scene& s = *(Scene*)h;
    for (uint64_t i = 0; i < n; i++) {
        Ray r; r.o = v3(rays[i].origin); r.d = v3(rays[i].direction); r.tnear = rays[i].tnear; r.tfar = rays[i].tfar;
        const Bvh4Out o = bvh4_trace(s, nodes, order, r, ignore ? ignore[i] : RGK_NO_TRIANGLE, false, counters);
        status[i] = o.deferred ? 1 : 0;
        hits[i].triangle = o.found ? o.hit.tri : RGK_NO_TRIANGLE; hits[i].t = o.hit.t; hits[i].a = o.hit.a; hits[i].b = o.hit.b; hits[i].c = o.hit.c;
    }
    return 0;
}
// shadow segments a -> b (3 floats each): visible[i] as Scene::Visibility would say, status[i] = 1 when deferred
int rgko_bvh4_shadow(void* h, const float* nodes, const uint32_t* order, const float* a, const float* b, uint64_t n, uint8_t* visible,
                     uint8_t* status, uint64_t* counters) {
    const Scene& s = *(Scene*)h;
    for (uint64_t i = 0; i < n; i++) {
        Ray r; r.o = v3(a + 3 * i);
        const V3 diff = v3(b + 3 * i) - r.o;
        r.d = normalize(diff);
        const float len = length(diff), e = s.epsilon * 20.0f;
        r.tnear = 0.0f + e; r.tfar = len - e;
        const Bvh4Out o = bvh4_trace(s, nodes, order, r, RGK_NO_TRIANGLE, true, counters);
        status[i] = o.deferred ? 1 : 0;
        visible[i] = o.found ? 0 : 1;
    }
    return 0;
}

// ---- unit-level probes (same layouts as the rgkref_* probes)
int rgko_bxdf_sample(void* h, uint32_t material, const float* Vi, const float* uv, const float* sample, uint64_t n, float* out) {
    const Scene& s = *(Scene*)h;
    for (uint64_t i = 0; i < n; i++) {
        V3 dir; RGB w; bool leak;
        bxdf_sample(s, material, v3(Vi + 3 * i), V2{uv[2 * i], uv[2 * i + 1]}, V2{sample[2 * i], sample[2 * i + 1]}, dir, w, leak);
        float* o = out + 7 * i; o[0] = dir.x; o[1] = dir.y; o[2] = dir.z; o[3] = w.r; o[4] = w.g; o[5] = w.b; o[6] = leak ? 1.0f : 0.0f;
    }
    return 0;
}
int rgko_bxdf_value(void* h, uint32_t material, const float* Vi, const float* Vr, const float* uv, uint64_t n, float* out) {
    const Scene& s = *(Scene*)h;
    for (uint64_t i = 0; i < n; i++) {
        RGB c = bxdf_value(s, material, v3(Vi + 3 * i), v3(Vr + 3 * i), V2{uv[2 * i], uv[2 * i + 1]});
        out[3 * i] = c.r; out[3 * i + 1] = c.g; out[3 * i + 2] = c.b;
    }
    return 0;
}
int rgko_texture_fetch(void* h, uint32_t tex, const float* uv, uint64_t n, float* out) {
    const Scene& s = *(Scene*)h;
    for (uint64_t i = 0; i < n; i++) {
        V2 p = V2{uv[2 * i], uv[2 * i + 1]};
        RGB c = tex_fetch(s, tex, p);
        out[5 * i] = c.r; out[5 * i + 1] = c.g; out[5 * i + 2] = c.b;
        tex_slopes(s, tex, p, out[5 * i + 3], out[5 * i + 4]);
    }
    return 0;
}
int rgko_random_light(void* h, const float* in, uint64_t n, float* out) {
    const Scene& s = *(Scene*)h;
    for (uint64_t i = 0; i < n; i++) {
        const float* q = in + 5 * i;
        Light l = random_light(s, V2{q[0], q[1]}, q[2], V2{q[3], q[4]});
        float* o = out + 12 * i;
        o[0] = (float)l.type; o[1] = l.pos.x; o[2] = l.pos.y; o[3] = l.pos.z; o[4] = l.color.r; o[5] = l.color.g; o[6] = l.color.b;
        o[7] = l.intensity; o[8] = l.size; o[9] = l.normal.x; o[10] = l.normal.y; o[11] = l.normal.z;
    }
    return 0;
}
int rgko_sky(void* h, const float* dir, uint64_t n, float* out) {
    const Scene& s = *(Scene*)h;
    for (uint64_t i = 0; i < n; i++) { RGB c = sky_radiance(s, v3(dir + 3 * i)); out[3 * i] = c.r; out[3 * i + 1] = c.g; out[3 * i + 2] = c.b; }
    return 0;
}

} // extern "C"
