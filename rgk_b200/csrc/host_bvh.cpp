// host_bvh.cpp -- the opt-in wide BVH (RGK_WIDE_BVH=1) next to the reference's kd-tree.
//
// Not a reference structure: RGKrt only has the kd-tree (src/scene.cpp:294-429).  The BVH is a *candidate generator*
// for the traversal kernels in bvh_device.cuh -- it finds the globally closest hit with the reference's own
// Triangle::TestIntersection arithmetic in 4-6x fewer dependent steps than the kd-tree, flags the rare rays
// (~2e-4) whose answer could depend on the kd-tree's per-leaf +-epsilon accept rule, and those are re-traced by the
// kd kernels (tests/bvh_study.py).  The kd-tree stays the authority for every result.
//
// Build: binned SAH (16 bins on the centroid of the longest axis) down to leaves of <= 4 triangles, then collapsed to
// 4 children per node by repeatedly opening the inner child with the largest surface.  Deterministic, single-threaded
// (O(n log n); 2 M triangles in ~2 s).
//
// Node layout (32 floats = 128 bytes = one cache line / L2 sector group):
//   [0..3] lo.x of children 0-3   [4..7] hi.x   [8..11] lo.y   [12..15] hi.y   [16..19] lo.z   [20..23] hi.z
//   [24..27] child codes (uint32 bits): inner = node index (< 2^31 - 1); leaf = 1<<31 | (count-1)<<29 | first slot
//   [28..31] zero.  Empty child: lo = hi = +inf (a box at infinity: the slab test gives an empty interval for every
//   direction -- an inverted box would NOT, min/max of the two plane distances un-inverts it), code 0x7fffffff.
// Boxes are the exact fp32 bounds of the triangles (no padding: the traversal pads per ray).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <stdexcept>
#include "rgk_internal.h"

namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; k++) { lo[k] = std::numeric_limits<float>::infinity(); hi[k] = -std::numeric_limits<float>::infinity(); } }
    void grow(const Box& b) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
    float half_area() const { const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2]; return x * y + y * z + z * x; }
};
struct Bin2 { Box box; int left, right, first, count; };      // count > 0: leaf over order[first, first + count)

struct BvhBuilder {
    const std::vector<float>* ev;       // ev[axis][2 i], [2 i + 1] = min, max of triangle i (host_scene.cpp)
    std::vector<uint32_t> order;
    std::vector<Bin2> bin;
    std::vector<float>* nodes;          // wide nodes out
    unsigned deepest = 0;

    Box tri_box(uint32_t t) const {
        Box b;
        for (int k = 0; k < 3; k++) { b.lo[k] = ev[k][2 * (size_t)t]; b.hi[k] = ev[k][2 * (size_t)t + 1]; }
        return b;
    }
    int build(int first, int count) {
        Bin2 n; n.box.reset(); n.left = n.right = -1; n.first = first; n.count = count;
        Box cb; cb.reset();
        for (int i = first; i < first + count; i++) {
            const Box b = tri_box(order[i]);
            n.box.grow(b);
            for (int k = 0; k < 3; k++) { const float c = 0.5f * b.lo[k] + 0.5f * b.hi[k]; cb.lo[k] = std::min(cb.lo[k], c); cb.hi[k] = std::max(cb.hi[k], c); }
        }
        const int me = (int)bin.size();
        bin.push_back(n);
        if (count <= 4) return me;
        int axis = 0;
        for (int k = 1; k < 3; k++) if (cb.hi[k] - cb.lo[k] > cb.hi[axis] - cb.lo[axis]) axis = k;
        int nl = 0;
        if (cb.hi[axis] > cb.lo[axis]) {
            constexpr int NB = 16;
            int cnt[NB] = {0};
            Box bb[NB];
            for (auto& b : bb) b.reset();
            const float base = cb.lo[axis], scale = (float)NB / (cb.hi[axis] - cb.lo[axis]);
            auto bin_of = [&](uint32_t t) {
                const float c = 0.5f * ev[axis][2 * (size_t)t] + 0.5f * ev[axis][2 * (size_t)t + 1];
                const int q = (int)((c - base) * scale);
                return q < 0 ? 0 : (q > NB - 1 ? NB - 1 : q);
            };
            for (int i = first; i < first + count; i++) { const int q = bin_of(order[i]); cnt[q]++; bb[q].grow(tri_box(order[i])); }
            // sweep: suffix boxes, then prefix
            Box suf[NB]; int sufc[NB];
            { Box acc; acc.reset(); int c = 0; for (int q = NB - 1; q >= 0; q--) { if (cnt[q]) acc.grow(bb[q]); c += cnt[q]; suf[q] = acc; sufc[q] = c; } }
            Box acc; acc.reset(); int c0 = 0, split = -1; float best = std::numeric_limits<float>::infinity();
            for (int sp = 1; sp < NB; sp++) {
                if (cnt[sp - 1]) acc.grow(bb[sp - 1]);
                c0 += cnt[sp - 1];
                if (c0 == 0 || sufc[sp] == 0) continue;
                const float cost = (float)c0 * acc.half_area() + (float)sufc[sp] * suf[sp].half_area();
                if (cost < best) { best = cost; split = sp; }
            }
            if (split > 0) {
                auto mid = std::stable_partition(order.begin() + first, order.begin() + first + count, [&](uint32_t t) { return bin_of(t) < split; });
                nl = (int)(mid - (order.begin() + first));
            }
        }
        if (nl <= 0 || nl >= count) nl = count / 2;          // identical centroids: halve by position (leaves must reach <= 4)
        const int l = build(first, nl), r = build(first + nl, count - nl);
        bin[me].left = l; bin[me].right = r; bin[me].count = 0;
        return me;
    }
    // wide node for the binary subtree `b`; returns its index
    uint32_t collapse(int b, unsigned depth) {
        if (depth > deepest) deepest = depth;
        int kids[4]; int nk = 1; kids[0] = b;
        if (bin[b].count == 0) { kids[0] = bin[b].left; kids[1] = bin[b].right; nk = 2; }
        while (nk < 4) {
            int pick = -1; float area = -1.0f;
            for (int i = 0; i < nk; i++) if (bin[kids[i]].count == 0) { const float a = bin[kids[i]].box.half_area(); if (a > area) { area = a; pick = i; } }
            if (pick < 0) break;
            const int open = kids[pick];
            kids[pick] = bin[open].left; kids[nk++] = bin[open].right;
        }
        const size_t me = nodes->size() / 32;
        nodes->resize(nodes->size() + 32, 0.0f);
        uint32_t code[4];
        for (int i = 0; i < 4; i++) {
            float* f = nodes->data() + 32 * me;
            if (i < nk) {
                const Bin2& k = bin[kids[i]];
                f[0 + i] = k.box.lo[0]; f[4 + i] = k.box.hi[0]; f[8 + i] = k.box.lo[1]; f[12 + i] = k.box.hi[1]; f[16 + i] = k.box.lo[2]; f[20 + i] = k.box.hi[2];
                code[i] = k.count > 0 ? (0x80000000u | ((uint32_t)(k.count - 1) << 29) | (uint32_t)k.first) : 0u;   // inner: patched below
            } else {
                const float inf = std::numeric_limits<float>::infinity();
                f[0 + i] = f[8 + i] = f[16 + i] = inf; f[4 + i] = f[12 + i] = f[20 + i] = inf;    // a box at +infinity
                code[i] = 0x7fffffffu;
            }
        }
        for (int i = 0; i < nk; i++) if (bin[kids[i]].count == 0) code[i] = collapse(kids[i], depth + 1);
        std::memcpy(nodes->data() + 32 * me + 24, code, 16);       // after the recursion: the vector may have moved
        return (uint32_t)me;
    }
};

} // namespace

void host_bvh_build(const std::vector<float> ev[3], uint32_t nt, HostScene& hs) {
    hs.bvh_nodes.clear(); hs.bvh_order.clear(); hs.bvh_depth = 0;
    if (nt == 0) return;
    if (nt >= (1u << 29)) throw std::runtime_error("wide BVH: more than 2^29 triangles");
    BvhBuilder b;
    b.ev = ev; b.nodes = &hs.bvh_nodes;
    b.order.resize(nt);
    for (uint32_t i = 0; i < nt; i++) b.order[i] = i;
    b.bin.reserve((size_t)nt);
    const int root = b.build(0, (int)nt);
    b.collapse(root, 1);
    hs.bvh_order.swap(b.order);
    hs.bvh_depth = b.deepest;
    // a node pushes at most 3 entries and continues into the 4th child
    if (3 * (size_t)hs.bvh_depth + 1 > RGK_STACK_CAP) { hs.bvh_nodes.clear(); hs.bvh_order.clear(); hs.bvh_depth = 0; }
}
