// host_bvh.cpp -- the wide BVH (RGK_TRAVERSAL_BVH, the default) next to the reference's kd-tree.
//
// Not a reference structure: RGKrt only has the kd-tree (src/scene.cpp:294-429).  The BVH is a *candidate generator*
// for the traversal kernels in bvh_device.cuh -- it finds the globally closest hit with the reference's own
// Triangle::TestIntersection arithmetic in 4-6x fewer dependent steps than the kd-tree, flags the rare rays
// (~2e-4) whose answer could depend on the kd-tree's per-leaf +-epsilon accept rule, and those are re-traced by the
// kd kernels (tests/bvh_study.py).  The kd-tree stays the authority for every result.
//
// Build: binary binned-SAH tree (32 centroid bins on each of the three axes) down to single triangles, then the SAH-optimal
// collapse of Ylitie, Karras and Laine (2017, section 3.1) into 4-wide nodes whose leaves hold <= 4 triangles (node visit
// cost 1, triangle cost 0.3).  Against the first version (16 bins on the longest axis, leaves of <= 4 fixed by the binary
// build, greedy largest-surface collapse) the CPU mirror counts 14 % fewer node visits and 27 % fewer exact tests per ray
// (tests/bvh_quality.py).  rgk_device_cfg::bvh_bins / bvh_all_axes / bvh_leaf_max / bvh_greedy_collapse / bvh_c_prim are study knobs; so is bvh_reinsert_iters = n
// (insertion-based optimisation of the binary tree: 2-3 % fewer node visits on the atrium stand-in, not monotonic in n -- off).  Deterministic,
// single-threaded (2 M triangles in ~5 s).
//
// Node layout (32 floats = 128 bytes = one cache line / L2 sector group):
//   [0..3] lo.x of children 0-3   [4..7] hi.x   [8..11] lo.y   [12..15] hi.y   [16..19] lo.z   [20..23] hi.z
//   [24..27] child codes (uint32 bits): inner = node index (< 2^31 - 1); leaf = 1<<31 | (count-1)<<29 | first slot
//   [28..31] zero.  Empty child: lo = hi = +inf (a box at infinity: the slab test gives an empty interval for every
//   direction -- an inverted box would NOT, min/max of the two plane distances un-inverts it), code 0x7fffffff.
// Boxes are the fp32 bounds host_scene.cpp hands in: the triangles' extents and the corners of the region TestIntersection
// accepts (no padding beyond that: the traversal pads per ray).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include "rgk_internal.h"

namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; k++) { lo[k] = std::numeric_limits<float>::infinity(); hi[k] = -std::numeric_limits<float>::infinity(); } }
    void grow(const Box& b) { for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
    float half_area() const { const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2]; return x * y + y * z + z * x; }
};
struct Bin2 { Box box; int left, right, first, count, total; };      // count > 0: leaf over order[first, first + count); total: triangles below
// optimal collapse (Ylitie, Karras, Laine 2017, 3.1): c[i-1] = least SAH cost of the subtree as a forest of <= i wide-BVH roots
struct Dp { float c[3]; uint8_t k[3]; uint8_t leaf, k4; };

struct BvhBuilder {
    const std::vector<float>* ev;       // ev[axis][2 i], [2 i + 1] = min, max of triangle i (host_scene.cpp)
    std::vector<uint32_t> order;
    std::vector<Bin2> bin;
    std::vector<float>* nodes;          // wide nodes out
    unsigned deepest = 0;

    int nbins = 32; bool all_axes = true; int leaf_max = 4;     // measured with tests/bvh_quality.py: 16 bins on the longest axis cost 14 % more node visits
    std::vector<int> scratch_cnt, scratch_sufc; std::vector<Box> scratch_box, scratch_suf;
    int bin_of(uint32_t t, int axis, float base, float scale) const {
        const float c = 0.5f * ev[axis][2 * (size_t)t] + 0.5f * ev[axis][2 * (size_t)t + 1];
        const int q = (int)((c - base) * scale);
        return q < 0 ? 0 : (q > nbins - 1 ? nbins - 1 : q);
    }
    Box tri_box(uint32_t t) const {
        Box b;
        for (int k = 0; k < 3; k++) { b.lo[k] = ev[k][2 * (size_t)t]; b.hi[k] = ev[k][2 * (size_t)t + 1]; }
        return b;
    }
    int build(int first, int count) {
        Bin2 n; n.box.reset(); n.left = n.right = -1; n.first = first; n.count = count; n.total = count;
        Box cb; cb.reset();
        for (int i = first; i < first + count; i++) {
            const Box b = tri_box(order[i]);
            n.box.grow(b);
            for (int k = 0; k < 3; k++) { const float c = 0.5f * b.lo[k] + 0.5f * b.hi[k]; cb.lo[k] = std::min(cb.lo[k], c); cb.hi[k] = std::max(cb.hi[k], c); }
        }
        const int me = (int)bin.size();
        bin.push_back(n);
        if (count <= leaf_max) return me;
        // binned SAH over every axis with centroid extent (nbins bins each); the split with the least
        // count_left * area_left + count_right * area_right wins
        int nl = 0, best_axis = -1, best_split = -1;
        float best = std::numeric_limits<float>::infinity(), best_base = 0.0f, best_scale = 0.0f;
        const int NB = nbins;
        std::vector<int>& cnt = scratch_cnt; std::vector<Box>& bb = scratch_box; std::vector<Box>& suf = scratch_suf; std::vector<int>& sufc = scratch_sufc;
        int longest = 0;
        for (int k = 1; k < 3; k++) if (cb.hi[k] - cb.lo[k] > cb.hi[longest] - cb.lo[longest]) longest = k;
        for (int axis = 0; axis < 3; axis++) {
            if (!all_axes && axis != longest) continue;
            if (!(cb.hi[axis] > cb.lo[axis])) continue;
            const float base = cb.lo[axis], scale = (float)NB / (cb.hi[axis] - cb.lo[axis]);
            for (int q = 0; q < NB; q++) { cnt[q] = 0; bb[q].reset(); }
            for (int i = first; i < first + count; i++) { const int q = bin_of(order[i], axis, base, scale); cnt[q]++; bb[q].grow(tri_box(order[i])); }
            { Box acc; acc.reset(); int c = 0; for (int q = NB - 1; q >= 0; q--) { if (cnt[q]) acc.grow(bb[q]); c += cnt[q]; suf[q] = acc; sufc[q] = c; } }
            Box acc; acc.reset(); int c0 = 0;
            for (int sp = 1; sp < NB; sp++) {
                if (cnt[sp - 1]) acc.grow(bb[sp - 1]);
                c0 += cnt[sp - 1];
                if (c0 == 0 || sufc[sp] == 0) continue;
                const float cost = (float)c0 * acc.half_area() + (float)sufc[sp] * suf[sp].half_area();
                if (cost < best) { best = cost; best_axis = axis; best_split = sp; best_base = base; best_scale = scale; }
            }
        }
        if (best_axis >= 0) {
            auto mid = std::stable_partition(order.begin() + first, order.begin() + first + count,
                                             [&](uint32_t t) { return bin_of(t, best_axis, best_base, best_scale) < best_split; });
            nl = (int)(mid - (order.begin() + first));
        }
        if (nl <= 0 || nl >= count) nl = count / 2;          // identical centroids: halve by position (leaves must reach <= 4)
        const int l = build(first, nl), r = build(first + nl, count - nl);
        bin[me].left = l; bin[me].right = r; bin[me].count = 0;      // total stays
        return me;
    }
    // ---- insertion-based optimisation of the binary tree (after Bittner, Hapala, Havran 2013): take an inner node out,
    // re-insert its two children where they add the least surface (branch and bound over induced + direct cost).  Needs
    // single-triangle leaves; the triangle order is re-linearised afterwards.  Returns the (possibly new) root.
    int find_best(int root, int X, std::vector<std::pair<float, int>>& heap) const {
        const Box bx = bin[X].box; const float ax = bx.half_area();
        float best = std::numeric_limits<float>::infinity(); int best_y = root;
        heap.clear(); heap.push_back({0.0f, root});
        auto cmp = [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first > b.first; };
        while (!heap.empty()) {
            std::pop_heap(heap.begin(), heap.end(), cmp);
            const float ci = heap.back().first; const int Y = heap.back().second; heap.pop_back();
            if (ci + ax >= best) break;
            Box u = bin[Y].box; u.grow(bx);
            const float tot = ci + u.half_area();
            if (tot < best) { best = tot; best_y = Y; }
            const float cc = tot - bin[Y].box.half_area();
            if (bin[Y].count == 0 && cc + ax < best) {
                heap.push_back({cc, bin[Y].left}); std::push_heap(heap.begin(), heap.end(), cmp);
                heap.push_back({cc, bin[Y].right}); std::push_heap(heap.begin(), heap.end(), cmp);
            }
        }
        return best_y;
    }
    int reinsert(int root, int iterations, float fraction) {
        const int n = (int)bin.size();
        std::vector<int> parent(n, -1);
        for (int i = 0; i < n; i++) if (bin[i].count == 0) { parent[bin[i].left] = i; parent[bin[i].right] = i; }
        auto refit_up = [&](int i) {
            while (i >= 0) { Box b = bin[bin[i].left].box; b.grow(bin[bin[i].right].box); bin[i].box = b; i = parent[i]; }
        };
        std::vector<int> cand; std::vector<std::pair<float, int>> heap;
        for (int it = 0; it < iterations; it++) {
            cand.clear();
            for (int i = 0; i < n; i++) if (bin[i].count == 0 && i != root && parent[i] != root) cand.push_back(i);
            std::sort(cand.begin(), cand.end(), [&](int a, int b) { const float x = bin[a].box.half_area(), y = bin[b].box.half_area(); return x > y || (x == y && a < b); });
            cand.resize(std::max<size_t>(1, (size_t)(cand.size() * fraction)));
            for (int N : cand) {
                const int P = parent[N];
                if (bin[N].count != 0 || N == root || P < 0 || P == root) continue;       // the tree changes during the pass
                const int G = parent[P], S = bin[P].left == N ? bin[P].right : bin[P].left;
                (bin[G].left == P ? bin[G].left : bin[G].right) = S; parent[S] = G; refit_up(G);
                int L = bin[N].left, R = bin[N].right;
                if (bin[L].box.half_area() < bin[R].box.half_area()) std::swap(L, R);
                const int slot[2] = {N, P}, sub[2] = {L, R};
                for (int q = 0; q < 2; q++) {
                    const int X = sub[q], Y = find_best(root, X, heap), NP = slot[q], YP = parent[Y];
                    bin[NP].left = Y; bin[NP].right = X; bin[NP].count = 0;
                    parent[Y] = NP; parent[X] = NP; parent[NP] = YP;
                    if (YP < 0) root = NP; else (bin[YP].left == Y ? bin[YP].left : bin[YP].right) = NP;
                    refit_up(NP);
                }
            }
        }
        // re-linearise: subtree = contiguous range of `order` again
        std::vector<uint32_t> fresh; fresh.reserve(order.size());
        std::vector<std::pair<int, int>> st; st.push_back({root, 0});
        while (!st.empty()) {
            const int i = st.back().first; const int phase = st.back().second; st.pop_back();
            if (bin[i].count > 0) { const uint32_t t = order[bin[i].first]; bin[i].first = (int)fresh.size(); bin[i].total = 1; fresh.push_back(t); continue; }
            if (phase == 0) { bin[i].first = (int)fresh.size(); st.push_back({i, 1}); st.push_back({bin[i].right, 0}); st.push_back({bin[i].left, 0}); }
            else bin[i].total = bin[bin[i].left].total + bin[bin[i].right].total;
        }
        order.swap(fresh);
        return root;
    }

    // ---- SAH-optimal collapse of the binary tree into 4-wide nodes with leaves of <= 4 triangles
    std::vector<Dp> dp;
    float c_node = 1.0f, c_prim = 0.3f;
    float forest(int n, int i) const { return dp[n].c[i - 1]; }
    float distribute(int l, int r, int j, uint8_t& kbest) const {
        float best = std::numeric_limits<float>::infinity(); kbest = 1;
        for (int k = 1; k < j; k++) {
            if (k > 3 || j - k > 3) continue;
            const float v = forest(l, k) + forest(r, j - k);
            if (v < best) { best = v; kbest = (uint8_t)k; }
        }
        return best;
    }
    void solve(int n) {
        Dp& d = dp[n];
        const float A = bin[n].box.half_area();
        if (bin[n].count > 0) { d.c[0] = d.c[1] = d.c[2] = A * (float)bin[n].count * c_prim; d.k[0] = d.k[1] = d.k[2] = 0; d.leaf = 1; d.k4 = 0; return; }
        const int l = bin[n].left, r = bin[n].right;
        solve(l); solve(r);
        Dp& e = dp[n];
        const float c_leaf = bin[n].total <= 4 ? A * (float)bin[n].total * c_prim : std::numeric_limits<float>::infinity();
        uint8_t k4; const float c_int = distribute(l, r, 4, k4) + A * c_node;
        e.leaf = c_leaf <= c_int ? 1 : 0; e.k4 = k4; e.c[0] = e.leaf ? c_leaf : c_int; e.k[0] = 0;
        for (int i = 2; i <= 3; i++) {
            uint8_t k; const float v = distribute(l, r, i, k);
            if (v < e.c[i - 2]) { e.c[i - 1] = v; e.k[i - 1] = k; } else { e.c[i - 1] = e.c[i - 2]; e.k[i - 1] = 0; }
        }
    }
    void collect(int n, int j, int* kids, int& nk) const {
        if (j == 1 || bin[n].count > 0) { kids[nk++] = n; return; }
        const uint8_t k = dp[n].k[j - 1];
        if (k == 0) { collect(n, j - 1, kids, nk); return; }
        collect(bin[n].left, k, kids, nk); collect(bin[n].right, j - k, kids, nk);
    }
    bool is_leaf_root(int n) const { return bin[n].count > 0 || dp[n].leaf; }
    uint32_t emit(int n, unsigned depth) {
        if (depth > deepest) deepest = depth;
        int kids[4]; int nk = 0;
        if (is_leaf_root(n)) kids[nk++] = n;
        else { collect(bin[n].left, dp[n].k4, kids, nk); collect(bin[n].right, 4 - dp[n].k4, kids, nk); }
        const size_t me = nodes->size() / 32;
        nodes->resize(nodes->size() + 32, 0.0f);
        uint32_t code[4];
        for (int i = 0; i < 4; i++) {
            float* f = nodes->data() + 32 * me;
            if (i < nk) {
                const Bin2& k = bin[kids[i]];
                f[0 + i] = k.box.lo[0]; f[4 + i] = k.box.hi[0]; f[8 + i] = k.box.lo[1]; f[12 + i] = k.box.hi[1]; f[16 + i] = k.box.lo[2]; f[20 + i] = k.box.hi[2];
                code[i] = is_leaf_root(kids[i]) ? (0x80000000u | ((uint32_t)(k.total - 1) << 29) | (uint32_t)k.first) : 0u;
            } else {
                const float inf = std::numeric_limits<float>::infinity();
                f[0 + i] = f[8 + i] = f[16 + i] = inf; f[4 + i] = f[12 + i] = f[20 + i] = inf;    // a box at +infinity
                code[i] = 0x7fffffffu;
            }
        }
        for (int i = 0; i < nk; i++) if (!is_leaf_root(kids[i])) code[i] = emit(kids[i], depth + 1);
        std::memcpy(nodes->data() + 32 * me + 24, code, 16);
        return (uint32_t)me;
    }

    // greedy alternative (rgk_device_cfg::bvh_greedy_collapse): wide node for the binary subtree `b`, opening the largest child first
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

void host_bvh_build(const std::vector<float> ev[3], uint32_t nt, const rgk_device_cfg& cfg, HostScene& hs) {
    hs.bvh_nodes.clear(); hs.bvh_order.clear(); hs.bvh_depth = 0;
    if (nt == 0) return;
    if (nt >= (1u << 29)) throw std::runtime_error("wide BVH: more than 2^29 triangles");
    BvhBuilder b;
    b.ev = ev; b.nodes = &hs.bvh_nodes;
    if (cfg.bvh_bins) b.nbins = (int)std::min(256u, std::max(4u, cfg.bvh_bins));
    b.all_axes = cfg.bvh_all_axes != 0;
    const bool greedy = cfg.bvh_greedy_collapse != 0;
    if (cfg.bvh_leaf_max) b.leaf_max = (int)std::min(4u, cfg.bvh_leaf_max);
    else if (!greedy) b.leaf_max = 1;                      // the collapse chooses the leaves (<= 4 triangles) itself
    if (cfg.bvh_c_prim > 0.0f) b.c_prim = cfg.bvh_c_prim;
    b.scratch_cnt.resize(b.nbins); b.scratch_sufc.resize(b.nbins); b.scratch_box.resize(b.nbins); b.scratch_suf.resize(b.nbins);
    b.order.resize(nt);
    for (uint32_t i = 0; i < nt; i++) b.order[i] = i;
    b.bin.reserve(2 * (size_t)nt);
    int root = b.build(0, (int)nt);
    if (cfg.bvh_reinsert_iters > 0 && b.leaf_max == 1 && nt > 8)
        root = b.reinsert(root, (int)cfg.bvh_reinsert_iters, cfg.bvh_reinsert_frac > 0.0f ? cfg.bvh_reinsert_frac : 0.25f);
    if (greedy) b.collapse(root, 1);
    else { b.dp.resize(b.bin.size()); b.solve(root); b.emit(root, 1); }
    hs.bvh_order.swap(b.order);
    hs.bvh_depth = b.deepest;
    // a node pushes at most 3 entries and continues into the 4th child
    // ... and the traversal addresses a node by a 32-bit byte offset (node << 7): at most 2^25 nodes
    if (3 * (size_t)hs.bvh_depth + 1 > RGK_STACK_CAP || hs.bvh_nodes.size() / 32 >= ((size_t)1 << 25)) { hs.bvh_nodes.clear(); hs.bvh_order.clear(); hs.bvh_depth = 0; }
}
