// device_on_host.cpp -- the product's traversal headers compiled for the host (device_shim.h) behind a small C interface:
// the kd-tree traversal in both control structures (trace_persistent / trace_phased) and the wide-BVH candidate pass with
// its deferral rules, driven by the SAME persistent-warp drivers the kernels use, over arrays produced by the product's own
// host scene commit (rgk_host_scene_*).  tests/test_device_on_host.py compares the results with the oracle.
//
// build (__graft_entry__.build): python tests/host_cpp/gen_host_sources.py rgk_b200/csrc build/host/gen, then
//        g++ -std=c++17 -O2 -ffp-contract=off -fPIC -shared -I/usr/local/cuda/include -Ibuild/host/gen -Iinclude
//        tests/host_cpp/device_on_host.cpp -o build/host/libdevice_on_host.so      (-ffp-contract=off = nvcc -fmad=false)
// -DRGK_DOH_MT: the same harness over device_shim_mt.h (a CUDA thread is a host thread, warps have 32 lanes, barriers are pthread
// barriers): libdevice_on_host_mt.so, whose doh_render_round runs whole rounds of the wavefront with real warps -- k_bin's
// shared-memory counting sort, the ballots of the queue compaction, the refill logic of the persistent traversal at its shipped
// thresholds (tests/test_wavefront_lanes_on_host.py).  Only the scene / round entry points are exported from that build.
#ifdef RGK_DOH_MT
#include "device_shim_mt.h"
#else
#include "device_shim.h"
#endif
#define DOH_API __attribute__((visibility("default")))
#include <algorithm>
#include <cstdio>
#include <string>
#include <vector>
#include "bvh_device.cuh"          // includes trace_device.cuh and rgk_internal.h
#include "probe_device.cuh"        // includes shade_device.cuh
#include "render_host.inc"         // render.cu with its launches rewritten (gen_host_sources.py): kernels + host loop

// the two helpers of api.cu that render.cu calls, for the host build
rgk_status rgk_fail(rgk_context* ctx, rgk_status s, const std::string& msg) { if (ctx) ctx->last_error = msg; return s; }
void* rgk_scratch(rgk_context* ctx, int slot, size_t bytes) {
    if (ctx->scratch_size[slot] >= bytes && ctx->scratch[slot]) return ctx->scratch[slot];
    std::free(ctx->scratch[slot]);
    ctx->scratch[slot] = std::calloc(std::max<size_t>(bytes, 256), 1); ctx->scratch_size[slot] = std::max<size_t>(bytes, 256);
    return ctx->scratch[slot];
}

namespace {
// the shading side of DevScene, prepared from the scene description the way rgk_scene_commit prepares its uploads
// (rgk_b200/csrc/api.cu); areal lights and scene constants come from the product's own host_scene_commit
struct ShadeScene {
    HostScene hs;
    std::vector<float4> positions, normals, tangents, texels, ltc_M[2];
    std::vector<float2> texcoords;
    std::vector<uint4> tri_shade;
    std::vector<DevMaterial> materials; std::vector<DevTexture> textures; std::vector<DevPointLight> point_lights;
    std::vector<float> ltc_amp[2];
    std::vector<uint2> nodes; std::vector<float4> ref_planes, ref_bounds, tri_isect, bvh_nodes, bvh_planes;
    BvhStats bvh_stats[2] = {};
    DevScene S{};
};
std::vector<float4> pad3(const float* src, uint32_t n) {
    std::vector<float4> v(n + 1);
    for (uint32_t i = 0; i < n; i++) v[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.0f);
    return v;
}
struct Scene {
    std::vector<uint2> nodes; std::vector<uint32_t> refs; std::vector<float4> ref_planes, ref_bounds, tri_isect, bvh_nodes, bvh_planes;
    std::vector<uint32_t> bvh_refs;
    DevScene S{};
};
}

extern "C" {

// nodes: 2 words / kd node; planes 4, records 12, bounds 4 floats / triangle; bvh_nodes 32 floats / wide node (may be empty)
void* doh_scene_create(const uint32_t* nodes, uint32_t n_nodes, const uint32_t* refs, uint32_t n_refs, const float* planes, const float* records,
                       const float* bounds, uint32_t n_tris, const float* bvh_nodes, uint32_t n_bvh_nodes, const uint32_t* bvh_order, float epsilon,
                       const float* bbox) {
    Scene* s = new Scene();
    s->nodes.resize(n_nodes);
    for (uint32_t i = 0; i < n_nodes; i++) s->nodes[i] = make_uint2(nodes[2 * i], nodes[2 * i + 1]);
    s->refs.assign(refs, refs + n_refs);
    s->ref_planes.resize(n_refs + 1); s->ref_bounds.resize(n_refs + 1);
    for (uint32_t j = 0; j < n_refs; j++) {
        const float* p = planes + 4 * (size_t)refs[j]; const float* b = bounds + 4 * (size_t)refs[j];
        s->ref_planes[j] = make_float4(p[0], p[1], p[2], p[3]); s->ref_bounds[j] = make_float4(b[0], b[1], b[2], b[3]);
    }
    s->tri_isect.resize(3 * (size_t)n_tris);
    std::memcpy(s->tri_isect.data(), records, 48 * (size_t)n_tris);
    DevScene& S = s->S;
    S.nodes = s->nodes.data(); S.refs = s->refs.data(); S.ref_planes = s->ref_planes.data(); S.ref_bounds = s->ref_bounds.data();
    S.tri_isect = s->tri_isect.data();
    if (n_bvh_nodes) {
        s->bvh_nodes.resize(8 * (size_t)n_bvh_nodes);
        std::memcpy(s->bvh_nodes.data(), bvh_nodes, 128 * (size_t)n_bvh_nodes);
        s->bvh_refs.assign(bvh_order, bvh_order + n_tris);
        s->bvh_planes.resize(n_tris);
        for (uint32_t j = 0; j < n_tris; j++) { const float* p = planes + 4 * (size_t)bvh_order[j]; s->bvh_planes[j] = make_float4(p[0], p[1], p[2], p[3]); }
        S.bvh_nodes = s->bvh_nodes.data(); S.bvh_refs = s->bvh_refs.data(); S.bvh_planes = s->bvh_planes.data();
    }
    S.n_nodes = n_nodes; S.n_refs = n_refs; S.n_triangles = n_tris; S.epsilon = epsilon;
    for (int k = 0; k < 6; k++) S.bb[k] = bbox[k];
    S.refill_threshold = 1;             // a one-lane warp refills after every ray
    return s;
}
void doh_scene_destroy(void* h) { delete (Scene*)h; }

DOH_API void* doh_shade_scene_create(const rgk_scene_desc* d, const rgk_device_cfg* cfg) {
    ShadeScene* s = new ShadeScene();
    try { host_scene_commit(d, nullptr, cfg ? *cfg : default_device_cfg(), s->hs); } catch (...) { delete s; return nullptr; }
    DevScene& D = s->S;
    s->positions = pad3(d->positions, d->n_vertices); s->normals = pad3(d->normals, d->n_vertices); s->tangents = pad3(d->tangents, d->n_vertices);
    s->texcoords.resize(d->n_vertices + 1);
    for (uint32_t i = 0; i < d->n_vertices; i++) s->texcoords[i] = make_float2(d->texcoords[2 * i], d->texcoords[2 * i + 1]);
    s->tri_shade.resize(s->hs.tri_shade.size() / 4 + 1);
    std::memcpy(s->tri_shade.data(), s->hs.tri_shade.data(), s->hs.tri_shade.size() * 4);
    s->materials.resize(d->n_materials + 1);
    std::memcpy(s->materials.data(), d->materials, sizeof(DevMaterial) * d->n_materials);
    s->textures.resize(d->n_textures + 1);
    for (uint32_t i = 0; i < d->n_textures; i++) {
        const rgk_texture& t = d->textures[i];
        DevTexture& o = s->textures[i];
        o.kind = t.kind; o.width = t.width; o.height = t.height; o._pad = 0;
        o.color[0] = t.color[0]; o.color[1] = t.color[1]; o.color[2] = t.color[2];
        o.offset = (uint32_t)s->texels.size();
        if (t.kind == 1) for (size_t k = 0; k < (size_t)t.width * t.height; k++) s->texels.push_back(make_float4(t.texels[3 * k], t.texels[3 * k + 1], t.texels[3 * k + 2], 0.0f));
    }
    s->texels.push_back(make_float4(0, 0, 0, 0));
    s->point_lights.resize(d->n_point_lights + 1);
    for (uint32_t i = 0; i < d->n_point_lights; i++) {
        const rgk_point_light& q = d->point_lights[i];
        for (int k = 0; k < 3; k++) { s->point_lights[i].pos[k] = q.position[k]; s->point_lights[i].color[k] = q.color[k]; }
        s->point_lights[i].intensity = q.intensity; s->point_lights[i].size = q.size;
    }
    const rgk_ltc_table* lt[2] = {&d->ltc_ggx, &d->ltc_beckmann};
    D.has_ltc = 1;
    for (int k = 0; k < 2; k++) {
        if (lt[k]->M && lt[k]->amplitude) {
            s->ltc_M[k].resize(4096 * 3); s->ltc_amp[k].assign(lt[k]->amplitude, lt[k]->amplitude + 4096);
            for (int i = 0; i < 4096; i++) {
                const float* m = lt[k]->M + 9 * i;
                s->ltc_M[k][3 * i] = make_float4(m[0], m[1], m[2], m[3]); s->ltc_M[k][3 * i + 1] = make_float4(m[4], m[5], m[6], m[7]);
                s->ltc_M[k][3 * i + 2] = make_float4(m[8], 0.0f, 0.0f, 0.0f);
            }
        } else { D.has_ltc = 0; s->ltc_M[k].resize(1); s->ltc_amp[k].resize(1); }
        D.ltc_M[k] = s->ltc_M[k].data(); D.ltc_amp[k] = s->ltc_amp[k].data();
    }
    if (s->hs.areal_lights.empty()) s->hs.areal_lights.resize(1);
    if (s->hs.areal_tris.empty()) s->hs.areal_tris.resize(1);
    D.positions = s->positions.data(); D.normals = s->normals.data(); D.tangents = s->tangents.data(); D.texcoords = s->texcoords.data();
    D.tri_shade = s->tri_shade.data(); D.materials = s->materials.data(); D.textures = s->textures.data(); D.texels = s->texels.data();
    D.point_lights = s->point_lights.data(); D.areal_lights = s->hs.areal_lights.data(); D.areal_tris = s->hs.areal_tris.data();
    D.sky_mode = d->sky.mode; D.sky_intensity = d->sky.intensity; D.sky_rotate = d->sky.rotate; D.sky_envmap = d->sky.envmap;
    for (int k = 0; k < 3; k++) D.sky_color[k] = d->sky.color[k];
    // traversal arrays, as rgk_scene_commit lays them out
    const HostScene& hs = s->hs;
    s->nodes.resize(hs.nodes.size() / 2);
    for (size_t i = 0; i < s->nodes.size(); i++) s->nodes[i] = make_uint2(hs.nodes[2 * i], hs.nodes[2 * i + 1]);
    s->ref_planes.resize(hs.refs.size() + 1); s->ref_bounds.resize(hs.refs.size() + 1);
    for (size_t j = 0; j < hs.refs.size(); j++) {
        const float* q = &hs.planes[4 * (size_t)hs.refs[j]]; const float* b = &hs.tri_bounds[4 * (size_t)hs.refs[j]];
        s->ref_planes[j] = make_float4(q[0], q[1], q[2], q[3]); s->ref_bounds[j] = make_float4(b[0], b[1], b[2], b[3]);
    }
    s->tri_isect.resize(hs.tri_isect.size() / 4 + 1);
    std::memcpy(s->tri_isect.data(), hs.tri_isect.data(), hs.tri_isect.size() * 4);
    D.nodes = s->nodes.data(); D.refs = hs.refs.data(); D.ref_planes = s->ref_planes.data(); D.ref_bounds = s->ref_bounds.data(); D.tri_isect = s->tri_isect.data();
    if (!hs.bvh_nodes.empty()) {
        s->bvh_nodes.resize(hs.bvh_nodes.size() / 4);
        std::memcpy(s->bvh_nodes.data(), hs.bvh_nodes.data(), hs.bvh_nodes.size() * 4);
        s->bvh_planes.resize(hs.bvh_order.size());
        for (size_t j = 0; j < hs.bvh_order.size(); j++) { const float* q = &hs.planes[4 * (size_t)hs.bvh_order[j]]; s->bvh_planes[j] = make_float4(q[0], q[1], q[2], q[3]); }
        D.bvh_nodes = s->bvh_nodes.data(); D.bvh_refs = hs.bvh_order.data(); D.bvh_planes = s->bvh_planes.data();
    }
    for (int k = 0; k < 6; k++) D.bb[k] = hs.info.bbox[k];
    D.n_nodes = hs.info.n_nodes; D.n_refs = hs.info.n_refs; D.refill_threshold = 1;
    const rgk_scene_info& in = s->hs.info;
    D.n_triangles = in.n_triangles; D.n_vertices = d->n_vertices; D.n_materials = d->n_materials; D.n_textures = d->n_textures;
    D.n_point_lights = d->n_point_lights; D.n_areal_lights = in.n_areal_lights;
    D.total_point_power = in.total_point_power; D.total_areal_power = in.total_areal_power; D.epsilon = in.epsilon;
    return s;
}
DOH_API void doh_shade_scene_destroy(void* h) { delete (ShadeScene*)h; }

// One RenderDriver round through render_round_impl -- the product's own host loop and kernels -- with caller-supplied
// sampler tables (RGK_SAMPLER_TABLES; t1[pixel][dim][set], t2[pixel][dim][set][2]).  out_bvh: rays through the wide-BVH
// kernels and how many of them were deferred to the kd arbiter (both 0 when the scene was committed with RGK_TRAVERSAL_KD).
// cfg: the rgk_device_cfg of the call (the tests switch k_bin off and refill one-lane warps after every ray).
DOH_API int doh_render_round(void* h, const rgk_device_cfg* cfg, const rgk_camera* cam, const rgk_render_params* p, const rgk_task* tasks, uint32_t n_tasks, uint32_t seedstart,
                     uint32_t seedcount_base, const float* t1, const float* t2, uint32_t n1d, uint32_t n2d, uint64_t n_pixels, float* rgb,
                     uint32_t* count, rgk_round_stats* stats, uint64_t* out_bvh) {
    ShadeScene* s = (ShadeScene*)h;
    rgk_context ctx;
    if (cfg) ctx.cfg = *cfg;
    ctx.dev = s->S; ctx.has_scene = true;
    ctx.dev.refill_threshold = 1;
    if (s->S.n_point_lights) ctx.first_point_light = s->point_lights[0];
    s->bvh_stats[0] = s->bvh_stats[1] = BvhStats{0, 0, 0, 0, 0};
    ctx.d_bvh_stats = s->S.bvh_nodes ? s->bvh_stats : nullptr;
    if (t1 && t2) {
        ctx.d_user_t1 = const_cast<float*>(t1); ctx.d_user_t2 = const_cast<float*>(t2);
        ctx.user_n1d = n1d; ctx.user_n2d = n2d; ctx.user_ss = host_sampler_set_size(p->multisample); ctx.user_npix = n_pixels;
    }
    rgk_round_stats local{};
    const rgk_status st = render_round_impl(&ctx, cam, p, tasks, n_tasks, seedstart, seedcount_base, rgb, count, stats ? stats : &local);
    if (st != RGK_OK) std::fprintf(stderr, "doh_render_round: %s\n", ctx.last_error.c_str());
    if (const char* px = std::getenv("DOH_DUMP_PIXEL")) {      // "x,y": the per-sample radiance sums of that pixel (last chunk of the call), for diffing two runs
        unsigned x = 0, y = 0;
        if (std::sscanf(px, "%u,%u", &x, &y) == 2 && ctx.paths) {
            const PathBuffers& B = *ctx.paths;
            const uint32_t want = x | (y << 16);
            const size_t npix = B.cap_pixels;
            for (size_t pos = 0; pos < npix; pos++)
                if (B.pix_xy[pos] == want) {
                    // (the chunk's pixel count is the stride of the sample index; single-tile calls: the tile's pixel count)
                    const size_t stride = (size_t)(tasks[0].x2 - tasks[0].x1) * (tasks[0].y2 - tasks[0].y1);
                    for (uint32_t smp = 0; smp < p->multisample; smp++) {
                        const float4 t = B.tot[(size_t)smp * stride + pos];
                        std::fprintf(stderr, "DUMP pos %zu sample %u slot %zu tot %.9g %.9g %.9g\n", pos, smp, (size_t)smp * stride + pos, t.x, t.y, t.z);
                    }
                    break;
                }
        }
    }
    if (out_bvh) { out_bvh[0] = s->bvh_stats[0].rays + s->bvh_stats[1].rays; out_bvh[1] = s->bvh_stats[0].ambiguous + s->bvh_stats[1].ambiguous; }
    free_path_buffers(&ctx);
    for (auto& q : ctx.scratch) if (q) std::free(q);
    return (int)st;
}
// rgk_probe on the host: the same probe_one the k_probe kernel runs per row
int doh_probe(void* h, uint32_t kind, uint32_t index, const float* in, uint64_t n, float* out) {
    const DevScene& S = ((ShadeScene*)h)->S;
    if (kind > RGK_PROBE_FRAME) return 1;
    for (uint64_t i = 0; i < n; i++) probe_one(S, kind, index, in, i, out);
    return 0;
}

static void store_hit(rgk_hit& out, bool found, const HitRec& h) {
    out.triangle = found ? h.tri : RGK_NO_TRIANGLE; out.t = found ? h.t : __int_as_float(0x7f800000);
    if (found) { out.a = 1.0f - h.alpha - h.beta; out.b = h.alpha; out.c = h.beta; }
    else { out.a = 0.0f; out.b = 0.0f; out.c = 0.0f; }
}

// variant 2 / 6: the kd traversal (status untouched); variant 4: the wide-BVH pass, status[i] = 1 for a deferred ray
int doh_closest(void* h, int variant, const rgk_ray* rays, const uint32_t* ignore, uint64_t n, rgk_hit* hits, uint8_t* status, uint64_t* counters) {
    const DevScene& S = ((Scene*)h)->S;
    unsigned long long work = 0; uint32_t done = 0;
    if (variant == 4) {
        if (!S.bvh_nodes) return 1;
        BvhCount cnt{0, 0, 0}; uint32_t deferred = 0;
        trace_bvh<false, true>(S, (uint32_t)n, &work, cnt, done, deferred,
            [&](uint32_t i, BvhTraverser<false, true>& T) {
                const rgk_ray& r = rays[i];
                return T.init(S, r.origin[0], r.origin[1], r.origin[2], r.direction[0], r.direction[1], r.direction[2], r.tnear, r.tfar, ignore ? ignore[i] : RGK_NO_TRIANGLE);
            },
            [&](uint32_t i, bool found, const HitRec& hr) { store_hit(hits[i], found, hr); status[i] = 0; },
            [&](uint32_t i) { status[i] = 1; });
        if (counters) { counters[0] = cnt.nodes; counters[1] = cnt.tests; }
        return done == n ? 0 : 2;
    }
    TravCount cnt{0, 0, 0, 0, 0, 0, 0};
    auto fetch = [&](uint32_t i, Traverser<false, true>& T) {
        const rgk_ray& r = rays[i];
        return T.init(S, r.origin[0], r.origin[1], r.origin[2], r.direction[0], r.direction[1], r.direction[2], r.tnear, r.tfar, ignore ? ignore[i] : RGK_NO_TRIANGLE);
    };
    auto commit = [&](uint32_t i, bool found, const HitRec& hr) { store_hit(hits[i], found, hr); };
    if (variant == 6) trace_phased<false, true>(S, (uint32_t)n, &work, cnt, done, fetch, commit);
    else trace_persistent<false, true>(S, (uint32_t)n, &work, cnt, done, fetch, commit);
    if (counters) { counters[0] = cnt.inner; counters[1] = cnt.leaf; counters[2] = cnt.refs; counters[3] = cnt.tests; counters[4] = cnt.exact; counters[5] = cnt.prefiltered; counters[6] = cnt.wrong; }
    return done == n ? 0 : 2;
}

int doh_shadow(void* h, int variant, const float* pa, const float* pb, uint64_t n, uint8_t* visible, uint8_t* status) {
    const DevScene& S = ((Scene*)h)->S;
    unsigned long long work = 0; uint32_t done = 0;
    auto init = [&](uint32_t i, auto& T) {          // Ray(from, to, eps) + Scene::Visibility, as k_trace_shadow does
        const float ax = pa[3 * (size_t)i], ay = pa[3 * (size_t)i + 1], az = pa[3 * (size_t)i + 2];
        const float ex = pb[3 * (size_t)i] - ax, ey = pb[3 * (size_t)i + 1] - ay, ez = pb[3 * (size_t)i + 2] - az;
        const float d2 = ex * ex + ey * ey + ez * ez;
        const float inv = 1.0f / sqrtf(d2), len = sqrtf(d2);
        const float e20 = S.epsilon * 20.0f;
        return T.init(S, ax, ay, az, ex * inv, ey * inv, ez * inv, 0.0f + e20, len - e20, RGK_NO_TRIANGLE);
    };
    if (variant == 4) {
        if (!S.bvh_nodes) return 1;
        BvhCount cnt{0, 0, 0}; uint32_t deferred = 0;
        trace_bvh<true, false>(S, (uint32_t)n, &work, cnt, done, deferred,
            [&](uint32_t i, BvhTraverser<true, false>& T) { return init(i, T); },
            [&](uint32_t i, bool found, const HitRec&) { visible[i] = found ? 0 : 1; status[i] = 0; },
            [&](uint32_t i) { status[i] = 1; });
        return done == n ? 0 : 2;
    }
    TravCount cnt{0, 0, 0, 0, 0, 0, 0};
    auto fetch = [&](uint32_t i, Traverser<true, false>& T) { return init(i, T); };
    auto commit = [&](uint32_t i, bool found, const HitRec&) { visible[i] = found ? 0 : 1; };
    if (variant == 6) trace_phased<true, false>(S, (uint32_t)n, &work, cnt, done, fetch, commit);
    else trace_persistent<true, false>(S, (uint32_t)n, &work, cnt, done, fetch, commit);
    return done == n ? 0 : 2;
}

}
