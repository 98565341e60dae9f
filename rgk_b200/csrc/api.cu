// api.cu -- the C ABI of librgk_b200.so (include/rgk_b200.h): context, scene upload,
// host-buffer wrappers around the kernels.  No CPU fallback anywhere: without a CUDA device
// rgk_context_create fails with RGK_ERR_NO_DEVICE.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include "rgk_internal.h"

rgk_status rgk_fail(rgk_context* ctx, rgk_status s, const std::string& msg) {
    if (ctx) ctx->last_error = msg;
    return s;
}

void* rgk_scratch(rgk_context* ctx, int slot, size_t bytes) {
    if (ctx->scratch_size[slot] >= bytes && ctx->scratch[slot]) return ctx->scratch[slot];
    if (ctx->scratch[slot]) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->scratch[slot]); ctx->scratch[slot] = nullptr; ctx->scratch_size[slot] = 0; }
    size_t want = std::max<size_t>(bytes, 256);
    void* p = nullptr;
    if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    ctx->scratch[slot] = p; ctx->scratch_size[slot] = want;
    return p;
}

namespace {

thread_local std::string g_create_error;

void free_scene(rgk_context* ctx) {
    for (void* p : ctx->scene_allocs) cudaFree(p);
    ctx->scene_allocs.clear();
    ctx->has_scene = false;
    ctx->dev = DevScene{};
    ctx->d_bvh_stats = nullptr;
}

template <class T>
rgk_status upload(rgk_context* ctx, const std::vector<T>& h, const T** d) {
    *d = nullptr;
    const size_t bytes = std::max<size_t>(h.size() * sizeof(T), 64);   // never empty: element 0 is always readable
    void* p = nullptr;
    RGK_CUDA(ctx, cudaMalloc(&p, bytes));
    ctx->scene_allocs.push_back(p);
    if (!h.empty()) RGK_CUDA(ctx, cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *d = (const T*)p;
    return RGK_OK;
}
#define UP(vec, ptr) do { rgk_status s_ = upload(ctx, vec, ptr); if (s_ != RGK_OK) { free_scene(ctx); return s_; } } while (0)

std::vector<float4> pad3(const float* src, uint32_t n) {
    std::vector<float4> v(n);
    for (uint32_t i = 0; i < n; i++) v[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.0f);
    return v;
}

} // namespace

rgk_device_cfg default_device_cfg() {
    rgk_device_cfg c;
    std::memset(&c, 0, sizeof c);
    c.struct_size = (uint32_t)sizeof(rgk_device_cfg);
    c.traversal = RGK_TRAVERSAL_BVH;
    c.bvh_bins = 32; c.bvh_all_axes = 1; c.bvh_reinsert_frac = 0.25f;
    c.chunk_paths = (uint64_t)128 << 20; c.table_bytes = (uint64_t)24 << 30; c.reverse_bytes = (uint64_t)8 << 30;
    c.refill_batch = 16; c.refill_coherent = 0; c.refill_incoherent = 24; c.refill_shadow = 12;
    c.binning = 1; c.bin_shadow_first = 1; c.bin_items = 2048; c.bin_min_frac = 0.25f; c.shade_path_order = 1;
    c.skip_null_shadow = 1; c.const_light = 1; c.arb_grid = 8;
    c.sampler_smem = 1; c.trace_threads = 128; c.kd_variant = 6; c.sampler_ctas_per_sm = 3;
    return c;
}

extern "C" {

uint32_t rgk_abi_version(void) { return RGK_ABI_VERSION; }

void rgk_device_cfg_init(rgk_device_cfg* cfg) { if (cfg) *cfg = default_device_cfg(); }

static const char* check_cfg(const rgk_device_cfg& c) {
    if (c.struct_size != sizeof(rgk_device_cfg)) return "rgk_device_cfg: struct_size mismatch (call rgk_device_cfg_init first)";
    if (c.traversal > RGK_TRAVERSAL_KD) return "rgk_device_cfg: unknown traversal";
    if (c.trace_threads != 64 && c.trace_threads != 128 && c.trace_threads != 256) return "rgk_device_cfg: trace_threads must be 64, 128 or 256";
    if (c.kd_variant != 2 && c.kd_variant != 6) return "rgk_device_cfg: kd_variant must be 2 or 6";
    if (c.refill_batch < 1 || c.refill_batch > 32 || c.refill_coherent > 32 || c.refill_incoherent < 1 || c.refill_incoherent > 32 ||
        c.refill_shadow < 1 || c.refill_shadow > 32) return "rgk_device_cfg: refill thresholds are lane counts (1..32)";
    if (c.bvh_leaf_max > 4) return "rgk_device_cfg: bvh_leaf_max > 4";
    if (c.sampler_ctas_per_sm < 1 || c.sampler_ctas_per_sm > 16) return "rgk_device_cfg: sampler_ctas_per_sm must be 1..16";
    if (c.sampler_kernel > 2 || c.sampler_slots > 32) return "rgk_device_cfg: sampler_kernel is 0..2, sampler_slots at most 32";
    if (c.bin_items < 64 || c.arb_grid < 1) return "rgk_device_cfg: bin_items >= 64, arb_grid >= 1";
    if (c.chunk_paths < 64 || c.table_bytes < ((uint64_t)1 << 20) || c.reverse_bytes < ((uint64_t)1 << 20)) return "rgk_device_cfg: memory sizes too small";
    return nullptr;
}

rgk_status rgk_context_configure(rgk_context* ctx, const rgk_device_cfg* cfg) {
    if (!ctx || !cfg) return RGK_ERR_INVALID;
    if (const char* why = check_cfg(*cfg)) return rgk_fail(ctx, RGK_ERR_INVALID, why);
    ctx->cfg = *cfg;
    ctx->dev.refill_threshold = cfg->refill_batch;
    return RGK_OK;
}
rgk_status rgk_context_get_cfg(const rgk_context* ctx, rgk_device_cfg* out) {
    if (!ctx || !out) return RGK_ERR_INVALID;
    *out = ctx->cfg;
    return RGK_OK;
}

const char* rgk_status_string(rgk_status s) {
    switch (s) {
    case RGK_OK: return "ok";
    case RGK_ERR_INVALID: return "invalid argument";
    case RGK_ERR_CUDA: return "CUDA error";
    case RGK_ERR_NOMEM: return "out of memory";
    case RGK_ERR_NO_DEVICE: return "no CUDA device (rgk_b200 has no CPU fallback)";
    case RGK_ERR_NO_SCENE: return "no scene committed";
    case RGK_ERR_UNSUPPORTED: return "unsupported";
    }
    return "unknown";
}

rgk_status rgk_context_create(int device, void* stream, rgk_context** out) {
    if (!out) return RGK_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        g_create_error = "no CUDA device visible: rgk_b200 has no CPU fallback";
        return RGK_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) { g_create_error = "device ordinal out of range"; return RGK_ERR_INVALID; }
    if (cudaSetDevice(device) != cudaSuccess) { g_create_error = cudaGetErrorString(cudaGetLastError()); return RGK_ERR_CUDA; }
    rgk_context* ctx = new rgk_context();
    ctx->device = device;
    if (stream) { ctx->stream = (cudaStream_t)stream; ctx->own_stream = false; }
    else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            g_create_error = cudaGetErrorString(cudaGetLastError()); delete ctx; return RGK_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    for (auto& ev : ctx->ev) cudaEventCreate(&ev);
    if (cudaMalloc((void**)&ctx->d_stats, 2 * sizeof(rgk_trav_stats)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(cudaGetLastError()); delete ctx; return RGK_ERR_NOMEM;
    }
    *out = ctx;
    return RGK_OK;
}

void rgk_context_destroy(rgk_context* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_path_buffers(ctx);
    free_scene(ctx);
    for (auto& p : ctx->scratch) if (p) cudaFree(p);
    if (ctx->d_stats) cudaFree(ctx->d_stats);
    if (ctx->d_user_t1) cudaFree(ctx->d_user_t1);
    if (ctx->d_user_t2) cudaFree(ctx->d_user_t2);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* rgk_last_error(const rgk_context* ctx) { return ctx ? ctx->last_error.c_str() : g_create_error.c_str(); }

rgk_status rgk_synchronize(rgk_context* ctx) {
    if (!ctx) return RGK_ERR_INVALID;
    RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RGK_OK;
}

rgk_status rgk_scene_commit(rgk_context* ctx, const rgk_scene_desc* d, const rgk_kdtree* tree) {
    if (!ctx || !d) return RGK_ERR_INVALID;
    RGK_CUDA(ctx, cudaSetDevice(ctx->device));
    RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    free_scene(ctx);
    try {
        host_scene_commit(d, tree, ctx->cfg, ctx->host);
    } catch (const std::exception& e) {
        return rgk_fail(ctx, RGK_ERR_INVALID, e.what());
    }
    const HostScene& hs = ctx->host;
    DevScene& D = ctx->dev;
    // geometry
    {
        std::vector<uint2> nodes(hs.nodes.size() / 2);
        for (size_t i = 0; i < nodes.size(); i++) nodes[i] = make_uint2(hs.nodes[2 * i], hs.nodes[2 * i + 1]);
        UP(nodes, &D.nodes);
        UP(hs.refs, &D.refs);
        {
            std::vector<float4> rp(hs.refs.size());
            for (size_t j = 0; j < rp.size(); j++) {
                const float* q = &hs.planes[4 * (size_t)hs.refs[j]];
                rp[j] = make_float4(q[0], q[1], q[2], q[3]);
            }
            UP(rp, &D.ref_planes);
            for (size_t j = 0; j < rp.size(); j++) {
                const float* q = &hs.tri_bounds[4 * (size_t)hs.refs[j]];
                rp[j] = make_float4(q[0], q[1], q[2], q[3]);
            }
            UP(rp, &D.ref_bounds);
        }
        if (!hs.bvh_nodes.empty()) {                 // wide BVH (RGK_TRAVERSAL_BVH)
            std::vector<float4> bn(hs.bvh_nodes.size() / 4);
            std::memcpy(bn.data(), hs.bvh_nodes.data(), hs.bvh_nodes.size() * 4);
            UP(bn, &D.bvh_nodes);
            UP(hs.bvh_order, &D.bvh_refs);
            std::vector<float4> bp(hs.bvh_order.size());
            for (size_t j = 0; j < bp.size(); j++) {
                const float* q = &hs.planes[4 * (size_t)hs.bvh_order[j]];
                bp[j] = make_float4(q[0], q[1], q[2], q[3]);
            }
            UP(bp, &D.bvh_planes);
            std::vector<BvhStats> z(2, BvhStats{0, 0, 0, 0, 0});
            const BvhStats* dz = nullptr;
            UP(z, &dz);
            ctx->d_bvh_stats = const_cast<BvhStats*>(dz);
        }
        std::vector<float4> rec(hs.tri_isect.size() / 4);
        std::memcpy(rec.data(), hs.tri_isect.data(), hs.tri_isect.size() * 4);
        UP(rec, &D.tri_isect);
        std::vector<uint4> sh(hs.tri_shade.size() / 4);
        std::memcpy(sh.data(), hs.tri_shade.data(), hs.tri_shade.size() * 4);
        UP(sh, &D.tri_shade);
        UP(pad3(d->positions, d->n_vertices), &D.positions);
        UP(pad3(d->normals, d->n_vertices), &D.normals);
        UP(pad3(d->tangents, d->n_vertices), &D.tangents);
        std::vector<float2> uv(d->n_vertices);
        for (uint32_t i = 0; i < d->n_vertices; i++) uv[i] = make_float2(d->texcoords[2 * i], d->texcoords[2 * i + 1]);
        UP(uv, &D.texcoords);
    }
    // materials / textures
    {
        static_assert(sizeof(DevMaterial) == sizeof(rgk_material), "material layout");
        std::vector<DevMaterial> mats(d->n_materials);
        std::memcpy(mats.data(), d->materials, sizeof(DevMaterial) * d->n_materials);
        UP(mats, &D.materials);
        std::vector<DevTexture> tex(d->n_textures);
        std::vector<float4> pool;
        for (uint32_t i = 0; i < d->n_textures; i++) {
            const rgk_texture& t = d->textures[i];
            DevTexture& o = tex[i];
            o.kind = t.kind; o.width = t.width; o.height = t.height; o._pad = 0;
            o.color[0] = t.color[0]; o.color[1] = t.color[1]; o.color[2] = t.color[2];
            o.offset = (uint32_t)pool.size();
            if (t.kind == 1) {
                if (!t.texels || t.width == 0 || t.height == 0) { free_scene(ctx); return rgk_fail(ctx, RGK_ERR_INVALID, "image texture without texels"); }
                const size_t n = (size_t)t.width * t.height;
                if (pool.size() + n > 0xFFFFFFFFull) { free_scene(ctx); return rgk_fail(ctx, RGK_ERR_INVALID, "texture pool too large"); }
                pool.reserve(pool.size() + n);
                for (size_t k = 0; k < n; k++) pool.push_back(make_float4(t.texels[3 * k], t.texels[3 * k + 1], t.texels[3 * k + 2], 0.0f));
            }
        }
        UP(tex, &D.textures);
        UP(pool, &D.texels);
    }
    // lights, sky, LTC
    {
        std::vector<DevPointLight> pl(d->n_point_lights);
        for (uint32_t i = 0; i < d->n_point_lights; i++) {
            const rgk_point_light& s = d->point_lights[i];
            for (int k = 0; k < 3; k++) { pl[i].pos[k] = s.position[k]; pl[i].color[k] = s.color[k]; }
            pl[i].intensity = s.intensity; pl[i].size = s.size;
        }
        UP(pl, &D.point_lights);
        ctx->first_point_light = pl.empty() ? DevPointLight{} : pl[0];
        UP(hs.areal_lights, &D.areal_lights);
        UP(hs.areal_tris, &D.areal_tris);
        const rgk_ltc_table* lt[2] = {&d->ltc_ggx, &d->ltc_beckmann};
        D.has_ltc = 1;
        for (int k = 0; k < 2; k++) {
            std::vector<float4> M; std::vector<float> A;
            if (lt[k]->M && lt[k]->amplitude) {
                M.resize(4096 * 3); A.assign(lt[k]->amplitude, lt[k]->amplitude + 4096);
                for (int i = 0; i < 4096; i++) {
                    const float* m = lt[k]->M + 9 * i;
                    M[3 * i] = make_float4(m[0], m[1], m[2], m[3]);
                    M[3 * i + 1] = make_float4(m[4], m[5], m[6], m[7]);
                    M[3 * i + 2] = make_float4(m[8], 0.0f, 0.0f, 0.0f);
                }
            } else D.has_ltc = 0;
            UP(M, &D.ltc_M[k]); UP(A, &D.ltc_amp[k]);
        }
        if (!D.has_ltc)
            for (uint32_t i = 0; i < d->n_materials; i++)
                if (d->materials[i].bxdf >= RGK_BXDF_LTC_BECKMANN) { free_scene(ctx); return rgk_fail(ctx, RGK_ERR_INVALID, "LTC material used but LTC tables not supplied"); }
        if (d->sky.mode == 1 && (d->sky.envmap < 0 || (uint32_t)d->sky.envmap >= d->n_textures)) { free_scene(ctx); return rgk_fail(ctx, RGK_ERR_INVALID, "sky envmap texture index out of range"); }
        D.sky_mode = d->sky.mode; D.sky_intensity = d->sky.intensity; D.sky_rotate = d->sky.rotate; D.sky_envmap = d->sky.envmap;
        for (int k = 0; k < 3; k++) D.sky_color[k] = d->sky.color[k];
    }
    D.n_nodes = hs.info.n_nodes; D.n_refs = hs.info.n_refs; D.n_triangles = hs.info.n_triangles; D.n_vertices = d->n_vertices;
    D.n_materials = d->n_materials; D.n_textures = d->n_textures; D.n_point_lights = d->n_point_lights;
    D.n_areal_lights = hs.info.n_areal_lights;
    D.total_point_power = hs.info.total_point_power; D.total_areal_power = hs.info.total_areal_power;
    D.epsilon = hs.info.epsilon;
    D.refill_threshold = ctx->cfg.refill_batch;
    for (int k = 0; k < 6; k++) D.bb[k] = hs.info.bbox[k];
    RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->has_scene = true;
    return RGK_OK;
}

rgk_status rgk_scene_get_info(const rgk_context* ctx, rgk_scene_info* out) {
    if (!ctx || !out) return RGK_ERR_INVALID;
    if (!ctx->has_scene) return RGK_ERR_NO_SCENE;
    *out = ctx->host.info;
    return RGK_OK;
}

rgk_status rgk_scene_get_kdtree(const rgk_context* ctx, uint32_t* nodes, uint32_t* refs) {
    if (!ctx || !nodes || !refs) return RGK_ERR_INVALID;
    if (!ctx->has_scene) return RGK_ERR_NO_SCENE;
    std::memcpy(nodes, ctx->host.nodes.data(), 4 * ctx->host.nodes.size());
    std::memcpy(refs, ctx->host.refs.data(), 4 * ctx->host.refs.size());
    return RGK_OK;
}

// ---- host-only scene commit (no device needed): what rgk_scene_commit computes before uploading ------------------
struct rgk_host_scene { HostScene hs; std::string error; };
thread_local std::string g_host_error;

rgk_status rgk_host_scene_create(const rgk_scene_desc* d, const rgk_kdtree* tree, const rgk_device_cfg* cfg, rgk_host_scene** out) {
    if (!d || !out) return RGK_ERR_INVALID;
    *out = nullptr;
    const rgk_device_cfg use = cfg ? *cfg : default_device_cfg();
    if (const char* why = check_cfg(use)) { g_host_error = why; return RGK_ERR_INVALID; }
    rgk_host_scene* h = new rgk_host_scene();
    try {
        host_scene_commit(d, tree, use, h->hs);
    } catch (const std::exception& e) {
        g_host_error = e.what();
        delete h;
        return RGK_ERR_INVALID;
    }
    *out = h;
    return RGK_OK;
}
void rgk_host_scene_destroy(rgk_host_scene* h) { delete h; }
const char* rgk_host_last_error(void) { return g_host_error.c_str(); }
rgk_status rgk_host_scene_get_info(const rgk_host_scene* h, rgk_scene_info* out) {
    if (!h || !out) return RGK_ERR_INVALID;
    *out = h->hs.info;
    return RGK_OK;
}
rgk_status rgk_host_scene_get_kdtree(const rgk_host_scene* h, uint32_t* nodes, uint32_t* refs) {
    if (!h || !nodes || !refs) return RGK_ERR_INVALID;
    std::memcpy(nodes, h->hs.nodes.data(), 4 * h->hs.nodes.size());
    std::memcpy(refs, h->hs.refs.data(), 4 * h->hs.refs.size());
    return RGK_OK;
}
// triangle intersection records (12 floats / triangle) and planes (4 floats / triangle)
rgk_status rgk_host_scene_get_records(const rgk_host_scene* h, float* planes, float* records) {
    if (!h) return RGK_ERR_INVALID;
    if (planes) std::memcpy(planes, h->hs.planes.data(), 4 * h->hs.planes.size());
    if (records) std::memcpy(records, h->hs.tri_isect.data(), 4 * h->hs.tri_isect.size());
    return RGK_OK;
}

// conservative 2-D bounds of every triangle (4 floats; the axis code rides in the two low mantissa bits of the first)
rgk_status rgk_host_scene_get_bounds(const rgk_host_scene* h, float* bounds) {
    if (!h || !bounds) return RGK_ERR_INVALID;
    std::memcpy(bounds, h->hs.tri_bounds.data(), 4 * h->hs.tri_bounds.size());
    return RGK_OK;
}

// the wide BVH (RGK_TRAVERSAL_BVH): sizes (all 0 when it is off), then nodes (32 floats each) and leaf order
rgk_status rgk_host_scene_get_bvh_size(const rgk_host_scene* h, uint32_t* n_nodes, uint32_t* n_slots, uint32_t* depth) {
    if (!h) return RGK_ERR_INVALID;
    if (n_nodes) *n_nodes = (uint32_t)(h->hs.bvh_nodes.size() / 32);
    if (n_slots) *n_slots = (uint32_t)h->hs.bvh_order.size();
    if (depth) *depth = h->hs.bvh_depth;
    return RGK_OK;
}
rgk_status rgk_host_scene_get_bvh(const rgk_host_scene* h, float* nodes, uint32_t* order) {
    if (!h || !nodes || !order) return RGK_ERR_INVALID;
    std::memcpy(nodes, h->hs.bvh_nodes.data(), 4 * h->hs.bvh_nodes.size());
    std::memcpy(order, h->hs.bvh_order.data(), 4 * h->hs.bvh_order.size());
    return RGK_OK;
}

// counters of the wide-BVH traversal launches since the previous call, closest-hit launches then any-hit launches: rays,
// ambiguous (re-traced through the kd-tree), wide nodes visited, exact triangle tests, leaf slots scanned (the last three only
// while rgk_render_set_counting is on).  Zeros when the BVH is off.
rgk_status rgk_bvh_stats(rgk_context* ctx, uint64_t out[10]) {
    if (!ctx || !out) return RGK_ERR_INVALID;
    for (int k = 0; k < 10; k++) out[k] = 0;
    if (!ctx->d_bvh_stats) return RGK_OK;
    BvhStats h[2] = {};
    RGK_CUDA(ctx, cudaMemcpyAsync(h, ctx->d_bvh_stats, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    RGK_CUDA(ctx, cudaMemsetAsync(ctx->d_bvh_stats, 0, sizeof(h), ctx->stream));
    RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 2; k++) { out[5 * k] = h[k].rays; out[5 * k + 1] = h[k].ambiguous; out[5 * k + 2] = h[k].nodes; out[5 * k + 3] = h[k].tests; out[5 * k + 4] = h[k].slots; }
    return RGK_OK;
}

// ---- traversal -----------------------------------------------------------
rgk_status rgk_trace_closest_device(rgk_context* ctx, const rgk_ray* d_rays, const uint32_t* d_ignore, uint64_t n,
                                    rgk_hit* d_hits, rgk_trav_stats* d_stats) {
    if (!ctx || (n && (!d_rays || !d_hits))) return RGK_ERR_INVALID;
    if (!ctx->has_scene) return rgk_fail(ctx, RGK_ERR_NO_SCENE, "rgk_scene_commit has not been called");
    return launch_trace_closest(ctx, d_rays, d_ignore, n, d_hits, d_stats);
}

rgk_status rgk_trace_shadow_device(rgk_context* ctx, const float* d_a, const float* d_b, uint64_t n, uint8_t* d_visible,
                                   rgk_trav_stats* d_stats) {
    if (!ctx || (n && (!d_a || !d_b || !d_visible))) return RGK_ERR_INVALID;
    if (!ctx->has_scene) return rgk_fail(ctx, RGK_ERR_NO_SCENE, "rgk_scene_commit has not been called");
    return launch_trace_shadow(ctx, d_a, d_b, n, d_visible, d_stats);
}

rgk_status rgk_trace_closest(rgk_context* ctx, const rgk_ray* rays, const uint32_t* ignore, uint64_t n, rgk_hit* hits,
                             rgk_trav_stats* stats) {
    if (!ctx || (n && (!rays || !hits))) return RGK_ERR_INVALID;
    if (!ctx->has_scene) return rgk_fail(ctx, RGK_ERR_NO_SCENE, "rgk_scene_commit has not been called");
    if (n == 0) { if (stats) std::memset(stats, 0, sizeof *stats); return RGK_OK; }
    RGK_CUDA(ctx, cudaSetDevice(ctx->device));
    rgk_ray* d_rays = (rgk_ray*)rgk_scratch(ctx, 0, n * sizeof(rgk_ray));
    rgk_hit* d_hits = (rgk_hit*)rgk_scratch(ctx, 1, n * sizeof(rgk_hit));
    uint32_t* d_ign = ignore ? (uint32_t*)rgk_scratch(ctx, 2, n * 4) : nullptr;
    if (!d_rays || !d_hits || (ignore && !d_ign)) return rgk_fail(ctx, RGK_ERR_NOMEM, "device scratch allocation failed");
    RGK_CUDA(ctx, cudaMemcpyAsync(d_rays, rays, n * sizeof(rgk_ray), cudaMemcpyHostToDevice, ctx->stream));
    if (ignore) RGK_CUDA(ctx, cudaMemcpyAsync(d_ign, ignore, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (stats) RGK_CUDA(ctx, cudaMemsetAsync(ctx->d_stats, 0, sizeof(rgk_trav_stats), ctx->stream));
    rgk_status s = launch_trace_closest(ctx, d_rays, d_ign, n, d_hits, stats ? ctx->d_stats : nullptr);
    if (s != RGK_OK) return s;
    RGK_CUDA(ctx, cudaMemcpyAsync(hits, d_hits, n * sizeof(rgk_hit), cudaMemcpyDeviceToHost, ctx->stream));
    if (stats) RGK_CUDA(ctx, cudaMemcpyAsync(stats, ctx->d_stats, sizeof(rgk_trav_stats), cudaMemcpyDeviceToHost, ctx->stream));
    RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RGK_OK;
}

rgk_status rgk_trace_shadow(rgk_context* ctx, const float* a, const float* b, uint64_t n, uint8_t* visible,
                            rgk_trav_stats* stats) {
    if (!ctx || (n && (!a || !b || !visible))) return RGK_ERR_INVALID;
    if (!ctx->has_scene) return rgk_fail(ctx, RGK_ERR_NO_SCENE, "rgk_scene_commit has not been called");
    if (n == 0) { if (stats) std::memset(stats, 0, sizeof *stats); return RGK_OK; }
    RGK_CUDA(ctx, cudaSetDevice(ctx->device));
    float* d_a = (float*)rgk_scratch(ctx, 0, n * 12);
    float* d_b = (float*)rgk_scratch(ctx, 1, n * 12);
    uint8_t* d_v = (uint8_t*)rgk_scratch(ctx, 2, n);
    if (!d_a || !d_b || !d_v) return rgk_fail(ctx, RGK_ERR_NOMEM, "device scratch allocation failed");
    RGK_CUDA(ctx, cudaMemcpyAsync(d_a, a, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    RGK_CUDA(ctx, cudaMemcpyAsync(d_b, b, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    if (stats) RGK_CUDA(ctx, cudaMemsetAsync(ctx->d_stats, 0, sizeof(rgk_trav_stats), ctx->stream));
    rgk_status s = launch_trace_shadow(ctx, d_a, d_b, n, d_v, stats ? ctx->d_stats : nullptr);
    if (s != RGK_OK) return s;
    RGK_CUDA(ctx, cudaMemcpyAsync(visible, d_v, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (stats) RGK_CUDA(ctx, cudaMemcpyAsync(stats, ctx->d_stats, sizeof(rgk_trav_stats), cudaMemcpyDeviceToHost, ctx->stream));
    RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RGK_OK;
}

// ---- camera / tasks / sampler ------------------------------------------------
// Camera::Camera, src/camera.cpp:7-24 (GLM normalize / cross formulas, no FMA on the host)
void rgk_camera_init(rgk_camera* c, const float pos[3], const float la[3], const float up[3], float yview, float xview,
                     int32_t xres, int32_t yres, float focus_plane, float lens_size) {
    struct V { float x, y, z; };
    auto sub = [](V a, V b) { return V{a.x - b.x, a.y - b.y, a.z - b.z}; };
    auto add = [](V a, V b) { return V{a.x + b.x, a.y + b.y, a.z + b.z}; };
    auto mul = [](V a, float s) { return V{a.x * s, a.y * s, a.z * s}; };
    auto smul = [](float s, V a) { return V{s * a.x, s * a.y, s * a.z}; };
    auto nrm = [&](V v) { const float d = v.x * v.x + v.y * v.y + v.z * v.z; return mul(v, 1.0f / std::sqrt(d)); };
    auto crs = [](V x, V y) { return V{x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; };
    const V origin{pos[0], pos[1], pos[2]}, lookat{la[0], la[1], la[2]}, upv{up[0], up[1], up[2]};
    const V direction = nrm(sub(lookat, origin));
    const V left = nrm(crs(upv, direction));
    const V cup = nrm(crs(left, direction));
    const V vx = mul(smul(-xview, left), focus_plane);
    const V vy = mul(smul(yview, cup), focus_plane);
    const V vs = sub(sub(add(origin, mul(direction, focus_plane)), smul(0.5f, vy)), smul(0.5f, vx));
    auto put = [](float* d, V v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; };
    put(c->origin, origin); put(c->lookat, lookat); put(c->direction, direction); put(c->cameraup, cup); put(c->cameraleft, left);
    put(c->viewscreen, vs); put(c->viewscreen_x, vx); put(c->viewscreen_y, vy);
    c->lens_size = lens_size; c->xsize = xres; c->ysize = yres;
}

rgk_status rgk_camera_rays(rgk_context* ctx, const rgk_camera* cam, uint32_t xres, uint32_t yres, const int32_t* xy,
                           const float* offsets, const float* lens, uint64_t n, rgk_ray* rays) {
    if (!ctx || !cam || (n && (!xy || !offsets || !rays))) return RGK_ERR_INVALID;
    if (cam->lens_size != 0.0f && !lens) return rgk_fail(ctx, RGK_ERR_INVALID, "lens samples required when lens_size != 0");
    if (n == 0) return RGK_OK;
    RGK_CUDA(ctx, cudaSetDevice(ctx->device));
    char* buf = (char*)rgk_scratch(ctx, 0, n * (8 + 8 + 8));
    rgk_ray* d_rays = (rgk_ray*)rgk_scratch(ctx, 1, n * sizeof(rgk_ray));
    if (!buf || !d_rays) return rgk_fail(ctx, RGK_ERR_NOMEM, "device scratch allocation failed");
    int32_t* d_xy = (int32_t*)buf; float* d_off = (float*)(buf + 8 * n); float* d_lens = (float*)(buf + 16 * n);
    RGK_CUDA(ctx, cudaMemcpyAsync(d_xy, xy, 8 * n, cudaMemcpyHostToDevice, ctx->stream));
    RGK_CUDA(ctx, cudaMemcpyAsync(d_off, offsets, 8 * n, cudaMemcpyHostToDevice, ctx->stream));
    if (lens) RGK_CUDA(ctx, cudaMemcpyAsync(d_lens, lens, 8 * n, cudaMemcpyHostToDevice, ctx->stream));
    rgk_status s = launch_camera_rays(ctx, cam, xres, yres, d_xy, d_off, lens ? d_lens : nullptr, n, d_rays);
    if (s != RGK_OK) return s;
    RGK_CUDA(ctx, cudaMemcpyAsync(rays, d_rays, n * sizeof(rgk_ray), cudaMemcpyDeviceToHost, ctx->stream));
    RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RGK_OK;
}

// GenerateTaskList, src/render_driver.cpp:30-46; RenderTask::midpoint, src/tracer.hpp:18.
// std::sort with the same comparator on the same initial order as the reference.
uint32_t rgk_generate_tasks(uint32_t tile, uint32_t xres, uint32_t yres, rgk_task* out, uint32_t capacity) {
    if (tile == 0) return 0;
    struct Item { rgk_task t; float mx, my; };
    std::vector<Item> items;
    for (uint32_t yp = 0; yp < yres; yp += tile)
        for (uint32_t xp = 0; xp < xres; xp += tile) {
            Item it;
            it.t = rgk_task{xp, std::min(xres, xp + tile), yp, std::min(yres, yp + tile)};
            it.mx = (it.t.x1 + it.t.x2) / 2.0f; it.my = (it.t.y1 + it.t.y2) / 2.0f;
            items.push_back(it);
        }
    const float cx = xres / 2.0f, cy = yres / 2.0f;
    auto len = [cx, cy](const Item& a) { const float dx = cx - a.mx, dy = cy - a.my; return std::sqrt(dx * dx + dy * dy); };
    std::sort(items.begin(), items.end(), [&](const Item& a, const Item& b) { return len(a) < len(b); });
    if (out) for (uint32_t i = 0; i < items.size() && i < capacity; i++) out[i] = items[i].t;
    return (uint32_t)items.size();
}

uint32_t rgk_sampler_set_size(uint32_t multisample) { return host_sampler_set_size(multisample); }

rgk_status rgk_sampler_tables(rgk_context* ctx, const uint32_t* seeds, uint32_t n_seeds, uint32_t ms, uint32_t n1d, uint32_t n2d,
                              float* out1d, float* out2d) {
    if (!ctx || !seeds || ms == 0 || n1d > 64 || n2d > 64 || (n1d && !out1d) || (n2d && !out2d)) return RGK_ERR_INVALID;
    if (n_seeds == 0) return RGK_OK;
    RGK_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t ss = host_sampler_set_size(ms);
    const size_t b1 = (size_t)n_seeds * n1d * ss * 4, b2 = (size_t)n_seeds * n2d * ss * 8;
    uint32_t* d_seeds = (uint32_t*)rgk_scratch(ctx, 0, (size_t)n_seeds * 4);
    float* d1 = (float*)rgk_scratch(ctx, 1, b1);
    float* d2 = (float*)rgk_scratch(ctx, 2, b2);
    if (!d_seeds || !d1 || !d2) return rgk_fail(ctx, RGK_ERR_NOMEM, "device scratch allocation failed");
    RGK_CUDA(ctx, cudaMemcpyAsync(d_seeds, seeds, (size_t)n_seeds * 4, cudaMemcpyHostToDevice, ctx->stream));
    rgk_status s = launch_sampler_tables(ctx, d_seeds, n_seeds, ms, n1d, n2d, d1, d2);
    if (s != RGK_OK) return s;
    // device layout is [dim][set][seed]; the ABI returns [seed][dim][set]
    std::vector<float> h1(b1 / 4), h2(b2 / 4);
    if (b1) RGK_CUDA(ctx, cudaMemcpyAsync(h1.data(), d1, b1, cudaMemcpyDeviceToHost, ctx->stream));
    if (b2) RGK_CUDA(ctx, cudaMemcpyAsync(h2.data(), d2, b2, cudaMemcpyDeviceToHost, ctx->stream));
    RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (uint32_t d = 0; d < n1d; d++)
        for (uint32_t k = 0; k < ss; k++)
            for (uint32_t i = 0; i < n_seeds; i++)
                out1d[((size_t)i * n1d + d) * ss + k] = h1[((size_t)d * ss + k) * n_seeds + i];
    for (uint32_t d = 0; d < n2d; d++)
        for (uint32_t k = 0; k < ss; k++)
            for (uint32_t i = 0; i < n_seeds; i++) {
                out2d[(((size_t)i * n2d + d) * ss + k) * 2] = h2[(((size_t)d * ss + k) * n_seeds + i) * 2];
                out2d[(((size_t)i * n2d + d) * ss + k) * 2 + 1] = h2[(((size_t)d * ss + k) * n_seeds + i) * 2 + 1];
            }
    return RGK_OK;
}

// ---- rendering --------------------------------------------------------------
static rgk_status check_render_args(rgk_context* ctx, const rgk_camera* cam, const rgk_render_params* p, const rgk_task* tasks,
                                    uint32_t n_tasks, const void* rgb, const void* count) {
    if (!ctx || !cam || !p || (n_tasks && !tasks) || !rgb || !count) return RGK_ERR_INVALID;
    if (!ctx->has_scene) return rgk_fail(ctx, RGK_ERR_NO_SCENE, "rgk_scene_commit has not been called");
    if (p->reverse > 16) return rgk_fail(ctx, RGK_ERR_UNSUPPORTED, "reverse > 16 light-path vertices");
    if (p->xres == 0 || p->yres == 0 || p->multisample == 0) return rgk_fail(ctx, RGK_ERR_INVALID, "xres, yres and multisample must be positive");
    // pixel coordinates travel packed as x | y << 16 (k_pixel_setup / k_raygen / k_finish)
    if (p->xres > 65535u || p->yres > 65535u) return rgk_fail(ctx, RGK_ERR_UNSUPPORTED, "image dimensions above 65535");
    if (p->sampler_mode > RGK_SAMPLER_FAST) return rgk_fail(ctx, RGK_ERR_INVALID, "unknown sampler_mode");
    for (uint32_t i = 0; i < n_tasks; i++)
        if (tasks[i].x1 > tasks[i].x2 || tasks[i].y1 > tasks[i].y2 || tasks[i].x2 > p->xres || tasks[i].y2 > p->yres)
            return rgk_fail(ctx, RGK_ERR_INVALID, "task outside the image");
    return RGK_OK;
}

rgk_status rgk_render_round_device(rgk_context* ctx, const rgk_camera* cam, const rgk_render_params* p, const rgk_task* tasks,
                                   uint32_t n_tasks, uint32_t seedstart, uint32_t seedcount_base, float* d_rgb, uint32_t* d_count,
                                   rgk_round_stats* stats) {
    rgk_status s = check_render_args(ctx, cam, p, tasks, n_tasks, d_rgb, d_count);
    if (s != RGK_OK) return s;
    RGK_CUDA(ctx, cudaSetDevice(ctx->device));
    return render_round_impl(ctx, cam, p, tasks, n_tasks, seedstart, seedcount_base, d_rgb, d_count, stats);
}

rgk_status rgk_render_round(rgk_context* ctx, const rgk_camera* cam, const rgk_render_params* p, const rgk_task* tasks,
                            uint32_t n_tasks, uint32_t seedstart, uint32_t seedcount_base, float* rgb_sum, uint32_t* count,
                            rgk_round_stats* stats) {
    rgk_status s = check_render_args(ctx, cam, p, tasks, n_tasks, rgb_sum, count);
    if (s != RGK_OK) return s;
    RGK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t npx = (size_t)p->xres * p->yres;
    float* d_rgb = (float*)rgk_scratch(ctx, 0, npx * 12);
    uint32_t* d_cnt = (uint32_t*)rgk_scratch(ctx, 1, npx * 4);
    if (!d_rgb || !d_cnt) return rgk_fail(ctx, RGK_ERR_NOMEM, "device framebuffer allocation failed");
    RGK_CUDA(ctx, cudaMemcpyAsync(d_rgb, rgb_sum, npx * 12, cudaMemcpyHostToDevice, ctx->stream));
    RGK_CUDA(ctx, cudaMemcpyAsync(d_cnt, count, npx * 4, cudaMemcpyHostToDevice, ctx->stream));
    s = render_round_impl(ctx, cam, p, tasks, n_tasks, seedstart, seedcount_base, d_rgb, d_cnt, stats);
    if (s != RGK_OK) return s;
    RGK_CUDA(ctx, cudaMemcpyAsync(rgb_sum, d_rgb, npx * 12, cudaMemcpyDeviceToHost, ctx->stream));
    RGK_CUDA(ctx, cudaMemcpyAsync(count, d_cnt, npx * 4, cudaMemcpyDeviceToHost, ctx->stream));
    RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RGK_OK;
}

// RenderDriver::RenderFrame, Rounds mode (src/render_driver.cpp:192-253): seedstart 42, seedcount running on.
rgk_status rgk_render_frame(rgk_context* ctx, const rgk_camera* cam, const rgk_render_params* p, uint32_t rounds,
                            float* rgb_sum, uint32_t* count, rgk_round_stats* stats) {
    if (!ctx || !cam || !p || !rgb_sum || !count) return RGK_ERR_INVALID;
    const uint32_t nt = rgk_generate_tasks(32, p->xres, p->yres, nullptr, 0);   // TILE_SIZE, src/global_config.hpp:8
    std::vector<rgk_task> tasks(nt);
    rgk_generate_tasks(32, p->xres, p->yres, tasks.data(), nt);
    rgk_status s = check_render_args(ctx, cam, p, tasks.data(), nt, rgb_sum, count);
    if (s != RGK_OK) return s;
    RGK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t npx = (size_t)p->xres * p->yres;
    float* d_rgb = (float*)rgk_scratch(ctx, 0, npx * 12);
    uint32_t* d_cnt = (uint32_t*)rgk_scratch(ctx, 1, npx * 4);
    if (!d_rgb || !d_cnt) return rgk_fail(ctx, RGK_ERR_NOMEM, "device framebuffer allocation failed");
    RGK_CUDA(ctx, cudaMemsetAsync(d_rgb, 0, npx * 12, ctx->stream));
    RGK_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, npx * 4, ctx->stream));
    rgk_round_stats total{};
    uint32_t seedcount = 0;
    for (uint32_t r = 0; r < rounds; r++) {
        rgk_round_stats rs{};
        s = render_round_impl(ctx, cam, p, tasks.data(), nt, 42u, seedcount, d_rgb, d_cnt, &rs);
        if (s != RGK_OK) return s;
        seedcount += nt;
        total.closest_rays += rs.closest_rays; total.shadow_rays += rs.shadow_rays; total.samples += rs.samples;
        total.kernel_launches += rs.kernel_launches; total.gpu_ms += rs.gpu_ms; total.trace_ms += rs.trace_ms;
        total.closest_ms += rs.closest_ms; total.shadow_ms += rs.shadow_ms; total.sampler_ms += rs.sampler_ms; total.shade_ms += rs.shade_ms;
        total.closest_launches += rs.closest_launches; total.shadow_launches += rs.shadow_launches;
        total.shadow_rays_skipped += rs.shadow_rays_skipped;
    }
    RGK_CUDA(ctx, cudaMemcpyAsync(rgb_sum, d_rgb, npx * 12, cudaMemcpyDeviceToHost, ctx->stream));
    RGK_CUDA(ctx, cudaMemcpyAsync(count, d_cnt, npx * 4, cudaMemcpyDeviceToHost, ctx->stream));
    RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (stats) *stats = total;
    return RGK_OK;
}

rgk_status rgk_accumulate_device(rgk_context* ctx, float* d_rgb_sum, const float* d_other_sum, uint64_t n_floats, uint32_t* d_count,
                                 const uint32_t* d_other_count, void* stream) {
    if (!ctx || (n_floats && (!d_rgb_sum || !d_other_sum)) || ((d_count == nullptr) != (d_other_count == nullptr))) return RGK_ERR_INVALID;
    if (n_floats == 0) return RGK_OK;
    RGK_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_accumulate(ctx, d_rgb_sum, d_other_sum, n_floats, d_count, d_other_count, stream ? (cudaStream_t)stream : ctx->stream);
}

rgk_status rgk_render_set_shard(rgk_context* ctx, uint32_t first, uint32_t stride) {
    if (!ctx || stride == 0 || first >= stride) return RGK_ERR_INVALID;
    ctx->shard_first = first; ctx->shard_stride = stride;
    return RGK_OK;
}
rgk_status rgk_render_set_counting(rgk_context* ctx, int enabled) {
    if (!ctx) return RGK_ERR_INVALID;
    ctx->counting = enabled != 0;
    return RGK_OK;
}
rgk_status rgk_render_get_trav_stats(const rgk_context* ctx, rgk_trav_stats* closest, rgk_trav_stats* shadow) {
    if (!ctx || !closest || !shadow) return RGK_ERR_INVALID;
    *closest = ctx->last_closest; *shadow = ctx->last_shadow;
    return RGK_OK;
}

rgk_status rgk_render_get_shade_stats(const rgk_context* ctx, uint64_t out[8]) {
    if (!ctx || !out) return RGK_ERR_INVALID;
    for (int k = 0; k < 8; k++) out[k] = ctx->last_shade[k];
    return RGK_OK;
}

rgk_status rgk_render_set_tables(rgk_context* ctx, uint32_t multisample, uint32_t n1d, uint32_t n2d, const float* t1d, const float* t2d,
                                 uint64_t n_pixels) {
    if (!ctx || multisample == 0 || n1d == 0 || n2d == 0 || n1d > 64 || n2d > 64 || !t1d || !t2d || n_pixels == 0) return RGK_ERR_INVALID;
    RGK_CUDA(ctx, cudaSetDevice(ctx->device));
    RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const uint32_t ss = host_sampler_set_size(multisample);
    if (ctx->d_user_t1) { cudaFree(ctx->d_user_t1); ctx->d_user_t1 = nullptr; }
    if (ctx->d_user_t2) { cudaFree(ctx->d_user_t2); ctx->d_user_t2 = nullptr; }
    const size_t b1 = (size_t)n_pixels * n1d * ss * 4, b2 = (size_t)n_pixels * n2d * ss * 8;
    RGK_CUDA(ctx, cudaMalloc((void**)&ctx->d_user_t1, b1));
    RGK_CUDA(ctx, cudaMalloc((void**)&ctx->d_user_t2, b2));
    RGK_CUDA(ctx, cudaMemcpy(ctx->d_user_t1, t1d, b1, cudaMemcpyHostToDevice));
    RGK_CUDA(ctx, cudaMemcpy(ctx->d_user_t2, t2d, b2, cudaMemcpyHostToDevice));
    ctx->user_n1d = n1d; ctx->user_n2d = n2d; ctx->user_ss = ss; ctx->user_npix = n_pixels;
    return RGK_OK;
}

} // extern "C"
