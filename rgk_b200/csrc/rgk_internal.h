// rgk_internal.h -- shared between the host layer (scene commit, driver) and the CUDA
// kernels of librgk_b200.so.  Not part of the public ABI (include/rgk_b200.h).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "rgk_b200.h"

// ------------------------------------------------------------------ device scene layout (HBM)
// All arrays are 16-byte aligned (cudaMalloc) and read through the read-only path.
//
//  nodes     uint2 / node     the reference's CompressedKdNode words (src/scene.hpp:212-253):
//                             .x = split_plane bits | first ref, .y = (other child | n refs) << 2 | kind
//  refs      u32 / reference  Scene::compressed_triangles
//  tri_isect 3 x float4 / triangle (48 B), the ray-independent part of
//            Triangle::TestIntersection (src/primitives.cpp:83,104-133,149), precomputed with
//            the reference's float op order:
//              [0] plane n.x n.y n.z d
//              [1] v0[i1] v0[i2] q1.x q1.y          (q1 = v1 - v0 projected on the dominant plane)
//              [2] q2.x q2.y denom flags             (denom = q2.y*q1.x - q2.x*q1.y;
//                                                     flags bits 0-1 = axis code, bit 2 = |q1.x| < eps)
//  tri_shade uint4 / triangle  va vb vc material   (touched by shading only)
//  positions/normals/tangents float4 / vertex, texcoords float2 / vertex
struct DevTexture {          // 32 B
    uint32_t kind, width, height, _pad;
    float color[3];
    uint32_t offset;         // first texel in the float4 texel pool
};
struct DevMaterial {         // 64 B (same fields as rgk_material)
    uint32_t bxdf, no_russian;
    float emission[3];
    float roughness, ior, amount;
    int32_t mix_a, mix_b, tex_diffuse, tex_color, tex_bump;
    uint32_t _pad[3];
};
struct DevPointLight { float pos[3]; float color[3]; float intensity; float size; };
struct DevArealLight { float power, total_area; float emission[3]; uint32_t first, count; uint32_t _pad; };
struct DevArealTri { float area; uint32_t tri; };

struct DevScene {
    const uint2* nodes;
    const uint32_t* refs;
    const float4* ref_planes;    // [n_refs] plane record (n.xyz, d) of triangle refs[j]: the leaf scan reads reference and plane
                                 // side by side instead of chasing refs[j] -> tri_isect[3*refs[j]] (one dependent load less)
    const float4* ref_bounds;    // [n_refs] conservative 2-D bounds of triangle refs[j] in its projection plane
                                 // (lo1 | axis code in the two low mantissa bits, hi1, lo2, hi2): pre-filter of the exact test
    const float4* tri_isect;
    const uint4* tri_shade;
    const float4* positions;
    const float4* normals;
    const float4* tangents;
    const float2* texcoords;
    const DevMaterial* materials;
    const DevTexture* textures;
    const float4* texels;
    const DevPointLight* point_lights;
    const DevArealLight* areal_lights;
    const DevArealTri* areal_tris;
    const float4* ltc_M[2];      // 3 x float4 per entry (9 floats + pad), [0] GGX [1] Beckmann
    const float* ltc_amp[2];
    uint32_t n_nodes, n_refs, n_triangles, n_vertices, n_materials, n_textures;
    uint32_t n_point_lights, n_areal_lights;
    float total_point_power, total_areal_power;
    float epsilon;
    float bb[6];
    uint32_t sky_mode; float sky_color[3]; float sky_intensity, sky_rotate; int32_t sky_envmap;
    uint32_t has_ltc;
    uint32_t refill_threshold;   // idle lanes of a warp that trigger a refill (rgk_device_cfg::refill_*)
    // wide BVH (rgk_device_cfg::traversal == RGK_TRAVERSAL_BVH, host_bvh.cpp / bvh_device.cuh); bvh_nodes == nullptr: kd-tree only
    const float4* bvh_nodes;     // 8 x float4 per node
    const uint32_t* bvh_refs;    // [n_triangles] triangle of leaf slot j
    const float4* bvh_planes;    // [n_triangles] plane record of the triangle in leaf slot j
};

#define RGK_STACK_CAP 64  // traversal stack entries per ray (tree depth <= log2(n)+8, src/scene.cpp:409)

// ------------------------------------------------------------------ host scene (product host layer)
struct HostScene {
    std::vector<uint32_t> nodes, refs;          // reference encoding
    std::vector<float> planes;                  // 4 / triangle
    std::vector<float> tri_isect;               // 12 / triangle
    std::vector<float> tri_bounds;              // 4 / triangle (see DevScene::ref_bounds)
    std::vector<uint32_t> tri_shade;            // 4 / triangle
    std::vector<DevArealLight> areal_lights;
    std::vector<DevArealTri> areal_tris;
    std::vector<float> bvh_nodes;               // 32 / wide node (host_bvh.cpp); empty on the kd-only traversal
    std::vector<uint32_t> bvh_order;            // triangle of leaf slot j
    unsigned bvh_depth = 0;
    uint32_t nan_prone_triangles = 0;           // triangles that keep the scene on the kd path (host_scene.cpp)
    rgk_scene_info info{};
};
// host_bvh.cpp: ev[axis][2 i], [2 i + 1] = min, max of triangle i along the axis
void host_bvh_build(const std::vector<float> ev[3], uint32_t nt, const rgk_device_cfg& cfg, HostScene& hs);
// Scene::Commit (src/scene.cpp:294-429) on the host. Throws std::runtime_error on bad input.
void host_scene_commit(const rgk_scene_desc* d, const rgk_kdtree* tree, const rgk_device_cfg& cfg, HostScene& out);
rgk_device_cfg default_device_cfg();          // api.cu: what rgk_device_cfg_init writes

// ------------------------------------------------------------------ context
// counters of the wide-BVH launches since the last call (device buffer ctx->d_bvh_stats[2]: [0] closest-hit launches, [1] any-hit):
// rays, rays deferred to the kd arbiter, and -- counting instantiations only -- wide nodes visited, exact tests, leaf slots scanned
struct BvhStats { unsigned long long rays, ambiguous, nodes, tests, slots; };
struct PathBuffers;   // render.cu
struct rgk_context {
    int device = 0;
    rgk_device_cfg cfg = default_device_cfg();   // rgk_context_configure
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string last_error;
    bool has_scene = false;
    HostScene host;
    DevScene dev{};
    std::vector<void*> scene_allocs;
    // scratch for host-buffer entry points
    void* scratch[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // 0-2 staging, 3 work counters, 4 deferred-ray list
    size_t scratch_size[5] = {0, 0, 0, 0, 0};
    rgk_trav_stats* d_stats = nullptr;
    PathBuffers* paths = nullptr;
    uint64_t launches = 0;
    // caller-supplied sampler tables (RGK_SAMPLER_TABLES)
    float* d_user_t1 = nullptr; float* d_user_t2 = nullptr; uint32_t user_n1d = 0, user_n2d = 0, user_ss = 0; uint64_t user_npix = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool counting = false;
    uint32_t shard_first = 0, shard_stride = 1;   // rgk_render_set_shard
    rgk_trav_stats last_closest{}, last_shadow{};
    unsigned long long last_shade[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // k_shade work counters of the last counting round
    BvhStats* d_bvh_stats = nullptr;              // wide-BVH counters (allocated with the scene when the BVH is on)
    DevPointLight first_point_light{};            // host copy of point light 0 (single fixed light: no per-path light records)
};

rgk_status rgk_fail(rgk_context* ctx, rgk_status s, const std::string& msg);
#define RGK_CUDA(ctx, call)                                                                         \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return rgk_fail(ctx, e_ == cudaErrorMemoryAllocation ? RGK_ERR_NOMEM : RGK_ERR_CUDA,   \
                            std::string(#call) + ": " + cudaGetErrorString(e_));                    \
    } while (0)

void* rgk_scratch(rgk_context* ctx, int slot, size_t bytes);  // grows a reusable device buffer (nullptr on failure)

// trace.cu
rgk_status launch_trace_closest(rgk_context* ctx, const rgk_ray* d_rays, const uint32_t* d_ignore, uint64_t n,
                                rgk_hit* d_hits, rgk_trav_stats* d_stats);
rgk_status launch_trace_shadow(rgk_context* ctx, const float* d_a, const float* d_b, uint64_t n,
                               uint8_t* d_visible, rgk_trav_stats* d_stats);
// render.cu
rgk_status render_round_impl(rgk_context* ctx, const rgk_camera* cam, const rgk_render_params* p, const rgk_task* tasks,
                             uint32_t n_tasks, uint32_t seedstart, uint32_t seedcount_base, float* d_rgb, uint32_t* d_count,
                             rgk_round_stats* stats);
rgk_status launch_camera_rays(rgk_context* ctx, const rgk_camera* cam, uint32_t xres, uint32_t yres, const int32_t* d_xy,
                              const float* d_off, const float* d_lens, uint64_t n, rgk_ray* d_rays);
rgk_status launch_sampler_tables(rgk_context* ctx, const uint32_t* d_seeds, uint32_t n_seeds, uint32_t multisample,
                                 uint32_t n1d, uint32_t n2d, float* d_out1, float* d_out2);
rgk_status launch_accumulate(rgk_context* ctx, float* d_dst, const float* d_src, uint64_t n_floats, uint32_t* d_cnt, const uint32_t* d_ocnt, cudaStream_t stream);
void free_path_buffers(rgk_context* ctx);
uint32_t host_sampler_set_size(uint32_t multisample);
