/* rgk_b200.h -- C ABI of librgk_b200.so: the B200 (sm_100a) implementation of
 * RGKrt's data-parallel hot path (kd-tree closest-hit / shadow traversal and the
 * unidirectional path-tracing bounce loop).
 *
 * The reference (Enhex/RGK) has no plugin or FFI interface; its seam is the C++
 * class boundary RenderDriver -> PathTracer -> Scene.  Every entry point below
 * names the reference interface it replaces (paths relative to the reference
 * tree).  Plain pointers and sizes only; the caller owns every host buffer, the
 * library owns device memory behind the opaque context; no exception crosses
 * this boundary -- every call returns an rgk_status and rgk_last_error() holds
 * the text.  There is no CPU fallback: without a CUDA device every compute
 * entry point fails with RGK_ERR_NO_DEVICE.
 */
#ifndef RGK_B200_H
#define RGK_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define RGK_ABI_VERSION 4   /* 2: rgk_trav_stats grew the pre-filter counters; rgk_host_scene_*; reverse > 0.  3: wide-BVH entry points.
                               4: struct rgk_device_cfg -- the library reads no environment variable; the wide BVH is the default traversal */

typedef enum rgk_status {
    RGK_OK = 0,
    RGK_ERR_INVALID = 1,     /* bad argument / inconsistent scene (reference: ConfigFileException, std::runtime_error) */
    RGK_ERR_CUDA = 2,        /* CUDA runtime error (text in rgk_last_error) */
    RGK_ERR_NOMEM = 3,
    RGK_ERR_NO_DEVICE = 4,   /* no CUDA device: the library never falls back to the CPU */
    RGK_ERR_NO_SCENE = 5,    /* compute call before rgk_scene_commit */
    RGK_ERR_UNSUPPORTED = 6  /* e.g. recursion-max beyond the 64 tabulated sampler dimensions */
} rgk_status;

typedef struct rgk_context rgk_context;

/* ---- scene description (host side, input) ------------------------------- */

/* BxDF kinds, src/bxdf/bxdf.cpp:63-84 */
enum {
    RGK_BXDF_DIFFUSE = 0, RGK_BXDF_MIX = 1, RGK_BXDF_DIELECTRIC = 2, RGK_BXDF_MIRROR = 3,
    RGK_BXDF_TRANSPARENT = 4, RGK_BXDF_LTC_BECKMANN = 5, RGK_BXDF_LTC_GGX = 6,
    RGK_BXDF_LTC_BECKMANN_DIFFUSE = 7, RGK_BXDF_LTC_GGX_DIFFUSE = 8
};

/* ReadableTexture, src/texture.hpp:10-80.  kind 0 = SolidTexture(color),
 * kind 1 = FileTexture(width x height, texels[(y*width+x)*3 + c], already
 * gamma-decoded / flipped by the loader, src/texture.cpp:189-321). */
typedef struct rgk_texture {
    uint32_t kind;
    uint32_t width, height;
    float color[3];
    const float* texels;
} rgk_texture;

/* Material + its BxDF, src/bxdf/bxdf.hpp:19-159.  Texture slots index
 * rgk_scene_desc::textures; -1 = EmptyTexture (only meaningful for tex_bump,
 * "no bump map"; an absent colour slot reads as black like EmptyTexture).
 *   diffuse:           tex_diffuse
 *   mirror/dielectric: tex_color (+ ior)
 *   ltc_*:             tex_color (specular), roughness; *_diffuse adds tex_diffuse
 *   mix:               mix_a, mix_b (material indices, must be < own index), amount */
typedef struct rgk_material {
    uint32_t bxdf;
    uint32_t no_russian;
    float emission[3];
    float roughness;
    float ior;
    float amount;
    int32_t mix_a, mix_b;
    int32_t tex_diffuse, tex_color, tex_bump;
    uint32_t _pad[3];
} rgk_material; /* 64 bytes */

/* One imported mesh / primitive: a run of triangles sharing one material.  A
 * mesh whose material emits becomes one ArealLight (src/scene.cpp:149-206,
 * 215-249), which is why the grouping is part of the input. */
typedef struct rgk_mesh {
    uint32_t first_triangle, n_triangles;
    uint32_t material;
    uint32_t _pad;
} rgk_mesh;

/* Light{FULL_SPHERE}, src/primitives.hpp:26-43, src/config.cpp:372-387 */
typedef struct rgk_point_light {
    float position[3];
    float color[3];
    float intensity;
    float size;
} rgk_point_light;

/* src/scene.hpp:122-133,164-172: mode 0 = SimpleRadiance(color*intensity),
 * mode 1 = lat-long envmap (texture index), rotate in degrees. */
typedef struct rgk_sky {
    uint32_t mode;
    float color[3];
    float intensity;
    float rotate;
    int32_t envmap;
} rgk_sky;

/* LTC lobe tables, src/LTC/ltc.hpp:6-35: 64x64 entries, index = alpha + theta*64;
 * M as 9 floats per entry in mat33::m order (the double table cast to float as
 * mat33::operator glm::mat3 does), amplitude one float per entry. */
typedef struct rgk_ltc_table {
    const float* M;
    const float* amplitude;
} rgk_ltc_table;

typedef struct rgk_scene_desc {
    uint32_t n_vertices;
    const float* positions;   /* 3*n_vertices  Scene::vertices  */
    const float* normals;     /* 3*n_vertices  Scene::normals   */
    const float* tangents;    /* 3*n_vertices  Scene::tangents  */
    const float* texcoords;   /* 2*n_vertices  Scene::texcoords */
    uint32_t n_triangles;
    const uint32_t* indices;  /* 3*n_triangles (va,vb,vc) */
    uint32_t n_meshes;
    const rgk_mesh* meshes;   /* cover [0,n_triangles) in order */
    uint32_t n_materials;
    const rgk_material* materials;
    uint32_t n_textures;
    const rgk_texture* textures;
    uint32_t n_point_lights;
    const rgk_point_light* point_lights;
    rgk_sky sky;
    rgk_ltc_table ltc_ggx, ltc_beckmann;
    uint32_t thinglass; /* nonzero selects FindIntersectKdOtherThanWithThinglass (result-identical) */
} rgk_scene_desc;

/* The flattened kd-tree in the reference's own encoding, src/scene.hpp:212-253:
 * nodes[2*i] = split_plane bits | triangles_start, nodes[2*i+1] = (other_child
 * | triangles_num) << 2 | kind; left child = i+1; preorder (src/scene.cpp:637-657). */
typedef struct rgk_kdtree {
    uint32_t n_nodes;
    const uint32_t* nodes;
    uint32_t n_refs;
    const uint32_t* refs;     /* Scene::compressed_triangles */
} rgk_kdtree;

typedef struct rgk_scene_info {
    float epsilon;            /* Scene::epsilon = 1e-5 * bbox diagonal, src/scene.cpp:390 */
    float bbox[6];            /* xBB.first, xBB.second, yBB..., zBB... (src/scene.cpp:393-395) */
    uint32_t n_nodes, n_refs, n_triangles, n_areal_lights;
    uint32_t max_depth;       /* deepest node of the committed tree */
    float total_point_power, total_areal_power;
} rgk_scene_info;

/* ---- rays --------------------------------------------------------------- */

/* Ray, src/ray.hpp:6-29 (near = 0, far = 10000 for directional rays). */
typedef struct rgk_ray {
    float origin[3];
    float direction[3];
    float tnear, tfar;
} rgk_ray;

#define RGK_NO_TRIANGLE 0xFFFFFFFFu
/* Intersection, src/primitives.hpp:98-110 (triangle pointer -> index). */
typedef struct rgk_hit {
    uint32_t triangle;
    float t, a, b, c;
} rgk_hit;

/* Work counters of one traversal batch (algorithmic bytes per ray =
 * 36 + 20 + 8*inner + 8*leaf + 4*refs + 48*tests). */
typedef struct rgk_trav_stats {
    uint64_t rays, inner, leaf, refs, tests;
    /* device-side only (0 from a CPU checker): of the `tests`, how many reached the exact fp64 TestIntersection
     * arithmetic, how many were settled before it by the conservative 2-D bounds pre-filter, and how many of those
     * the exact test would have accepted (must be 0: the pre-filter may only reject what the reference rejects). */
    uint64_t exact, prefiltered, prefilter_wrong;
} rgk_trav_stats;

/* ---- rendering ---------------------------------------------------------- */

/* Camera's public fields, src/camera.hpp:27-41. */
typedef struct rgk_camera {
    float origin[3], lookat[3], direction[3];
    float cameraup[3], cameraleft[3];
    float viewscreen[3], viewscreen_x[3], viewscreen_y[3];
    float lens_size;
    int32_t xsize, ysize;
} rgk_camera;

enum {
    RGK_SAMPLER_MT19937 = 0, /* bit-exact device replica of StratifiedSampler(seed,64,ms), src/sampler.cpp:85-116 */
    RGK_SAMPLER_TABLES = 1,  /* caller supplies per-pixel sample tables (rgk_render_set_tables) */
    RGK_SAMPLER_FAST = 2     /* counter-based stratified+permuted sampler: same distribution, different sequence */
};

/* PathTracer ctor arguments, src/path_tracer.hpp:10-21 / src/render_driver.cpp:164-173 */
typedef struct rgk_render_params {
    uint32_t xres, yres;
    uint32_t multisample;
    uint32_t depth;           /* recursion-max */
    float clamp;
    float russian;
    float bumpmap_scale;
    uint32_t force_fresnell;  /* stored, never used (as in the reference) */
    uint32_t reverse;         /* light-path vertices of the bidirectional mode (src/path_tracer.cpp:336-398,462-480); 0 = unidirectional */
    uint32_t sampler_mode;
} rgk_render_params;

/* RenderTask, src/tracer.hpp:14-24 */
typedef struct rgk_task {
    uint32_t x1, x2, y1, y2;
} rgk_task;

typedef struct rgk_round_stats {
    uint64_t closest_rays;    /* PathTracer raycount (src/path_tracer.cpp:126) */
    uint64_t shadow_rays;     /* Scene::Visibility calls (not counted upstream) */
    uint64_t samples;         /* camera samples traced */
    uint64_t kernel_launches; /* kernels launched by the library during the call */
    float gpu_ms;             /* device time of the call, CUDA events on the context stream */
    float trace_ms;           /* of which closest-hit + shadow traversal kernels */
    float closest_ms;         /* closest-hit traversal kernels only */
    float shadow_ms;          /* shadow traversal (+ NEE resolve) kernels only */
    float sampler_ms;         /* sampler table generation */
    float shade_ms;           /* raygen + shade + finish */
    uint32_t closest_launches, shadow_launches;
    uint64_t shadow_rays_skipped; /* Visibility calls of the reference not traced because their direct term is exactly 0
                                     (the result cannot depend on them); shadow_rays counts traced rays only */
} rgk_round_stats;

/* ---- device configuration ----------------------------------------------- */

/* How the library runs the path on the device.  The reference has no counterpart (its only knobs are the JSON
 * config's, src/config.cpp, which arrive here as rgk_render_params); this replaces what round 1 read from RGK_*
 * environment variables -- the library itself reads none.  Fill with rgk_device_cfg_init, change what you need, pass to
 * rgk_context_configure BEFORE rgk_scene_commit (the traversal structure is built at commit) or to
 * rgk_host_scene_create.  `traversal` selects the acceleration structure; every other field is scheduling or memory
 * sizing and NEVER changes a result (tests/test_gpu_fullsize.py renders under several settings and compares bits). */
enum {
    RGK_TRAVERSAL_BVH = 0,   /* default: 4-wide BVH candidate pass for every ray + the reference's kd-tree as arbiter for the
                                rays whose answer could depend on the kd rule (~1e-3 of them); bit-identical to RGK_TRAVERSAL_KD */
    RGK_TRAVERSAL_KD = 1     /* the reference's kd-tree (src/scene_intersect.cpp:211-327) for every ray */
};
typedef struct rgk_device_cfg {
    uint32_t struct_size;        /* sizeof(rgk_device_cfg), set by rgk_device_cfg_init */
    uint32_t traversal;          /* RGK_TRAVERSAL_* */
    /* host build */
    uint32_t build_threads;      /* kd-tree build threads; 0 = all cores (at most 64) */
    uint32_t bvh_bins;           /* binned-SAH bins per axis of the wide-BVH build (32) */
    uint32_t bvh_all_axes;       /* 1: try the three axes at every split (default), 0: the longest only */
    uint32_t bvh_leaf_max;       /* 0: the SAH-optimal collapse chooses the leaves (<= 4 triangles); 1-4: binary leaves of at most this size */
    uint32_t bvh_greedy_collapse;/* 1: greedy collapse to 4 children instead of the SAH-optimal one */
    float    bvh_c_prim;         /* SAH cost of one triangle test relative to one node visit; 0 = built-in */
    uint32_t bvh_reinsert_iters; /* insertion-based re-optimisation passes over the binary tree (0 = off) */
    float    bvh_reinsert_frac;  /* fraction of the nodes each pass re-inserts (0.25) */
    /* memory sizing (bytes / paths); 0 = built-in, sized for 180 GB of HBM */
    uint64_t chunk_paths;        /* paths per chunk of a round (128 Mi) */
    uint64_t table_bytes;        /* sampler tables + generator states per chunk (24 GiB) */
    uint64_t reverse_bytes;      /* vertex storage of the bidirectional mode per chunk (8 GiB) */
    /* wavefront scheduling */
    uint32_t refill_batch;       /* idle lanes of a warp that trigger a refill in the batch entry points rgk_trace_* (16) */
    uint32_t refill_coherent;    /* ... camera-ray launches of a round (0 = built-in: 24 with the BVH, 32 kd) */
    uint32_t refill_incoherent;  /* ... bounce launches (24) */
    uint32_t refill_shadow;      /* ... shadow launches after the first bounce (12) */
    uint32_t binning;            /* 1: direction-binned continuation / shadow queues (k_bin), 0: atomic compaction in path order */
    uint32_t bin_shadow_first;   /* 1: also bin the shadow rays of the camera-ray hit points */
    uint32_t bin_items;          /* path slots per reordering group (2048) */
    float    bin_min_frac;       /* bounces with fewer live paths than this fraction of the chunk are not binned (0.25) */
    uint32_t shade_path_order;   /* 1: k_shade gathers its per-path state through the path-ordered twin of the queue */
    uint32_t skip_null_shadow;   /* 1: Visibility queries whose direct term is exactly 0 are not traced (counted in shadow_rays_skipped) */
    uint32_t const_light;        /* 1: a scene whose only light is one point light of size 0 keeps it in launch constants */
    uint32_t arb_grid;           /* CTAs per SM of the kd arbiter launches (8) */
    uint32_t bvh_shadow_nosort;  /* 1: any-hit rays enter BVH children in slot order */
    uint32_t bvh_closest_nearest;/* 1: closest-hit rays enter the nearest child first, the others in slot order */
    uint32_t sampler_smem;       /* 1: sampler tables of small set sizes are shuffled in shared memory */
    uint32_t trace_threads;      /* CTA size of the batch entry points: 64, 128 (default) or 256 */
    uint32_t kd_variant;         /* control structure of the kd traversal in the batch entry points: 6 phased (default), 2 per-lane */
    uint32_t sampler_ctas_per_sm;/* persistent CTAs per SM of the sampler kernel; each owns one 312 KB generator-state block, and
                                    SMs x this x 312 KB is its whole generator-state footprint (3: what 64 KB of shared-memory tables per CTA allows) */
    uint32_t sampler_kernel;     /* 0: the warp-per-pixel table builder (generator state in shared memory) for set sizes above 16 that leave
                                    it 16 warps per SM, else the thread-per-pixel one; 1: always thread per pixel; 2: always warp per pixel */
    uint32_t sampler_slots;      /* table slots per warp of the warp-per-pixel builder (0 = built-in: about two pixels' worth) */
    uint32_t _reserved[5];
} rgk_device_cfg;

/* ---- entry points ------------------------------------------------------- */

uint32_t rgk_abi_version(void);
const char* rgk_status_string(rgk_status s);

/* Replaces: nothing upstream (process-wide CPU state).  device = CUDA ordinal;
 * stream = a cudaStream_t to run on (NULL: the library creates its own). */
rgk_status rgk_context_create(int device, void* stream, rgk_context** out);
void rgk_context_destroy(rgk_context* ctx);
const char* rgk_last_error(const rgk_context* ctx);
/* Defaults (see the struct); rgk_context_configure validates and stores a copy -- fields of the host-build and
 * traversal groups take effect at the next rgk_scene_commit, the others at the next call. */
void rgk_device_cfg_init(rgk_device_cfg* cfg);
rgk_status rgk_context_configure(rgk_context* ctx, const rgk_device_cfg* cfg);
rgk_status rgk_context_get_cfg(const rgk_context* ctx, rgk_device_cfg* out);

/* Replaces Scene::Commit (src/scene.cpp:294-429): planes, areal lights, epsilon,
 * bbox, SAH kd-tree build + Compress on the host, flatten to the device layout,
 * upload.  If tree != NULL the given flattened tree is used instead of building
 * one (parity mode: the reference's own arrays). */
rgk_status rgk_scene_commit(rgk_context* ctx, const rgk_scene_desc* desc, const rgk_kdtree* tree);
rgk_status rgk_scene_get_info(const rgk_context* ctx, rgk_scene_info* out);
/* Copies out the committed tree in the reference encoding (sizes from rgk_scene_info). */
rgk_status rgk_scene_get_kdtree(const rgk_context* ctx, uint32_t* nodes, uint32_t* refs);

/* Host-only half of rgk_scene_commit (no device needed): Triangle::CalculatePlane (src/primitives.cpp:24-36), areal
 * lights, epsilon, bbox (src/scene.cpp:294-429) and the SAH kd-tree build + Compress (src/scene.cpp:431-657), the
 * build forked over host threads (rgk_device_cfg::build_threads, default = all cores; arrays are byte-identical to the
 * sequential reference procedure).  get_records copies the 4-float plane and the 12-float intersection record of
 * every triangle (either pointer may be NULL).  Errors: RGK_ERR_INVALID + rgk_host_last_error() (thread-local). */
typedef struct rgk_host_scene rgk_host_scene;
rgk_status rgk_host_scene_create(const rgk_scene_desc* desc, const rgk_kdtree* tree, const rgk_device_cfg* cfg /* NULL: defaults */,
                                 rgk_host_scene** out);
void rgk_host_scene_destroy(rgk_host_scene* hs);
const char* rgk_host_last_error(void);
rgk_status rgk_host_scene_get_info(const rgk_host_scene* hs, rgk_scene_info* out);
rgk_status rgk_host_scene_get_kdtree(const rgk_host_scene* hs, uint32_t* nodes, uint32_t* refs);
rgk_status rgk_host_scene_get_records(const rgk_host_scene* hs, float* planes, float* records);
/* The per-triangle bounds of the device's pre-filter (4 floats: lo1 | axis code in the two low mantissa bits, hi1, lo2,
 * hi2, in the triangle's projection plane), for checking their conservativeness without a GPU. */
rgk_status rgk_host_scene_get_bounds(const rgk_host_scene* hs, float* bounds);

/* The wide BVH (rgk_device_cfg::traversal == RGK_TRAVERSAL_BVH, the default; no reference counterpart -- RGKrt only has the kd-tree,
 * src/scene.cpp:294-429).  A candidate generator: the traversal entry points find the globally closest hit through it with
 * the same Triangle::TestIntersection arithmetic, and every ray whose answer could depend on the kd-tree's per-leaf
 * +-epsilon accept rule (more than one hit within 2 epsilon of the closest, or a hit within epsilon of the ray's ends) is
 * re-traced through the kd-tree, which stays the authority.  Sizes are 0 when it is off, when the tree came out too deep for the
 * traversal stack, or when the scene has a triangle whose exact test can produce NaN barycentrics (rgk_scene_info-independent;
 * such scenes stay on the kd-tree, still on the GPU).
 * nodes: 32 floats each -- lo.x[4] hi.x[4] lo.y[4] hi.y[4] lo.z[4] hi.z[4], 4 child codes (uint32 bits: inner = node
 * index, leaf = 1<<31 | (count-1)<<29 | first slot, empty = 0x7fffffff), 4 zeros; order[slot] = triangle. */
rgk_status rgk_host_scene_get_bvh_size(const rgk_host_scene* hs, uint32_t* n_nodes, uint32_t* n_slots, uint32_t* depth);
rgk_status rgk_host_scene_get_bvh(const rgk_host_scene* hs, float* nodes, uint32_t* order);
/* Counters of the wide-BVH traversal launches since the previous call (synchronises the stream), out[0..4] for the
 * closest-hit launches and out[5..9] for the any-hit (shadow) launches: rays, ambiguous rays handed to the kd-tree, wide
 * nodes visited, exact triangle tests, leaf slots scanned (the last three only while rgk_render_set_counting is on: the
 * algorithmic bytes per BVH ray are 36 in + 20 (1) out + 128 * nodes + 20 * slots + 32 * tests).  All zero when the BVH is off. */
rgk_status rgk_bvh_stats(rgk_context* ctx, uint64_t out[10]);

/* Replaces Scene::FindIntersectKdOtherThan (src/scene_intersect.cpp:211-327) over a
 * batch; ignore[i] = triangle index to skip or RGK_NO_TRIANGLE (then it is
 * Scene::FindIntersectKd, :4-116); ignore may be NULL.  Host buffers. */
rgk_status rgk_trace_closest(rgk_context* ctx, const rgk_ray* rays, const uint32_t* ignore,
                             uint64_t n, rgk_hit* hits, rgk_trav_stats* stats);
/* Replaces Scene::Visibility(a,b) (src/scene.cpp:670-673): visible[i] = 1 iff no
 * triangle is hit on the segment a_i -> b_i shortened by 20*epsilon at both ends. */
rgk_status rgk_trace_shadow(rgk_context* ctx, const float* a, const float* b,
                            uint64_t n, uint8_t* visible, rgk_trav_stats* stats);
/* Same two queries on device-resident buffers (bench / pipelines); asynchronous on
 * the context stream.  d_stats may be NULL. */
rgk_status rgk_trace_closest_device(rgk_context* ctx, const rgk_ray* d_rays, const uint32_t* d_ignore,
                                    uint64_t n, rgk_hit* d_hits, rgk_trav_stats* d_stats);
rgk_status rgk_trace_shadow_device(rgk_context* ctx, const float* d_a, const float* d_b,
                                   uint64_t n, uint8_t* d_visible, rgk_trav_stats* d_stats);

/* Replaces Camera::Camera (src/camera.cpp:7-24). */
void rgk_camera_init(rgk_camera* cam, const float pos[3], const float lookat[3], const float up[3],
                     float yview, float xview, int32_t xres, int32_t yres, float focus_plane, float lens_size);
/* Replaces Camera::GetPixelRay / GetPixelRayLens (src/camera.cpp:32-46) for a batch of
 * (x, y, subpixel offset, lens sample); lens may be NULL when lens_size == 0. Host buffers. */
rgk_status rgk_camera_rays(rgk_context* ctx, const rgk_camera* cam, uint32_t xres, uint32_t yres,
                           const int32_t* xy, const float* offsets, const float* lens,
                           uint64_t n, rgk_ray* rays);

/* Replaces GenerateTaskList (src/render_driver.cpp:30-46): tile_size x tile_size tiles
 * sorted by distance of their midpoint to the image centre.  Returns the count;
 * writes at most capacity tasks. */
uint32_t rgk_generate_tasks(uint32_t tile_size, uint32_t xres, uint32_t yres, rgk_task* out, uint32_t capacity);

/* Replaces StratifiedSampler(seed, 64, multisample) table construction
 * (src/sampler.cpp:85-116) for a batch of seeds, on the device, for tests:
 * out1d[seed][dim][set] (n1d dims), out2d[seed][dim][set][2] (n2d dims),
 * set_size = rgk_sampler_set_size(multisample). */
uint32_t rgk_sampler_set_size(uint32_t multisample);
rgk_status rgk_sampler_tables(rgk_context* ctx, const uint32_t* seeds, uint32_t n_seeds, uint32_t multisample,
                              uint32_t n1d, uint32_t n2d, float* out1d, float* out2d);

/* Replaces RenderDriver::RenderRound (src/render_driver.cpp:144-190): every task i is
 * rendered by a PathTracer seeded seedstart + seedcount_base + i (:160,173), pixels
 * y-major/x-minor with the per-pixel seed bump (src/tracer.cpp:8-9,
 * src/path_tracer.cpp:47), and (sum of samples, multisample) is added to
 * rgb_sum[(y*xres+x)*3+c] / count[y*xres+x] (src/tracer.cpp:18,
 * src/texture.cpp:342-348).  Host framebuffer; rgb_sum/count are read-modify-written. */
rgk_status rgk_render_round(rgk_context* ctx, const rgk_camera* cam, const rgk_render_params* params,
                            const rgk_task* tasks, uint32_t n_tasks, uint32_t seedstart, uint32_t seedcount_base,
                            float* rgb_sum, uint32_t* count, rgk_round_stats* stats);
/* Same with a device-resident framebuffer (multi-GPU reduce, bench); asynchronous. */
rgk_status rgk_render_round_device(rgk_context* ctx, const rgk_camera* cam, const rgk_render_params* params,
                                   const rgk_task* tasks, uint32_t n_tasks, uint32_t seedstart, uint32_t seedcount_base,
                                   float* d_rgb_sum, uint32_t* d_count, rgk_round_stats* stats);
/* Replaces RenderDriver::RenderFrame's Rounds loop (src/render_driver.cpp:192-253):
 * task list, seedstart = 42, seedcount running across rounds; framebuffer zeroed
 * first; (sum,count) returned un-normalised. */
rgk_status rgk_render_frame(rgk_context* ctx, const rgk_camera* cam, const rgk_render_params* params,
                            uint32_t rounds, float* rgb_sum, uint32_t* count, rgk_round_stats* stats);
/* RGK_SAMPLER_TABLES: the caller supplies the sample tables (what OfflineSampler holds per pixel, src/sampler.hpp:
 * 62-72) for the pixels of the following rgk_render_round* calls, in task order then y-major/x-minor:
 * t1d[pixel][dim][set], t2d[pixel][dim][set][2], set = 0 .. rgk_sampler_set_size(multisample)-1.  Needs at least
 * 1 + depth 1-D dims and 4 (+1 with a lens) + depth 2-D dims.  Copied to the device at once. */
rgk_status rgk_render_set_tables(rgk_context* ctx, uint32_t multisample, uint32_t n1d, uint32_t n2d,
                                 const float* t1d, const float* t2d, uint64_t n_pixels);

/* Replaces EXRTexture::Accumulate (src/texture.cpp:403-412) for device-resident buffers: d_rgb_sum[i] += d_other_sum[i]
 * for n_floats floats and, when both count pointers are given, d_count[j] += d_other_count[j] for n_floats / 3 pixels.
 * `stream` = the cudaStream_t to run on (NULL: the context stream) -- multi-GPU drivers run it on their communication
 * stream right after the NCCL reduce of a round, so that it overlaps the next round (rgk_b200/host/rgk_render_multi.cpp). */
rgk_status rgk_accumulate_device(rgk_context* ctx, float* d_rgb_sum, const float* d_other_sum, uint64_t n_floats,
                                 uint32_t* d_count, const uint32_t* d_other_count, void* stream);

/* Tile sharding across GPUs: of every task list given to rgk_render_round*, this context renders only
 * tasks first, first+stride, first+2*stride, ...; each keeps the seed of its position in the full list, so the union
 * over `stride` contexts equals the unsharded round pixel for pixel.  Default (0, 1) = everything. */
rgk_status rgk_render_set_shard(rgk_context* ctx, uint32_t first, uint32_t stride);

/* Traversal work counters of the rendering calls: when enabled, the closest-hit and shadow
 * kernels of rgk_render_round* count visited nodes / references / triangle tests (slower; off by default,
 * never on in a timed run) and rgk_render_get_trav_stats returns the totals of the last rendering call. */
rgk_status rgk_render_set_counting(rgk_context* ctx, int enabled);
rgk_status rgk_render_get_trav_stats(const rgk_context* ctx, rgk_trav_stats* closest, rgk_trav_stats* shadow);

/* Work counters of the shading kernel over the last rendering call made while counting was enabled (unidirectional mode):
 * out[0] surface vertices shaded (one GeneratePath iteration, src/path_tracer.cpp:122-302), out[1] sky vertices, out[2] image
 * texels the vertices needed (4 per bilinear fetch, 3 per bump-slope fetch), out[3] LTC lobe evaluations (BxDF::value and the
 * specular branch of BxDF::sample; mix materials not counted), out[4] light evaluations (NEE set-ups), out[5] continuation
 * rays produced, out[6..7] reserved.  The algorithmic bytes of one vertex are path-state read + written + 96
 * (vertex attributes) + 64 (material) + 16 * texels + 208 * LTC evaluations. */
rgk_status rgk_render_get_shade_stats(const rgk_context* ctx, uint64_t out[8]);

/* Parity probe (no reference counterpart): evaluates ONE device shading function on n caller-supplied inputs so
 * tests can compare it with the CPU path function by function.  Host buffers; float rows per item:
 *   RGK_PROBE_BXDF_SAMPLE  index = material  in Vi[3] uv[2] sample[2]          out dir[3] weight[3] may_leak   (BxDF::sample)
 *   RGK_PROBE_BXDF_VALUE   index = material  in Vi[3] Vr[3] uv[2]               out rgb                          (BxDF::value)
 *   RGK_PROBE_TEXTURE      index = texture   in uv[2]                           out rgb slope_right slope_bottom (src/texture.cpp:35-102)
 *   RGK_PROBE_RANDOM_LIGHT                   in choice[2] light tri_sample[2]   out type pos[3] colour[3] intensity size normal[3]
 *   RGK_PROBE_SKY                            in dir[3]                          out rgb                          (Scene::GetSkyboxRay)
 *   RGK_PROBE_FRAME                          in normal[3] v[3]                  out toLocal(v)[3] toGlobal(toLocal(v))[3] (SystemTransform) */
enum { RGK_PROBE_BXDF_SAMPLE = 0, RGK_PROBE_BXDF_VALUE = 1, RGK_PROBE_TEXTURE = 2, RGK_PROBE_RANDOM_LIGHT = 3, RGK_PROBE_SKY = 4,
       RGK_PROBE_FRAME = 5 };
rgk_status rgk_probe(rgk_context* ctx, uint32_t kind, uint32_t index, const float* in, uint64_t n, float* out);

/* Blocks until the context stream is idle. */
rgk_status rgk_synchronize(rgk_context* ctx);

#ifdef __cplusplus
}
#endif
#endif /* RGK_B200_H */
