// ref_harness.cpp -- C-ABI wrapper around the UNMODIFIED RGKrt sources.
//
// TEST INFRASTRUCTURE.  This file is compiled together with the reference's own
// translation units (taken where they lie under /root/reference/src, see
// oracle/Makefile) into oracle/_ref/librgk_ref.so.  It feeds a scene pack
// (include/rgk_b200.h : rgk_scene_desc) into the reference through its own
// loader entry points (Scene::RegisterMaterial, Scene::LoadAiSceneMeshes with
// plain-struct stand-ins for the assimp types, Scene::AddPointLight,
// Scene::Commit) and then runs the reference's own hot path
// (Scene::FindIntersectKdOtherThan, Scene::Visibility, StratifiedSampler,
// Camera, RenderDriver::RenderRound -> PathTracer::Render) on caller buffers.
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load
// the resulting library.  GLM itself is absent from the container: the reference
// is compiled against oracle/ref_build/shim/glm (GLM's published formulas).
//
// `private`/`protected` are re-defined for the reference headers below so that
// the flattened kd-tree (Scene::compressed_array, src/scene.hpp:140-145), the
// framebuffer vectors (EXRTexture::data/count, src/texture.hpp:112-115) and
// RenderDriver::RenderRound (src/render_driver.hpp:22) can be reached without
// editing the reference.  GCC's layout does not depend on access specifiers.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <vector>
#include <assimp/scene.h>
#include "json/json.h"

#define private public
#define protected public
#include "scene.hpp"
#include "camera.hpp"
#include "sampler.hpp"
#include "path_tracer.hpp"
#include "render_driver.hpp"
#include "bxdf/bxdf.hpp"
#include "texture.hpp"
#include "out.hpp"
#undef private
#undef protected

#include "rgk_b200.h"
#include "../../integration/gpu_bridge.hpp"   // the RGKrt-side binding, compiled here against the reference's own headers

// src/render_driver.cpp:30 (free function, not declared in a header)
std::vector<RenderTask> GenerateTaskList(unsigned int tile_size, unsigned int xres, unsigned int yres, glm::vec2 middle);

namespace {

struct HarnessConfig : public Config {
    Camera GetCamera(float) const override { throw std::runtime_error("unused"); }
    void InstallLights(Scene&) const override {}
    void InstallScene(Scene&) const override {}
    void InstallMaterials(Scene&) const override {}
    void InstallSky(Scene&) const override {}
    void PerformPostCheck() const override {}
};

struct RefScene {
    Scene scene;
    std::vector<std::shared_ptr<ReadableTexture>> textures;
    std::vector<std::shared_ptr<Material>> materials;
};

std::shared_ptr<ReadableTexture> get_tex(RefScene& rs, int32_t id) {
    if (id < 0) return std::make_shared<EmptyTexture>();
    return rs.textures.at(id);
}

Camera make_camera(const rgk_camera* c) {
    Camera cam(glm::vec3(0, 0, 0), glm::vec3(0, 0, -1), glm::vec3(0, 1, 0), 1.0f, 1.0f, 1, 1);
    auto v = [](const float* p) { return glm::vec3(p[0], p[1], p[2]); };
    cam.origin = v(c->origin); cam.lookat = v(c->lookat); cam.direction = v(c->direction);
    cam.cameraup = v(c->cameraup); cam.cameraleft = v(c->cameraleft);
    cam.viewscreen = v(c->viewscreen); cam.viewscreen_x = v(c->viewscreen_x); cam.viewscreen_y = v(c->viewscreen_y);
    cam.lens_size = c->lens_size; cam.xsize = c->xsize; cam.ysize = c->ysize;
    return cam;
}

void put3(float* d, const glm::vec3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

} // namespace

extern "C" {

const char* rgkref_describe(void) {
    return "RGKrt reference sources (unmodified, /root/reference/src) + GLM-formula shim; libstdc++ <random>";
}

void* rgkref_scene_create(const rgk_scene_desc* d) {
    out::verbosity_level = 0;
    RefScene* rs = new RefScene();
    try {
        Scene& s = rs->scene;
        // Textures: FileTexture(int,int)+SetPixel (src/texture.hpp:35,40) / Scene::CreateSolidTexture
        for (uint32_t i = 0; i < d->n_textures; i++) {
            const rgk_texture& t = d->textures[i];
            if (t.kind == 0) {
                rs->textures.push_back(s.CreateSolidTexture(Color(t.color[0], t.color[1], t.color[2])));
            } else {
                auto ft = std::make_shared<FileTexture>((int)t.width, (int)t.height);
                for (uint32_t y = 0; y < t.height; y++)
                    for (uint32_t x = 0; x < t.width; x++) {
                        const float* p = t.texels + 3 * ((size_t)y * t.width + x);
                        ft->SetPixel(x, y, Color(p[0], p[1], p[2]));
                    }
                rs->textures.push_back(ft);
            }
        }
        // Materials: public fields of Material / BxDF* (src/bxdf/bxdf.hpp:19-159)
        for (uint32_t i = 0; i < d->n_materials; i++) {
            const rgk_material& m = d->materials[i];
            auto mat = std::make_shared<Material>();
            mat->name = "m" + std::to_string(i);
            mat->emission = Radiance(m.emission[0], m.emission[1], m.emission[2]);
            mat->no_russian = m.no_russian != 0;
            if (m.tex_bump >= 0) mat->bumpmap = get_tex(*rs, m.tex_bump);
            switch (m.bxdf) {
            case RGK_BXDF_DIFFUSE: {
                auto b = std::make_unique<BxDFDiffuse>(); b->diffuse = get_tex(*rs, m.tex_diffuse);
                mat->bxdf = std::move(b); break; }
            case RGK_BXDF_MIX: {
                auto b = std::make_unique<BxDFMix>();
                b->m1 = rs->materials.at(m.mix_a); b->m2 = rs->materials.at(m.mix_b); b->amt1 = m.amount;
                mat->bxdf = std::move(b); break; }
            case RGK_BXDF_DIELECTRIC: {
                auto b = std::make_unique<BxDFDielectric>(); b->ior = m.ior; b->color = get_tex(*rs, m.tex_color);
                mat->bxdf = std::move(b); break; }
            case RGK_BXDF_MIRROR: {
                auto b = std::make_unique<BxDFMirror>(); b->color = get_tex(*rs, m.tex_color);
                mat->bxdf = std::move(b); break; }
            case RGK_BXDF_TRANSPARENT: {
                mat->bxdf = std::make_unique<BxDFTransparent>(); break; }
            case RGK_BXDF_LTC_BECKMANN: {
                auto b = std::make_unique<BxDFLTC<LTC::Beckmann>>(); b->roughness = m.roughness; b->color = get_tex(*rs, m.tex_color);
                mat->bxdf = std::move(b); break; }
            case RGK_BXDF_LTC_GGX: {
                auto b = std::make_unique<BxDFLTC<LTC::GGX>>(); b->roughness = m.roughness; b->color = get_tex(*rs, m.tex_color);
                mat->bxdf = std::move(b); break; }
            case RGK_BXDF_LTC_BECKMANN_DIFFUSE: {
                auto b = std::make_unique<BxDFLTCDiffuse<LTC::Beckmann>>(); b->roughness = m.roughness;
                b->color = get_tex(*rs, m.tex_color); b->diffuse = get_tex(*rs, m.tex_diffuse);
                mat->bxdf = std::move(b); break; }
            case RGK_BXDF_LTC_GGX_DIFFUSE: {
                auto b = std::make_unique<BxDFLTCDiffuse<LTC::GGX>>(); b->roughness = m.roughness;
                b->color = get_tex(*rs, m.tex_color); b->diffuse = get_tex(*rs, m.tex_diffuse);
                mat->bxdf = std::move(b); break; }
            default: throw std::runtime_error("bad bxdf kind");
            }
            rs->materials.push_back(mat);
            s.RegisterMaterial(mat, true);
        }
        // Geometry: one aiMesh holding every vertex, per rgk_mesh one aiMesh would
        // re-base indices; instead each mesh gets the full vertex arrays only once:
        // mesh 0 carries all vertices, later meshes carry none and index with a
        // negative offset is impossible -> so build one aiMesh per rgk_mesh with the
        // vertex sub-range it uses (indices inside a mesh must be contiguous-range).
        std::vector<aiMaterial> aimats(d->n_materials);
        std::vector<aiMaterial*> aimatp(d->n_materials);
        for (uint32_t i = 0; i < d->n_materials; i++) { aimats[i].name = "m" + std::to_string(i); aimatp[i] = &aimats[i]; }
        std::vector<aiMesh> meshes(d->n_meshes);
        std::vector<aiMesh*> meshp(d->n_meshes);
        std::vector<std::vector<aiVector3D>> vp(d->n_meshes), vn(d->n_meshes), vt(d->n_meshes), vuv(d->n_meshes);
        std::vector<std::vector<aiFace>> faces(d->n_meshes);
        std::vector<std::vector<unsigned int>> fidx(d->n_meshes);
        uint32_t vcursor = 0;
        for (uint32_t mi = 0; mi < d->n_meshes; mi++) {
            const rgk_mesh& rm = d->meshes[mi];
            // vertex range used by this mesh
            uint32_t lo = 0xFFFFFFFFu, hi = 0;
            for (uint32_t t = 0; t < rm.n_triangles; t++)
                for (int k = 0; k < 3; k++) {
                    uint32_t v = d->indices[3 * (rm.first_triangle + t) + k];
                    lo = std::min(lo, v); hi = std::max(hi, v);
                }
            if (rm.n_triangles == 0) { lo = vcursor; hi = vcursor ? vcursor - 1 : 0; }
            if (lo != vcursor) throw std::runtime_error("scene pack: meshes must use consecutive, disjoint vertex ranges");
            uint32_t nv = (rm.n_triangles == 0) ? 0 : hi - lo + 1;
            vp[mi].resize(nv); vn[mi].resize(nv); vt[mi].resize(nv); vuv[mi].resize(nv);
            for (uint32_t v = 0; v < nv; v++) {
                const uint32_t g = lo + v;
                vp[mi][v] = aiVector3D{d->positions[3 * g], d->positions[3 * g + 1], d->positions[3 * g + 2]};
                vn[mi][v] = aiVector3D{d->normals[3 * g], d->normals[3 * g + 1], d->normals[3 * g + 2]};
                vt[mi][v] = aiVector3D{d->tangents[3 * g], d->tangents[3 * g + 1], d->tangents[3 * g + 2]};
                vuv[mi][v] = aiVector3D{d->texcoords[2 * g], d->texcoords[2 * g + 1], 0.0f};
            }
            fidx[mi].resize(3 * rm.n_triangles); faces[mi].resize(rm.n_triangles);
            for (uint32_t t = 0; t < rm.n_triangles; t++) {
                for (int k = 0; k < 3; k++) fidx[mi][3 * t + k] = d->indices[3 * (rm.first_triangle + t) + k] - lo;
                faces[mi][t].mNumIndices = 3; faces[mi][t].mIndices = &fidx[mi][3 * t];
            }
            aiMesh& am = meshes[mi];
            am.mNumVertices = nv; am.mNumFaces = rm.n_triangles; am.mMaterialIndex = rm.material;
            am.mVertices = vp[mi].data(); am.mNormals = vn[mi].data(); am.mTangents = vt[mi].data();
            am.mTextureCoords[0] = vuv[mi].data(); am.mFaces = faces[mi].data();
            meshp[mi] = &am;
            vcursor += nv;
        }
        if (vcursor != d->n_vertices) throw std::runtime_error("scene pack: unused trailing vertices");
        std::vector<unsigned int> node_meshes(d->n_meshes);
        for (uint32_t i = 0; i < d->n_meshes; i++) node_meshes[i] = i;
        aiNode root; root.mNumMeshes = d->n_meshes; root.mMeshes = node_meshes.data();
        aiScene ais; ais.mNumMeshes = d->n_meshes; ais.mMeshes = meshp.data();
        ais.mNumMaterials = d->n_materials; ais.mMaterials = aimatp.data(); ais.mRootNode = &root;
        s.LoadAiSceneMeshes(&ais, glm::mat4(), "");  // src/scene.cpp:67 (identity transform: exact pass-through)
        // Lights (src/config.cpp:372-387)
        for (uint32_t i = 0; i < d->n_point_lights; i++) {
            const rgk_point_light& pl = d->point_lights[i];
            Light l(Light::Type::FULL_SPHERE);
            l.pos = glm::vec3(pl.position[0], pl.position[1], pl.position[2]);
            l.color = Radiance(pl.color[0], pl.color[1], pl.color[2]);
            l.intensity = pl.intensity; l.size = pl.size;
            l.normal = glm::vec3(0, 0, 0);
            s.AddPointLight(l);
        }
        // Sky (src/scene.hpp:122-132)
        if (d->sky.mode == 0) {
            s.SetSkyboxColor(Color(d->sky.color[0], d->sky.color[1], d->sky.color[2]), d->sky.intensity);
        } else {
            s.skybox_mode = Scene::Envmap;
            s.skybox_texture = rs->textures.at(d->sky.envmap);
            s.skybox_intensity = d->sky.intensity;
            s.skybox_rotate = d->sky.rotate;
        }
        if (d->thinglass && !rs->materials.empty()) s.thinglass.insert(rs->materials[0]);
        s.Commit();  // src/scene.cpp:294
    } catch (const std::exception& e) {
        std::cerr << "rgkref_scene_create: " << e.what() << std::endl;
        delete rs;
        return nullptr;
    }
    return rs;
}

void rgkref_scene_destroy(void* h) { delete (RefScene*)h; }

static void tree_depth(const CompressedKdNode* arr, uint32_t i, uint32_t d, uint32_t& mx) {
    mx = std::max(mx, d);
    if (arr[i].IsLeaf()) return;
    tree_depth(arr, i + 1, d + 1, mx);
    tree_depth(arr, arr[i].GetOtherChildIndex(), d + 1, mx);
}

int rgkref_scene_get_info(void* h, rgk_scene_info* o) {
    Scene& s = ((RefScene*)h)->scene;
    o->epsilon = s.epsilon;
    o->bbox[0] = s.xBB.first; o->bbox[1] = s.xBB.second; o->bbox[2] = s.yBB.first;
    o->bbox[3] = s.yBB.second; o->bbox[4] = s.zBB.first; o->bbox[5] = s.zBB.second;
    o->n_nodes = s.compressed_array_size; o->n_refs = s.compressed_triangles_size;
    o->n_triangles = s.n_triangles; o->n_areal_lights = s.areal_lights.size();
    o->max_depth = 0;
    if (s.compressed_array_size) tree_depth(s.compressed_array, 0, 0, o->max_depth);
    o->total_point_power = s.total_point_power; o->total_areal_power = s.total_areal_power;
    return 0;
}

int rgkref_scene_get_kdtree(void* h, uint32_t* nodes, uint32_t* refs) {
    Scene& s = ((RefScene*)h)->scene;
    static_assert(sizeof(CompressedKdNode) == 8, "CompressedKdNode is 8 bytes");
    std::memcpy(nodes, s.compressed_array, 8ull * s.compressed_array_size);
    std::memcpy(refs, s.compressed_triangles, 4ull * s.compressed_triangles_size);
    return 0;
}

// Triangle planes as computed by Triangle::CalculatePlane (src/primitives.cpp:24-36)
int rgkref_scene_get_planes(void* h, float* planes) {
    Scene& s = ((RefScene*)h)->scene;
    for (uint32_t i = 0; i < s.n_triangles; i++) {
        planes[4 * i] = s.triangles[i].p.x; planes[4 * i + 1] = s.triangles[i].p.y;
        planes[4 * i + 2] = s.triangles[i].p.z; planes[4 * i + 3] = s.triangles[i].p.w;
    }
    return 0;
}

static void run_parallel(uint64_t n, int nthreads, const std::function<void(uint64_t, uint64_t)>& f) {
    if (nthreads <= 1 || n < 1024) { f(0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) {
        uint64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
        th.emplace_back([=, &f] { f(lo, hi); });
    }
    for (auto& t : th) t.join();
}

int rgkref_trace_closest(void* h, const rgk_ray* rays, const uint32_t* ignore, uint64_t n, rgk_hit* hits, int nthreads) {
    const Scene& s = ((RefScene*)h)->scene;
    run_parallel(n, nthreads, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; i++) {
            Ray r;
            r.origin = glm::vec3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]);
            r.direction = glm::vec3(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]);
            r.near = rays[i].tnear; r.far = rays[i].tfar;
            const Triangle* ign = (ignore && ignore[i] != RGK_NO_TRIANGLE) ? &s.triangles[ignore[i]] : nullptr;
            Intersection is = s.thinglass.size() == 0 ? s.FindIntersectKdOtherThan(r, ign)
                                                      : s.FindIntersectKdOtherThanWithThinglass(r, ign);
            if (is.triangle) {
                hits[i].triangle = (uint32_t)(is.triangle - s.triangles);
                hits[i].t = is.t; hits[i].a = is.a; hits[i].b = is.b; hits[i].c = is.c;
            } else {
                hits[i].triangle = RGK_NO_TRIANGLE; hits[i].t = is.t; hits[i].a = hits[i].b = hits[i].c = 0.0f;
            }
        }
    });
    return 0;
}

int rgkref_trace_shadow(void* h, const float* a, const float* b, uint64_t n, uint8_t* visible, int nthreads) {
    const Scene& s = ((RefScene*)h)->scene;
    run_parallel(n, nthreads, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; i++) {
            const glm::vec3 pa(a[3 * i], a[3 * i + 1], a[3 * i + 2]), pb(b[3 * i], b[3 * i + 1], b[3 * i + 2]);
            if (s.thinglass.size() == 0) visible[i] = s.Visibility(pa, pb) ? 1 : 0;
            else { ThinglassIsections ti; visible[i] = s.VisibilityWithThinglass(pa, pb, ti) ? 1 : 0; }   // as src/path_tracer.cpp:431-432 selects
        }
    });
    return 0;
}

void rgkref_camera_init(rgk_camera* c, const float pos[3], const float la[3], const float up[3],
                        float yview, float xview, int32_t xres, int32_t yres, float focus_plane, float lens_size) {
    Camera cam(glm::vec3(pos[0], pos[1], pos[2]), glm::vec3(la[0], la[1], la[2]), glm::vec3(up[0], up[1], up[2]),
               yview, xview, xres, yres, focus_plane, lens_size);
    put3(c->origin, cam.origin); put3(c->lookat, cam.lookat); put3(c->direction, cam.direction);
    put3(c->cameraup, cam.cameraup); put3(c->cameraleft, cam.cameraleft);
    put3(c->viewscreen, cam.viewscreen); put3(c->viewscreen_x, cam.viewscreen_x); put3(c->viewscreen_y, cam.viewscreen_y);
    c->lens_size = cam.lens_size; c->xsize = cam.xsize; c->ysize = cam.ysize;
}

int rgkref_camera_rays(const rgk_camera* c, uint32_t xres, uint32_t yres, const int32_t* xy, const float* offsets,
                       const float* lens, uint64_t n, rgk_ray* rays) {
    Camera cam = make_camera(c);
    for (uint64_t i = 0; i < n; i++) {
        glm::vec2 off(offsets[2 * i], offsets[2 * i + 1]);
        Ray r = cam.IsSimple() ? cam.GetPixelRay(xy[2 * i], xy[2 * i + 1], xres, yres, off)
                               : cam.GetPixelRayLens(xy[2 * i], xy[2 * i + 1], xres, yres, off,
                                                     glm::vec2(lens[2 * i], lens[2 * i + 1]));
        put3(rays[i].origin, r.origin); put3(rays[i].direction, r.direction);
        rays[i].tnear = r.near; rays[i].tfar = r.far;
    }
    return 0;
}

uint32_t rgkref_generate_tasks(uint32_t tile, uint32_t xres, uint32_t yres, rgk_task* out, uint32_t cap) {
    std::vector<RenderTask> t = GenerateTaskList(tile, xres, yres, glm::vec2(xres / 2.0f, yres / 2.0f));
    for (uint32_t i = 0; i < t.size() && i < cap; i++)
        out[i] = rgk_task{t[i].xrange_start, t[i].xrange_end, t[i].yrange_start, t[i].yrange_end};
    return (uint32_t)t.size();
}

uint32_t rgkref_sampler_set_size(uint32_t ms) {
    StratifiedSampler s(1, 1, ms);
    return s.set_size;
}

// StratifiedSampler tables through the public Advance/Get1D/Get2D interface.
int rgkref_sampler_tables(const uint32_t* seeds, uint32_t n_seeds, uint32_t ms, uint32_t n1d, uint32_t n2d,
                          float* out1d, float* out2d) {
    for (uint32_t si = 0; si < n_seeds; si++) {
        StratifiedSampler s(seeds[si], 64, ms);
        const uint32_t set_size = s.set_size;
        for (uint32_t set = 0; set < set_size; set++) {
            s.Advance();
            for (uint32_t d = 0; d < n1d; d++) out1d[((size_t)si * n1d + d) * set_size + set] = s.Get1D();
            for (uint32_t d = 0; d < n2d; d++) {
                glm::vec2 v = s.Get2D();
                out2d[(((size_t)si * n2d + d) * set_size + set) * 2] = v.x;
                out2d[(((size_t)si * n2d + d) * set_size + set) * 2 + 1] = v.y;
            }
        }
    }
    return 0;
}

// RenderDriver::RenderRound (src/render_driver.cpp:144-190), the reference's own
// ctpl-threaded round, with `concurrency` worker threads.
int rgkref_render_round(void* h, const rgk_camera* c, const rgk_render_params* p, const rgk_task* tasks, uint32_t n_tasks,
                        uint32_t seedstart, uint32_t seedcount_base, float* rgb_sum, uint32_t* count,
                        rgk_round_stats* stats, int concurrency) {
    RefScene* rs = (RefScene*)h;
    Camera cam = make_camera(c);
    auto cfg = std::make_shared<HarnessConfig>();
    cfg->xres = p->xres; cfg->yres = p->yres; cfg->multisample = p->multisample; cfg->recursion_level = p->depth;
    cfg->clamp = p->clamp; cfg->russian = p->russian; cfg->bumpmap_scale = p->bumpmap_scale;
    cfg->force_fresnell = p->force_fresnell != 0; cfg->reverse = p->reverse;
    std::vector<RenderTask> tl;
    for (uint32_t i = 0; i < n_tasks; i++)
        tl.push_back(RenderTask(p->xres, p->yres, tasks[i].x1, tasks[i].x2, tasks[i].y1, tasks[i].y2));
    EXRTexture total(p->xres, p->yres);
    RenderDriver::ResetCounters();
    unsigned int seedcount = seedcount_base;
    auto t0 = std::chrono::high_resolution_clock::now();
    RenderDriver::RenderRound(rs->scene, cfg, cam, tl, seedcount, (int)seedstart, std::max(1, concurrency), total);
    auto t1 = std::chrono::high_resolution_clock::now();
    const size_t npx = (size_t)p->xres * p->yres;
    for (size_t i = 0; i < npx; i++) {
        rgb_sum[3 * i] += total.data[i].r; rgb_sum[3 * i + 1] += total.data[i].g; rgb_sum[3 * i + 2] += total.data[i].b;
        count[i] += total.count[i];
    }
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        stats->closest_rays = RenderDriver::rays_done.load();
        uint64_t px = 0;
        for (uint32_t i = 0; i < n_tasks; i++) px += (uint64_t)(tasks[i].x2 - tasks[i].x1) * (tasks[i].y2 - tasks[i].y1);
        stats->samples = px * p->multisample;
        stats->gpu_ms = std::chrono::duration<float, std::milli>(t1 - t0).count();  // wall-clock ms of the CPU round
    }
    return 0;
}

// ---- unit-level probes of the shading functions (for oracle/GPU unit parity) ----

// BxDF::sample (src/bxdf/bxdf.hpp:41): out = dir[3], spectrum[3], may_leak
int rgkref_bxdf_sample(void* h, uint32_t material, const float* Vi, const float* uv, const float* sample,
                       uint64_t n, float* out) {
    RefScene* rs = (RefScene*)h;
    const Material& m = *rs->materials.at(material);
    for (uint64_t i = 0; i < n; i++) {
        glm::vec3 dir; Spectrum sp; bool leak;
        std::tie(dir, sp, leak) = m.bxdf->sample(glm::vec3(Vi[3 * i], Vi[3 * i + 1], Vi[3 * i + 2]),
                                                 glm::vec2(uv[2 * i], uv[2 * i + 1]),
                                                 glm::vec2(sample[2 * i], sample[2 * i + 1]), false);
        float* o = out + 7 * i;
        o[0] = dir.x; o[1] = dir.y; o[2] = dir.z; o[3] = sp.r; o[4] = sp.g; o[5] = sp.b; o[6] = leak ? 1.0f : 0.0f;
    }
    return 0;
}

// BxDF::value (src/bxdf/bxdf.hpp:40)
int rgkref_bxdf_value(void* h, uint32_t material, const float* Vi, const float* Vr, const float* uv, uint64_t n, float* out) {
    RefScene* rs = (RefScene*)h;
    const Material& m = *rs->materials.at(material);
    for (uint64_t i = 0; i < n; i++) {
        Spectrum sp = m.bxdf->value(glm::vec3(Vi[3 * i], Vi[3 * i + 1], Vi[3 * i + 2]),
                                    glm::vec3(Vr[3 * i], Vr[3 * i + 1], Vr[3 * i + 2]),
                                    glm::vec2(uv[2 * i], uv[2 * i + 1]), false);
        out[3 * i] = sp.r; out[3 * i + 1] = sp.g; out[3 * i + 2] = sp.b;
    }
    return 0;
}

// ReadableTexture::GetPixelInterpolated / GetSlopeRight / GetSlopeBottom (src/texture.cpp:35-102): out = rgb, right, bottom
int rgkref_texture_fetch(void* h, uint32_t tex, const float* uv, uint64_t n, float* out) {
    RefScene* rs = (RefScene*)h;
    const ReadableTexture& t = *rs->textures.at(tex);
    for (uint64_t i = 0; i < n; i++) {
        glm::vec2 p(uv[2 * i], uv[2 * i + 1]);
        Color c = t.GetPixelInterpolated(p);
        out[5 * i] = c.r; out[5 * i + 1] = c.g; out[5 * i + 2] = c.b;
        out[5 * i + 3] = t.GetSlopeRight(p); out[5 * i + 4] = t.GetSlopeBottom(p);
    }
    return 0;
}

// Scene::GetRandomLight (src/scene.cpp:686-745): in = choice[2], light_sample, tri_sample[2];
// out = type, pos[3], color[3], intensity, size, normal[3]  (12 floats)
int rgkref_random_light(void* h, const float* in, uint64_t n, float* out) {
    const Scene& s = ((RefScene*)h)->scene;
    for (uint64_t i = 0; i < n; i++) {
        const float* q = in + 5 * i;
        Light l = s.GetRandomLight(glm::vec2(q[0], q[1]), q[2], glm::vec2(q[3], q[4]), false);
        float* o = out + 12 * i;
        o[0] = (float)l.type; put3(o + 1, l.pos); o[4] = l.color.r; o[5] = l.color.g; o[6] = l.color.b;
        o[7] = l.intensity; o[8] = (l.type == Light::FULL_SPHERE) ? l.size : 0.0f;
        if (l.type == Light::HEMISPHERE) put3(o + 9, l.normal); else o[9] = o[10] = o[11] = 0.0f;
    }
    return 0;
}

// Scene::GetSkyboxRay (src/scene.cpp:748-763)
int rgkref_sky(void* h, const float* dir, uint64_t n, float* out) {
    const Scene& s = ((RefScene*)h)->scene;
    for (uint64_t i = 0; i < n; i++) {
        Radiance r = s.GetSkyboxRay(glm::vec3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]));
        out[3 * i] = r.r; out[3 * i + 1] = r.g; out[3 * i + 2] = r.b;
    }
    return 0;
}

// The LTC tables cast to float exactly as mat33::operator glm::mat3 does (src/LTC/ltc.hpp:6-9):
// which = 0 GGX, 1 Beckmann; M = 4096*9 floats in mat33::m order, amp = 4096 floats.
int rgkref_ltc_tables(int which, float* M, float* amp) {
    const LTCdef& l = which == 0 ? LTC::GGX : LTC::Beckmann;
    for (int i = 0; i < l.size * l.size; i++) {
        for (int k = 0; k < 9; k++) M[9 * i + k] = (float)l.tabM[i].m[k];
        amp[i] = l.tabAmplitude[i];
    }
    return l.size;
}


// The bridge's Upload / RenderRound call into librgk_b200.so; this checker library must not depend on the product, so
// the five entry points they use are satisfied by inert weak definitions here (the bridge is compiled in full, only
// Describe() is executed by the tests).
__attribute__((weak)) rgk_status rgk_context_create(int, void*, rgk_context**) { return RGK_ERR_NO_DEVICE; }
__attribute__((weak)) void rgk_context_destroy(rgk_context*) {}
__attribute__((weak)) const char* rgk_last_error(const rgk_context*) { return "checker build: librgk_b200 is not linked"; }
__attribute__((weak)) rgk_status rgk_scene_commit(rgk_context*, const rgk_scene_desc*, const rgk_kdtree*) { return RGK_ERR_NO_DEVICE; }
__attribute__((weak)) rgk_status rgk_render_round(rgk_context*, const rgk_camera*, const rgk_render_params*, const rgk_task*, uint32_t, uint32_t, uint32_t,
                                                  float*, uint32_t*, rgk_round_stats*) { return RGK_ERR_NO_DEVICE; }

// The integration bridge (integration/gpu_bridge.hpp) run on this reference Scene: returns a heap-allocated bridge whose
// `desc` is what a RGKrt build would hand to rgk_scene_commit.  Test-only.
void* rgkref_bridge_describe(void* h) {
    RefScene* rs = (RefScene*)h;
    RgkGpuBridge* b = new RgkGpuBridge();
    try { b->Describe(rs->scene); } catch (const std::exception& e) { std::cerr << "bridge: " << e.what() << std::endl; delete b; return nullptr; }
    return b;
}
const rgk_scene_desc* rgkref_bridge_desc(void* b) { return &((RgkGpuBridge*)b)->desc; }
void rgkref_bridge_destroy(void* b) { delete (RgkGpuBridge*)b; }

} // extern "C"
