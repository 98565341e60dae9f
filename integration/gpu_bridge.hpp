// gpu_bridge.hpp -- the binding a RGKrt maintainer adds to call librgk_b200, as real code.
//
// Include it from a RGKrt translation unit AFTER RGKrt's own headers (scene.hpp, bxdf/bxdf.hpp, texture.hpp, LTC/ltc.hpp,
// camera.hpp, render_driver.hpp, config.hpp) and "rgk_b200.h".  It reads a few members RGKrt keeps private
// (Scene::materials, Scene::skybox_*, FileTexture::data/xsize/ysize, SolidTexture::color, EXRTexture::data/count): in
// RGKrt add `friend struct RgkGpuBridge;` to those four classes (the test build of this file re-defines `private`
// instead, see oracle/ref_build/ref_harness.cpp).  Nothing in RGKrt is changed otherwise.
//
//   RgkGpuBridge bridge;                 // once, after scene.Commit() (src/main.cpp:217)
//   bridge.Upload(scene);                // Scene -> rgk_scene_desc -> rgk_scene_commit (the library rebuilds the same kd-tree)
//   bridge.RenderRound(scene, cfg, camera, tasks, seedcount, seedstart, total_ob);   // body of RenderDriver::RenderRound
//
// tests/test_oracle_vs_ref.py checks, in the container that has the reference, that Describe() turns the reference's
// own Scene back into exactly the scene pack it was loaded from and that the library's host commit of that description
// yields the reference's own flattened kd-tree.
#pragma once
#include <map>
#include <memory>
#include <stdexcept>
#include <vector>

struct RgkGpuBridge {
    // ---- everything rgk_scene_desc points into
    std::vector<float> positions, normals, tangents, texcoords;
    std::vector<uint32_t> indices;
    std::vector<rgk_mesh> meshes;
    std::vector<rgk_material> materials;
    std::vector<rgk_texture> textures;
    std::vector<std::vector<float>> texels;
    std::vector<rgk_point_light> lights;
    std::vector<float> ggx_M, ggx_amp, beck_M, beck_amp;
    rgk_scene_desc desc{};
    rgk_context* ctx = nullptr;

    ~RgkGpuBridge() { if (ctx) rgk_context_destroy(ctx); }

    // ReadableTexture -> index into `textures` (-1 for EmptyTexture / null)
    int32_t texture_index(const std::shared_ptr<ReadableTexture>& t, std::map<const ReadableTexture*, int32_t>& seen) {
        if (!t || t->Empty()) return -1;
        auto it = seen.find(t.get());
        if (it != seen.end()) return it->second;
        rgk_texture r{};
        texels.emplace_back();
        if (const FileTexture* f = dynamic_cast<const FileTexture*>(t.get())) {
            r.kind = 1; r.width = f->xsize; r.height = f->ysize;
            std::vector<float>& px = texels.back();
            px.resize(3 * (size_t)f->xsize * f->ysize);
            for (size_t i = 0; i < f->data.size(); i++) { px[3 * i] = f->data[i].r; px[3 * i + 1] = f->data[i].g; px[3 * i + 2] = f->data[i].b; }
        } else if (const SolidTexture* s = dynamic_cast<const SolidTexture*>(t.get())) {
            r.kind = 0; r.color[0] = s->color.r; r.color[1] = s->color.g; r.color[2] = s->color.b;
        } else throw std::runtime_error("unknown texture class");
        textures.push_back(r);
        return seen[t.get()] = (int32_t)textures.size() - 1;
    }

    // Scene (after Commit) -> rgk_scene_desc
    const rgk_scene_desc& Describe(const Scene& s) {
        positions.clear(); normals.clear(); tangents.clear(); texcoords.clear(); indices.clear(); meshes.clear();
        materials.clear(); textures.clear(); texels.clear(); lights.clear();
        for (unsigned i = 0; i < s.n_vertices; i++) {
            for (int k = 0; k < 3; k++) { positions.push_back(s.vertices[i][k]); normals.push_back(s.normals[i][k]); tangents.push_back(s.tangents[i][k]); }
            texcoords.push_back(s.texcoords[i].x); texcoords.push_back(s.texcoords[i].y);
        }
        // materials in registration order; a mix material's children were registered before it
        std::map<const Material*, int32_t> mat_index;
        std::map<const ReadableTexture*, int32_t> seen;
        for (const std::shared_ptr<Material>& mp : s.materials) mat_index[mp.get()] = (int32_t)mat_index.size();
        for (const std::shared_ptr<Material>& mp : s.materials) {
            const Material& m = *mp;
            rgk_material r{};
            r.no_russian = m.no_russian ? 1u : 0u;
            r.emission[0] = m.emission.r; r.emission[1] = m.emission.g; r.emission[2] = m.emission.b;
            r.mix_a = r.mix_b = r.tex_diffuse = r.tex_color = -1;
            r.tex_bump = texture_index(m.bumpmap, seen);
            const BxDF* b = m.bxdf.get();
            if (const BxDFDiffuse* d = dynamic_cast<const BxDFDiffuse*>(b)) { r.bxdf = RGK_BXDF_DIFFUSE; r.tex_diffuse = texture_index(d->diffuse, seen); }
            else if (const BxDFMix* x = dynamic_cast<const BxDFMix*>(b)) {
                r.bxdf = RGK_BXDF_MIX; r.amount = x->amt1; r.mix_a = mat_index.at(x->m1.get()); r.mix_b = mat_index.at(x->m2.get());
            }
            else if (const BxDFDielectric* e = dynamic_cast<const BxDFDielectric*>(b)) { r.bxdf = RGK_BXDF_DIELECTRIC; r.ior = e->ior; r.tex_color = texture_index(e->color, seen); }
            else if (const BxDFMirror* mi = dynamic_cast<const BxDFMirror*>(b)) { r.bxdf = RGK_BXDF_MIRROR; r.tex_color = texture_index(mi->color, seen); }
            else if (dynamic_cast<const BxDFTransparent*>(b)) r.bxdf = RGK_BXDF_TRANSPARENT;
            else if (const BxDFLTCDiffuseBase* ld = dynamic_cast<const BxDFLTCDiffuseBase*>(b)) {
                r.bxdf = dynamic_cast<const BxDFLTCDiffuse<LTC::GGX>*>(b) ? RGK_BXDF_LTC_GGX_DIFFUSE : RGK_BXDF_LTC_BECKMANN_DIFFUSE;
                r.roughness = ld->roughness; r.tex_color = texture_index(ld->color, seen); r.tex_diffuse = texture_index(ld->diffuse, seen);
            }
            else if (const BxDFLTCBase* l = dynamic_cast<const BxDFLTCBase*>(b)) {
                r.bxdf = dynamic_cast<const BxDFLTC<LTC::GGX>*>(b) ? RGK_BXDF_LTC_GGX : RGK_BXDF_LTC_BECKMANN;
                r.roughness = l->roughness; r.tex_color = texture_index(l->color, seen);
            }
            else throw std::runtime_error("material \"" + m.name + "\" has a BxDF this library does not know");
            materials.push_back(r);
        }
        // triangles; one rgk_mesh per run of equal (material, areal light): an emissive mesh is one ArealLight (src/scene.cpp:149-206)
        std::vector<int32_t> light_of(s.n_triangles, -1);
        for (size_t k = 0; k < s.areal_lights.size(); k++)
            for (const auto& ta : s.areal_lights[k].second.triangles_with_areas) light_of[ta.second] = (int32_t)k;
        for (unsigned i = 0; i < s.n_triangles; i++) {
            const Triangle& t = s.triangles[i];
            indices.push_back(t.va); indices.push_back(t.vb); indices.push_back(t.vc);
            const uint32_t mat = (uint32_t)mat_index.at(t.mat);
            if (meshes.empty() || meshes.back().material != mat || light_of[i] != light_of[i - 1]) meshes.push_back(rgk_mesh{i, 0, mat, 0});
            meshes.back().n_triangles++;
        }
        for (const Light& l : s.pointlights) {
            rgk_point_light p{};
            for (int k = 0; k < 3; k++) p.position[k] = l.pos[k];
            p.color[0] = l.color.r; p.color[1] = l.color.g; p.color[2] = l.color.b;
            p.intensity = l.intensity; p.size = l.size;
            lights.push_back(p);
        }
        rgk_sky sky{};
        if (s.skybox_mode == Scene::SimpleRadiance) {
            sky.mode = 0; sky.color[0] = s.skybox_color.r; sky.color[1] = s.skybox_color.g; sky.color[2] = s.skybox_color.b;
            sky.intensity = s.skybox_intensity; sky.envmap = -1;
        } else {
            sky.mode = 1; sky.intensity = s.skybox_intensity; sky.rotate = s.skybox_rotate; sky.envmap = texture_index(s.skybox_texture, seen);
        }
        for (size_t i = 0; i < textures.size(); i++) textures[i].texels = textures[i].kind == 1 ? texels[i].data() : nullptr;
        // LTC fits: the double tables cast to float as mat33::operator glm::mat3 does (src/LTC/ltc.hpp:6-9)
        auto tables = [](const LTCdef& def, std::vector<float>& M, std::vector<float>& A) {
            const int n = def.size * def.size;
            M.resize(9 * (size_t)n); A.resize(n);
            for (int i = 0; i < n; i++) { for (int k = 0; k < 9; k++) M[9 * i + k] = (float)def.tabM[i].m[k]; A[i] = def.tabAmplitude[i]; }
        };
        tables(LTC::GGX, ggx_M, ggx_amp); tables(LTC::Beckmann, beck_M, beck_amp);
        desc = rgk_scene_desc{};
        desc.n_vertices = s.n_vertices; desc.positions = positions.data(); desc.normals = normals.data(); desc.tangents = tangents.data(); desc.texcoords = texcoords.data();
        desc.n_triangles = s.n_triangles; desc.indices = indices.data();
        desc.n_meshes = (uint32_t)meshes.size(); desc.meshes = meshes.data();
        desc.n_materials = (uint32_t)materials.size(); desc.materials = materials.data();
        desc.n_textures = (uint32_t)textures.size(); desc.textures = textures.data();
        desc.n_point_lights = (uint32_t)lights.size(); desc.point_lights = lights.data();
        desc.sky = sky;
        desc.ltc_ggx.M = ggx_M.data(); desc.ltc_ggx.amplitude = ggx_amp.data();
        desc.ltc_beckmann.M = beck_M.data(); desc.ltc_beckmann.amplitude = beck_amp.data();
        desc.thinglass = s.thinglass.empty() ? 0u : 1u;
        return desc;
    }

    // after scene.Commit(): no CPU fallback -- without a CUDA device this throws
    void Upload(const Scene& s, int device = 0) {
        if (!ctx && rgk_context_create(device, nullptr, &ctx) != RGK_OK) throw std::runtime_error(rgk_last_error(nullptr));
        if (rgk_scene_commit(ctx, &Describe(s), nullptr) != RGK_OK) throw std::runtime_error(rgk_last_error(ctx));
    }

    // the body of RenderDriver::RenderRound (src/render_driver.cpp:154-189)
    rgk_round_stats RenderRound(std::shared_ptr<Config> cfg, const Camera& camera, const std::vector<RenderTask>& tasks,
                                unsigned int& seedcount, const int seedstart, EXRTexture& total_ob) {
        rgk_camera cam{};
        auto c3 = [](float* d, const glm::vec3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; };
        c3(cam.origin, camera.origin); c3(cam.lookat, camera.lookat); c3(cam.direction, camera.direction);
        c3(cam.cameraup, camera.cameraup); c3(cam.cameraleft, camera.cameraleft);
        c3(cam.viewscreen, camera.viewscreen); c3(cam.viewscreen_x, camera.viewscreen_x); c3(cam.viewscreen_y, camera.viewscreen_y);
        cam.lens_size = camera.lens_size; cam.xsize = camera.xsize; cam.ysize = camera.ysize;
        rgk_render_params p{};
        p.xres = cfg->xres; p.yres = cfg->yres; p.multisample = cfg->multisample; p.depth = cfg->recursion_level; p.clamp = cfg->clamp;
        p.russian = cfg->russian; p.bumpmap_scale = cfg->bumpmap_scale; p.force_fresnell = cfg->force_fresnell ? 1u : 0u;
        p.reverse = cfg->reverse; p.sampler_mode = RGK_SAMPLER_MT19937;
        std::vector<rgk_task> t;
        for (const RenderTask& r : tasks) t.push_back(rgk_task{r.xrange_start, r.xrange_end, r.yrange_start, r.yrange_end});
        rgk_round_stats st{};
        // EXRTexture::data is vector<Radiance{r,g,b}> (12 bytes each) and ::count vector<unsigned>: the ABI's framebuffer
        if (rgk_render_round(ctx, &cam, &p, t.data(), (uint32_t)t.size(), (uint32_t)seedstart, seedcount,
                             &total_ob.data[0].r, total_ob.count.data(), &st) != RGK_OK)
            throw std::runtime_error(rgk_last_error(ctx));
        seedcount += (unsigned int)tasks.size();      // what the per-task `seedcount++` did (src/render_driver.cpp:160)
        return st;
    }
};
