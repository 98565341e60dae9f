// rgk_b200_host.hpp -- C++ host layer above the C ABI (include/rgk_b200.h), header-only, C++17, no dependencies.
//
// It mirrors the part of RGKrt's host interface that surrounds the hot path, with the reference's names, argument
// meaning and error behaviour, so that code written against RGKrt's classes reads the same against this library:
//
//   rgkb::Camera        Camera::Camera                          src/camera.cpp:7-24, public fields src/camera.hpp:27-41
//   rgkb::RenderTask    RenderTask / GenerateTaskList            src/tracer.hpp:14-24, src/render_driver.cpp:30-46
//   rgkb::EXRTexture    EXRTexture (sum, count) accumulator      src/texture.hpp:83-118, src/texture.cpp:334-412
//   rgkb::Scene         Scene::Commit / FindIntersectKdOtherThan / Visibility   src/scene.hpp, src/scene.cpp:294-429,670-673
//   rgkb::PathTracer    PathTracer(scene, camera, xres, yres, multisample, depth, clamp, russian, bumpmap_scale,
//                       force_fresnell, reverse, samplerSeed) + Tracer::Render(task, output, pixel_count, ray_count)
//                                                                 src/path_tracer.hpp:10-21, src/tracer.cpp:6-37
//   rgkb::RenderDriver  RenderRound / RenderFrame (Rounds and Timed modes, progressive Normalize().Write())
//                                                                 src/render_driver.cpp:144-253
//   rgkb::PackFile      the on-disk scene pack (RGKPACK1) written by rgk_b200.scene.ScenePack.save: what the reference
//                       gets from assimp + its texture loaders, already decoded
//
// Differences, all forced by the device: a Scene is bound to one GPU context; RenderRound issues ONE library call per
// round instead of one PathTracer per tile on a thread pool (results are identical pixel for pixel, see
// rgk_render_round); errors are std::runtime_error carrying rgk_last_error (the reference throws at load time and
// aborts on asserts); there is no CPU fallback -- without a CUDA device the Scene constructor throws.
#ifndef RGK_B200_HOST_HPP
#define RGK_B200_HOST_HPP
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "rgk_b200.h"

namespace rgkb {

struct Radiance {                       // src/radiance.hpp: three floats, r g b
    float r = 0.0f, g = 0.0f, b = 0.0f;
};

// ---------------------------------------------------------------------------------------------- EXRTexture
class EXRTexture {
public:
    explicit EXRTexture(int xsize = 0, int ysize = 0) : xsize(xsize), ysize(ysize), data((size_t)xsize * ysize), count((size_t)xsize * ysize, 0u) {}
    void AddPixel(int x, int y, Radiance c, unsigned int n = 1) {          // src/texture.cpp:342-348
        Radiance& d = data[(size_t)y * xsize + x];
        d.r += c.r; d.g += c.g; d.b += c.b;
        count[(size_t)y * xsize + x] += n;
    }
    Radiance GetPixel(int x, int y) const {                                 // :349-354
        const size_t n = (size_t)y * xsize + x;
        if (count[n] == 0) return Radiance();
        const float c = (float)count[n];
        Radiance out; out.r = data[n].r / c; out.g = data[n].g / c; out.b = data[n].b / c;
        return out;
    }
    // A positive value scales the texture; a non-positive one picks 1 / (largest channel of any pixel)   (:375-401)
    EXRTexture Normalize(float val) const {
        EXRTexture out(xsize, ysize);
        out.data = data; out.count = count;
        if (val <= 0.0f) {
            float m = 0.0f;
            for (int y = 0; y < ysize; y++)
                for (int x = 0; x < xsize; x++) {
                    const Radiance q = GetPixel(x, y);
                    m = std::max(m, q.r); m = std::max(m, q.g); m = std::max(m, q.b);
                }
            val = 1.0f / m;
        }
        for (Radiance& d : out.data) { d.r *= val; d.g *= val; d.b *= val; }
        return out;
    }
    void Accumulate(const EXRTexture& other) {                              // :403-412
        if (xsize != other.xsize || ysize != other.ysize) throw std::runtime_error("EXRTexture::Accumulate: size mismatch");
        for (size_t i = 0; i < data.size(); i++) {
            data[i].r += other.data[i].r; data[i].g += other.data[i].g; data[i].b += other.data[i].b;
            count[i] += other.count[i];
        }
    }
    // OpenEXR scanline file, RGBA half like Imf::RgbaOutputFile writes (:356-373), uncompressed.
    bool Write(const std::string& path) const;
    // Raw accumulators (sum, count) for checkpoint / resume: "RGKACC01", xsize, ysize, rounds.
    bool WriteRaw(const std::string& path, uint32_t rounds_done) const;
    static EXRTexture ReadRaw(const std::string& path, uint32_t* rounds_done);

    int XSize() const { return xsize; }
    int YSize() const { return ysize; }
    // the library's framebuffer layout is exactly this object's: rgb_sum[(y*xres+x)*3+c], count[y*xres+x]
    float* sum_ptr() { return data.empty() ? nullptr : &data[0].r; }
    uint32_t* count_ptr() { return count.data(); }
    const std::vector<Radiance>& Data() const { return data; }
    const std::vector<uint32_t>& Count() const { return count; }

    static uint16_t FloatToHalf(float f);                                    // Imath half(float): round to nearest even
private:
    int xsize, ysize;
    std::vector<Radiance> data;
    std::vector<uint32_t> count;
};
static_assert(sizeof(Radiance) == 12, "Radiance must be three packed floats");

inline uint16_t EXRTexture::FloatToHalf(float f) {
    uint32_t x; std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const int32_t e = (int32_t)((x >> 23) & 0xffu) - 127 + 15;
    uint32_t m = x & 0x007fffffu;
    if (((x >> 23) & 0xffu) == 0xffu) return (uint16_t)(sign | 0x7c00u | (m ? (0x0200u | (m >> 13)) : 0u));   // inf / nan
    if (e >= 31) return (uint16_t)(sign | 0x7c00u);                                                                // overflow
    if (e <= 0) {                                                                                                   // subnormal / zero
        if (e < -10) return (uint16_t)sign;
        m |= 0x00800000u;
        const int shift = 14 - e;
        const uint32_t half = m >> shift, rem = m & ((1u << shift) - 1u), mid = 1u << (shift - 1);
        return (uint16_t)(sign | (half + ((rem > mid || (rem == mid && (half & 1u))) ? 1u : 0u)));
    }
    const uint32_t half = ((uint32_t)e << 10) | (m >> 13), rem = m & 0x1fffu;
    return (uint16_t)(sign | (half + ((rem > 0x1000u || (rem == 0x1000u && (half & 1u))) ? 1u : 0u)));   // carry into the exponent is right
}

inline bool EXRTexture::Write(const std::string& path) const {
    std::ofstream f(path, std::ios::binary | std::ios::trunc);
    if (!f) return false;
    std::string h;
    auto put = [&](const void* p, size_t n) { h.append((const char*)p, n); };
    auto str = [&](const char* s) { h.append(s, std::strlen(s) + 1); };
    auto i32 = [&](int32_t v) { put(&v, 4); };
    auto f32 = [&](float v) { put(&v, 4); };
    auto attr = [&](const char* name, const char* type, int32_t size) { str(name); str(type); i32(size); };
    const unsigned char magic[8] = {0x76, 0x2f, 0x31, 0x01, 2, 0, 0, 0};
    put(magic, 8);
    attr("channels", "chlist", 4 * 18 + 1);
    for (const char* ch : {"A", "B", "G", "R"}) { str(ch); i32(1 /*HALF*/); const unsigned char lin[4] = {0, 0, 0, 0}; put(lin, 4); i32(1); i32(1); }
    h.push_back('\0');
    attr("compression", "compression", 1); h.push_back('\0');
    attr("dataWindow", "box2i", 16); i32(0); i32(0); i32(xsize - 1); i32(ysize - 1);
    attr("displayWindow", "box2i", 16); i32(0); i32(0); i32(xsize - 1); i32(ysize - 1);
    attr("lineOrder", "lineOrder", 1); h.push_back('\0');
    attr("pixelAspectRatio", "float", 4); f32(1.0f);
    attr("screenWindowCenter", "v2f", 8); f32(0.0f); f32(0.0f);
    attr("screenWindowWidth", "float", 4); f32(1.0f);
    h.push_back('\0');
    const uint64_t line_bytes = 8 + (uint64_t)xsize * 4 * 2;
    uint64_t off = h.size() + (uint64_t)ysize * 8;
    for (int y = 0; y < ysize; y++) { put(&off, 8); off += line_bytes; }
    f.write(h.data(), (std::streamsize)h.size());
    std::vector<uint16_t> row((size_t)xsize * 4);
    for (int y = 0; y < ysize; y++) {
        for (int x = 0; x < xsize; x++) {
            const Radiance q = GetPixel(x, y);
            row[x] = FloatToHalf(1.0f); row[(size_t)xsize + x] = FloatToHalf(q.b);
            row[(size_t)2 * xsize + x] = FloatToHalf(q.g); row[(size_t)3 * xsize + x] = FloatToHalf(q.r);
        }
        const int32_t yy = y, sz = (int32_t)(row.size() * 2);
        f.write((const char*)&yy, 4); f.write((const char*)&sz, 4);
        f.write((const char*)row.data(), sz);
    }
    return (bool)f;
}

inline bool EXRTexture::WriteRaw(const std::string& path, uint32_t rounds_done) const {
    const std::string tmp = path + ".tmp";
    {
        std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
        if (!f) return false;
        const uint32_t hdr[3] = {(uint32_t)xsize, (uint32_t)ysize, rounds_done};
        f.write("RGKACC01", 8); f.write((const char*)hdr, 12);
        f.write((const char*)data.data(), (std::streamsize)(data.size() * sizeof(Radiance)));
        f.write((const char*)count.data(), (std::streamsize)(count.size() * 4));
        if (!f) return false;
    }
    return std::rename(tmp.c_str(), path.c_str()) == 0;      // a crash never leaves a half-written checkpoint
}

inline EXRTexture EXRTexture::ReadRaw(const std::string& path, uint32_t* rounds_done) {
    std::ifstream f(path, std::ios::binary);
    char magic[8]; uint32_t hdr[3];
    if (!f || !f.read(magic, 8) || std::memcmp(magic, "RGKACC01", 8) != 0 || !f.read((char*)hdr, 12))
        throw std::runtime_error("not an RGKACC01 accumulator file: " + path);
    EXRTexture t((int)hdr[0], (int)hdr[1]);
    if (!f.read((char*)t.data.data(), (std::streamsize)(t.data.size() * sizeof(Radiance))) ||
        !f.read((char*)t.count.data(), (std::streamsize)(t.count.size() * 4)))
        throw std::runtime_error("truncated accumulator file: " + path);
    if (rounds_done) *rounds_done = hdr[2];
    return t;
}

// ---------------------------------------------------------------------------------------------- RenderTask
struct RenderTask {                     // src/tracer.hpp:14-24
    unsigned int xres = 0, yres = 0;
    unsigned int xrange_start = 0, xrange_end = 0, yrange_start = 0, yrange_end = 0;
    float midpoint_x = 0.0f, midpoint_y = 0.0f;
};
// GenerateTaskList(tile_size, xres, yres, midpoint = image centre), src/render_driver.cpp:30-46
inline std::vector<RenderTask> GenerateTaskList(unsigned int tile_size, unsigned int xres, unsigned int yres) {
    const uint32_t cap = ((xres + tile_size - 1) / tile_size) * ((yres + tile_size - 1) / tile_size);
    std::vector<rgk_task> t(cap ? cap : 1);
    const uint32_t n = rgk_generate_tasks(tile_size, xres, yres, t.data(), cap);
    std::vector<RenderTask> out(n);
    for (uint32_t i = 0; i < n; i++) {
        RenderTask& r = out[i];
        r.xres = xres; r.yres = yres;
        r.xrange_start = t[i].x1; r.xrange_end = t[i].x2; r.yrange_start = t[i].y1; r.yrange_end = t[i].y2;
        r.midpoint_x = (t[i].x1 + t[i].x2) / 2.0f; r.midpoint_y = (t[i].y1 + t[i].y2) / 2.0f;
    }
    return out;
}

// ---------------------------------------------------------------------------------------------- Camera
class Camera : public rgk_camera {      // the nine public fields of src/camera.hpp:27-41, as the ABI struct
public:
    Camera() { std::memset(static_cast<rgk_camera*>(this), 0, sizeof(rgk_camera)); }
    Camera(const float pos[3], const float la[3], const float up[3], float yview, float xview, int xres, int yres,
           float focus_plane = 1.0f, float ls = 0.0f) {
        rgk_camera_init(this, pos, la, up, yview, xview, xres, yres, focus_plane, ls);
    }
    bool IsSimple() const { return lens_size == 0.0f; }                    // src/camera.hpp:22
};

// ---------------------------------------------------------------------------------------------- Config (render part)
enum class RenderLimitMode { Rounds, Timed };                                // src/config.hpp
struct Config {                         // the fields of src/config.hpp:25-56 that RenderDriver reads
    unsigned int xres = 0, yres = 0;
    unsigned int multisample = 1, recursion_level = 40;
    float clamp = 10000000.0f, russian = 0.74f, bumpmap_scale = 1.0f;
    bool force_fresnell = false;
    unsigned int reverse = 0;
    RenderLimitMode render_limit_mode = RenderLimitMode::Rounds;
    unsigned int render_rounds = 1;
    float render_minutes = 0.0f;
    float output_scale = -1.0f;
    uint32_t sampler_mode = RGK_SAMPLER_MT19937;
};

// ---------------------------------------------------------------------------------------------- Scene
struct Ray { float origin[3]; float direction[3]; float tnear = 0.0f; float tfar = 10000.0f; };   // src/ray.hpp
struct Intersection { uint32_t triangle = RGK_NO_TRIANGLE; float t = 0, a = 0, b = 0, c = 0; };   // src/primitives.hpp:98-110

class Scene {
public:
    explicit Scene(int device = 0, void* stream = nullptr) {
        const rgk_status s = rgk_context_create(device, stream, &ctx);
        if (s != RGK_OK) throw std::runtime_error(std::string("rgk_context_create: ") + rgk_status_string(s) + ": " + rgk_last_error(nullptr));
    }
    ~Scene() { if (ctx) rgk_context_destroy(ctx); }
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;
    // Scene::Commit (src/scene.cpp:294-429): planes, areal lights, epsilon, bbox, kd-tree; tree != nullptr installs a
    // given flattened tree instead.  Throws like the reference's loader does on inconsistent input.
    // traversal: RGK_TRAVERSAL_BVH (default: 4-wide BVH candidate pass, the kd-tree as arbiter) or RGK_TRAVERSAL_KD (the
    // reference's kd-tree for every ray); same results either way
    void Commit(const rgk_scene_desc& desc, const rgk_kdtree* tree = nullptr, uint32_t traversal = RGK_TRAVERSAL_BVH) {
        rgk_device_cfg cfg;
        check(rgk_context_get_cfg(ctx, &cfg), "rgk_context_get_cfg");
        if (cfg.traversal != traversal) { cfg.traversal = traversal; Configure(cfg); }
        check(rgk_scene_commit(ctx, &desc, tree), "rgk_scene_commit");
        check(rgk_scene_get_info(ctx, &info), "rgk_scene_get_info");
    }
    // rgk_device_cfg: scheduling / memory sizing of the device path (before Commit for the traversal and build fields)
    void Configure(const rgk_device_cfg& cfg) { check(rgk_context_configure(ctx, &cfg), "rgk_context_configure"); }
    // Scene::FindIntersectKdOtherThan (src/scene_intersect.cpp:211-327); ignored = RGK_NO_TRIANGLE: FindIntersectKd
    Intersection FindIntersectKdOtherThan(const Ray& r, uint32_t ignored = RGK_NO_TRIANGLE) const {
        rgk_ray rr; std::memcpy(rr.origin, r.origin, 12); std::memcpy(rr.direction, r.direction, 12); rr.tnear = r.tnear; rr.tfar = r.tfar;
        rgk_hit h;
        check(rgk_trace_closest(ctx, &rr, &ignored, 1, &h, nullptr), "rgk_trace_closest");
        Intersection i; i.triangle = h.triangle; i.t = h.t; i.a = h.a; i.b = h.b; i.c = h.c;
        return i;
    }
    bool Visibility(const float a[3], const float b[3]) const {              // src/scene.cpp:670-673
        uint8_t v = 0;
        check(rgk_trace_shadow(ctx, a, b, 1, &v, nullptr), "rgk_trace_shadow");
        return v != 0;
    }
    rgk_context* context() const { return ctx; }
    rgk_scene_info info{};
    float epsilon() const { return info.epsilon; }
    void check(rgk_status s, const char* what) const {
        if (s != RGK_OK) throw std::runtime_error(std::string(what) + ": " + rgk_status_string(s) + ": " + rgk_last_error(ctx));
    }
private:
    rgk_context* ctx = nullptr;
};

// ---------------------------------------------------------------------------------------------- PathTracer
class Tracer {                          // src/tracer.hpp:27-63
public:
    Tracer(const Scene& scene, const Camera& camera, unsigned int xres, unsigned int yres, unsigned int multisample, float bumpmap_scale = 10.0f)
        : scene(scene), camera(camera), xres(xres), yres(yres), multisample(multisample), bumpmap_scale(bumpmap_scale) {}
    virtual ~Tracer() {}
    // Renders the tile; adds (sum of the pixel's samples, multisample) to the output buffer (src/tracer.cpp:6-37)
    virtual void Render(const RenderTask& task, EXRTexture* output, std::atomic<int>& pixel_count, std::atomic<unsigned int>& ray_count) = 0;
protected:
    const Scene& scene;
    const Camera& camera;
    unsigned int xres, yres, multisample;
    float bumpmap_scale;
};

class PathTracer : public Tracer {      // src/path_tracer.hpp:10-21
public:
    PathTracer(const Scene& scene, const Camera& camera, unsigned int xres, unsigned int yres, unsigned int multisample,
               unsigned int depth, float clamp, float russian, float bumpmap_scale, bool force_fresnell, unsigned int reverse,
               unsigned int samplerSeed, uint32_t sampler_mode = RGK_SAMPLER_MT19937)
        : Tracer(scene, camera, xres, yres, multisample, bumpmap_scale), samplerSeed(samplerSeed) {
        params.xres = xres; params.yres = yres; params.multisample = multisample; params.depth = depth; params.clamp = clamp;
        params.russian = russian; params.bumpmap_scale = bumpmap_scale; params.force_fresnell = force_fresnell ? 1u : 0u;
        params.reverse = reverse; params.sampler_mode = sampler_mode;
    }
    void Render(const RenderTask& task, EXRTexture* output, std::atomic<int>& pixel_count, std::atomic<unsigned int>& ray_count) override {
        const rgk_task t = {task.xrange_start, task.xrange_end, task.yrange_start, task.yrange_end};
        rgk_round_stats st;
        // one tile: the tracer's own seed is the tile seed (seedstart + c of src/render_driver.cpp:160,173)
        scene.check(rgk_render_round(scene.context(), &camera, &params, &t, 1, samplerSeed, 0, output->sum_ptr(), output->count_ptr(), &st), "rgk_render_round");
        pixel_count += (int)((task.xrange_end - task.xrange_start) * (task.yrange_end - task.yrange_start));
        ray_count += (unsigned int)st.closest_rays;
        last_stats = st;
    }
    rgk_round_stats last_stats{};
private:
    rgk_render_params params{};
    unsigned int samplerSeed;
};

// ---------------------------------------------------------------------------------------------- RenderDriver
class RenderDriver {
public:
    static constexpr unsigned int TILE_SIZE = 32;                            // src/render_driver.hpp
    // counters of the reference's monitor thread (src/render_driver.cpp:18-28), per driver instead of global
    std::atomic<int> pixels_done{0};
    std::atomic<unsigned int> rays_done{0};
    std::atomic<int> rounds_done{0};
    uint64_t shadow_rays_done = 0, samples_done = 0;
    double gpu_ms = 0.0;

    // RenderDriver::RenderRound (src/render_driver.cpp:144-190): task i gets the seed seedstart + seedcount + i; the
    // whole round is one library call (concurrency is the device's business).
    void RenderRound(const Scene& scene, const Config& cfg, const Camera& camera, const std::vector<RenderTask>& tasks,
                     unsigned int& seedcount, const int seedstart, unsigned int /*concurrency*/, EXRTexture& total_ob) {
        std::vector<rgk_task> t(tasks.size());
        for (size_t i = 0; i < tasks.size(); i++) t[i] = {tasks[i].xrange_start, tasks[i].xrange_end, tasks[i].yrange_start, tasks[i].yrange_end};
        rgk_render_params p{};
        p.xres = cfg.xres; p.yres = cfg.yres; p.multisample = cfg.multisample; p.depth = cfg.recursion_level; p.clamp = cfg.clamp;
        p.russian = cfg.russian; p.bumpmap_scale = cfg.bumpmap_scale; p.force_fresnell = cfg.force_fresnell ? 1u : 0u;
        p.reverse = cfg.reverse; p.sampler_mode = cfg.sampler_mode;
        rgk_round_stats st;
        scene.check(rgk_render_round(scene.context(), &camera, &p, t.data(), (uint32_t)t.size(), (uint32_t)seedstart, seedcount,
                                     total_ob.sum_ptr(), total_ob.count_ptr(), &st), "rgk_render_round");
        seedcount += (unsigned int)tasks.size();
        pixels_done += (int)(cfg.xres * cfg.yres);
        rays_done += (unsigned int)st.closest_rays;
        shadow_rays_done += st.shadow_rays; samples_done += st.samples; gpu_ms += st.gpu_ms;
        rounds_done++;
    }

    // RenderDriver::RenderFrame (src/render_driver.cpp:192-253): Rounds or Timed mode; after every round the normalised
    // image is written to output_file (progressive output).  Additions: when checkpoint_file is not
    // empty the raw (sum, count) accumulators and the number of finished rounds are saved after every round, and with
    // resume = true an existing checkpoint is continued -- the seeds of round k depend only on k, so the result is
    // bit-identical to an uninterrupted run.
    EXRTexture RenderFrame(const Scene& scene, const Config& cfg, const Camera& camera, const std::string& output_file,
                           const std::string& checkpoint_file = "", bool resume = false) {
        pixels_done = 0; rays_done = 0; rounds_done = 0; shadow_rays_done = 0; samples_done = 0; gpu_ms = 0.0;
        EXRTexture total_ob((int)cfg.xres, (int)cfg.yres);
        const std::vector<RenderTask> tasks = GenerateTaskList(TILE_SIZE, cfg.xres, cfg.yres);
        unsigned int seedcount = 0, seedstart = 42, first_round = 0;
        if (resume && !checkpoint_file.empty()) {
            std::ifstream probe(checkpoint_file, std::ios::binary);
            if (probe) {
                uint32_t done = 0;
                EXRTexture saved = EXRTexture::ReadRaw(checkpoint_file, &done);
                if (saved.XSize() != (int)cfg.xres || saved.YSize() != (int)cfg.yres) throw std::runtime_error("checkpoint resolution differs from the configuration");
                total_ob = saved;
                first_round = done;
                seedcount = done * (unsigned int)tasks.size();          // what `seedcount++` per task had reached (:160)
                rounds_done = (int)done;
            }
        }
        if (!output_file.empty() && first_round == 0) total_ob.Write(output_file);      // :200
        const auto start = std::chrono::high_resolution_clock::now();
        auto one_round = [&]() {
            RenderRound(scene, cfg, camera, tasks, seedcount, (int)seedstart, 0, total_ob);
            if (!checkpoint_file.empty()) total_ob.WriteRaw(checkpoint_file, (uint32_t)rounds_done.load());
            if (!output_file.empty()) total_ob.Normalize(cfg.output_scale).Write(output_file);
        };
        switch (cfg.render_limit_mode) {
        case RenderLimitMode::Rounds:
            for (unsigned int roundno = first_round; roundno < cfg.render_rounds; roundno++) one_round();
            break;
        case RenderLimitMode::Timed:
            while (true) {
                const auto now = std::chrono::high_resolution_clock::now();
                const float minutes = std::chrono::duration_cast<std::chrono::seconds>(now - start).count() / 60.0f;
                if (minutes >= cfg.render_minutes) break;
                one_round();
            }
            break;
        }
        return total_ob;
    }
};

// ---------------------------------------------------------------------------------------------- scene pack file
// RGKPACK1 (little endian), written by rgk_b200/scene.py: ScenePack.save.  Everything the reference obtains from its
// JSON config, assimp and the image loaders, already decoded: vertex arrays, triangles, mesh ranges, rgk_material and
// rgk_point_light records verbatim, textures as float RGB, sky, both LTC tables, then the render configuration and
// the Camera constructor arguments.
class PackFile {
public:
    explicit PackFile(const std::string& path) {
        std::ifstream f(path, std::ios::binary);
        if (!f) throw std::runtime_error("cannot open scene pack " + path);
        auto rd = [&](void* p, size_t n) { if (n && !f.read((char*)p, (std::streamsize)n)) throw std::runtime_error("truncated scene pack " + path); };
        char magic[8]; rd(magic, 8);
        if (std::memcmp(magic, "RGKPACK1", 8) != 0) throw std::runtime_error("not an RGKPACK1 file: " + path);
        uint32_t h[8]; rd(h, sizeof h);
        const uint32_t nv = h[0], nt = h[1], nm = h[2], nmat = h[3], ntex = h[4], npl = h[5];
        positions.resize(3 * (size_t)nv); normals.resize(3 * (size_t)nv); tangents.resize(3 * (size_t)nv); texcoords.resize(2 * (size_t)nv);
        indices.resize(3 * (size_t)nt); meshes.resize(nm); materials.resize(nmat); textures.resize(ntex); texels.resize(ntex); lights.resize(npl);
        rd(positions.data(), positions.size() * 4); rd(normals.data(), normals.size() * 4); rd(tangents.data(), tangents.size() * 4);
        rd(texcoords.data(), texcoords.size() * 4); rd(indices.data(), indices.size() * 4);
        rd(meshes.data(), meshes.size() * sizeof(rgk_mesh)); rd(materials.data(), materials.size() * sizeof(rgk_material));
        for (uint32_t i = 0; i < ntex; i++) {
            uint32_t k[3]; float c[3]; rd(k, 12); rd(c, 12);
            rgk_texture& t = textures[i];
            t.kind = k[0]; t.width = k[1]; t.height = k[2]; std::memcpy(t.color, c, 12); t.texels = nullptr;
            if (t.kind == 1) { texels[i].resize(3 * (size_t)t.width * t.height); rd(texels[i].data(), texels[i].size() * 4); }
        }
        for (uint32_t i = 0; i < ntex; i++) if (textures[i].kind == 1) textures[i].texels = texels[i].data();
        rd(lights.data(), lights.size() * sizeof(rgk_point_light));
        rd(&sky, sizeof sky);
        for (auto* v : {&ggx_M, &beck_M}) v->resize(4096 * 9);
        for (auto* v : {&ggx_amp, &beck_amp}) v->resize(4096);
        rd(ggx_M.data(), ggx_M.size() * 4); rd(ggx_amp.data(), ggx_amp.size() * 4); rd(beck_M.data(), beck_M.size() * 4); rd(beck_amp.data(), beck_amp.size() * 4);
        uint32_t c[8]; rd(c, sizeof c);
        float cf[4]; rd(cf, sizeof cf);
        config.xres = c[0]; config.yres = c[1]; config.multisample = c[2]; config.recursion_level = c[3]; config.render_rounds = c[4];
        config.force_fresnell = c[5] != 0; config.reverse = c[6];
        config.clamp = cf[0]; config.russian = cf[1]; config.bumpmap_scale = cf[2]; config.output_scale = cf[3];
        rd(cam_args, sizeof cam_args);
        thinglass = h[6];
    }
    rgk_scene_desc desc() const {
        rgk_scene_desc d{};
        d.n_vertices = (uint32_t)(positions.size() / 3); d.positions = positions.data(); d.normals = normals.data();
        d.tangents = tangents.data(); d.texcoords = texcoords.data();
        d.n_triangles = (uint32_t)(indices.size() / 3); d.indices = indices.data();
        d.n_meshes = (uint32_t)meshes.size(); d.meshes = meshes.data();
        d.n_materials = (uint32_t)materials.size(); d.materials = materials.data();
        d.n_textures = (uint32_t)textures.size(); d.textures = textures.data();
        d.n_point_lights = (uint32_t)lights.size(); d.point_lights = lights.data();
        d.sky = sky;
        d.ltc_ggx.M = ggx_M.data(); d.ltc_ggx.amplitude = ggx_amp.data();
        d.ltc_beckmann.M = beck_M.data(); d.ltc_beckmann.amplitude = beck_amp.data();
        d.thinglass = thinglass;
        return d;
    }
    Camera camera() const {                 // pos, lookat, up, yview, xview, focus_plane, lens_size (ConfigJSON::GetCamera)
        return Camera(cam_args, cam_args + 3, cam_args + 6, cam_args[9], cam_args[10], (int)config.xres, (int)config.yres, cam_args[11], cam_args[12]);
    }
    Config config;
    std::vector<float> positions, normals, tangents, texcoords;
    std::vector<uint32_t> indices;
    std::vector<rgk_mesh> meshes;
    std::vector<rgk_material> materials;
    std::vector<rgk_texture> textures;
    std::vector<std::vector<float>> texels;
    std::vector<rgk_point_light> lights;
    rgk_sky sky{};
    std::vector<float> ggx_M, ggx_amp, beck_M, beck_amp;
    float cam_args[13] = {0};
    uint32_t thinglass = 0;
};

} // namespace rgkb
#endif // RGK_B200_HOST_HPP
