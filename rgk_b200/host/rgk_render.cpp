// rgk_render -- minimal C++ host driver over include/rgk_b200_host.hpp: loads an RGKPACK1 scene pack, commits it to
// the GPU and runs RenderDriver::RenderFrame (rounds or timed, progressive EXR output, checkpoint / resume).
// It is the C++ counterpart of what RGKrt's main() does after its config and scene are loaded (src/main.cpp:217-247);
// argument parsing, the progress monitor and animation stay out of scope.
//
//   rgk_render scene.rgkpack out.exr [--rounds N | --minutes M] [--checkpoint file] [--resume] [--raw file] [--kd]
//   (--kd: the reference's kd-tree for every ray instead of the wide-BVH candidate pass + kd-tree arbiter; same pixels, slower)
//              [--tiles]   (render tile by tile through PathTracer::Render, the reference's granularity)
#include <cstdio>
#include <cstdlib>
#include <string>
#include "rgk_b200_host.hpp"

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: rgk_render scene.rgkpack out.exr [--rounds N | --minutes M] [--checkpoint f] [--resume] [--raw f] [--tiles] [--kd]\n"); return 2; }
    try {
        rgkb::PackFile pack(argv[1]);
        const std::string out = argv[2];
        std::string checkpoint, raw;
        bool resume = false, tiles = false, kd = false;
        rgkb::Config cfg = pack.config;
        for (int i = 3; i < argc; i++) {
            const std::string a = argv[i];
            if (a == "--rounds" && i + 1 < argc) { cfg.render_rounds = (unsigned)std::atoi(argv[++i]); cfg.render_limit_mode = rgkb::RenderLimitMode::Rounds; }
            else if (a == "--minutes" && i + 1 < argc) { cfg.render_minutes = (float)std::atof(argv[++i]); cfg.render_limit_mode = rgkb::RenderLimitMode::Timed; }
            else if (a == "--checkpoint" && i + 1 < argc) checkpoint = argv[++i];
            else if (a == "--raw" && i + 1 < argc) raw = argv[++i];
            else if (a == "--resume") resume = true;
            else if (a == "--tiles") tiles = true;
            else if (a == "--kd") kd = true;
            else if (a == "--bvh") kd = false;            // the default since ABI 4
            else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
        }
        rgkb::Scene scene(0);
        const rgk_scene_desc desc = pack.desc();
        scene.Commit(desc, nullptr, kd ? RGK_TRAVERSAL_KD : RGK_TRAVERSAL_BVH);
        const rgkb::Camera camera = pack.camera();
        rgkb::EXRTexture total(0, 0);
        rgkb::RenderDriver driver;
        if (tiles) {
            // the reference's own loop: one PathTracer per tile, seeded seedstart + seedcount++ (src/render_driver.cpp:158-184)
            total = rgkb::EXRTexture((int)cfg.xres, (int)cfg.yres);
            const auto tasks = rgkb::GenerateTaskList(rgkb::RenderDriver::TILE_SIZE, cfg.xres, cfg.yres);
            unsigned int seedcount = 0; const unsigned int seedstart = 42;
            for (unsigned int r = 0; r < cfg.render_rounds; r++)
                for (const rgkb::RenderTask& task : tasks) {
                    rgkb::PathTracer rt(scene, camera, task.xres, task.yres, cfg.multisample, cfg.recursion_level, cfg.clamp, cfg.russian,
                                        cfg.bumpmap_scale, cfg.force_fresnell, cfg.reverse, seedstart + seedcount++);
                    rgkb::EXRTexture output_buffer((int)cfg.xres, (int)cfg.yres);
                    rt.Render(task, &output_buffer, driver.pixels_done, driver.rays_done);
                    total.Accumulate(output_buffer);
                }
            total.Normalize(cfg.output_scale).Write(out);
        } else {
            total = driver.RenderFrame(scene, cfg, camera, out, checkpoint, resume);
        }
        if (!raw.empty()) total.WriteRaw(raw, (uint32_t)driver.rounds_done.load());
        std::printf("{\"triangles\": %u, \"rounds\": %d, \"closest_rays\": %u, \"shadow_rays\": %llu, \"samples\": %llu, \"gpu_ms\": %.3f}\n",
                    scene.info.n_triangles, driver.rounds_done.load(), driver.rays_done.load(), (unsigned long long)driver.shadow_rays_done,
                    (unsigned long long)driver.samples_done, driver.gpu_ms);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "rgk_render: %s\n", e.what());
        return 1;
    }
}
