// Host-only checks of include/rgk_b200_host.hpp (no GPU): prints one JSON object that tests/test_host_cpp.py
// compares with the CPU oracle / numpy.  usage: host_checks <pack file> <out dir>
#include <cstdio>
#include <string>
#include "rgk_b200_host.hpp"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const std::string dir = argv[2];
    std::printf("{");
    // --- EXRTexture: AddPixel / Accumulate / GetPixel / Normalize (src/texture.cpp:334-412)
    rgkb::EXRTexture a(4, 3), b(4, 3);
    for (int y = 0; y < 3; y++)
        for (int x = 0; x < 4; x++) {
            rgkb::Radiance c; c.r = 0.25f * x + y; c.g = 1.5f * y; c.b = 0.125f * (x + 1);
            a.AddPixel(x, y, c, 2);
            if ((x + y) % 2) { rgkb::Radiance d; d.r = 1.0f; d.g = 2.0f; d.b = 3.0f; b.AddPixel(x, y, d, 5); }
        }
    a.Accumulate(b);
    const rgkb::Radiance p = a.GetPixel(1, 2), z = rgkb::EXRTexture(2, 2).GetPixel(0, 0);
    std::printf("\"pixel_1_2\": [%.9g, %.9g, %.9g], \"count_1_2\": %u, \"empty\": [%g, %g, %g], ", p.r, p.g, p.b, a.Count()[2 * 4 + 1], z.r, z.g, z.b);
    const rgkb::EXRTexture n = a.Normalize(-1.0f), s = a.Normalize(0.5f);
    float m = 0.0f;
    for (int y = 0; y < 3; y++) for (int x = 0; x < 4; x++) { const rgkb::Radiance q = n.GetPixel(x, y); m = std::max(m, std::max(q.r, std::max(q.g, q.b))); }
    std::printf("\"normalized_max\": %.9g, \"scaled_0_0_b\": %.9g, ", m, s.GetPixel(0, 0).b);
    a.Write(dir + "/a.exr");
    a.WriteRaw(dir + "/a.acc", 7);
    uint32_t rounds = 0;
    const rgkb::EXRTexture back = rgkb::EXRTexture::ReadRaw(dir + "/a.acc", &rounds);
    bool same = rounds == 7 && back.Count() == a.Count();
    for (size_t i = 0; same && i < a.Data().size(); i++) same = a.Data()[i].r == back.Data()[i].r && a.Data()[i].g == back.Data()[i].g && a.Data()[i].b == back.Data()[i].b;
    std::printf("\"raw_roundtrip\": %s, ", same ? "true" : "false");
    std::printf("\"half\": [%u, %u, %u, %u, %u, %u, %u], ", rgkb::EXRTexture::FloatToHalf(1.0f), rgkb::EXRTexture::FloatToHalf(-2.5f), rgkb::EXRTexture::FloatToHalf(65504.0f),
                rgkb::EXRTexture::FloatToHalf(1e6f), rgkb::EXRTexture::FloatToHalf(5.9604645e-8f), rgkb::EXRTexture::FloatToHalf(0.33333334f), rgkb::EXRTexture::FloatToHalf(1.0009766f + 0.00048828f));
    // --- task list (src/render_driver.cpp:30-46)
    const auto tasks = rgkb::GenerateTaskList(32, 200, 100);
    std::printf("\"tasks\": [");
    for (size_t i = 0; i < tasks.size(); i++) std::printf("%s[%u, %u, %u, %u]", i ? ", " : "", tasks[i].xrange_start, tasks[i].xrange_end, tasks[i].yrange_start, tasks[i].yrange_end);
    std::printf("], ");
    // --- pack file + camera
    rgkb::PackFile pack(argv[1]);
    const rgk_scene_desc d = pack.desc();
    const rgkb::Camera cam = pack.camera();
    double possum = 0; for (float v : pack.positions) possum += v;
    std::printf("\"pack\": {\"n_vertices\": %u, \"n_triangles\": %u, \"n_meshes\": %u, \"n_materials\": %u, \"n_textures\": %u, \"possum\": %.9g, \"xres\": %u, \"multisample\": %u, "
                "\"depth\": %u, \"russian\": %.9g, \"clamp\": %.9g, \"emission3\": %.9g}, ",
                d.n_vertices, d.n_triangles, d.n_meshes, d.n_materials, d.n_textures, possum, pack.config.xres, pack.config.multisample, pack.config.recursion_level,
                pack.config.russian, pack.config.clamp, d.n_materials > 3 ? d.materials[3].emission[0] : -1.0f);
    std::printf("\"camera\": {\"origin\": [%.9g, %.9g, %.9g], \"viewscreen\": [%.9g, %.9g, %.9g], \"viewscreen_x\": [%.9g, %.9g, %.9g], \"viewscreen_y\": [%.9g, %.9g, %.9g], \"simple\": %s}, ",
                cam.origin[0], cam.origin[1], cam.origin[2], cam.viewscreen[0], cam.viewscreen[1], cam.viewscreen[2], cam.viewscreen_x[0], cam.viewscreen_x[1], cam.viewscreen_x[2],
                cam.viewscreen_y[0], cam.viewscreen_y[1], cam.viewscreen_y[2], cam.IsSimple() ? "true" : "false");
    // --- host-only commit through the C ABI with the pack's description
    rgk_host_scene* hs = nullptr;
    const rgk_status st = rgk_host_scene_create(&d, nullptr, nullptr, &hs);
    rgk_scene_info info{};
    if (st == RGK_OK) rgk_host_scene_get_info(hs, &info);
    std::printf("\"host_scene\": {\"status\": %d, \"n_nodes\": %u, \"n_refs\": %u, \"epsilon\": %.9g}, ", (int)st, info.n_nodes, info.n_refs, info.epsilon);
    rgk_host_scene_destroy(hs);
    // --- no device: the Scene constructor throws, nothing falls back to the CPU
    std::string err = "no exception";
    try { rgkb::Scene scene(0); err = "constructed"; } catch (const std::exception& e) { err = e.what(); }
    for (char& c : err) if (c == '"' || c == '\\' || c == '\n') c = ' ';
    std::printf("\"scene_ctor\": \"%s\"}\n", err.c_str());
    return 0;
}
