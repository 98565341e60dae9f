// rgk_render_multi -- native multi-GPU driver: one host thread and one rgk_context per GPU of the node, rounds dealt
// round-robin (GPU g renders rounds g, g + N, ...; seedcount of round r is r * tasks, src/render_driver.cpp:160,222),
// device-resident partial framebuffers, and ONE ncclReduce per super-round (N rounds) over NVLink into GPU 0
// (SURVEY 8e).  The library itself stays NCCL-free: the collective is issued here, on the stream each context was
// created with, so it is ordered after that GPU's round without a host synchronisation.
//
//   rgk_render_multi scene.rgkpack out.exr --gpus N [--rounds R] [--raw file] [--bvh]
//
// build: g++ -std=c++17 -O2 -Iinclude -I/usr/local/cuda/include rgk_b200/host/rgk_render_multi.cpp -o rgk_render_multi
//        -Lrgk_b200 -lrgk_b200 -L/usr/local/cuda/lib64 -lcudart -lnccl -lpthread
#include <cuda_runtime.h>
#include <nccl.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>
#include "rgk_b200_host.hpp"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
#define NK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) throw std::runtime_error(std::string(#call) + ": " + ncclGetErrorString(r_)); } while (0)

struct Gpu {
    int dev = 0;
    cudaStream_t stream = nullptr;
    rgk_context* ctx = nullptr;
    float* d_sum = nullptr; uint32_t* d_cnt = nullptr;
    ncclComm_t comm = nullptr;
    rgk_round_stats stats{};
    uint64_t closest = 0, shadow = 0, samples = 0;
    double gpu_ms = 0.0;
    std::string error;
};

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: rgk_render_multi scene.rgkpack out.exr --gpus N [--rounds R] [--raw f] [--bvh]\n"); return 2; }
    try {
        rgkb::PackFile pack(argv[1]);
        const std::string out = argv[2];
        std::string raw;
        int n = 1;
        rgkb::Config cfg = pack.config;
        for (int i = 3; i < argc; i++) {
            const std::string a = argv[i];
            if (a == "--gpus" && i + 1 < argc) n = std::atoi(argv[++i]);
            else if (a == "--rounds" && i + 1 < argc) cfg.render_rounds = (unsigned)std::atoi(argv[++i]);
            else if (a == "--raw" && i + 1 < argc) raw = argv[++i];
            else if (a == "--bvh") setenv("RGK_WIDE_BVH", "1", 1);          // read by every rgk_scene_commit below
            else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
        }
        int have = 0;
        CK(cudaGetDeviceCount(&have));
        if (n < 1 || n > have) throw std::runtime_error("--gpus " + std::to_string(n) + " but " + std::to_string(have) + " CUDA devices are visible");
        const rgk_scene_desc desc = pack.desc();
        const rgkb::Camera camera = pack.camera();
        const std::vector<rgkb::RenderTask> tasks = rgkb::GenerateTaskList(rgkb::RenderDriver::TILE_SIZE, cfg.xres, cfg.yres);
        std::vector<rgk_task> t(tasks.size());
        for (size_t i = 0; i < tasks.size(); i++) t[i] = {tasks[i].xrange_start, tasks[i].xrange_end, tasks[i].yrange_start, tasks[i].yrange_end};
        rgk_render_params p{};
        p.xres = cfg.xres; p.yres = cfg.yres; p.multisample = cfg.multisample; p.depth = cfg.recursion_level; p.clamp = cfg.clamp;
        p.russian = cfg.russian; p.bumpmap_scale = cfg.bumpmap_scale; p.force_fresnell = cfg.force_fresnell ? 1u : 0u;
        p.reverse = cfg.reverse; p.sampler_mode = cfg.sampler_mode;
        const size_t npx = (size_t)cfg.xres * cfg.yres;

        std::vector<Gpu> gpus(n);
        std::vector<int> devs(n);
        std::vector<ncclComm_t> comms(n);
        for (int g = 0; g < n; g++) devs[g] = g;
        NK(ncclCommInitAll(comms.data(), n, devs.data()));
        for (int g = 0; g < n; g++) {                       // scene replicated on every GPU (read-only, a few hundred MB at most)
            Gpu& G = gpus[g];
            G.dev = g; G.comm = comms[g];
            CK(cudaSetDevice(g));
            CK(cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking));
            if (rgk_context_create(g, G.stream, &G.ctx) != RGK_OK) throw std::runtime_error(std::string("rgk_context_create: ") + rgk_last_error(nullptr));
            if (rgk_scene_commit(G.ctx, &desc, nullptr) != RGK_OK) throw std::runtime_error(std::string("rgk_scene_commit: ") + rgk_last_error(G.ctx));
            CK(cudaMalloc((void**)&G.d_sum, npx * 3 * sizeof(float)));
            CK(cudaMalloc((void**)&G.d_cnt, npx * sizeof(uint32_t)));
            CK(cudaMemsetAsync(G.d_sum, 0, npx * 3 * sizeof(float), G.stream));
            CK(cudaMemsetAsync(G.d_cnt, 0, npx * sizeof(uint32_t), G.stream));
        }
        const unsigned rounds = cfg.render_rounds, super_rounds = (rounds + n - 1) / n;
        auto worker = [&](int g) {
            Gpu& G = gpus[g];
            try {
                CK(cudaSetDevice(G.dev));
                for (unsigned s = 0; s < super_rounds; s++) {
                    const unsigned r = s * n + g;
                    if (r < rounds) {
                        if (rgk_render_round_device(G.ctx, &camera, &p, t.data(), (uint32_t)t.size(), 42u, r * (uint32_t)t.size(), G.d_sum, G.d_cnt, &G.stats) != RGK_OK)
                            throw std::runtime_error(std::string("rgk_render_round_device: ") + rgk_last_error(G.ctx));
                        G.closest += G.stats.closest_rays; G.shadow += G.stats.shadow_rays; G.samples += G.stats.samples; G.gpu_ms += G.stats.gpu_ms;
                    }
                    // one reduce per super-round: GPU 0 renders into the running total and receives the others' rounds in place
                    NK(ncclGroupStart());
                    NK(ncclReduce(G.d_sum, G.d_sum, npx * 3, ncclFloat32, ncclSum, 0, G.comm, G.stream));
                    NK(ncclReduce(G.d_cnt, G.d_cnt, npx, ncclUint32, ncclSum, 0, G.comm, G.stream));
                    NK(ncclGroupEnd());
                    if (g != 0) {
                        CK(cudaMemsetAsync(G.d_sum, 0, npx * 3 * sizeof(float), G.stream));
                        CK(cudaMemsetAsync(G.d_cnt, 0, npx * sizeof(uint32_t), G.stream));
                    }
                }
                CK(cudaStreamSynchronize(G.stream));
            } catch (const std::exception& e) { G.error = e.what(); }
        };
        std::vector<std::thread> th;
        for (int g = 0; g < n; g++) th.emplace_back(worker, g);
        for (auto& x : th) x.join();
        for (const Gpu& G : gpus) if (!G.error.empty()) throw std::runtime_error("GPU " + std::to_string(G.dev) + ": " + G.error);

        rgkb::EXRTexture total((int)cfg.xres, (int)cfg.yres);
        CK(cudaSetDevice(0));
        CK(cudaMemcpy(total.sum_ptr(), gpus[0].d_sum, npx * 3 * sizeof(float), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(total.count_ptr(), gpus[0].d_cnt, npx * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        total.Normalize(cfg.output_scale).Write(out);
        if (!raw.empty()) total.WriteRaw(raw, rounds);
        uint64_t closest = 0, shadow = 0, samples = 0; double ms = 0.0;
        for (const Gpu& G : gpus) { closest += G.closest; shadow += G.shadow; samples += G.samples; ms = std::max(ms, G.gpu_ms); }
        std::printf("{\"gpus\": %d, \"rounds\": %u, \"closest_rays\": %llu, \"shadow_rays\": %llu, \"samples\": %llu, \"max_gpu_ms\": %.3f}\n", n, rounds,
                    (unsigned long long)closest, (unsigned long long)shadow, (unsigned long long)samples, ms);
        for (Gpu& G : gpus) {
            cudaSetDevice(G.dev);
            rgk_context_destroy(G.ctx); cudaFree(G.d_sum); cudaFree(G.d_cnt); cudaStreamDestroy(G.stream); ncclCommDestroy(G.comm);
        }
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "rgk_render_multi: %s\n", e.what());
        return 1;
    }
}
