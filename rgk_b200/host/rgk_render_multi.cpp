// rgk_render_multi -- native multi-GPU driver: one host thread and one rgk_context per GPU of the node, rounds dealt
// round-robin (GPU g renders rounds g, g + N, ...; seedcount of round r is r * tasks, src/render_driver.cpp:160,222),
// device-resident partial framebuffers, and ONE ncclReduce per super-round (N rounds) over NVLink into GPU 0
// (SURVEY 8e).  The library itself stays NCCL-free: the collective is issued here.
//
// Overlap: every GPU renders super-round s into partial buffer s % 2 on its render stream while the reduce of buffer
// (s - 1) % 2 runs on a second (communication) stream; events order "render done -> reduce" and "reduce done -> buffer
// reused".  Only the colour sums travel: the sample count of a pixel is analytic (every round adds `multisample` to every
// pixel of the frame, src/tracer.cpp:18), so GPU 0 fills it in at the end.
// Failure handling: a GPU whose round fails votes `failed` at the super-round's host barrier, and every thread leaves
// before anyone enters the collective; an error inside the collective aborts all communicators (ncclCommAbort), which
// releases the peers blocked in it.  Either way main() reports the error instead of hanging.
//
//   rgk_render_multi scene.rgkpack out.exr --gpus N [--rounds R] [--raw file] [--kd] [--fail-gpu G (test hook)]
//
// build: g++ -std=c++17 -O2 -Iinclude -I/usr/local/cuda/include rgk_b200/host/rgk_render_multi.cpp -o rgk_render_multi
//        -Lrgk_b200 -lrgk_b200 -L/usr/local/cuda/lib64 -lcudart -lnccl -lpthread
#include <cuda_runtime.h>
#include <nccl.h>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>
#include "rgk_b200_host.hpp"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
#define NK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) throw std::runtime_error(std::string(#call) + ": " + ncclGetErrorString(r_)); } while (0)

// reusable host barrier for the GPU threads (C++17: no std::barrier)
class HostBarrier {
    std::mutex m; std::condition_variable cv; int n, waiting = 0; unsigned long generation = 0;
public:
    explicit HostBarrier(int n_) : n(n_) {}
    void arrive_and_wait() {
        std::unique_lock<std::mutex> lk(m);
        const unsigned long gen = generation;
        if (++waiting == n) { waiting = 0; ++generation; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != generation; });
    }
};

struct Gpu {
    int dev = 0;
    cudaStream_t stream = nullptr, comm_stream = nullptr;
    cudaEvent_t rendered[2] = {nullptr, nullptr}, reduced[2] = {nullptr, nullptr};
    rgk_context* ctx = nullptr;
    float* d_part[2] = {nullptr, nullptr};     // this GPU's colour sums of super-round s at [s % 2]
    float* d_sum = nullptr;                    // GPU 0: the running total
    uint32_t* d_cnt = nullptr;                 // scratch for the library's per-round counts (not reduced: analytic)
    ncclComm_t comm = nullptr;
    rgk_round_stats stats{};
    uint64_t closest = 0, shadow = 0, samples = 0;
    double gpu_ms = 0.0;
    std::string error;
};

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: rgk_render_multi scene.rgkpack out.exr --gpus N [--rounds R] [--raw f] [--kd]\n"); return 2; }
    try {
        rgkb::PackFile pack(argv[1]);
        const std::string out = argv[2];
        std::string raw;
        int n = 1, fail_gpu = -1;
        uint32_t traversal = RGK_TRAVERSAL_BVH;
        rgkb::Config cfg = pack.config;
        for (int i = 3; i < argc; i++) {
            const std::string a = argv[i];
            if (a == "--gpus" && i + 1 < argc) n = std::atoi(argv[++i]);
            else if (a == "--rounds" && i + 1 < argc) cfg.render_rounds = (unsigned)std::atoi(argv[++i]);
            else if (a == "--raw" && i + 1 < argc) raw = argv[++i];
            else if (a == "--kd") traversal = RGK_TRAVERSAL_KD;
            else if (a == "--bvh") traversal = RGK_TRAVERSAL_BVH;      // the default since ABI 4
            else if (a == "--fail-gpu" && i + 1 < argc) fail_gpu = std::atoi(argv[++i]);   // test hook: that GPU's second round fails
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
        float* d_red[2] = {nullptr, nullptr};               // GPU 0: where the reduce of buffer b lands before it is added to the total
        for (int g = 0; g < n; g++) {                       // scene replicated on every GPU (read-only, a few hundred MB at most)
            Gpu& G = gpus[g];
            G.dev = g; G.comm = comms[g];
            CK(cudaSetDevice(g));
            CK(cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking));
            CK(cudaStreamCreateWithFlags(&G.comm_stream, cudaStreamNonBlocking));
            for (int b = 0; b < 2; b++) {
                CK(cudaEventCreateWithFlags(&G.rendered[b], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&G.reduced[b], cudaEventDisableTiming));
                CK(cudaMalloc((void**)&G.d_part[b], npx * 3 * sizeof(float)));
            }
            if (rgk_context_create(g, G.stream, &G.ctx) != RGK_OK) throw std::runtime_error(std::string("rgk_context_create: ") + rgk_last_error(nullptr));
            rgk_device_cfg dc; rgk_device_cfg_init(&dc); dc.traversal = traversal;
            if (rgk_context_configure(G.ctx, &dc) != RGK_OK) throw std::runtime_error(std::string("rgk_context_configure: ") + rgk_last_error(G.ctx));
            if (rgk_scene_commit(G.ctx, &desc, nullptr) != RGK_OK) throw std::runtime_error(std::string("rgk_scene_commit: ") + rgk_last_error(G.ctx));
            CK(cudaMalloc((void**)&G.d_cnt, npx * sizeof(uint32_t)));
            if (g == 0) {
                CK(cudaMalloc((void**)&G.d_sum, npx * 3 * sizeof(float)));
                CK(cudaMemsetAsync(G.d_sum, 0, npx * 3 * sizeof(float), G.comm_stream));
                for (int b = 0; b < 2; b++) CK(cudaMalloc((void**)&d_red[b], npx * 3 * sizeof(float)));
            }
        }
        const unsigned rounds = cfg.render_rounds, super_rounds = (rounds + n - 1) / n;
        HostBarrier barrier(n);
        std::atomic<bool> failed{false}, aborted{false};
        auto abort_all = [&] {                              // releases every peer blocked in (or about to enter) a collective
            if (!aborted.exchange(true)) for (ncclComm_t c : comms) ncclCommAbort(c);
        };
        auto worker = [&](int g) {
            Gpu& G = gpus[g];
            bool in_collective = false;
            try {
                CK(cudaSetDevice(G.dev));
                for (unsigned s = 0; s < super_rounds; s++) {
                    const unsigned r = s * n + g;
                    const int b = (int)(s & 1u);
                    try {
                        CK(cudaStreamWaitEvent(G.stream, G.reduced[b], 0));           // buffer b was last sent two super-rounds ago
                        CK(cudaMemsetAsync(G.d_part[b], 0, npx * 3 * sizeof(float), G.stream));
                        if (r < rounds) {
                            if (g == fail_gpu && s == 1) throw std::runtime_error("--fail-gpu test hook");
                            if (rgk_render_round_device(G.ctx, &camera, &p, t.data(), (uint32_t)t.size(), 42u, r * (uint32_t)t.size(), G.d_part[b], G.d_cnt, &G.stats) != RGK_OK)
                                throw std::runtime_error(std::string("rgk_render_round_device: ") + rgk_last_error(G.ctx));
                            G.closest += G.stats.closest_rays; G.shadow += G.stats.shadow_rays; G.samples += G.stats.samples; G.gpu_ms += G.stats.gpu_ms;
                        }
                        CK(cudaEventRecord(G.rendered[b], G.stream));
                    } catch (const std::exception& e) { G.error = e.what(); failed = true; }
                    // vote: nobody enters the collective of a super-round in which some GPU failed
                    barrier.arrive_and_wait();
                    if (failed) break;
                    // one reduce per super-round, on the communication stream: overlaps the next super-round's rendering
                    in_collective = true;
                    CK(cudaStreamWaitEvent(G.comm_stream, G.rendered[b], 0));
                    NK(ncclReduce(G.d_part[b], g == 0 ? d_red[b] : nullptr, npx * 3, ncclFloat32, ncclSum, 0, G.comm, G.comm_stream));
                    if (g == 0 && rgk_accumulate_device(G.ctx, G.d_sum, d_red[b], npx * 3, nullptr, nullptr, G.comm_stream) != RGK_OK)      // EXRTexture::Accumulate
                        throw std::runtime_error(std::string("rgk_accumulate_device: ") + rgk_last_error(G.ctx));
                    CK(cudaEventRecord(G.reduced[b], G.comm_stream));
                    in_collective = false;
                }
                if (!failed) { CK(cudaStreamSynchronize(G.stream)); CK(cudaStreamSynchronize(G.comm_stream)); }
            } catch (const std::exception& e) {
                if (G.error.empty()) G.error = e.what();
                failed = true;
                if (in_collective) abort_all();
            }
        };
        std::vector<std::thread> th;
        for (int g = 0; g < n; g++) th.emplace_back(worker, g);
        for (auto& x : th) x.join();
        for (const Gpu& G : gpus) if (!G.error.empty()) { abort_all(); throw std::runtime_error("GPU " + std::to_string(G.dev) + ": " + G.error); }

        rgkb::EXRTexture total((int)cfg.xres, (int)cfg.yres);
        CK(cudaSetDevice(0));
        CK(cudaMemcpy(total.sum_ptr(), gpus[0].d_sum, npx * 3 * sizeof(float), cudaMemcpyDeviceToHost));
        // every round adds `multisample` samples to every pixel (src/tracer.cpp:18): the count needs no collective
        for (size_t i = 0; i < npx; i++) total.count_ptr()[i] = rounds * cfg.multisample;
        total.Normalize(cfg.output_scale).Write(out);
        if (!raw.empty()) total.WriteRaw(raw, rounds);
        uint64_t closest = 0, shadow = 0, samples = 0; double ms = 0.0;
        for (const Gpu& G : gpus) { closest += G.closest; shadow += G.shadow; samples += G.samples; ms = std::max(ms, G.gpu_ms); }
        std::printf("{\"gpus\": %d, \"rounds\": %u, \"closest_rays\": %llu, \"shadow_rays\": %llu, \"samples\": %llu, \"max_gpu_ms\": %.3f}\n", n, rounds,
                    (unsigned long long)closest, (unsigned long long)shadow, (unsigned long long)samples, ms);
        for (Gpu& G : gpus) {
            cudaSetDevice(G.dev);
            rgk_context_destroy(G.ctx); cudaFree(G.d_sum); cudaFree(G.d_cnt);
            for (int b = 0; b < 2; b++) { cudaFree(G.d_part[b]); cudaEventDestroy(G.rendered[b]); cudaEventDestroy(G.reduced[b]); }
            cudaStreamDestroy(G.stream); cudaStreamDestroy(G.comm_stream); ncclCommDestroy(G.comm);
        }
        for (int b = 0; b < 2; b++) cudaFree(d_red[b]);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "rgk_render_multi: %s\n", e.what());
        return 1;
    }
}
