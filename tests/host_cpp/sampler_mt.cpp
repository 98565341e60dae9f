// sampler_mt.cpp -- render.cu compiled for the host with 32-lane warps made of real threads (device_shim_mt.h), behind one C
// function: the sampler-table entry point (launch_sampler_tables -> k_sampler_warp / k_sampler_mt).  tests/
// test_sampler_lanes_on_host.py compares the tables with the oracle's bit for bit; tools/tsan_lanes_on_host.sh runs the same
// under ThreadSanitizer and AddressSanitizer.  Test infrastructure only.
// (-fvisibility=hidden -fno-gnu-unique: the one-lane build libdevice_on_host.so compiles the same sources into the same symbol
// names; loaded into one process, the two must not share inline variables)
//   g++ -std=c++17 -O2 -ffp-contract=off -fPIC -shared -pthread -fvisibility=hidden -fno-gnu-unique -I/usr/local/cuda/include -Ibuild/host/gen32 -Iinclude -Itests/host_cpp \
//       tests/host_cpp/sampler_mt.cpp -o build/host/libsampler_mt.so
#include "device_shim_mt.h"
#include <cstdio>
#include <string>
#include "bvh_device.cuh"
#include "probe_device.cuh"
#include "render_host.inc"

rgk_status rgk_fail(rgk_context* ctx, rgk_status s, const std::string& msg) { if (ctx) ctx->last_error = msg; return s; }
void* rgk_scratch(rgk_context* ctx, int slot, size_t bytes) {
    if (ctx->scratch_size[slot] >= bytes && ctx->scratch[slot]) return ctx->scratch[slot];
    std::free(ctx->scratch[slot]);
    ctx->scratch[slot] = std::calloc(std::max<size_t>(bytes, 256), 1); ctx->scratch_size[slot] = std::max<size_t>(bytes, 256);
    return ctx->scratch[slot];
}

extern "C" {
// tables of n seeds: out1[n1d][set size][n], out2[n2d][set size][n][2] (the layout of rgk_sampler_tables' device buffers);
// kernel: rgk_device_cfg::sampler_kernel (2 = the warp-per-pixel builder), slots: ::sampler_slots.  Returns the status.
__attribute__((visibility("default")))
int doh32_sampler_tables(const uint32_t* seeds, uint32_t n, uint32_t multisample, uint32_t n1d, uint32_t n2d, float* out1, float* out2,
                         uint32_t kernel, uint32_t slots, const char** error) {
    static rgk_context ctx;
    static std::string err;
    ctx.cfg = rgk_device_cfg{};
    ctx.cfg.sampler_kernel = kernel; ctx.cfg.sampler_slots = slots; ctx.cfg.sampler_smem = 1; ctx.cfg.sampler_ctas_per_sm = 1;
    ctx.stream = nullptr; ctx.device = 0;
    const rgk_status st = launch_sampler_tables(&ctx, seeds, n, multisample, n1d, n2d, out1, out2);
    err = ctx.last_error;
    if (error) *error = err.c_str();
    return (int)st;
}
}
