// probe.cu -- rgk_probe: evaluates single shading functions of the device code on caller-supplied inputs, so that
// the parity tests can compare them one by one with the CPU oracle (SURVEY 7 "per-function unit tests at 1e-5 rel").
// Not part of the reference's API; test entry point only, never on the timed path.
#include "probe_device.cuh"

namespace {
__global__ void k_probe(DevScene S, uint32_t kind, uint32_t index, const float* __restrict__ in, uint64_t n, float* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    probe_one(S, kind, index, in, i, out);
}
const uint32_t IN_W[] = {7, 8, 2, 5, 3, 6}, OUT_W[] = {7, 3, 5, 12, 3, 6};
}

extern "C" rgk_status rgk_probe(rgk_context* ctx, uint32_t kind, uint32_t index, const float* in, uint64_t n, float* out) {
    if (!ctx || kind > RGK_PROBE_FRAME || (n && (!in || !out))) return RGK_ERR_INVALID;
    if (!ctx->has_scene) return rgk_fail(ctx, RGK_ERR_NO_SCENE, "rgk_scene_commit has not been called");
    if ((kind <= RGK_PROBE_BXDF_VALUE && index >= ctx->dev.n_materials) || (kind == RGK_PROBE_TEXTURE && index >= ctx->dev.n_textures))
        return rgk_fail(ctx, RGK_ERR_INVALID, "probe index out of range");
    if (n == 0) return RGK_OK;
    RGK_CUDA(ctx, cudaSetDevice(ctx->device));
    float* d_in = (float*)rgk_scratch(ctx, 0, n * IN_W[kind] * 4);
    float* d_out = (float*)rgk_scratch(ctx, 1, n * OUT_W[kind] * 4);
    if (!d_in || !d_out) return rgk_fail(ctx, RGK_ERR_NOMEM, "device scratch allocation failed");
    RGK_CUDA(ctx, cudaMemcpyAsync(d_in, in, n * IN_W[kind] * 4, cudaMemcpyHostToDevice, ctx->stream));
    k_probe<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->dev, kind, index, d_in, n, d_out);
    ctx->launches++;
    RGK_CUDA(ctx, cudaGetLastError());
    RGK_CUDA(ctx, cudaMemcpyAsync(out, d_out, n * OUT_W[kind] * 4, cudaMemcpyDeviceToHost, ctx->stream));
    RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RGK_OK;
}
