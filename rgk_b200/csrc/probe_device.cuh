// probe_device.cuh -- one row of rgk_probe: a single shading function of the device code evaluated on caller-supplied
// inputs (probe.cu launches it per row; tests/host_cpp/device_on_host.cpp compiles it for the host).
#pragma once
#include "shade_device.cuh"

__device__ __forceinline__ void probe_one(const DevScene& S, uint32_t kind, uint32_t index, const float* __restrict__ in, uint64_t i, float* __restrict__ out) {
    switch (kind) {
    case RGK_PROBE_BXDF_SAMPLE: {          // in: Vi[3] uv[2] sample[2]; out: dir[3] weight[3] may_leak
        const float* q = in + 7 * i; float* o = out + 7 * i;
        V3 dir; RGB w; bool leak;
        bxdf_sample(S, index, v3(q), V2{q[3], q[4]}, V2{q[5], q[6]}, dir, w, leak);
        o[0] = dir.x; o[1] = dir.y; o[2] = dir.z; o[3] = w.r; o[4] = w.g; o[5] = w.b; o[6] = leak ? 1.0f : 0.0f;
        break; }
    case RGK_PROBE_BXDF_VALUE: {           // in: Vi[3] Vr[3] uv[2]; out: rgb
        const float* q = in + 8 * i; float* o = out + 3 * i;
        const RGB c = bxdf_value(S, index, v3(q), v3(q + 3), V2{q[6], q[7]});
        o[0] = c.r; o[1] = c.g; o[2] = c.b;
        break; }
    case RGK_PROBE_TEXTURE: {              // in: uv[2]; out: rgb, slope right, slope bottom
        const float* q = in + 2 * i; float* o = out + 5 * i;
        const RGB c = tex_fetch(S, (int32_t)index, V2{q[0], q[1]});
        o[0] = c.r; o[1] = c.g; o[2] = c.b;
        tex_slopes(S, (int32_t)index, V2{q[0], q[1]}, o[3], o[4]);
        break; }
    case RGK_PROBE_RANDOM_LIGHT: {         // in: choice[2] light_sample tri_sample[2]; out: type pos[3] colour[3] intensity size normal[3]
        const float* q = in + 5 * i; float* o = out + 12 * i;
        const LightRec l = random_light(S, V2{q[0], q[1]}, q[2], V2{q[3], q[4]});
        o[0] = (float)l.type; o[1] = l.pos.x; o[2] = l.pos.y; o[3] = l.pos.z; o[4] = l.color.r; o[5] = l.color.g; o[6] = l.color.b;
        o[7] = l.intensity; o[8] = l.size; o[9] = l.normal.x; o[10] = l.normal.y; o[11] = l.normal.z;
        break; }
    case RGK_PROBE_SKY: {                  // in: dir[3]; out: rgb
        const float* q = in + 3 * i; float* o = out + 3 * i;
        const RGB c = sky_radiance(S, v3(q));
        o[0] = c.r; o[1] = c.g; o[2] = c.b;
        break; }
    case RGK_PROBE_FRAME: {                // in: normal[3] v[3]; out: toLocal(v)[3] toGlobal(toLocal(v))[3]
        const float* q = in + 6 * i; float* o = out + 6 * i;
        const Frame f = system_transform_z(v3(q));
        const V3 l = qrot(f.g2l, v3(q + 3)), g = qrot(f.l2g, l);
        o[0] = l.x; o[1] = l.y; o[2] = l.z; o[3] = g.x; o[4] = g.y; o[5] = g.z;
        break; }
    }
}
