// shade_device.cuh -- device-side shading functions of the bounce loop: local frames,
// sampling warps, textures, LTC lobes, the nine live BxDFs, light picking and the sky.
// Each function names the reference code it implements; arithmetic keeps the reference's
// float operation order (this TU is compiled with -fmad=false).  CUDA's sinf/cosf/acosf/
// asinf/atan2f differ from glibc's in the last ulp, which is why shading parity is
// tolerance-based (1e-5 relative per function, RMSE for images) and not bit-exact.
#pragma once
#include "rgk_internal.h"

struct V3 { float x, y, z; };
struct V2 { float x, y; };
struct RGB { float r, g, b; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 v3(const float4& f) { return V3{f.x, f.y, f.z}; }
__device__ __forceinline__ V3 v3(const float* p) { return V3{p[0], p[1], p[2]}; }
__device__ __forceinline__ RGB rgb(float r, float g, float b) { return RGB{r, g, b}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }   // GLM: tmp=a*b; x+y+z
__device__ __forceinline__ float length(V3 v) { return sqrtf(dot(v, v)); }
__device__ __forceinline__ V3 normalize(V3 v) { return v * (1.0f / sqrtf(dot(v, v))); }             // v * inversesqrt
__device__ __forceinline__ V3 cross(V3 x, V3 y) { return v3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
__device__ __forceinline__ float gmax(float a, float b) { return (a < b) ? b : a; }                // glm::max
__device__ __forceinline__ float gmin(float a, float b) { return (b < a) ? b : a; }                // glm::min
__device__ __forceinline__ float gangle(V3 x, V3 y) { return acosf(gmin(gmax(dot(x, y), -1.0f), 1.0f)); }
#define RGK_PI_F 3.14159265358979323846264338327950288f

struct M3 { V3 c0, c1, c2; };
__device__ __forceinline__ V3 mul(const M3& m, V3 v) {
    return v3(m.c0.x * v.x + m.c1.x * v.y + m.c2.x * v.z,
              m.c0.y * v.x + m.c1.y * v.y + m.c2.y * v.z,
              m.c0.z * v.x + m.c1.z * v.y + m.c2.z * v.z);
}
__device__ __forceinline__ float det3(const M3& m) {
    return + m.c0.x * (m.c1.y * m.c2.z - m.c2.y * m.c1.z)
           - m.c1.x * (m.c0.y * m.c2.z - m.c2.y * m.c0.z)
           + m.c2.x * (m.c0.y * m.c1.z - m.c1.y * m.c0.z);
}
__device__ __forceinline__ M3 inv3(const M3& m) {   // GLM compute_inverse<mat3>
    const float a00 = m.c0.x, a01 = m.c0.y, a02 = m.c0.z, a10 = m.c1.x, a11 = m.c1.y, a12 = m.c1.z, a20 = m.c2.x, a21 = m.c2.y, a22 = m.c2.z;
    const float ood = 1.0f / (+ a00 * (a11 * a22 - a21 * a12) - a10 * (a01 * a22 - a21 * a02) + a20 * (a01 * a12 - a11 * a02));
    M3 I;
    I.c0.x = + (a11 * a22 - a21 * a12) * ood; I.c1.x = - (a10 * a22 - a20 * a12) * ood; I.c2.x = + (a10 * a21 - a20 * a11) * ood;
    I.c0.y = - (a01 * a22 - a21 * a02) * ood; I.c1.y = + (a00 * a22 - a20 * a02) * ood; I.c2.y = - (a00 * a21 - a20 * a01) * ood;
    I.c0.z = + (a01 * a12 - a11 * a02) * ood; I.c1.z = - (a00 * a12 - a10 * a02) * ood; I.c2.z = + (a00 * a11 - a10 * a01) * ood;
    return I;
}

struct Quat { float w, x, y, z; };
__device__ __forceinline__ V3 qrot(const Quat& q, V3 v) {   // GLM quat * vec3
    const V3 qv = v3(q.x, q.y, q.z);
    const V3 uv = cross(qv, v);
    const V3 uuv = cross(qv, uv);
    return v + ((uv * q.w) + uuv) * 2.0f;
}
// RotationBetweenVectors(start, +Z), src/glm.cpp:3-33, and its inverse (SystemTransform, src/glm.hpp:18-35)
struct Frame { Quat g2l, l2g; };
__device__ __forceinline__ Frame system_transform_z(V3 start) {
    start = normalize(start);
    const V3 dest = v3(0.0f, 0.0f, 1.0f);
    const float cosTheta = dot(start, dest);
    Quat q;
    if (cosTheta < -1 + 0.001f) {
        V3 axis = cross(v3(0.0f, 1.0f, 0.0f), start);
        if ((double)length(axis) < 0.01) axis = cross(v3(1.0f, 0.0f, 0.0f), start);
        axis = normalize(axis);
        const float a = RGK_PI_F;
        const float s = sinf(a * 0.5f);
        const V3 vs = axis * s;
        q = Quat{cosf(a * 0.5f), vs.x, vs.y, vs.z};
    } else {
        const V3 axis = cross(start, dest);
        const float s = sqrtf((1 + cosTheta) * 2);
        const float invs = 1 / s;
        q = Quat{s * 0.5f, axis.x * invs, axis.y * invs, axis.z * invs};
    }
    Frame f; f.g2l = q;
    const float d = (q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w);   // dot(quat,quat)
    f.l2g = Quat{q.w / d, -q.x / d, -q.y / d, -q.z / d};                 // conjugate / dot
    return f;
}

// ---- RandomUtils (src/random_utils.hpp)
__device__ __forceinline__ V2 disc_uniform(V2 s) {
    const float r = sqrtf(s.x);
    const float a = (float)((double)(s.y * 2.0f) * 3.14159265358979323846);
    return V2{r * sinf(a), r * cosf(a)};
}
__device__ __forceinline__ V3 hemi_cos_z(V2 s) {
    const V2 p = disc_uniform(s);
    const float z = sqrtf(gmax(0.00001f, 1 - p.x * p.x - p.y * p.y));
    return v3(p.x, p.y, z);
}
// Sample2DToHemisphereCosine (y up) and Sample2DToHemisphereCosineDirected = RotationFromY(direction) * it
// (src/random_utils.hpp:32-47,76-78, src/glm.cpp:36-60): the direction in which a light path leaves its light
__device__ __forceinline__ V3 hemi_cos_directed(V2 s, V3 direction) {
    const V2 p = disc_uniform(s);
    const float y = sqrtf(gmax(0.00001f, 1 - p.x * p.x - p.y * p.y));
    const V3 r = v3(p.x, y, p.y);
    const V3 dest = normalize(direction);
    const float cosTheta = dest.y;
    Quat q;
    if (cosTheta < -1 + 0.00001f) {
        const float a = RGK_PI_F;
        const float sn = sinf(a * 0.5f);
        q = Quat{cosf(a * 0.5f), 1.0f * sn, 0.0f * sn, 0.0f * sn};
    } else {
        const V3 axis = cross(v3(0.0f, 1.0f, 0.0f), dest);
        const float sq = sqrtf((1 + cosTheta) * 2);
        const float invs = 1 / sq;
        q = Quat{sq * 0.5f, axis.x * invs, axis.y * invs, axis.z * invs};
    }
    return qrot(q, r);
}
__device__ __forceinline__ V3 sphere_uniform(V2 s) {
    const float z = s.x * 2.0f - 1.0f;
    const float a = (float)((double)s.y * 6.283185);
    const float r = sqrtf(1 - z * z);
    return v3(r * cosf(a), r * sinf(a), z);
}
__device__ __forceinline__ bool decide_and_rescale(float& sample, float probability) {
    if (probability == 0.0f) return false;
    if (probability == 1.0f) return true;
    if (sample < probability) { sample /= probability; return true; }
    sample = (sample - probability) / (1.0f - probability);
    return false;
}

// What the BxDF functions read of the scene, as a value: the functions below are templates over the scene type so that an
// out-of-line callee (the mix-material walk) can be handed these seven pointers in registers.  Passing `const DevScene&` to a
// __noinline__ function makes the compiler copy the whole kernel-parameter struct (~300 B) into the caller's local memory --
// round 1's k_shade paid a 592-byte stack frame and a call sequence per vertex for that.
struct ShadeTables {
    const DevMaterial* materials; const DevTexture* textures; const float4* texels;
    const float4* ltc_M[2]; const float* ltc_amp[2];
};
__device__ __forceinline__ ShadeTables shade_tables(const DevScene& S) {
    ShadeTables t; t.materials = S.materials; t.textures = S.textures; t.texels = S.texels;
    t.ltc_M[0] = S.ltc_M[0]; t.ltc_M[1] = S.ltc_M[1]; t.ltc_amp[0] = S.ltc_amp[0]; t.ltc_amp[1] = S.ltc_amp[1];
    return t;
}
__device__ __forceinline__ ShadeTables shade_tables(const ShadeTables& S) { return S; }

// ---- textures (src/texture.cpp:35-102; manual fp32 bilinear)
__device__ __forceinline__ float frepeat(float x) { return x - floorf(x); }
// Address part and arithmetic part of FileTexture::GetPixelInterpolated (src/texture.cpp:35-76), split so that several
// fetches of one vertex can issue all their loads before any of them is consumed.
struct BilinearTaps { uint32_t i00, i01, i10, i11; float fx, fy; };
__device__ __forceinline__ BilinearTaps bilinear_taps(const DevTexture& t, V2 uv) {
    const int W = (int)t.width, H = (int)t.height;
    const float x = frepeat(uv.x) * t.width - 0.5f, y = frepeat(uv.y) * t.height - 0.5f;
    float ix0f, iy0f;
    BilinearTaps a;
    a.fx = modff(x, &ix0f); a.fy = modff(y, &iy0f);
    int ix0 = (int)ix0f, iy0 = (int)iy0f;
    const int ix1 = (ix0 != W - 1) ? ix0 + 1 : ix0;
    const int iy1 = (iy0 != H - 1) ? iy0 + 1 : iy0;
    if (ix0 == -1) ix0 = 0;
    if (iy0 == -1) iy0 = 0;
    a.i00 = t.offset + (uint32_t)(iy0 * W + ix0); a.i01 = t.offset + (uint32_t)(iy0 * W + ix1);
    a.i10 = t.offset + (uint32_t)(iy1 * W + ix0); a.i11 = t.offset + (uint32_t)(iy1 * W + ix1);
    return a;
}
__device__ __forceinline__ RGB bilinear_mix(const BilinearTaps& a, float4 c00, float4 c01, float4 c10, float4 c11) {
    const float fy = 1.0f - a.fy, fx = 1.0f - a.fx;
    const float gx = 1.0f - fx, gy = 1.0f - fy;
    const RGB c0s = rgb(fx * c00.x + gx * c01.x, fx * c00.y + gx * c01.y, fx * c00.z + gx * c01.z);
    const RGB c1s = rgb(fx * c10.x + gx * c11.x, fx * c10.y + gx * c11.y, fx * c10.z + gx * c11.z);
    return rgb(fy * c0s.r + gy * c1s.r, fy * c0s.g + gy * c1s.g, fy * c0s.b + gy * c1s.b);
}
template <class SC>
__device__ __forceinline__ RGB tex_fetch(const SC& S, int32_t id, V2 uv) {
    if (id < 0) return rgb(0.0f, 0.0f, 0.0f);
    const DevTexture t = S.textures[id];
    if (t.kind == 0) return rgb(t.color[0], t.color[1], t.color[2]);
    const BilinearTaps a = bilinear_taps(t, uv);
    return bilinear_mix(a, __ldg(S.texels + a.i00), __ldg(S.texels + a.i01), __ldg(S.texels + a.i10), __ldg(S.texels + a.i11));
}
// FileTexture::GetSlopeRight / GetSlopeBottom (src/texture.cpp:78-102)
struct SlopeTaps { uint32_t ih, ir, ib; };
__device__ __forceinline__ SlopeTaps slope_taps(const DevTexture& t, V2 uv) {
    const int W = (int)t.width, H = (int)t.height;
    int x = (int)(frepeat(uv.x) * t.width - 0.5f), y = (int)(frepeat(uv.y) * t.height - 0.5f);
    const int x2 = (x != W - 1) ? x + 1 : x, y2 = (y != H - 1) ? y + 1 : y;
    if (x == -1) x = 0;
    if (y == -1) y = 0;
    SlopeTaps a;
    a.ih = t.offset + (uint32_t)(y * W + x); a.ir = t.offset + (uint32_t)(y * W + x2); a.ib = t.offset + (uint32_t)(y2 * W + x);
    return a;
}
__device__ __forceinline__ void slope_mix(float4 h, float4 r, float4 b, float& right, float& bottom) {
    const float here = (h.x + h.y + h.z) / 3;
    right = here - (r.x + r.y + r.z) / 3;
    bottom = here - (b.x + b.y + b.z) / 3;
}
template <class SC>
__device__ __forceinline__ void tex_slopes(const SC& S, int32_t id, V2 uv, float& right, float& bottom) {
    right = 0.0f; bottom = 0.0f;
    if (id < 0) return;
    const DevTexture t = S.textures[id];
    if (t.kind == 0) return;
    const SlopeTaps a = slope_taps(t, uv);
    slope_mix(__ldg(S.texels + a.ih), __ldg(S.texels + a.ir), __ldg(S.texels + a.ib), right, bottom);
}

// ---- LTC (src/LTC/ltc.cpp:20-143); N is always +Z (BxDFUpVector)
template <class SC>
__device__ __forceinline__ void ltc_bilinear(const SC& S, int which, float theta, float alpha, M3& M, float& amp) {
    float t = gmax(0.0f, gmin(1.0f, theta / (0.5f * 3.14159f)));
    float a = gmax(0.0f, gmin(1.0f, sqrtf(alpha)));
    if (t >= 1.0f) t = 0.999f;
    if (a >= 1.0f) a = 0.999f;
    const int sz = 63;
    const int t1 = (int)floorf(t * sz), t2 = t1 + 1, a1 = (int)floorf(a * sz), a2 = a1 + 1;
    const float dt1 = t * sz - t1, dt2 = t2 - t * sz, da1 = a * sz - a1, da2 = a2 - a * sz;
    const float4* Mt = S.ltc_M[which];
    const float* At = S.ltc_amp[which];
    const int i11 = a1 + t1 * 64, i12 = a2 + t1 * 64, i21 = a1 + t2 * 64, i22 = a2 + t2 * 64;
    float r[9];
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const float4 m11 = __ldg(Mt + 3 * i11 + q), m12 = __ldg(Mt + 3 * i12 + q), m21 = __ldg(Mt + 3 * i21 + q), m22 = __ldg(Mt + 3 * i22 + q);
        const float e11[4] = {m11.x, m11.y, m11.z, m11.w}, e12[4] = {m12.x, m12.y, m12.z, m12.w};
        const float e21[4] = {m21.x, m21.y, m21.z, m21.w}, e22[4] = {m22.x, m22.y, m22.z, m22.w};
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (4 * q + k < 9) r[4 * q + k] = e11[k] * dt2 * da2 + e12[k] * dt2 * da1 + e21[k] * dt1 * da2 + e22[k] * dt1 * da1;
    }
    M.c0 = v3(r[0], r[1], r[2]); M.c1 = v3(r[3], r[4], r[5]); M.c2 = v3(r[6], r[7], r[8]);
    amp = __ldg(At + i11) * dt2 * da2 + __ldg(At + i12) * dt2 * da1 + __ldg(At + i21) * dt1 * da2 + __ldg(At + i22) * dt1 * da1;
}
// LTC::GetPDF(ltc, N, Vr, Vi, alpha) -- parameter names as in src/LTC/ltc.cpp:59
template <class SC>
__device__ __forceinline__ float ltc_pdf(const SC& S, int which, V3 Vr, V3 Vi, float alpha) {
    const V3 N = v3(0.0f, 0.0f, 1.0f);
    const V3 tangent = cross(N, Vi), Vi_cast = cross(tangent, N);
    M3 rot; rot.c0 = Vi_cast; rot.c1 = tangent; rot.c2 = N;
    const M3 unrot = inv3(rot);
    const V3 Vr3 = mul(unrot, Vr);
    const float theta = gangle(Vi, N);
    M3 M; float amp; ltc_bilinear(S, which, theta, alpha, M, amp);
    const M3 invM = inv3(M);
    const V3 p = normalize(mul(invM, Vr3));
    const V3 L_ = mul(M, p);
    const float l = length(L_);
    const float detM = det3(M);
    const float J = detM / (l * l * l);
    const float D = 1.0f / 3.14159f * gmax(0.0f, p.z);
    return amp * D / J;
}
template <class SC>
__device__ __forceinline__ V3 ltc_random(const SC& S, int which, V3 Vi, float roughness, V3 rnd) {
    const V3 N = v3(0.0f, 0.0f, 1.0f);
    const V3 tangent = cross(N, Vi), Vi_cast = cross(tangent, N);
    M3 rot; rot.c0 = Vi_cast; rot.c1 = tangent; rot.c2 = N;
    const float theta = gangle(Vi, N);
    M3 M; float amp; ltc_bilinear(S, which, gmax(theta, RGK_PI_F / 4.0f), roughness, M, amp);
    V3 q = mul(M, rnd);
    if (q.z < 0.0001f) q.z = 0.0001f;
    q = mul(rot, q);
    return normalize(q);
}

// ---- BxDFs (src/bxdf/bxdf.hpp:107-159, src/bxdf/bxdf.cpp:192-423)
__device__ __forceinline__ void fresnel_dielectric(float eta, float cosTheta, float& R, float& cosT) {
    if (cosTheta < 0.0f) { eta = 1.0f / eta; cosTheta = -cosTheta; }
    const float s2 = eta * eta * (1.0f - cosTheta * cosTheta);
    if (s2 > 1.0f) { R = 1.0f; cosT = 0.0f; return; }
    const float ct = sqrtf(gmax(1.0f - s2, 0.0f));
    const float Rs = (eta * cosTheta - ct) / (eta * cosTheta + ct);
    const float Rp = (eta * ct - cosTheta) / (eta * ct + cosTheta);
    R = 0.5f * (Rs * Rs + Rp * Rp); cosT = ct;
}
// The two colour textures of a vertex's material, fetched once per vertex: BxDF::value (NEE) and BxDF::sample
// (continuation) of the reference each call GetPixelInterpolated with the same texture and uv, which returns the same
// value; `have` is false for the children of a mix material (their textures differ from the top-level one's).
struct TexPre { bool have; RGB diffuse, color; uint32_t taps; };   // taps: image texels this vertex needed (counting rounds)
template <class SC>
__device__ __forceinline__ RGB tex_diffuse_of(const SC& S, const DevMaterial& m, V2 uv, const TexPre& pre) {
    return pre.have ? pre.diffuse : tex_fetch(S, m.tex_diffuse, uv);
}
template <class SC>
__device__ __forceinline__ RGB tex_color_of(const SC& S, const DevMaterial& m, V2 uv, const TexPre& pre) {
    return pre.have ? pre.color : tex_fetch(S, m.tex_color, uv);
}
// Everything a vertex reads from textures, in one batch: the three descriptors first, then all eleven texels (four
// diffuse, four colour, three bump) are requested before any is used -- the shading kernel is bound by the latency of
// its dependent loads, not by their number.  Unused taps read texel 0 (always allocated) and are discarded; the
// arithmetic is the very code of tex_fetch / tex_slopes.  `have` stays false for mix materials.
template <class SC>
__device__ __forceinline__ TexPre vertex_textures(const SC& S, const DevMaterial& m, V2 uv, float& right, float& bottom) {
    TexPre pre; pre.have = m.bxdf != RGK_BXDF_MIX; pre.diffuse = rgb(0, 0, 0); pre.color = rgb(0, 0, 0); pre.taps = 0u;
    right = 0.0f; bottom = 0.0f;
    const bool uses_diffuse = m.bxdf == RGK_BXDF_DIFFUSE || m.bxdf == RGK_BXDF_LTC_BECKMANN_DIFFUSE || m.bxdf == RGK_BXDF_LTC_GGX_DIFFUSE;
    const bool uses_color = m.bxdf != RGK_BXDF_DIFFUSE && m.bxdf != RGK_BXDF_TRANSPARENT && m.bxdf != RGK_BXDF_MIX;
    const bool want_d = uses_diffuse && m.tex_diffuse >= 0, want_c = uses_color && m.tex_color >= 0, want_b = m.tex_bump >= 0;
    const DevTexture td = S.textures[want_d ? m.tex_diffuse : 0], tc = S.textures[want_c ? m.tex_color : 0], tb = S.textures[want_b ? m.tex_bump : 0];
    const bool img_d = want_d && td.kind == 1u, img_c = want_c && tc.kind == 1u, img_b = want_b && tb.kind == 1u;
    BilinearTaps ad = bilinear_taps(td, uv), ac = bilinear_taps(tc, uv);
    SlopeTaps ab = slope_taps(tb, uv);
    if (!img_d) { ad.i00 = 0u; ad.i01 = 0u; ad.i10 = 0u; ad.i11 = 0u; }
    if (!img_c) { ac.i00 = 0u; ac.i01 = 0u; ac.i10 = 0u; ac.i11 = 0u; }
    if (!img_b) { ab.ih = 0u; ab.ir = 0u; ab.ib = 0u; }
    const float4* __restrict__ px = S.texels;
    const float4 d00 = __ldg(px + ad.i00), d01 = __ldg(px + ad.i01), d10 = __ldg(px + ad.i10), d11 = __ldg(px + ad.i11);
    const float4 c00 = __ldg(px + ac.i00), c01 = __ldg(px + ac.i01), c10 = __ldg(px + ac.i10), c11 = __ldg(px + ac.i11);
    const float4 bh = __ldg(px + ab.ih), br = __ldg(px + ab.ir), bb = __ldg(px + ab.ib);
    if (img_d) pre.diffuse = bilinear_mix(ad, d00, d01, d10, d11);
    else if (want_d) pre.diffuse = rgb(td.color[0], td.color[1], td.color[2]);
    if (img_c) pre.color = bilinear_mix(ac, c00, c01, c10, c11);
    else if (want_c) pre.color = rgb(tc.color[0], tc.color[1], tc.color[2]);
    if (img_b) slope_mix(bh, br, bb, right, bottom);
    pre.taps = (img_d ? 4u : 0u) + (img_c ? 4u : 0u) + (img_b ? 3u : 0u);
    return pre;
}

// value of a non-mix material
template <class SC>
static __device__ __forceinline__ RGB bxdf_value_leaf(const SC& S, const DevMaterial& m, V3 Vi, V3 Vr, V2 uv, TexPre pre) {
    switch (m.bxdf) {
    case RGK_BXDF_DIFFUSE: {
        if (Vi.z <= 0 || Vr.z <= 0) return rgb(0, 0, 0);
        const RGB c = tex_diffuse_of(S, m, uv, pre); return rgb(c.r / RGK_PI_F, c.g / RGK_PI_F, c.b / RGK_PI_F); }
    case RGK_BXDF_MIRROR: {
        const V3 refl = v3(-Vi.x, -Vi.y, Vi.z);
        if (fabsf(dot(refl, Vr) - 1) < 0.0001f) return tex_color_of(S, m, uv, pre);
        return rgb(0, 0, 0); }
    case RGK_BXDF_DIELECTRIC: {
        const float eta = (Vi.z < 0) ? m.ior : (float)(1.0 / (double)m.ior);
        float R, cosT; fresnel_dielectric(eta, Vi.z, R, cosT);
        const RGB c = tex_color_of(S, m, uv, pre);
        if (Vi.z * Vr.z > 0) {
            const V3 refl = v3(-Vi.x, -Vi.y, Vi.z);
            if (fabsf(dot(Vr, refl) - 1) < 0.001f) return rgb(c.r * R, c.g * R, c.b * R);
            return rgb(0, 0, 0);
        }
        const V3 refr = v3(-Vi.x * eta, -Vi.y * eta, (Vi.z > 0) ? -cosT : cosT);
        const float T = 1.0f - R;
        if (fabsf(dot(Vr, refr) - 1) < 0.001f) return rgb(c.r * T, c.g * T, c.b * T);
        return rgb(0, 0, 0); }
    case RGK_BXDF_TRANSPARENT: {
        const V3 inv = v3(-Vi.x, -Vi.y, -Vi.z);
        if (fabsf(dot(inv, Vr) - 1) < 0.0001f) return rgb(1, 1, 1);
        return rgb(0, 0, 0); }
    case RGK_BXDF_LTC_BECKMANN: case RGK_BXDF_LTC_GGX: {
        if (Vi.z <= 0 || Vr.z <= 0) return rgb(0, 0, 0);
        const RGB c = tex_color_of(S, m, uv, pre);
        const float p = ltc_pdf(S, m.bxdf == RGK_BXDF_LTC_GGX ? 0 : 1, Vi, Vr, m.roughness);
        return rgb(p * c.r, p * c.g, p * c.b); }
    case RGK_BXDF_LTC_BECKMANN_DIFFUSE: case RGK_BXDF_LTC_GGX_DIFFUSE: {
        if (Vi.z <= 0 || Vr.z <= 0) return rgb(0, 0, 0);
        const RGB diff = tex_diffuse_of(S, m, uv, pre), spec = tex_color_of(S, m, uv, pre);
        const float p = ltc_pdf(S, m.bxdf == RGK_BXDF_LTC_GGX_DIFFUSE ? 0 : 1, Vi, Vr, m.roughness);
        return rgb(p * spec.r + diff.r / RGK_PI_F, p * spec.g + diff.g / RGK_PI_F, p * spec.b + diff.b / RGK_PI_F); }
    }
    return rgb(0, 0, 0);
}
// BxDFMix::value (src/bxdf/bxdf.cpp:235-239) is a binary tree of lerps; evaluated without recursion by an
// explicit post-order walk (mix children always precede the mix material, so depth is bounded; cap 8).
// Kept out of line: mix materials are rare, and their walk needs indexable stacks (local memory) and a second copy of the
// nine-way leaf -- inlined into k_shade they cost every vertex registers, instruction-cache footprint and a call
// sequence around the common leaf.
static __device__ __noinline__ RGB bxdf_value_mix(const ShadeTables S, uint32_t mi, V3 Vi, V3 Vr, V2 uv) {
    TexPre none; none.have = false; none.diffuse = rgb(0, 0, 0); none.color = rgb(0, 0, 0); none.taps = 0u;
    // stack of (material, state): state 0 = visit a, 1 = visit b, 2 = combine
    uint32_t st_m[8]; int st_s[8]; RGB val[9]; int sp = 0, vp = 0;
    st_m[0] = mi; st_s[0] = 0; sp = 1;
    while (sp > 0) {
        const uint32_t cur = st_m[sp - 1];
        const DevMaterial cm = S.materials[cur];
        if (cm.bxdf != RGK_BXDF_MIX) { val[vp++] = bxdf_value_leaf(S, cm, Vi, Vr, uv, none); --sp; continue; }
        const int s = st_s[sp - 1];
        if (s == 0) { st_s[sp - 1] = 1; if (sp < 8) { st_m[sp] = (uint32_t)cm.mix_a; st_s[sp] = 0; ++sp; } else val[vp++] = rgb(0, 0, 0); }
        else if (s == 1) { st_s[sp - 1] = 2; if (sp < 8) { st_m[sp] = (uint32_t)cm.mix_b; st_s[sp] = 0; ++sp; } else val[vp++] = rgb(0, 0, 0); }
        else {
            const RGB b = val[--vp], a = val[--vp];
            const float w = cm.amount, v = 1.0f - cm.amount;
            val[vp++] = rgb(w * a.r + v * b.r, w * a.g + v * b.g, w * a.b + v * b.b);
            --sp;
        }
    }
    return val[0];
}
template <class SC>
__device__ __forceinline__ RGB bxdf_value(const SC& S, uint32_t mi, const DevMaterial& m, V3 Vi, V3 Vr, V2 uv, const TexPre& pre) {
    if (m.bxdf != RGK_BXDF_MIX) return bxdf_value_leaf(S, m, Vi, Vr, uv, pre);
    return bxdf_value_mix(shade_tables(S), mi, Vi, Vr, uv);
}
template <class SC>
__device__ __forceinline__ RGB bxdf_value(const SC& S, uint32_t mi, V3 Vi, V3 Vr, V2 uv) {
    const DevMaterial m = S.materials[mi];
    TexPre none; none.have = false; none.diffuse = rgb(0, 0, 0); none.color = rgb(0, 0, 0); none.taps = 0u;
    return bxdf_value(S, mi, m, Vi, Vr, uv, none);
}
// BxDF::sample: returns local direction, weight and may_leak
template <class SC>
__device__ __forceinline__ void bxdf_sample(const SC& S, DevMaterial m, V3 Vi, V2 uv, V2 sample, V3& dir, RGB& w, bool& may_leak, TexPre pre) {
    for (int guard = 0; m.bxdf == RGK_BXDF_MIX && guard < 16; guard++) {        // BxDFMix::sample, src/bxdf/bxdf.cpp:241-249
        m = S.materials[decide_and_rescale(sample.x, m.amount) ? m.mix_a : m.mix_b];
        pre.have = false;
    }
    may_leak = false;
    switch (m.bxdf) {
    case RGK_BXDF_DIFFUSE:
        if (Vi.z <= 0) { dir = v3(0, 1, 0); w = rgb(0, 0, 0); return; }
        dir = hemi_cos_z(sample); w = tex_diffuse_of(S, m, uv, pre); return;
    case RGK_BXDF_MIRROR:
        dir = v3(-Vi.x, -Vi.y, Vi.z); w = tex_color_of(S, m, uv, pre); return;
    case RGK_BXDF_DIELECTRIC: {
        const float eta = (Vi.z < 0) ? m.ior : (float)(1.0 / (double)m.ior);
        float R, cosT; fresnel_dielectric(eta, fabsf(Vi.z), R, cosT);
        const RGB c = tex_color_of(S, m, uv, pre);
        if (decide_and_rescale(sample.x, R)) { dir = v3(-Vi.x, -Vi.y, Vi.z); w = c; return; }
        cosT = fabsf(cosT);
        dir = v3(-Vi.x * eta, -Vi.y * eta, (Vi.z > 0) ? -cosT : cosT); w = c; may_leak = true; return; }
    case RGK_BXDF_TRANSPARENT:
        dir = v3(-Vi.x, -Vi.y, -Vi.z); w = rgb(1, 1, 1); may_leak = true; return;
    case RGK_BXDF_LTC_BECKMANN: case RGK_BXDF_LTC_GGX: {
        V3 v = hemi_cos_z(sample);
        v = ltc_random(S, m.bxdf == RGK_BXDF_LTC_GGX ? 0 : 1, Vi, m.roughness, v);
        dir = v;
        if (v.z <= 0) { w = rgb(0, 0, 0); return; }
        w = tex_color_of(S, m, uv, pre); return; }
    case RGK_BXDF_LTC_BECKMANN_DIFFUSE: case RGK_BXDF_LTC_GGX_DIFFUSE: {
        const RGB diff = tex_diffuse_of(S, m, uv, pre), spec = tex_color_of(S, m, uv, pre);
        const float dp = diff.r + diff.g + diff.b, sp = spec.r + spec.g + spec.b;
        const float prob = dp / (dp + sp + 0.0001f);
        if (decide_and_rescale(sample.x, prob)) {
            if (Vi.z <= 0) { dir = v3(0, 1, 0); w = rgb(0, 0, 0); return; }
            dir = hemi_cos_z(sample); w = diff; return;
        }
        V3 v = hemi_cos_z(sample);
        v = ltc_random(S, m.bxdf == RGK_BXDF_LTC_GGX_DIFFUSE ? 0 : 1, Vi, m.roughness, v);
        dir = v;
        if (v.z <= 0) { w = rgb(0, 0, 0); return; }
        w = spec; return; }
    }
    dir = v3(0, 1, 0); w = rgb(0, 0, 0);
}

template <class SC>
__device__ __forceinline__ void bxdf_sample(const SC& S, uint32_t mi, V3 Vi, V2 uv, V2 sample, V3& dir, RGB& w, bool& may_leak) {
    TexPre none; none.have = false; none.diffuse = rgb(0, 0, 0); none.color = rgb(0, 0, 0); none.taps = 0u;
    bxdf_sample(S, S.materials[mi], Vi, uv, sample, dir, w, may_leak, none);
}

// ---- lights (src/scene.cpp:686-745, src/primitives.cpp:61-73) and sky (src/scene.cpp:748-763)
struct LightRec { int type; bool valid; V3 pos; RGB color; float intensity; float size; V3 normal; };
__device__ __forceinline__ LightRec random_light(const DevScene& S, V2 choice, float light_sample, V2 tri_sample) {
    LightRec none; none.type = 0; none.valid = false; none.pos = v3(0, 0, 0); none.color = rgb(0, 0, 0); none.intensity = 0; none.size = 0; none.normal = v3(0, 0, 0);
    const float total = S.total_point_power + S.total_areal_power;
    if (total <= 0.0f) return none;
    float q = choice.x * total;
    if (q < S.total_point_power) {
        for (uint32_t i = 0; i < S.n_point_lights; i++) {
            const DevPointLight p = S.point_lights[i];
            q -= p.intensity * 4.0f * RGK_PI_F;
            if (q <= 0.0f) {
                LightRec l; l.type = 0; l.valid = true; l.pos = v3(p.pos); l.color = rgb(p.color[0], p.color[1], p.color[2]);
                l.intensity = p.intensity; l.size = p.size; l.normal = v3(0, 0, 0); return l;
            }
        }
        return none;
    }
    q = choice.y * S.total_areal_power;
    for (uint32_t i = 0; i < S.n_areal_lights; i++) {
        const DevArealLight al = S.areal_lights[i];
        q -= al.power;
        if (q <= 0.0f) {
            float p = light_sample * al.total_area;
            for (uint32_t j = 0; j < al.count; j++) {
                const DevArealTri at = S.areal_tris[al.first + j];
                p -= at.area;
                if (p <= 0.0f) {
                    const uint4 tv = __ldg(S.tri_shade + at.tri);
                    V2 r = tri_sample;
                    const V3 a = v3(__ldg(S.positions + tv.x)), c = v3(__ldg(S.positions + tv.y)), b = v3(__ldg(S.positions + tv.z));
                    const V3 Va = a - c, Vb = b - c;
                    if (r.x + r.y > 1.0f) { r.x = 1.0f - r.x; r.y = 1.0f - r.y; }
                    LightRec l; l.type = 1; l.valid = true; l.pos = c + r.x * Va + r.y * Vb;
                    l.color = rgb(al.emission[0], al.emission[1], al.emission[2]); l.intensity = 1.0f; l.size = 0.0f;
                    l.normal = v3(__ldg(S.normals + tv.x)); return l;
                }
            }
            return none;
        }
    }
    return none;
}
__device__ __forceinline__ RGB sky_radiance(const DevScene& S, V3 dir) {
    if (S.sky_mode == 0) return rgb(S.sky_color[0] * S.sky_intensity, S.sky_color[1] * S.sky_intensity, S.sky_color[2] * S.sky_intensity);
    const float alpha = asinf(dir.y);
    float beta = -atan2f(dir.x, dir.z);
    beta += S.sky_rotate * 0.0174533f;
    const float x = beta / (2.0f * RGK_PI_F) + 0.5f, y = alpha / RGK_PI_F + 0.5f;
    const RGB c = tex_fetch(S, S.sky_envmap, V2{x, y});
    return rgb(c.r * S.sky_intensity, c.g * S.sky_intensity, c.b * S.sky_intensity);
}
