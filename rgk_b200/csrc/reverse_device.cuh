// reverse_device.cuh -- the bidirectional part of PathTracer::TracePath (reverse > 0, src/path_tracer.cpp:336-398,
// 462-480): a light path of up to `reverse` vertices leaves the sample's light, every light vertex is connected to the
// camera (a "side effect": radiance splatted into another pixel with count 0, src/tracer.cpp:20-26) and to every vertex
// of the camera path.  Included by render.cu; uses its PathBuffers / RenderConst / SamplerView / push_queue.
//
// Unlike the unidirectional case a sample's terms cannot be accumulated as they appear: a camera vertex's radiance
// (direct light + connections + emission) is clamped as a whole, and the light path draws its sampler dimensions AFTER
// the camera path (one 2-D dimension per camera vertex), so the camera path has to be finished and kept first.
//   camera bounces : k_closest -> k_shade_rev<false> (vertex record + NEE set-up) -> k_shadow_rev (direct term)
//   k_lightgen     : light ray (Sample2DToHemisphereCosineDirected), light_at_path_start, first light dimension
//   light bounces  : k_closest -> k_shade_rev<true> (light vertex record, russian = -1, depth = reverse)
//   k_connect_camera (splats, atomics), k_connect_vertices (V x L visibility + two BxDF values each), k_assemble
// This mode favours a direct transcription over speed: connection rays are traced one per thread.
#pragma once

constexpr uint32_t VR_INFINITY = 0x80000000u;   // flag in vr_pos.w (material index in the low bits)

// the geometric half of one GeneratePath vertex (src/path_tracer.cpp:150-236); false = the reference returns early
struct VertexGeo { V3 pos, faceN, lightN; V2 uv; uint32_t mat_id; DevMaterial mat; TexPre pre; };
__device__ __forceinline__ bool vertex_geometry(const DevScene& S, const RenderConst& R, uint32_t tri, float4 hit, V3 ro, V3 rd, VertexGeo& g) {
    const uint4 tv = __ldg(S.tri_shade + tri);
    const float ia = 1.0f - hit.y - hit.z, ib = hit.y, ic = hit.z;
    g.pos = ro + hit.x * rd;
    const V3 nA = v3(__ldg(S.normals + tv.x)), nB = v3(__ldg(S.normals + tv.y)), nC = v3(__ldg(S.normals + tv.z));
    V3 faceN = ia * nA + ib * nB + ic * nC;
    if (isnan(faceN.x)) { faceN = nA; if (isnan(faceN.x)) { faceN = nB; if (isnan(faceN.x)) { faceN = nC; if (isnan(faceN.x)) return false; } } }
    if (length(faceN) <= 0.0f) return false;
    faceN = normalize(faceN);
    g.faceN = faceN;
    g.mat_id = tv.w;
    g.mat = S.materials[tv.w];
    const float2 ta = __ldg(S.texcoords + tv.x), tb = __ldg(S.texcoords + tv.y), tc = __ldg(S.texcoords + tv.z);
    g.uv = V2{ia * ta.x + ib * tb.x + ic * tc.x, ia * ta.y + ib * tb.y + ic * tc.y};
    g.lightN = faceN;
    float right, bottom;
    g.pre = vertex_textures(S, g.mat, g.uv, right, bottom);
    if (g.mat.tex_bump >= 0) {
        V3 tangent = ia * v3(__ldg(S.tangents + tv.x)) + ib * v3(__ldg(S.tangents + tv.y)) + ic * v3(__ldg(S.tangents + tv.z));
        if (!(tangent.x * tangent.x + tangent.y * tangent.y + tangent.z * tangent.z < 0.001f)) {
            tangent = normalize(tangent);
            const V3 bitangent = normalize(cross(faceN, tangent));
            const V3 tangent2 = cross(bitangent, faceN);
            g.lightN = normalize(faceN + (tangent2 * right + bitangent * bottom) * R.bump_scale);
            if (isnan(g.lightN.x)) g.lightN = faceN;
        }
    }
    return true;
}

// Scene::Visibility(a, b) for one segment, one thread (src/scene.cpp:670-673): same ray as k_shadow builds
__device__ __forceinline__ bool segment_visible(const DevScene& S, V3 a, V3 b) {
    const float ex = b.x - a.x, ey = b.y - a.y, ez = b.z - a.z;
    const float d2 = ex * ex + ey * ey + ez * ez;
    const float inv = 1.0f / sqrtf(d2), len = sqrtf(d2);
    const float e20 = S.epsilon * 20.0f;
    Traverser<true, false> T;
    TravStack K;
    TravCount c{0, 0, 0, 0, 0, 0, 0};
    if (!T.init(S, a.x, a.y, a.z, ex * inv, ey * inv, ez * inv, 0.0f + e20, len - e20, RGK_NO_TRIANGLE)) return true;
    for (;;) {
        const uint2 w = T.descend(S, K, c);
        if (T.leaf(S, w, c)) return false;
        if (!T.pop(K)) return true;
    }
}

// One vertex of a camera path (LIGHT = false) or of a light path (LIGHT = true) in bidirectional mode.
template <bool LIGHT>
__global__ void __launch_bounds__(128, 4)
k_shade_rev(DevScene S, RenderConst R, SamplerView smp, PathBuffers B, ReverseBuffers V, const uint32_t* __restrict__ queue, uint32_t count,
            uint32_t* __restrict__ next_queue, uint32_t* __restrict__ shadow_queue, unsigned long long* counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool cont = false, shadow = false, null_shadow = false;
    uint32_t slot = 0;
    if (i < count) {
        slot = queue ? __ldg(queue + i) : i;
        const float4 hit = B.hit[slot];
        const uint32_t tri = __float_as_uint(hit.w);
        const V3 ro = v3(B.ray_o[slot]), rd = v3(B.ray_d[slot]);
        const float4 cum4 = B.cum[slot];
        uint32_t n = __float_as_uint(cum4.w) + 1u;                 // 1-based index of this vertex on its path
        const size_t idx = (size_t)(n - 1u) * R.npaths + slot;
        const uint32_t pixel = slot % R.npix, set = slot / R.npix;
        const uint32_t seed = B.pix_seed[pixel];
        const RGB contribution = rgb(cum4.x, cum4.y, cum4.z);
        const V3 Vr = -rd;
        const uint32_t max_depth = LIGHT ? R.reverse : R.depth;
        if (tri == RGK_NO_TRIANGLE) {
            if (!LIGHT) {                                          // sky vertex: kept for k_assemble (contribution * sky, no clamp)
                const RGB sky = sky_radiance(S, Vr);
                V.vr_pos[idx] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(VR_INFINITY));
                V.vr_here[idx] = make_float4(sky.r, sky.g, sky.b, 0.0f);
                V.vr_con[idx] = make_float4(contribution.r, contribution.g, contribution.b, 0.0f);
                V.nverts[slot] = n;
            }                                                      // an infinity vertex of the light path takes part in nothing
        } else {
            VertexGeo g;
            if (vertex_geometry(S, R, tri, hit, ro, rd, g)) {
                const Frame fr = system_transform_z(g.lightN);
                const V3 VrL = qrot(fr.g2l, Vr);
                if (LIGHT) {
                    const float4 ls = V.lstart[slot];
                    V.lr_pos[idx] = make_float4(g.pos.x, g.pos.y, g.pos.z, __uint_as_float(g.mat_id));
                    V.lr_nrm[idx] = make_float4(g.lightN.x, g.lightN.y, g.lightN.z, 0.0f);
                    V.lr_vr[idx] = make_float4(Vr.x, Vr.y, Vr.z, 0.0f);
                    V.lr_uv[idx] = make_float4(g.uv.x, g.uv.y, 0.0f, 0.0f);
                    V.lr_lfs[idx] = make_float4(contribution.r * ls.x, contribution.g * ls.y, contribution.b * ls.z, 0.0f);   // p.contribution * light_at_path_start
                    V.nlverts[slot] = n;
                } else {
                    RGB emis = rgb(0.0f, 0.0f, 0.0f);
                    if (dot(g.faceN, Vr) > 0) emis = rgb(g.mat.emission[0], g.mat.emission[1], g.mat.emission[2]);
                    V.vr_pos[idx] = make_float4(g.pos.x, g.pos.y, g.pos.z, __uint_as_float(g.mat_id));
                    V.vr_nrm[idx] = make_float4(g.lightN.x, g.lightN.y, g.lightN.z, 0.0f);
                    V.vr_vr[idx] = make_float4(Vr.x, Vr.y, Vr.z, 0.0f);
                    V.vr_uv[idx] = make_float4(g.uv.x, g.uv.y, 0.0f, 0.0f);
                    V.vr_con[idx] = make_float4(contribution.r, contribution.g, contribution.b, 0.0f);
                    V.vr_here[idx] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    V.vr_emis[idx] = make_float4(emis.r, emis.g, emis.b, 0.0f);
                    V.nverts[slot] = n;
                    // direct light (src/path_tracer.cpp:427-460); the visibility test runs in k_shadow_rev
                    const float4 lp4 = B.light_pos[slot];
                    const uint32_t lflags = __float_as_uint(lp4.w);
                    if (lflags & 1u) {
                        const V3 lpos = v3(lp4);
                        const float4 lc = B.light_col[slot];
                        const V3 Vi = normalize(lpos - g.pos);
                        const RGB f = bxdf_value(S, g.mat_id, g.mat, qrot(fr.g2l, Vi), VrL, g.uv, g.pre);
                        const V3 dlt = lpos - g.pos;
                        const float G = fabsf(dot(g.lightN, Vi)) / dot(dlt, dlt);
                        float df = 1.0f;
                        if (lflags & 2u) df = gmax(0.0f, dot(-Vi, v3(B.light_nrm[slot])));
                        const float k = lc.w * df;
                        const RGB inc = rgb(lc.x * k, lc.y * k, lc.z * k);
                        const RGB direct = rgb(inc.r * (G * f.r), inc.g * (G * f.g), inc.b * (G * f.b));
                        if (!(R.skip_null_shadow && direct.r == 0.0f && direct.g == 0.0f && direct.b == 0.0f)) {
                            B.sh_pos[slot] = make_float4(g.pos.x, g.pos.y, g.pos.z, __uint_as_float(n));
                            B.sh_direct[slot] = make_float4(direct.r, direct.g, direct.b, 0.0f);
                            shadow = true;
                        } else null_shadow = true;
                    }
                }
                // ---- continuation (src/path_tracer.cpp:238-302); the light path runs with russian = -1
                if (n < max_depth) {
                    const uint32_t dim = LIGHT ? V.d2base[slot] + (n - 1u) : R.base2 + (n - 1u);
                    const V2 sample = smp.get2d(pixel, seed, set, dim);
                    V3 dir; RGB tcf; bool may_leak;
                    bxdf_sample(S, g.mat, VrL, g.uv, sample, dir, tcf, may_leak, g.pre);
                    const bool inside = dir.z < 0;
                    dir = qrot(fr.l2g, dir);
                    if (!(dot(dir, g.faceN) * dot(Vr, g.faceN) > 0) && !may_leak) n += 10000u;
                    const float russian = LIGHT ? -1.0f : R.russian;
                    const float rcoef = (!g.mat.no_russian && russian > 0.0f && n > 1u) ? 1.0f / russian : 1.0f;
                    RGB cum = rgb(rcoef * contribution.r, rcoef * contribution.g, rcoef * contribution.b);
                    cum = rgb(tcf.r * cum.r, tcf.g * cum.g, tcf.b * cum.b);
                    cont = true;
                    if (gmax(gmax(cum.r, cum.g), cum.b) < 0.001f) cont = false;
                    if (cont && !g.mat.no_russian && russian >= 0.0f) {
                        const uint32_t c1 = B.cur1[slot];
                        B.cur1[slot] = c1 + 1u;
                        if (smp.get1d(pixel, seed, set, c1 < R.n1d ? c1 : 0u) > russian) cont = false;
                    }
                    if (cont && n > max_depth) cont = false;
                    if (cont && !(n < max_depth)) cont = false;
                    if (cont) {
                        const V3 no = g.pos + g.faceN * S.epsilon * 10.0f * (inside ? -1.0f : 1.0f);
                        const V3 nd = normalize(normalize(dir));
                        B.ray_o[slot] = make_float4(no.x, no.y, no.z, 0.0f);
                        B.ray_d[slot] = make_float4(nd.x, nd.y, nd.z, 0.0f);
                        B.cum[slot] = make_float4(cum.r, cum.g, cum.b, __uint_as_float(n));
                        B.last_tri[slot] = tri;
                    }
                }
            }
        }
    }
    push_queue(next_queue, counters + C_NEXT, cont, slot);
    if (!LIGHT) {
        push_queue(shadow_queue, counters + C_SHADOW, shadow, slot);
        const unsigned m = __ballot_sync(0xffffffffu, null_shadow);
        if (m && (threadIdx.x & 31) == 0) atomicAdd(counters + C_SHADOW_SKIPPED, (unsigned long long)__popc(m));
    }
}

// direct term of a camera vertex: visibility only, the sum is assembled (and clamped) in k_assemble
__global__ void __launch_bounds__(TRACE_THREADS, 8)
k_shadow_rev(DevScene S, RenderConst R, PathBuffers B, ReverseBuffers V, const uint32_t* __restrict__ queue, uint32_t count, unsigned long long* work) {
    TravCount cnt{0, 0, 0, 0, 0, 0, 0};
    uint32_t mine = 0;
    trace_rays<6, true, false>(S, count, work, cnt, mine,
        [&](uint32_t i, Traverser<true, false>& T) {
            const uint32_t slot = __ldg(queue + i);
            const float4 a = B.light_pos[slot], b = B.sh_pos[slot];
            const float ex = b.x - a.x, ey = b.y - a.y, ez = b.z - a.z;
            const float d2 = ex * ex + ey * ey + ez * ez;
            const float inv = 1.0f / sqrtf(d2), len = sqrtf(d2);
            const float e20 = S.epsilon * 20.0f;
            return T.init(S, a.x, a.y, a.z, ex * inv, ey * inv, ez * inv, 0.0f + e20, len - e20, RGK_NO_TRIANGLE);
        },
        [&](uint32_t i, bool blocked, const HitRec&) {
            const uint32_t slot = __ldg(queue + i);
            const uint32_t n = __float_as_uint(B.sh_pos[slot].w);
            const float4 dr = B.sh_direct[slot];
            V.vr_here[(size_t)(n - 1u) * R.npaths + slot] = make_float4(blocked ? 0.0f : dr.x, blocked ? 0.0f : dr.y, blocked ? 0.0f : dr.z, 0.0f);
        });
}

// the light path's first ray (src/path_tracer.cpp:336-349) and what it carries (:360-364)
__global__ void k_lightgen(DevScene S, RenderConst R, SamplerView smp, PathBuffers B, ReverseBuffers V, uint32_t* __restrict__ queue, unsigned long long* counters) {
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool go = false;
    if (slot < R.npaths) {
        V.nlverts[slot] = 0u;
        const float4 lp4 = B.light_pos[slot];
        const uint32_t lflags = __float_as_uint(lp4.w);
        if (lflags & 1u) {
            const uint32_t pixel = (uint32_t)(slot % R.npix), set = (uint32_t)(slot / R.npix);
            const uint32_t seed = B.pix_seed[pixel];
            const uint32_t d_areal = 1u + R.lens;
            const V2 areal = smp.get2d(pixel, seed, set, d_areal), lightdir = smp.get2d(pixel, seed, set, d_areal + 1u);
            const V3 normal = v3(B.light_nrm[slot]);
            V3 mdir;
            float dfac = 1.0f;
            if (lflags & 2u) { mdir = hemi_cos_directed(lightdir, normal); dfac = gmax(0.0f, dot(mdir, normal)); }   // HEMISPHERE
            else mdir = hemi_cos_directed(lightdir, normalize(sphere_uniform(areal)));                                // FULL_SPHERE
            const float4 lc = B.light_col[slot];
            const float k0 = lc.w * dfac;
            V.lstart[slot] = make_float4(lc.x * k0, lc.y * k0, lc.z * k0, 0.0f);
            const V3 o = v3(lp4) + S.epsilon * normal * 100.0f;
            const V3 d = normalize(mdir);
            B.ray_o[slot] = make_float4(o.x, o.y, o.z, 0.0f);
            B.ray_d[slot] = make_float4(d.x, d.y, d.z, 0.0f);
            B.cum[slot] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(0u));
            B.last_tri[slot] = RGK_NO_TRIANGLE;
            // every non-infinity camera vertex drew one 2-D sample (src/path_tracer.cpp:238), the light path continues there
            const uint32_t nv = V.nverts[slot];
            uint32_t drawn = nv;
            if (nv && (__float_as_uint(V.vr_pos[(size_t)(nv - 1u) * R.npaths + slot].w) & VR_INFINITY)) drawn--;
            V.d2base[slot] = R.base2 + drawn;
            go = true;
        }
    }
    push_queue(queue, counters + C_NEXT, go, (uint32_t)slot);
}

// side effects: light vertex -> camera (src/path_tracer.cpp:377-397); one thread per (path, light vertex)
__global__ void __launch_bounds__(128, 4)
k_connect_camera(DevScene S, RenderConst R, PathBuffers B, ReverseBuffers V, float* __restrict__ fb, unsigned long long* counters) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t slot = t % R.npaths;
    const uint32_t b = (uint32_t)(t / R.npaths);
    uint32_t traced = 0;
    if (b < R.reverse && b < V.nlverts[slot]) {
        const size_t idx = (size_t)b * R.npaths + slot;
        const float4 p4 = V.lr_pos[idx];
        const V3 pos = v3(p4), camerapos = v3(V.cam_o[slot]);
        traced = 1;
        if (segment_visible(S, pos, camerapos)) {
            const V3 direction = normalize(pos - camerapos);
            const V3 lightN = v3(V.lr_nrm[idx]), Vr = v3(V.lr_vr[idx]);
            const float4 uv4 = V.lr_uv[idx], lfs = V.lr_lfs[idx];
            const Frame fr = system_transform_z(lightN);
            const RGB f = bxdf_value(S, __float_as_uint(p4.w), qrot(fr.g2l, Vr), qrot(fr.g2l, -direction), V2{uv4.x, uv4.y});
            RGB q = rgb(lfs.x * f.r, lfs.y * f.g, lfs.z * f.b);
            const V3 dlt = camerapos - pos;
            const float G = gmax(0.0f, dot(lightN, -direction)) / dot(dlt, dlt);
            if (G >= 0.00001f && !isnan(q.r)) {
                q = rgb(q.r * G, q.g * G, q.b * G);
                // Camera::GetCoordsFromDirection, src/camera.cpp:48-83
                const V3 N = v3(R.cam.direction);
                const float qd = dot(direction, N);
                if (!((double)qd < 0.0001)) {
                    const float tt = dot(v3(R.cam.viewscreen) - v3(R.cam.origin), N) / qd;
                    if (!(tt <= 0)) {
                        const V3 p = v3(R.cam.origin) + direction * tt;
                        const V3 v1 = v3(R.cam.viewscreen_x), v2 = v3(R.cam.viewscreen_y);
                        const V3 vp = p - v3(R.cam.viewscreen);
                        const float plen = length(vp);
                        const float c1 = plen * dot(normalize(vp), normalize(v1)), c2 = plen * dot(normalize(vp), normalize(v2));
                        const float xr = c1 / length(v1), yr = c2 / length(v2);
                        if (!(xr < 0.0f || xr > 1.0f || yr < 0.0f || yr > 1.0f)) {
                            const int x2 = (int)(R.cam.xsize * xr), y2 = (int)(R.cam.ysize * yr);
                            if (x2 < R.cam.xsize && y2 < R.cam.ysize) {      // (ratio == 1 addresses one pixel past the row upstream)
                                float* px = fb + 3 * ((size_t)y2 * R.xres + x2);
                                atomicAdd(px, q.r); atomicAdd(px + 1, q.g); atomicAdd(px + 2, q.b);
                            }
                        }
                    }
                }
            }
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, traced != 0);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(counters + C_SHADOW, (unsigned long long)__popc(m));
}

// "Reverse light": camera vertex x light vertex (src/path_tracer.cpp:462-480); one thread per (path, camera vertex)
__global__ void __launch_bounds__(128, 4)
k_connect_vertices(DevScene S, RenderConst R, ReverseBuffers V, unsigned long long* counters) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t slot = t % R.npaths;
    const uint32_t n = (uint32_t)(t / R.npaths);
    uint32_t traced = 0;
    if (n < R.depth && n < V.nverts[slot]) {
        const size_t idx = (size_t)n * R.npaths + slot;
        const float4 p4 = V.vr_pos[idx];
        const uint32_t nl = V.nlverts[slot];
        if (!(__float_as_uint(p4.w) & VR_INFINITY) && nl) {
            const V3 pos = v3(p4), lightN = v3(V.vr_nrm[idx]), Vr = v3(V.vr_vr[idx]);
            const float4 uv4 = V.vr_uv[idx];
            const Frame fr = system_transform_z(lightN);
            const V3 VrL = qrot(fr.g2l, Vr);
            float4 here = V.vr_here[idx];
            for (uint32_t b = 0; b < nl; b++) {
                const size_t li = (size_t)b * R.npaths + slot;
                const float4 l4 = V.lr_pos[li];
                const V3 lpos = v3(l4);
                traced++;
                if (!segment_visible(S, lpos, pos)) continue;
                const V3 light_to_p = normalize(pos - lpos), p_to_light = -light_to_p;
                const V3 lN = v3(V.lr_nrm[li]), lVr = v3(V.lr_vr[li]);
                const float4 luv = V.lr_uv[li], lfs = V.lr_lfs[li];
                const Frame lf = system_transform_z(lN);
                const RGB f_light = bxdf_value(S, __float_as_uint(l4.w), qrot(lf.g2l, light_to_p), qrot(lf.g2l, lVr), V2{luv.x, luv.y});
                const RGB f_point = bxdf_value(S, __float_as_uint(p4.w), VrL, qrot(fr.g2l, p_to_light), V2{uv4.x, uv4.y});
                const V3 dlt = lpos - pos;
                const float G = fabsf(dot(lightN, p_to_light)) / dot(dlt, dlt);
                RGB ff = rgb(f_point.r * f_light.r, f_point.g * f_light.g, f_point.b * f_light.b);
                ff = rgb(G * ff.r, G * ff.g, G * ff.b);
                here.x += lfs.x * ff.r; here.y += lfs.y * ff.g; here.z += lfs.z * ff.b;
            }
            V.vr_here[idx] = here;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) traced += __shfl_xor_sync(0xffffffffu, traced, o);
    if (traced && (threadIdx.x & 31) == 0) atomicAdd(counters + C_SHADOW, (unsigned long long)traced);
}

// "Calculate light transmitted over view path" (src/path_tracer.cpp:400-500): per vertex direct + connections + emission,
// clamped, times the vertex's contribution; sky vertices unclamped
__global__ void k_assemble(RenderConst R, PathBuffers B, ReverseBuffers V) {
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= R.npaths) return;
    const uint32_t nv = V.nverts[slot];
    float tr = 0.0f, tg = 0.0f, tb = 0.0f;
    for (uint32_t n = 0; n < nv; n++) {
        const size_t idx = (size_t)n * R.npaths + slot;
        const float4 c = V.vr_con[idx], h = V.vr_here[idx];
        if (__float_as_uint(V.vr_pos[idx].w) & VR_INFINITY) { tr += h.x * c.x; tg += h.y * c.y; tb += h.z * c.z; continue; }
        const float4 e = V.vr_emis[idx];
        float r = h.x + e.x, g = h.y + e.y, b = h.z + e.z;
        if (r > R.clamp) r = R.clamp;
        if (g > R.clamp) g = R.clamp;
        if (b > R.clamp) b = R.clamp;
        tr += r * c.x; tg += g * c.y; tb += b * c.z;
    }
    B.tot[slot] = make_float4(tr, tg, tb, 0.0f);
}
