// render.cu -- the unidirectional bounce loop as a wavefront of kernels
// (replaces PathTracer::RenderPixel / TracePath / GeneratePath, src/path_tracer.cpp:42-512,
// Tracer::Render, src/tracer.cpp:6-37, and the per-task body of RenderDriver::RenderRound,
// src/render_driver.cpp:158-184).
//
// One call renders a list of tasks (tiles).  Tiles are processed in chunks sized to HBM; a
// chunk holds every multisample of its pixels: path slot = sample * npix + pixel, so that
// adjacent threads are adjacent pixels of one tile row (coherent primary rays, coalesced
// state and sampler-table access).  Per chunk:
//   pixel_setup -> sampler tables (device mt19937 replica) -> raygen (+ light pick)
//   repeat per bounce: closest-hit traversal -> shade (vertex set-up, NEE set-up, BxDF
//   sample, termination, warp-aggregated compaction into the next queue) -> shadow traversal
//   fused with the NEE resolve (visibility * direct + emission, clamp, accumulate)
//   finish: per pixel, samples summed in order and added to the framebuffer.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include "trace_device.cuh"
#include "bvh_device.cuh"
#include "shade_device.cuh"
#ifdef RGK_DOH_DEBUG
#include <cstdio>
static long long doh_debug_slot = std::getenv("DOH_SLOT") ? std::atoll(std::getenv("DOH_SLOT")) : -1;
#endif

// ------------------------------------------------------------------ buffers
// vertex storage of the bidirectional mode (reverse_device.cuh)
struct ReverseBuffers {
    float4 *cam_o = nullptr, *lstart = nullptr;                      // [slot] camera ray origin; light_at_path_start
    uint32_t *nverts = nullptr, *nlverts = nullptr, *d2base = nullptr;   // [slot] camera / light vertex records; first light 2-D dim
    // camera vertex n (0-based) of path slot at [n * npaths + slot]
    float4 *vr_pos = nullptr, *vr_nrm = nullptr, *vr_vr = nullptr, *vr_uv = nullptr, *vr_con = nullptr, *vr_here = nullptr, *vr_emis = nullptr;
    // light vertex b of path slot at [b * npaths + slot]
    float4 *lr_pos = nullptr, *lr_nrm = nullptr, *lr_vr = nullptr, *lr_uv = nullptr, *lr_lfs = nullptr;
    size_t cap_paths = 0, cap_cam = 0, cap_light = 0;
};
struct PathBuffers {
    size_t cap_paths = 0, cap_pixels = 0, cap_t1 = 0, cap_t2 = 0, cap_tiles = 0, cap_mt = 0, cap_bounces = 0;
    float4 *ray_o = nullptr, *ray_d = nullptr, *hit = nullptr, *cum = nullptr, *tot = nullptr;
    float4 *light_pos = nullptr, *light_col = nullptr, *light_nrm = nullptr;
    float4 *sh_pos = nullptr, *sh_direct = nullptr, *sh_emis = nullptr, *sh_contrib = nullptr;
    uint32_t *last_tri = nullptr, *cur1 = nullptr, *queue_a = nullptr, *queue_b = nullptr, *queue_s = nullptr;
    uint32_t *queue_ua = nullptr, *queue_ub = nullptr;    // the live queue once more in path order (k_shade reads its state through it)
    uint8_t *key_next = nullptr, *key_shadow = nullptr;   // per-slot direction bin of the continuation / shadow ray, 0xFF = none (binned queues)
    uint32_t *pix_xy = nullptr, *pix_seed = nullptr, *pix_src = nullptr;   // pix_src: index of the pixel in call order (task order, y-major)
    uint32_t *mt_state = nullptr;
    float *t1 = nullptr; float2 *t2 = nullptr;
    uint4 *tiles = nullptr; uint2 *tiles2 = nullptr;
    unsigned long long *counters = nullptr;   // device
    unsigned long long *shade_counts = nullptr;   // device, 8 words: work counters of k_shade in counting rounds (rgk_render_get_shade_stats)
    unsigned long long *h_counters = nullptr; // pinned: H_CHUNKS chunks x cap_bounces bounces x C_COUNT words, then cap_bounces words of early read-backs
    std::vector<cudaEvent_t>* term_events = nullptr;   // host only: completion of bounce b's early read-back
    struct EventPool* events = nullptr;       // host only
    ReverseBuffers* reverse = nullptr;        // host only (bidirectional mode)
};

// One block of counters per bounce (C_COUNT words, zeroed once per chunk): the queue lengths a bounce produces are read by the
// next kernels ON THE DEVICE (its own shadow launch, the next bounce), so the host never waits for them inside a chunk.
enum { C_NEXT = 0, C_SHADOW = 1, C_WORK_A = 2, C_WORK_B = 3, C_SHADOW_SKIPPED = 4, C_NEXT_U = 5,
       C_ARB_WORK = 8, C_ARB_COUNT = 9, C_ARB_SWORK = 10, C_ARB_SCOUNT = 11, C_COUNT = 16 };
// What a kernel of bounce b needs to know about a queue length: a device word (written by the previous kernels) or, where the
// host knows it (camera rays, the bidirectional mode), a value.  Binning decisions that depend on a length are taken on the
// device from the same words: a bounce is binned when its live paths number at least bin_thresh.
struct QueueLen {
    const unsigned long long* ptr; uint32_t val;
    __device__ __forceinline__ uint32_t get() const { return ptr ? (uint32_t)*ptr : val; }
};

struct RenderConst {
    rgk_camera cam;
    uint32_t xres, yres, ms, depth;
    float clamp, russian, bump_scale;
    uint32_t set_size, n1d, n2d, base2, sampler_mode, lens, skip_null_shadow;
    uint32_t binning;       // what MAY be binned this bounce (bit 0 continuation rays, bit 1 shadow rays): k_shade writes direction-bin keys and
                            // k_bin builds the queues (coherence reordering) if the bounce has at least bin_thresh live paths
    uint32_t bin_thresh;
    uint32_t first_bounce;  // 1: k_shade of the camera rays' hit points in the unidirectional loop -- throughput (1, 1, 1), no bounce done,
                            // Russian-roulette cursor 1, no radiance yet: the initial path state is known, k_raygen does not store it and
                            // this launch does not load it (48 B per path less through HBM)
    uint32_t npix;          // pixels in the chunk
    uint32_t const_light;   // 1: the scene's only light is one point light of size 0 -- every sample picks the same light record
    float4 cl_pos, cl_col;  //    (position + flags, colour + intensity), read from here instead of per-path arrays
    uint32_t count_shade;   // 1: counting round, k_shade adds its work counters to PathBuffers::shade_counts
    uint32_t reverse;       // light path length (bidirectional mode, reverse_device.cuh); 0 = unidirectional
    uint32_t npaths;        // npix * ms
};

namespace {

template <class T> bool alloc_dev(T** p, size_t n) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    return cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)) == cudaSuccess;
}

// ------------------------------------------------------------------ sampler: device replica of StratifiedSampler
// (src/sampler.cpp:5-36,85-116) over libstdc++'s mt19937 / generate_canonical / uniform_int_distribution (Lemire)
// / std::shuffle (pairwise) -- SURVEY Appendix C.  One thread per pixel; the 624-word generator state lives in
// global memory, interleaved across threads (state[k * stride + thread]) so every access is coalesced.
constexpr int MT_LANES = 128;   // threads per CTA of the sampler kernels = width of a state block

struct MT {
    uint32_t* st;      // this thread's column of its CTA's state block: word k at st[k * MT_LANES] (32-bit offsets, coalesced)
    int k0, cur;
    uint32_t buf[8];
    uint32_t la, lb;   // first generation only: seed-recurrence words k0 and k0 + 397, kept in registers
    bool gen0;
    static __device__ __forceinline__ uint32_t lcg(uint32_t prev, uint32_t i) { return 1812433253u * (prev ^ (prev >> 30)) + i; }
    // mt19937::seed(s): word_0 = s, word_i = 1812433253 * (word_{i-1} ^ (word_{i-1} >> 30)) + i.  The seeded state is never
    // written out: its words are only inputs of the first twist, and that reads them in ascending order at two places
    // (k and k + 397), so two running registers replace the store pass and most loads of the first generation.
    __device__ void seed(uint32_t s) {
        la = s; lb = s;
        for (uint32_t i = 1; i <= 397; i++) lb = lcg(lb, i);
        gen0 = true; k0 = 0; cur = 8;
    }
    // The twist is done incrementally, eight words at a time, right before their tempered values are handed out (the
    // standard implementation twists all 624 words when the block is exhausted; word k only depends on old k, old k+1
    // -- new word 0 for k = 623 -- and word k+397 mod 624, old for k < 227 and already renewed for k >= 227, so doing
    // it in ascending batches yields identical words).  A batch first loads its 17 inputs, then stores its 8 outputs:
    // one memory round trip per eight draws instead of one per draw.  Only the batches at k0 = 224 (the k+397 window
    // crosses the end of the state) and k0 = 616 (word k+1 of the last word is word 0) need per-word wrap-around;
    // all others address their words at constant offsets from one pointer.
    __device__ void fill() {
        uint32_t own[9], far[8];
        uint32_t* p = st + k0 * MT_LANES;
        if (gen0) {
            own[0] = la;
#pragma unroll
            for (int j = 1; j < 9; j++) own[j] = lcg(own[j - 1], (uint32_t)(k0 + j));
            la = own[8];
            if (k0 == 616) own[8] = st[0];                              // word "624" of the last batch is the NEW word 0
            if (k0 < 224) {
                far[0] = lb;
#pragma unroll
                for (int j = 1; j < 8; j++) far[j] = lcg(far[j - 1], (uint32_t)(k0 + 397 + j));
                lb = lcg(far[7], (uint32_t)(k0 + 397 + 8));
            } else if (k0 == 224) {                                     // old words 621..623, then new words 0..4
                far[0] = lb; far[1] = lcg(far[0], 622u); far[2] = lcg(far[1], 623u);
#pragma unroll
                for (int j = 3; j < 8; j++) far[j] = st[(j - 3) * MT_LANES];
            } else {
                const uint32_t* pf = p - 227 * MT_LANES;
#pragma unroll
                for (int j = 0; j < 8; j++) far[j] = pf[j * MT_LANES];
            }
        } else if (k0 != 224 && k0 != 616) {
            const uint32_t* pf = (k0 < 224) ? p + 397 * MT_LANES : p - 227 * MT_LANES;
#pragma unroll
            for (int j = 0; j < 9; j++) own[j] = p[j * MT_LANES];
#pragma unroll
            for (int j = 0; j < 8; j++) far[j] = pf[j * MT_LANES];
        } else {
#pragma unroll
            for (int j = 0; j < 9; j++) { const int k = k0 + j; own[j] = st[(k == 624 ? 0 : k) * MT_LANES]; }
#pragma unroll
            for (int j = 0; j < 8; j++) { const int k = k0 + j + 397; far[j] = st[(k < 624 ? k : k - 624) * MT_LANES]; }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t y0 = (own[j] & 0x80000000u) | (own[j + 1] & 0x7fffffffu);
            uint32_t y = far[j] ^ (y0 >> 1) ^ ((y0 & 1u) ? 0x9908b0dfu : 0u);
            p[j * MT_LANES] = y;
            y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
            buf[j] = y;
        }
        if (k0 + 8 == 624) { k0 = 0; gen0 = false; } else k0 += 8;
        cur = 0;
    }
    __device__ uint32_t next() {
        if (cur == 8) fill();
        return buf[cur++];
    }
    // generate_canonical<float,24>: one 32-bit draw / 2^32, result clamped below 1
    __device__ float canonical() {
        float r = __uint2float_rn(next()) / 4294967296.0f;
        if (r >= 1.0f) r = 0.99999994f;   // nextafterf(1, 0)
        return r;
    }
    __device__ float uniform_real(float a, float b) { return canonical() * (b - a) + a; }
    __device__ uint32_t lemire(uint32_t range) {
        unsigned long long product = (unsigned long long)next() * (unsigned long long)range;
        uint32_t low = (uint32_t)product;
        if (low < range) {
            const uint32_t threshold = (0u - range) % range;
            while (low < threshold) { product = (unsigned long long)next() * (unsigned long long)range; low = (uint32_t)product; }
        }
        return (uint32_t)(product >> 32);
    }
};

// I = index type of (entry * stride): 32-bit for the shared-memory tables, 64-bit for the global ones
template <class T, class I>
__device__ __forceinline__ void dev_swap(T* base, I stride, uint32_t i, uint32_t j) {
    const T a = base[(I)i * stride], b = base[(I)j * stride];
    base[(I)i * stride] = b; base[(I)j * stride] = a;
}
// std::shuffle, pairwise variant; `store` false only advances the generator
template <class T, class I>
__device__ void dev_shuffle(T* base, I stride, uint32_t n, MT& g, bool store) {
    uint32_t i = 1;
    if ((n % 2) == 0) { const uint32_t d = g.lemire(2); if (store) dev_swap(base, stride, i, d); i++; }
    while (i != n) {
        const uint32_t swap_range = i + 1;
        const uint32_t x = g.lemire(swap_range * (swap_range + 1));
        const uint32_t p1 = x / (swap_range + 1), p2 = x % (swap_range + 1);
        if (store) { dev_swap(base, stride, i, p1); dev_swap(base, stride, i + 1, p2); }
        i += 2;
    }
}

// Tables: t1[(dim * ss + set) * npix + pixel], t2 likewise (float2).
// SMEM = true: the two tables of the dimension being built live in shared memory, lane-interleaved
// (entry k of thread t at [k * blockDim + t]: every swap of the shuffle is bank-conflict free whatever its random
// position), and are copied out coalesced once shuffled.  SMEM = false (set sizes too large for shared memory):
// built in place in the global table, one scratch dimension appended for dims nobody reads.
// Persistent CTAs: the grid is a few CTAs per SM and every CTA walks the pixel groups blockIdx.x, blockIdx.x + gridDim.x, ...
// with ONE generator-state block (624 x MT_LANES words, 312 KB) that it reuses for every group.  The live state of the whole
// launch is then grid x 312 KB (~90 MB at two CTAs per SM) instead of 2.5 KB for every pixel of the chunk (5 GB at 1080p): it
// stays in the 126 MB L2, and the state words -- each read twice and written once per 624 draws -- stop travelling to HBM.
// The finished tables are written with streaming stores so that they do not push the state out of L2.
template <bool SMEM>
__global__ void k_sampler_mt(const uint32_t* __restrict__ seeds, uint32_t npix, uint32_t ss, uint32_t sq, uint32_t n1d, uint32_t n2d,
                             float* __restrict__ t1, float2* __restrict__ t2, uint32_t* __restrict__ state) {
    extern __shared__ float sm[];
    const float len1 = 1.0f / (float)ss, len2 = 1.0f / (float)sq;
    const uint32_t last_dim = max(n1d, n2d);   // dims >= last_dim are never read: stop there (the stream is not reused)
    typedef typename std::conditional<SMEM, uint32_t, size_t>::type I;
    const uint32_t ngroups = (npix + MT_LANES - 1) / MT_LANES;
  for (uint32_t group = blockIdx.x; group < ngroups; group += gridDim.x) {
    uint32_t p = group * MT_LANES + threadIdx.x;
    const bool live = p < npix;
    if (SMEM) __syncwarp();                    // the previous group's last copy-out has left the warp's shared tables
    // SMEM: every lane of a warp goes through the same warp barriers -- a lane past the end of the chunk builds the last
    // pixel once more and keeps nothing (__syncwarp needs every non-exited lane of its mask)
    if (!live) { if (!SMEM) continue; p = npix - 1u; }
    MT g; g.st = state + (size_t)blockIdx.x * (624 * MT_LANES) + threadIdx.x;
    g.seed(seeds[p]);
    // Shared-memory tables are private to a warp (entry k of lane l at [k * 32 + l] of the warp's region: every swap of
    // the shuffle is bank-conflict free whatever its random position).  The 2-D table reuses the storage of the 1-D one:
    // a dimension's 1-D table is built, shuffled and copied out before its 2-D table is started, so only one of them is
    // live (8 bytes x set size per thread instead of 12); the warp barriers keep a straggling lane's 1-D entries from
    // being overwritten by its neighbours' 2-D ones.
    const I bd = (I)32;
    float* wbase = sm + (size_t)(threadIdx.x >> 5) * ss * 64;
    float* s1 = wbase + (threadIdx.x & 31);                        // ss floats
    float2* s2 = reinterpret_cast<float2*>(wbase) + (threadIdx.x & 31);   // ss float2 (8-byte lanes: 2-way, still cheap)
    for (uint32_t dim = 0; dim < last_dim; dim++) {
        const bool keep1 = live && dim < n1d, keep2 = live && dim < n2d;
        float* out1 = t1 + ((size_t)(keep1 ? dim : n1d) * ss) * npix + p;
        float* a = SMEM ? s1 : out1;
        const I as = SMEM ? bd : (I)npix;
        for (uint32_t k = 0; k < ss; k++) {
            const float begin = (float)k / (float)ss;
            const float v = begin + g.uniform_real(0.0f, len1);
            if (keep1) a[(I)k * as] = v;
        }
        dev_shuffle(a, as, ss, g, keep1);
        if (SMEM && keep1) { float* o = out1; for (uint32_t k = 0; k < ss; k++, o += npix) __stcs(o, s1[k * (uint32_t)bd]); }
        if (SMEM) __syncwarp();
        float2* out2 = t2 + ((size_t)(keep2 ? dim : n2d) * ss) * npix + p;
        float2* b = SMEM ? s2 : out2;
        for (uint32_t sy = 0; sy < sq; sy++)
            for (uint32_t sx = 0; sx < sq; sx++) {
                const float bx = (float)sx / (float)sq, by = (float)sy / (float)sq;
                const float x = bx + g.uniform_real(0.0f, len2);
                const float y = by + g.uniform_real(0.0f, len2);
                if (keep2) b[(I)(sy * sq + sx) * as] = make_float2(x, y);
            }
        dev_shuffle(b, as, ss, g, keep2);
        if (SMEM && keep2) { float2* o = out2; for (uint32_t k = 0; k < ss; k++, o += npix) __stcs(o, s2[k * (uint32_t)bd]); }
        if (SMEM) __syncwarp();
    }
  }
}

// RGK_SAMPLER_TABLES: caller tables [pixel in call order][dim][set] -> chunk layout [dim][set][pixel position]
__global__ void k_tables_from_user(const float* __restrict__ u1, const float* __restrict__ u2, uint32_t un1, uint32_t un2,
                                   const uint32_t* __restrict__ pix_src, uint32_t npix, uint32_t ss, uint32_t n1d, uint32_t n2d,
                                   float* __restrict__ t1, float2* __restrict__ t2) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const size_t src = pix_src[p];
    for (uint32_t d = 0; d < n1d && d < un1; d++)
        for (uint32_t k = 0; k < ss; k++) t1[((size_t)d * ss + k) * npix + p] = u1[(src * un1 + d) * ss + k];
    for (uint32_t d = 0; d < n2d && d < un2; d++)
        for (uint32_t k = 0; k < ss; k++) {
            const float* q = u2 + ((src * un2 + d) * ss + k) * 2;
            t2[((size_t)d * ss + k) * npix + p] = make_float2(q[0], q[1]);
        }
}

// ---- warp-cooperative builder (the default where the set fits): one warp per pixel, the generator state in shared memory.
// The thread-per-pixel kernel above streams its 2.5 KB of state per pixel through L2 / HBM for every generation (12 B per
// draw: 54 GB per 1080p x 64 spp round, half of the DRAM peak for 16 ms); here the 624 state words of the pixel a warp is
// working on live in the warp's shared memory, the twist runs 32 words per step across the lanes (word k needs old k, old
// k + 1 and word k + 397 mod 624, which is old for k < 227 and renewed at least 227 words earlier otherwise: ascending
// 32-word steps in place give the standard generator's words), and a phase of the table construction takes its draws 32
// at a time:
//   * seeding is a 624-step dependent chain (no parallel form), so it is done for 32 pixels at once, one per lane; the
//     seeded states go through a 32 x 33 transposing tile to a per-warp scratch block in global memory (80 KB, written
//     and read back once, coalesced), and the warp then builds the 32 pixels one after the other;
//   * tables nobody reads (the caller's keep masks) are not built: their strata draws are skipped without being looked at
//     (only the twists happen), their shuffle draws are only checked for a Lemire rejection (which would shift the stream);
//   * strata: lane j computes entry k0 + j from draw j; shuffle draws: lane j turns draw j into the swap partners of the
//     pair of positions it stands for (std::shuffle's two-swaps-per-draw loop); a rejected draw ends the batch after the
//     lanes before it, and the next batch starts at that pair with the following draw;
//   * the swaps themselves are sequential (position i is exchanged with an earlier, random one): they are applied from
//     the stored partner lists when the warp's table slots are full, one table per lane, then the finished tables are
//     written out.  A slot is a table of one (pixel, dimension): several pixels share the slots so that more lanes have a
//     table to shuffle.
// Every decision above is warp-uniform.  WL = 1 in the host build of tests/host_cpp (one-lane warps): same code, batches of one.
#ifndef RGK_WARP_LANES
#define RGK_WARP_LANES 32
#endif
constexpr uint32_t WL = RGK_WARP_LANES;
constexpr uint32_t SW_WARPS = 8;           // warps per CTA of k_sampler_warp
#ifndef RGK_SAMPLER_X
#define RGK_SAMPLER_X 0     // timing experiments only (results wrong): 1 no seeding, 2 no swaps, 4 no table write-out, 8 no twist, 16 no shuffle draws, 32 no state load
#endif
constexpr uint32_t SW_TILE = WL * (WL + 1u);   // words of the transposing tile of the seeding phase (aliases state + slots)

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
    return y;
}

// One wide step of the twist: words [128 S, 128 S + 128) of the generation, four consecutive words per lane (one 128-bit load for
// old k .. k + 3, old k + 4 for the pair of the fourth word, the four k + 397-mod-624 words one by one -- 397 and 227 are odd, they
// never line up for a wide load -- and one 128-bit store): 38 instructions per 128 words instead of 54 in 32-word steps.  Word k
// needs old k, old k + 1 (new 0 for k = 623) and word k + 397 mod 624, which is old for k < 227 and was renewed by an earlier
// step otherwise; reads and writes of a step are separated by a warp barrier.
template <int S>
__device__ __forceinline__ void mt_twist_wide(uint32_t* __restrict__ st, uint32_t lane) {
    constexpr uint32_t base = 128u * S;
    const uint32_t k = base + 4u * lane;
    const bool on = k < 624u;                 // the last step covers 112 words: lanes 0 .. 27
    uint4 a = make_uint4(0u, 0u, 0u, 0u);
    uint32_t a4 = 0u, c[4] = {0u, 0u, 0u, 0u};
    if (on) {
        a = *reinterpret_cast<const uint4*>(st + k);
        a4 = st[k + 4u < 624u ? k + 4u : 0u];
#pragma unroll
        for (uint32_t j = 0; j < 4u; j++) {
            const uint32_t kj = k + j;
            c[j] = st[(base + 127u < 227u || (base < 227u && kj < 227u)) ? kj + 397u : kj - 227u];
        }
    }
    __syncwarp();
    if (on) {
        const uint32_t own[5] = {a.x, a.y, a.z, a.w, a4};
        uint32_t out[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t y0 = (own[j] & 0x80000000u) | (own[j + 1] & 0x7fffffffu);
            out[j] = c[j] ^ (y0 >> 1) ^ ((y0 & 1u) ? 0x9908b0dfu : 0u);
        }
        *reinterpret_cast<uint4*>(st + k) = make_uint4(out[0], out[1], out[2], out[3]);
    }
    __syncwarp();
}
// step s of the current generation (one copy in the kernel: the callers only pay a call)
static __device__ __noinline__ void mt_twist_step(uint32_t* st, uint32_t lane, uint32_t s) {
    switch (s) {
    case 0: mt_twist_wide<0>(st, lane); break;
    case 1: mt_twist_wide<1>(st, lane); break;
    case 2: mt_twist_wide<2>(st, lane); break;
    case 3: mt_twist_wide<3>(st, lane); break;
    default: mt_twist_wide<4>(st, lane); break;
    }
}

struct WarpMT {
    uint32_t* st;        // the pixel's 624 state words (shared memory, private to the warp)
    uint32_t pos;        // draws taken from the current generation, warp-uniform; 624 = exhausted (the state of a fresh mt19937)
    uint32_t ren;        // words of the current generation already renewed: st[0, ren) new, st[ren, 624) still the previous generation
    uint32_t lane;
    // The twist is lazy: a generation is renewed in ascending steps only as far as draws are asked for, so the unused tail of a
    // pixel's last generation is never computed.  32-lane warps: wide steps of 128 words; the one-lane warps of the host build
    // (tests/host_cpp): word by word, the textbook loop.
    __device__ __forceinline__ void renew(uint32_t upto) {
        if (RGK_SAMPLER_X & 8) { ren = upto; return; }
        if (WL == 32u) {
            while (ren < upto) { mt_twist_step(st, lane, ren >> 7); ren = min(ren + 128u, 624u); }
            return;
        }
        while (ren < upto) {
            const uint32_t k = ren + lane;
            uint32_t a = 0, b = 0, c = 0;
            if (k < 624u) { a = st[k]; b = st[k + 1u < 624u ? k + 1u : 0u]; c = st[k < 227u ? k + 397u : k - 227u]; }
            __syncwarp();
            if (k < 624u) {
                const uint32_t y0 = (a & 0x80000000u) | (b & 0x7fffffffu);
                st[k] = c ^ (y0 >> 1) ^ ((y0 & 1u) ? 0x9908b0dfu : 0u);
            }
            __syncwarp();
            ren = min(ren + WL, 624u);
        }
    }
    __device__ __forceinline__ void next_generation() { renew(624u); pos = 0u; ren = 0u; }
    // the next m <= WL draws: lane j < m gets the j-th
    __device__ __forceinline__ uint32_t take(uint32_t m) {
        uint32_t w = 0;
        const uint32_t first = min(m, 624u - pos);          // of them in the current generation
        renew(pos + first);
        if (lane < first) w = st[pos + lane];
        if (first < m) {
            next_generation();
            renew(m - first);
            if (lane >= first && lane < m) w = st[lane - first];
            pos = m - first;
        } else pos += m;
        return mt_temper(w);
    }
    __device__ __forceinline__ void skip(uint32_t n) {
        pos += n;
        while (pos > 624u) { const uint32_t over = pos - 624u; next_generation(); pos = over; }
    }
    // The draws of std::shuffle over n elements (n / 2 accepted ones: one swap for the second element of an even n, then two
    // swaps per draw) turned into partner[i] = the position element i is exchanged with, i = 1 .. n - 1 in that order.
    // keep false: only the stream position is kept right (rejections).
    __device__ __forceinline__ void shuffle_draws(uint32_t n, uint16_t* partner, bool keep) {
        const uint32_t odd = n & 1u, nd = n / 2u;
        if (RGK_SAMPLER_X & 16) { skip(nd); return; }
        uint32_t t = 0;
        while (t < nd) {
            if (pos == 624u) next_generation();
            const uint32_t m = min(min(WL, nd - t), 624u - pos);
            renew(pos + m);
            const uint32_t tt = t + lane, i = 2u * tt + odd;
            uint32_t hi = 0;
            bool rej = false;
            if (lane < m) {
                const uint32_t w = mt_temper(st[pos + lane]);
                const uint32_t range = (odd || tt) ? (i + 1u) * (i + 2u) : 2u;
                const unsigned long long product = (unsigned long long)w * (unsigned long long)range;     // Lemire, as MT::lemire
                const uint32_t low = (uint32_t)product;
                hi = (uint32_t)(product >> 32);
                rej = low < range && low < (0u - range) % range;
            }
            const unsigned rm = __ballot_sync(0xffffffffu, rej);
            const uint32_t acc = rm ? (uint32_t)(__ffs((int)rm) - 1) : m;
            if (keep && lane < acc) {
                if (!odd && tt == 0u) partner[1] = (uint16_t)hi;
                else { const uint32_t p1 = hi / (i + 2u); partner[i] = (uint16_t)p1; partner[i + 1u] = (uint16_t)(hi - p1 * (i + 2u)); }
            }
            pos += rm ? acc + 1u : m;       // the rejected draw is used up too; the pair it stood for draws again
            t += acc;
        }
    }
};

// uniform_real_distribution<float>(0, len) of an already drawn word: generate_canonical, then * (len - 0) + 0
__device__ __forceinline__ float mt_real_of(uint32_t word, float len) {
    float r = __uint2float_rn(word) / 4294967296.0f;
    if (r >= 1.0f) r = 0.99999994f;
    return r * (len - 0.0f) + 0.0f;
}

// slot g of a warp: float2 data[ss] (a 1-D table uses .x), then uint16 partner[ss]; slot_words apart (even; = 2 mod 32 so that
// the lanes of the swap phase, one slot each, start in different banks).  meta[g] = q | dim << 8 | 2-D << 16 | not built << 31
// (a warp whose pixel lies past the end of the chunk keeps the slot count of its CTA but builds nothing).
// The copy-out is done by the whole CTA: its warps hold the same table (q, dim) of SW_WARPS consecutive pixels, so entry k of
// all of them is one run of 32 / 64 bytes in the global table -- whole sectors.  (A warp writing its own pixel alone puts 4 or 8
// bytes into each of 32 sectors per store: 660 M partial-sector writes per 1080p x 64 spp round, which L2 takes at about one per
// slice and clock -- that, not the generator, bounded the kernel: 11.1 ms with, 7.3 ms without the copy-out.)
__device__ __noinline__ void sampler_flush(uint32_t* warp0, uint32_t warp_words, uint32_t nslots, uint32_t filled, uint32_t slot_words, uint32_t ss,
                                           uint32_t npix, uint32_t pix0, uint32_t w, uint32_t lane, float* __restrict__ t1, float2* __restrict__ t2) {
    uint32_t* slots = warp0 + w * warp_words + 624u;
    const uint32_t* meta = slots + nslots * slot_words;
    __syncwarp();
    // The exchanges of std::shuffle are sequential by definition (element i goes to an earlier, random place p_i); the first
    // `serial` of them are applied that way, one table per lane.  Beyond that the partners of 32 consecutive exchanges seldom
    // meet -- two exchanges of a batch touch a common element only if they have the same partner or one's partner is the other's
    // own position (expected 500 / i pairs per batch at position i) -- so the warp applies a batch in a few WAVES: an exchange
    // waits for the earlier batch-mates it shares an element with (its wave = 1 + theirs), exchanges of one wave touch disjoint
    // elements and go together.  Same result as the sequential order; a 256-entry table takes ~25 steps instead of 255.
    const uint32_t serial = (WL == 32u) ? min(ss, 64u) : ss;
    if (!(RGK_SAMPLER_X & 2) && lane < filled && !(meta[lane] >> 31)) {
        float2* d = reinterpret_cast<float2*>(slots + lane * slot_words);
        const uint16_t* pr = reinterpret_cast<const uint16_t*>(slots + lane * slot_words + 2u * ss);
        // partners two at a time (positions 2 k and 2 k + 1 share a word of the list), the next pair requested before the current
        // exchanges (the list is not touched by them)
        if (serial > 1u) { const uint32_t j = pr[1]; const float2 a = d[1], b = d[j]; d[1] = b; d[j] = a; }
        const uint32_t* pr2 = reinterpret_cast<const uint32_t*>(pr);
        uint32_t pn = serial > 2u ? pr2[1] : 0u;
        for (uint32_t i = 2; i < serial; i += 2u) {
            const uint32_t pp = pn;
            pn = pr2[(i + 2u < serial) ? (i >> 1) + 1u : (i >> 1)];
            {
                const uint32_t j = pp & 0xffffu;
                const float2 a = d[i], b = d[j];
                d[i] = b; d[j] = a;
            }
            if (i + 1u < serial) {
                const uint32_t j = pp >> 16;
                const float2 a = d[i + 1u], b = d[j];
                d[i + 1u] = b; d[j] = a;
            }
        }
    }
#if defined(__CUDA_ARCH__) || defined(__CUDA_ARCH_EMULATED_LANES__)      // (the latter: the 32-lane host build of tests/host_cpp/device_shim_mt.h)
    if (serial < ss && !(RGK_SAMPLER_X & 2)) {
        __syncwarp();
        for (uint32_t g = 0; g < filled; g++) {
            if (meta[g] >> 31) continue;
            float2* d = reinterpret_cast<float2*>(slots + g * slot_words);
            const uint16_t* pr = reinterpret_cast<const uint16_t*>(slots + g * slot_words + 2u * ss);
            for (uint32_t i0 = serial; i0 < ss; i0 += 32u) {
                const uint32_t i = i0 + lane;
                const bool on = i < ss;
                const uint32_t pt = on ? pr[i] : 0xffff0000u + lane;              // lanes past the end: a partner nobody shares
                const unsigned same = __match_any_sync(0xffffffffu, pt);
                const unsigned before = same & ((1u << lane) - 1u);               // earlier batch-mates with my partner
                const int prev_same = before ? 31 - __clz((int)before) : -1;
                const int dep_own = (on && pt >= i0 && pt < i) ? (int)(pt - i0) : -1;   // my partner is an earlier batch-mate's own position
                uint32_t wave = 0;
                for (;;) {
                    const uint32_t wa = __shfl_sync(0xffffffffu, wave, prev_same < 0 ? 0 : prev_same);
                    const uint32_t wb = __shfl_sync(0xffffffffu, wave, dep_own < 0 ? 0 : dep_own);
                    uint32_t nw = wave;
                    if (prev_same >= 0) nw = max(nw, wa + 1u);
                    if (dep_own >= 0) nw = max(nw, wb + 1u);
                    const bool changed = nw != wave;
                    wave = nw;
                    if (!__any_sync(0xffffffffu, changed)) break;
                }
                const uint32_t waves = __reduce_max_sync(0xffffffffu, wave) + 1u;
                for (uint32_t wv = 0; wv < waves; wv++) {
                    if (on && wave == wv && pt != i) { const float2 a = d[i], b = d[pt]; d[i] = b; d[pt] = a; }
                    __syncwarp();
                }
            }
        }
    }
#endif
    if (WL == 1u) {          // host build of the tests: the "warps" of a CTA run one after the other, each writes its own pixel
        for (uint32_t g = 0; g < filled; g++) {
            const uint32_t m = meta[g];
            if (m >> 31) continue;
            const uint32_t pix = pix0 + (m & 0xffu) * SW_WARPS + w, dim = (m >> 8) & 0xffu;
            const float2* d = reinterpret_cast<const float2*>(slots + g * slot_words);
            if ((m >> 16) & 1u) { float2* o = t2 + ((size_t)dim * ss) * npix + pix; for (uint32_t k = lane; k < ss; k += WL) __stcs(o + (size_t)k * npix, d[k]); }
            else { float* o = t1 + ((size_t)dim * ss) * npix + pix; for (uint32_t k = lane; k < ss; k += WL) __stcs(o + (size_t)k * npix, d[k].x); }
        }
        return;
    }
    __syncthreads();         // every warp's slots are final
    for (uint32_t g = 0; g < ((RGK_SAMPLER_X & 4) ? 0u : filled); g++) {
        const uint32_t m = meta[g];
        const uint32_t pixbase = pix0 + (m & 0xffu) * SW_WARPS, dim = (m >> 8) & 0xffu;
        const bool two = (m >> 16) & 1u;
        const uint32_t* src0 = warp0 + 624u + g * slot_words;
        for (uint32_t e = threadIdx.x; e < ss * SW_WARPS; e += SW_WARPS * WL) {
            const uint32_t k = e / SW_WARPS, wp = e % SW_WARPS;
            if (pixbase + wp >= npix) continue;
            const float2 v = *reinterpret_cast<const float2*>(src0 + wp * warp_words + 2u * k);
            const size_t at = ((size_t)dim * ss + k) * npix + pixbase + wp;
            if (two) __stcs(t2 + at, v); else __stcs(t1 + at, v.x);
        }
    }
    __syncthreads();         // before the slots are filled again
}

#ifndef RGK_SAMPLER_WARP_MINB
#define RGK_SAMPLER_WARP_MINB 4
#endif
// Shared memory: [ss + sq floats: the stratum origins k / ss and s / sq, computed once per CTA with the reference's division]
// [per warp: 624 state words | nslots slots | nslots meta words].  sq_magic = ceil(2^32 / sq): c / sq = umulhi(c, sq_magic)
// exactly for c < ss (c * sq < 2^32 is checked by the launcher).
__global__ void __launch_bounds__(SW_WARPS * WL, RGK_SAMPLER_WARP_MINB)
k_sampler_warp(const uint32_t* __restrict__ seeds, uint32_t npix, uint32_t ss, uint32_t sq, uint32_t sq_magic, uint32_t ndims, uint64_t keep1m, uint64_t keep2m,
               float* __restrict__ t1, float2* __restrict__ t2, uint32_t* __restrict__ scratch, uint32_t nslots, uint32_t slot_words, uint32_t warp_words,
               uint32_t ppw) {
    // ppw: pixels a warp seeds at once and then builds one after the other: WL where the chunk has enough pixels to keep every
    // resident warp busy that way, fewer (idle lanes in the seeding phase only) for small images
    extern __shared__ uint32_t swm[];
    const uint32_t lane = threadIdx.x % WL, w = threadIdx.x / WL;
    float* begin1 = reinterpret_cast<float*>(swm);
    float* begin2 = begin1 + ss;
    // (written by the CTA's first warp alone: the one-lane host build of the tests runs the "warps" of a CTA one after the other,
    // the first one first)
    if (w == 0u) {
        for (uint32_t k = lane; k < ss; k += WL) begin1[k] = (float)k / (float)ss;
        for (uint32_t k = lane; k < sq; k += WL) begin2[k] = (float)k / (float)sq;
    }
    __syncthreads();
    uint32_t* warp0 = swm + ((ss + sq + 3u) & ~3u);                          // 16-byte aligned: the wide twist loads 128 bits
    uint32_t* wsm = warp0 + w * warp_words;
    uint32_t* slots = wsm + 624u;
    uint32_t* meta = slots + nslots * slot_words;
    uint32_t* my_scratch = scratch + ((size_t)blockIdx.x * SW_WARPS + w) * (WL * 624u);
    WarpMT g; g.st = wsm; g.lane = lane; g.pos = 624u; g.ren = 624u;
    const float len1 = 1.0f / (float)ss, len2 = 1.0f / (float)sq;
    const uint32_t per_block = SW_WARPS * ppw, nblocks = (npix + per_block - 1u) / per_block;
    for (uint32_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        // a CTA owns per_block consecutive pixels; its warps walk them side by side (warp w: pixels pix0 + q * SW_WARPS + w), so the
        // sectors of a table row are completed by the CTA's warps at about the same time
        const uint32_t pix0 = blk * per_block;
        if (!(RGK_SAMPLER_X & 1))
        {   // ---- mt19937::seed for WL pixels, lane q its q-th: word_i = 1812433253 * (word_{i-1} ^ (word_{i-1} >> 30)) + i
            const uint32_t pix = pix0 + lane * SW_WARPS + w;
            uint32_t x = (lane < ppw && pix < npix) ? seeds[pix] : 0u;
            for (uint32_t c0 = 0; c0 < 624u; c0 += WL) {
                const uint32_t nw = min(WL, 624u - c0);
#pragma unroll 8
                for (uint32_t j = 0; j < nw; j++) {
                    if (c0 + j) x = MT::lcg(x, c0 + j);
                    wsm[j * (WL + 1u) + lane] = x;
                }
                __syncwarp();
                if (lane < nw) {
#pragma unroll 8
                    for (uint32_t q = 0; q < ppw; q++) __stcg(my_scratch + q * 624u + c0 + lane, wsm[lane * (WL + 1u) + q]);
                }
                __syncwarp();
            }
        }
        uint32_t filled = 0;
        for (uint32_t q = 0; q < ppw; q++) {
            if (pix0 + q * SW_WARPS >= npix) break;                      // no warp of the CTA has a pixel in this row
            const bool valid = pix0 + q * SW_WARPS + w < npix;          // (the other warps still need this one at their flushes)
            if (valid) {
                __syncwarp();          // the previous pixel's last draws have been read by every lane (set size 1: no ballot follows them)
                if (!(RGK_SAMPLER_X & 32)) for (uint32_t k = lane; k < 624u; k += WL) wsm[k] = __ldcg(my_scratch + q * 624u + k);
#ifdef __CUDA_ARCH__
                if (q + 1u < ppw && lane < 20u) asm volatile("prefetch.global.L2 [%0];" :: "l"(my_scratch + (q + 1u) * 624u + lane * 32u));
#endif
                __syncwarp();
            }
            g.pos = 624u; g.ren = 624u;
            // table tb: the 1-D (even) / 2-D (odd) table of dimension tb / 2; nothing after the last table anybody reads
            const uint32_t ntab = 2u * ndims - (((keep2m >> (ndims - 1u)) & 1ull) ? 0u : 1u);
            for (uint32_t tb = 0; tb < ntab; tb++) {
                const uint32_t dim = tb >> 1, two = tb & 1u;
                const bool keep = ((two ? keep2m : keep1m) >> dim) & 1ull;
                const uint32_t ndraws = ss << two;
                if (!keep) { if (valid) { g.skip(ndraws); g.shuffle_draws(ss, nullptr, false); } continue; }
                if (filled == nslots) { sampler_flush(warp0, warp_words, nslots, filled, slot_words, ss, npix, pix0, w, lane, t1, t2); filled = 0; }
                if (valid) {
                    float* d = reinterpret_cast<float*>(slots + filled * slot_words);
                    for (uint32_t j0 = 0; j0 < ndraws; j0 += WL) {
                        const uint32_t m = min(WL, ndraws - j0);
                        const uint32_t word = g.take(m);
                        const uint32_t j = j0 + lane;
                        if (lane < m) {
                            if (two) {                  // cell c = sy * sq + sx takes draws 2 c (x) and 2 c + 1 (y)
                                const uint32_t c = j >> 1, sy = __umulhi(c, sq_magic), sx = c - sy * sq;
                                d[j] = begin2[(j & 1u) ? sy : sx] + mt_real_of(word, len2);
                            } else d[2u * j] = begin1[j] + mt_real_of(word, len1);
                        }
                    }
                    g.shuffle_draws(ss, reinterpret_cast<uint16_t*>(slots + filled * slot_words + 2u * ss), true);
                }
                if (lane == 0) meta[filled] = q | (dim << 8) | (two << 16) | (valid ? 0u : 0x80000000u);
                filled++;
            }
        }
        sampler_flush(warp0, warp_words, nslots, filled, slot_words, ss, npix, pix0, w, lane, t1, t2);     // the seeding tile reuses the slots
    }
}

// launch geometry of the table generation.  kind 0: thread per pixel (k_sampler_mt), 1: warp per pixel (k_sampler_warp)
struct SamplerPlan { uint32_t kind, grid, nslots, slot_words, warp_words, ppw; size_t smem, scratch_words; };
static SamplerPlan sampler_plan(int device, const rgk_device_cfg& cfg, uint32_t npix, uint32_t ss, uint32_t kept_per_pixel) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    SamplerPlan p{};
    // Small sets (<= 16: a handful of draws per dimension, less than one generation per pixel) stay with the thread-per-pixel
    // kernel, which never materialises the seeded state and whose lanes are all busy; measured at 1080p, thread vs warp kernel:
    // 1 spp 0.5 / 3.1 ms, 4 spp 1.1 / 4.7, 16 spp 3.8 / 5.3, 36 spp 10.4 / 7.1, 64 spp 16.3 / 8.7, 256 spp (540p) 37.4 / 9.0.
    if (cfg.sampler_kernel == 2u || (cfg.sampler_kernel == 0u && ss > 16u)) {
        uint32_t sw = 2u * ss + (ss + 1u) / 2u;
        sw += sw & 1u;
        while (sw % 32u != 2u) sw += 2u;
        // slots: the tables of about two pixels, so that more lanes have one to shuffle (all of them are flushed at the end of a
        // 32-pixel block anyway), while five CTAs still fit an SM
        // (sets above 64 apply most of a shuffle warp-wide, wave by wave: one slot keeps the lanes busy and leaves room for more warps)
        uint32_t want = cfg.sampler_slots ? cfg.sampler_slots : (ss > 64u ? 1u : std::max(1u, std::min(2u * std::max(1u, kept_per_pixel), 8u)));
        want = std::min(want, WL);
        int per = 0;
        for (uint32_t ns = want; ns >= 1u; ns--) {
            uint32_t ww = std::max(SW_TILE, 624u + ns * sw + ns);
            ww = (ww + 3u) & ~3u;
            const size_t smem = ((size_t)ww * SW_WARPS + ((ss + (uint32_t)std::lround(std::sqrt((double)ss)) + 3u) & ~3u)) * 4u;
            if (smem > 200u * 1024u) continue;
            cudaFuncSetAttribute(k_sampler_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            per = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_sampler_warp, (int)(SW_WARPS * WL), smem);
            // at least 32 warps per SM; with one slot left, 16 (below that the thread-per-pixel kernel, unless this one is forced)
            const int need = cfg.sampler_slots ? 1 : (ns > 1u ? 4 : (cfg.sampler_kernel == 2u ? 1 : 2));
            if (per >= need) { p.nslots = ns; p.slot_words = sw; p.warp_words = ww; p.smem = smem; break; }
        }
        if (p.nslots) {
            const uint32_t resident_warps = (uint32_t)(sms * per) * SW_WARPS;
            p.ppw = std::max(1u, std::min(WL, (npix + resident_warps - 1u) / resident_warps));
            const uint32_t blocks = (npix + SW_WARPS * p.ppw - 1u) / (SW_WARPS * p.ppw);
            p.kind = 1u;
            p.grid = std::max(1u, std::min(blocks, (uint32_t)(sms * per)));
            p.scratch_words = (size_t)p.grid * SW_WARPS * WL * 624u;
            return p;
        }
    }
    const uint32_t groups = (npix + MT_LANES - 1) / MT_LANES;
    p.kind = 0u;
    p.grid = std::max(1u, std::min(groups, (uint32_t)sms * std::max(1u, cfg.sampler_ctas_per_sm)));
    p.scratch_words = (size_t)p.grid * MT_LANES * 624u;
    return p;
}
// keep masks: bit d = the 1-D / 2-D table of dimension d is read by somebody (the thread-per-pixel kernel builds them all)
static void launch_sampler_mt(cudaStream_t stream, const SamplerPlan& plan, bool use_smem, const uint32_t* seeds, uint32_t npix, uint32_t ss, uint32_t sq, uint32_t n1d, uint32_t n2d,
                              float* t1, float2* t2, uint32_t* state, uint64_t keep1m = ~0ull, uint64_t keep2m = ~0ull) {
    if (plan.kind == 1u) {
        keep1m &= n1d >= 64u ? ~0ull : ((1ull << n1d) - 1ull);
        keep2m &= n2d >= 64u ? ~0ull : ((1ull << n2d) - 1ull);
        uint32_t ndims = 0;
        for (uint32_t d = 0; d < 64u; d++) if (((keep1m | keep2m) >> d) & 1ull) ndims = d + 1u;
        if (!ndims) return;
        cudaFuncSetAttribute(k_sampler_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
        const uint32_t sq_magic = sq > 1u ? (uint32_t)((0x100000000ull + sq - 1u) / sq) : 0u;      // sq == 1: c is always 0
        k_sampler_warp<<<plan.grid, SW_WARPS * WL, plan.smem, stream>>>(seeds, npix, ss, sq, sq_magic, ndims, keep1m, keep2m, t1, t2, state, plan.nslots, plan.slot_words, plan.warp_words, plan.ppw);
        return;
    }
    // per device (function attributes are), so set on every launch rather than once per process: contexts on several
    // GPUs may live in one process
    cudaFuncSetAttribute(k_sampler_mt<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    // shared-memory tables only while two 128-thread CTAs still fit an SM (set size <= 66); beyond that the CTA count
    // per SM drops to one or two warps and the latency of the generator-state loads is no longer hidden (measured:
    // 2.4 s vs 1.2 s per 1080p x 256 spp x depth-40 round), so larger sets are shuffled in place in global memory
    if (use_smem) {
        const size_t bytes = (size_t)MT_LANES * ss * 8;
        if (bytes <= 100 * 1024) {
            k_sampler_mt<true><<<plan.grid, MT_LANES, bytes, stream>>>(seeds, npix, ss, sq, n1d, n2d, t1, t2, state);
            return;
        }
    }
    k_sampler_mt<false><<<plan.grid, MT_LANES, 0, stream>>>(seeds, npix, ss, sq, n1d, n2d, t1, t2, state);
}

// Counter-based sampler with the same structure (jittered strata visited in a per-(pixel,dim) random order),
// no tables, no sequential generator: RGK_SAMPLER_FAST.  Same distribution, different sequence.
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}
__device__ __forceinline__ uint32_t permute(uint32_t i, uint32_t l, uint32_t p) {   // Kensler's bijection on [0,l)
    uint32_t w = l - 1; w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    do {
        i ^= p; i *= 0xe170893du; i ^= p >> 16; i ^= (i & w) >> 4; i ^= p >> 8; i *= 0x0929eb3fu; i ^= p >> 23;
        i ^= (i & w) >> 1; i *= 1 | p >> 27; i *= 0x6935fa69u; i ^= (i & w) >> 11; i *= 0x74dcb303u; i ^= (i & w) >> 2;
        i *= 0x9e501cc3u; i ^= (i & w) >> 2; i *= 0xc860a3dfu; i &= w; i ^= i >> 5;
    } while (i >= l);
    return (i + p) % l;
}
__device__ __forceinline__ float u01(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

struct SamplerView {
    const float* t1; const float2* t2; uint32_t npix, ss, sq, mode;
    __device__ __forceinline__ float get1d(uint32_t pixel, uint32_t seed, uint32_t set, uint32_t dim) const {
        if (mode != RGK_SAMPLER_FAST) return __ldg(t1 + ((size_t)dim * ss + set) * npix + pixel);
        const uint32_t key = hash32(seed ^ hash32(dim * 2u + 1u));
        const uint32_t k = permute(set, ss, key);
        return ((float)k + u01(hash32(key ^ (set * 0x9e3779b9u + 0x85ebca6bu)))) / (float)ss;
    }
    __device__ __forceinline__ V2 get2d(uint32_t pixel, uint32_t seed, uint32_t set, uint32_t dim) const {
        if (mode != RGK_SAMPLER_FAST) { const float2 v = __ldg(t2 + ((size_t)dim * ss + set) * npix + pixel); return V2{v.x, v.y}; }
        const uint32_t key = hash32(seed ^ hash32(dim * 2u + 2u));
        const uint32_t k = permute(set, ss, key);
        const uint32_t sx = k % sq, sy = k / sq;
        const uint32_t h = hash32(key ^ (set * 0x9e3779b9u + 0xc2b2ae35u));
        return V2{((float)sx + u01(h)) / (float)sq, ((float)sy + u01(hash32(h))) / (float)sq};
    }
};

// ------------------------------------------------------------------ kernels
// tiles[i] = (x1, x2, y1, y2); tiles2[i] = (first pixel of the tile inside the chunk, tile seed)
__global__ void k_pixel_setup(const uint4* __restrict__ tiles, const uint2* __restrict__ tiles2,
                              uint32_t* __restrict__ pix_xy, uint32_t* __restrict__ pix_seed, uint32_t* __restrict__ pix_src,
                              uint32_t first_call_pixel) {
    const uint4 t = tiles[blockIdx.x];
    const uint2 u = tiles2[blockIdx.x];
    const uint32_t w = t.y - t.x, n = w * (t.w - t.z);
    const bool full = (w == 32u) && (t.w - t.z == 32u);
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
        const uint32_t lx = k % w, ly = k / w;
        // position of the pixel inside the chunk: full 32x32 tiles are laid out as 8x4-pixel blocks so that one warp
        // of primary rays covers a compact footprint; the pixel's seed still follows the reference's y-major order
        const uint32_t pos = full ? ((((ly >> 2) << 2) + (lx >> 3)) << 5) + ((ly & 3u) << 3) + (lx & 7u) : k;
        pix_xy[u.x + pos] = (t.x + lx) | ((t.z + ly) << 16);
        // PathTracer::RenderPixel: samplerSeed += 0x42424242 before every pixel, y-major / x-minor (src/tracer.cpp:8-9)
        pix_seed[u.x + pos] = u.y + (k + 1u) * 0x42424242u;
        pix_src[u.x + pos] = first_call_pixel + u.x + k;
    }
}

// Camera::GetPixelRay / GetPixelRayLens (src/camera.cpp:26-46) + Ray(from, dir) (src/ray.hpp:9-13)
__device__ __forceinline__ void camera_ray(const rgk_camera& c, int x, int y, uint32_t xres, uint32_t yres, V2 off, V2 lens, V3& o, V3& d) {
    const float fx = ((float)x + off.x) / (float)xres, fy = ((float)y + off.y) / (float)yres;
    const V3 p = v3(c.viewscreen) + fx * v3(c.viewscreen_x) + fy * v3(c.viewscreen_y);
    o = v3(c.origin);
    if (c.lens_size != 0.0f) {
        const V2 dsk = disc_uniform(lens);
        const V2 lo = V2{dsk.x * c.lens_size, dsk.y * c.lens_size};
        o = o + lo.x * v3(c.cameraleft) + lo.y * v3(c.cameraup);
    }
    d = normalize(p - o);
}

__global__ void k_camera_rays(rgk_camera cam, uint32_t xres, uint32_t yres, const int32_t* __restrict__ xy, const float* __restrict__ off,
                              const float* __restrict__ lens, uint64_t n, rgk_ray* __restrict__ rays) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    V3 o, d;
    camera_ray(cam, xy[2 * i], xy[2 * i + 1], xres, yres, V2{off[2 * i], off[2 * i + 1]}, lens ? V2{lens[2 * i], lens[2 * i + 1]} : V2{0, 0}, o, d);
    rgk_ray r; r.origin[0] = o.x; r.origin[1] = o.y; r.origin[2] = o.z; r.direction[0] = d.x; r.direction[1] = d.y; r.direction[2] = d.z;
    r.tnear = 0.0f; r.tfar = 10000.0f;
    rays[i] = r;
}

// RenderPixel's per-sample prologue (src/path_tracer.cpp:53-61) and TracePath's light pick (:315-322,337-346)
__global__ void k_raygen(DevScene S, RenderConst R, SamplerView smp, PathBuffers B, float4* __restrict__ cam_o) {
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t npaths = (size_t)R.npix * R.ms;
    if (slot >= npaths) return;
    const uint32_t pixel = (uint32_t)(slot % R.npix), set = (uint32_t)(slot / R.npix);
    const uint32_t xy = B.pix_xy[pixel], seed = B.pix_seed[pixel];
    uint32_t d2 = 0;
    const V2 coords = smp.get2d(pixel, seed, set, d2++);
    V2 lens = V2{0, 0};
    if (R.lens) lens = smp.get2d(pixel, seed, set, d2++);
    V3 o, d;
    camera_ray(R.cam, (int)(xy & 0xffffu), (int)(xy >> 16), R.xres, R.yres, coords, lens, o, d);
    B.ray_o[slot] = make_float4(o.x, o.y, o.z, 0.0f);
    if (cam_o) cam_o[slot] = make_float4(o.x, o.y, o.z, 0.0f);             // camerapos of the sample (bidirectional mode)
    B.ray_d[slot] = make_float4(d.x, d.y, d.z, 0.0f);
    if (R.reverse) {         // (the unidirectional loop knows this initial state: RenderConst::first_bounce; camera rays ignore no triangle)
        B.cum[slot] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(0u));      // .w = n (bounces done)
        B.tot[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        B.last_tri[slot] = RGK_NO_TRIANGLE;
        B.cur1[slot] = 1u;
    }
    if (R.const_light) return;      // GetRandomLight can only return that one light, unjittered (size 0): kept in RenderConst
    const V2 areal = smp.get2d(pixel, seed, set, d2++);         // (read here, not above: with the fixed light these tables are not even built)
    d2++;                                                       // lightdir_sample: drawn, unused when reverse == 0
    const V2 choice = smp.get2d(pixel, seed, set, d2++);
    const float ls = smp.get1d(pixel, seed, set, 0);
    LightRec L = random_light(S, choice, ls, areal);
    if (L.valid && L.type == 0) { const V3 dir = sphere_uniform(areal); L.pos = L.pos + L.size * dir; }
    B.light_pos[slot] = make_float4(L.pos.x, L.pos.y, L.pos.z, __uint_as_float((L.valid ? 1u : 0u) | ((uint32_t)L.type << 1)));
    B.light_col[slot] = make_float4(L.color.r, L.color.g, L.color.b, L.intensity);
    B.light_nrm[slot] = make_float4(L.normal.x, L.normal.y, L.normal.z, 0.0f);
}

constexpr int TRACE_THREADS = 128;
#ifndef RGK_INCOH_MINB
#define RGK_INCOH_MINB 10   // CTAs/SM the incoherent-bounce instantiations are compiled for (48 registers; 12 = 40 registers spills)
#endif
#ifndef RGK_COH_MINB
#define RGK_COH_MINB 9      // camera rays: 56 registers
#endif
#ifndef RGK_CLOSEST_BVH_MINB
#define RGK_CLOSEST_BVH_MINB 8   // CTAs/SM the wide-BVH closest-hit kernel is compiled for: 64 registers, 8 B spilled (per headline round: 6 CTAs / 78
                                 // registers 31.75 ms, 7 / 72: 30.8, 8 / 64: 29.8, 9 / 56: 30.8, 10 / 48: 33.7)
#endif
#ifndef RGK_CLOSEST_BVH_MINB_CAMERA
#define RGK_CLOSEST_BVH_MINB_CAMERA RGK_CLOSEST_BVH_MINB
#endif
#ifndef RGK_SHADOW_BVH_MINB
#define RGK_SHADOW_BVH_MINB 8    // likewise the any-hit kernel (6 CTAs / 72 registers 14.1 ms, 8 / 64: 13.5, 9: 13.6, 10: 14.3)
#endif
#ifndef RGK_RENDER_VARIANT
#define RGK_RENDER_VARIANT 6   // phase-synchronised traversal (trace_device.cuh)
#endif

template <bool COUNT>
__device__ __forceinline__ void flush_counts(const TravCount& c, uint32_t nrays, rgk_trav_stats* stats) {
    if (!COUNT) return;
    unsigned long long v[8] = {nrays, c.inner, c.leaf, c.refs, c.tests, c.exact, c.prefiltered, c.wrong};
#pragma unroll
    for (int k = 0; k < 8; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd((unsigned long long*)&stats->rays, v[0]); atomicAdd((unsigned long long*)&stats->inner, v[1]);
        atomicAdd((unsigned long long*)&stats->leaf, v[2]); atomicAdd((unsigned long long*)&stats->refs, v[3]);
        atomicAdd((unsigned long long*)&stats->tests, v[4]); atomicAdd((unsigned long long*)&stats->exact, v[5]);
        atomicAdd((unsigned long long*)&stats->prefiltered, v[6]); atomicAdd((unsigned long long*)&stats->prefilter_wrong, v[7]);
    }
}

// closest-hit over the live queue (queue == nullptr: identity), persistent warps
template <bool COUNT, int MINB>
__global__ void __launch_bounds__(TRACE_THREADS, MINB)
k_closest(DevScene S, PathBuffers B, const uint32_t* __restrict__ queue, QueueLen len, unsigned long long* work, rgk_trav_stats* stats) {
    const uint32_t count = len.get();
    TravCount cnt{0, 0, 0, 0, 0, 0, 0};
    uint32_t mine = 0;
    trace_rays<RGK_RENDER_VARIANT, false, COUNT>(S, count, work, cnt, mine,
        [&](uint32_t i, Traverser<false, COUNT>& T) {
            const uint32_t slot = queue ? __ldg(queue + i) : i;
            const float4 o = B.ray_o[slot], d = B.ray_d[slot];
            return T.init(S, o.x, o.y, o.z, d.x, d.y, d.z, 0.0f, 10000.0f, queue ? B.last_tri[slot] : RGK_NO_TRIANGLE);
        },
        [&](uint32_t i, bool found, const HitRec& h) {
            const uint32_t slot = queue ? __ldg(queue + i) : i;
#ifdef RGK_DOH_DEBUG
            if ((long long)slot == doh_debug_slot) {
                const float4 o = B.ray_o[slot], d = B.ray_d[slot];
                std::fprintf(stderr, "CLOSEST slot %u o %.9g %.9g %.9g d %.9g %.9g %.9g -> %s tri %u t %.9g a %.9g b %.9g\n", slot, o.x, o.y, o.z, d.x, d.y, d.z,
                             found ? "hit" : "miss", h.tri, h.t, h.alpha, h.beta);
            }
#endif
            B.hit[slot] = make_float4(h.t, h.alpha, h.beta, __uint_as_float(found ? h.tri : RGK_NO_TRIANGLE));
        });
    flush_counts<COUNT>(cnt, mine, stats);
}

// shadow traversal fused with the NEE resolve (src/path_tracer.cpp:431-460,485-496)
template <bool COUNT, int MINB>
__global__ void __launch_bounds__(TRACE_THREADS, MINB)
k_shadow(DevScene S, PathBuffers B, const uint32_t* __restrict__ queue, QueueLen len, float clampv, unsigned long long* work, rgk_trav_stats* stats,
         uint32_t const_light, float4 cl_pos) {
    const uint32_t count = len.get();
    TravCount cnt{0, 0, 0, 0, 0, 0, 0};
    uint32_t mine = 0;
    trace_rays<RGK_RENDER_VARIANT, true, COUNT>(S, count, work, cnt, mine,
        [&](uint32_t i, Traverser<true, COUNT>& T) {
            const uint32_t slot = __ldg(queue + i);
            const float4 a = const_light ? cl_pos : B.light_pos[slot], b = B.sh_pos[slot];
            const float ex = b.x - a.x, ey = b.y - a.y, ez = b.z - a.z;
            const float d2 = ex * ex + ey * ey + ez * ez;
            const float inv = 1.0f / sqrtf(d2), len = sqrtf(d2);
            const float e20 = S.epsilon * 20.0f;
            return T.init(S, a.x, a.y, a.z, ex * inv, ey * inv, ez * inv, 0.0f + e20, len - e20, RGK_NO_TRIANGLE);
        },
        [&](uint32_t i, bool blocked, const HitRec&) {
            const uint32_t slot = __ldg(queue + i);
#ifdef RGK_DOH_DEBUG
            if ((long long)slot == doh_debug_slot) {
                const float4 a = const_light ? cl_pos : B.light_pos[slot], b = B.sh_pos[slot];
                std::fprintf(stderr, "SHADOW slot %u from %.9g %.9g %.9g to %.9g %.9g %.9g -> %s\n", slot, a.x, a.y, a.z, b.x, b.y, b.z, blocked ? "blocked" : "visible");
            }
#endif
            const float4 dr = B.sh_direct[slot];
            if (__float_as_uint(dr.w) == 0u) {                 // vertex without emission: dr = min(direct, clamp) * contribution
                if (blocked) return;
                float4 t = B.tot[slot];
                t.x += dr.x; t.y += dr.y; t.z += dr.z;
                B.tot[slot] = t;
                return;
            }
            const float4 em = B.sh_emis[slot], cb = B.sh_contrib[slot];
            float hr = blocked ? 0.0f : dr.x, hg = blocked ? 0.0f : dr.y, hb = blocked ? 0.0f : dr.z;
            hr += em.x; hg += em.y; hb += em.z;
            if (hr > clampv) hr = clampv;
            if (hg > clampv) hg = clampv;
            if (hb > clampv) hb = clampv;
            float4 t = B.tot[slot];
            t.x += hr * cb.x; t.y += hg * cb.y; t.z += hb * cb.z;
            B.tot[slot] = t;
        });
    flush_counts<COUNT>(cnt, mine, stats);
}

// ---- wide BVH (RGK_TRAVERSAL_BVH, bvh_device.cuh): the same fetch / commit as k_closest and k_shadow, run through the
// BVH; rays whose answer could depend on the kd rule are not committed but appended (as path slots) to `arb`, and the
// *_arb kernels then run the kd traversal over that list (its length is read on the device).  A path's shadow resolve has
// side effects (B.tot), so it is committed exactly once: by the BVH pass or by the arbiter pass.
struct ClosestIO {
    PathBuffers B;
    __device__ __forceinline__ void commit(uint32_t slot, bool found, const HitRec& h) const {
#ifdef RGK_DOH_DEBUG      // host build of the tests only: follow one path slot through its bounces
        if ((long long)slot == doh_debug_slot) {
            const float4 o = B.ray_o[slot], d = B.ray_d[slot];
            std::fprintf(stderr, "CLOSEST slot %u o %.9g %.9g %.9g d %.9g %.9g %.9g -> %s tri %u t %.9g a %.9g b %.9g\n", slot, o.x, o.y, o.z, d.x, d.y, d.z,
                         found ? "hit" : "miss", h.tri, h.t, h.alpha, h.beta);
        }
#endif
        B.hit[slot] = make_float4(h.t, h.alpha, h.beta, __uint_as_float(found ? h.tri : RGK_NO_TRIANGLE));
    }
};
struct ShadowIO {
    PathBuffers B; float clampv; uint32_t const_light; float4 cl_pos;
    template <class T>
    __device__ __forceinline__ bool fetch(const DevScene& S, uint32_t slot, T& tr) const {
        const float4 a = const_light ? cl_pos : B.light_pos[slot], b = B.sh_pos[slot];
        const float ex = b.x - a.x, ey = b.y - a.y, ez = b.z - a.z;
        const float d2 = ex * ex + ey * ey + ez * ez;
        const float inv = 1.0f / sqrtf(d2), len = sqrtf(d2);
        const float e20 = S.epsilon * 20.0f;
        return tr.init(S, a.x, a.y, a.z, ex * inv, ey * inv, ez * inv, 0.0f + e20, len - e20, RGK_NO_TRIANGLE);
    }
    __device__ __forceinline__ void commit(uint32_t slot, bool blocked) const {       // the NEE resolve of k_shadow
#ifdef RGK_DOH_DEBUG
        if ((long long)slot == doh_debug_slot) {
            const float4 a = const_light ? cl_pos : B.light_pos[slot], b = B.sh_pos[slot];
            std::fprintf(stderr, "SHADOW slot %u from %.9g %.9g %.9g to %.9g %.9g %.9g -> %s\n", slot, a.x, a.y, a.z, b.x, b.y, b.z, blocked ? "blocked" : "visible");
        }
#endif
        const float4 dr = B.sh_direct[slot];
        if (__float_as_uint(dr.w) == 0u) {
            if (blocked) return;
            float4 t = B.tot[slot];
            t.x += dr.x; t.y += dr.y; t.z += dr.z;
            B.tot[slot] = t;
            return;
        }
        const float4 em = B.sh_emis[slot], cb = B.sh_contrib[slot];
        float hr = blocked ? 0.0f : dr.x, hg = blocked ? 0.0f : dr.y, hb = blocked ? 0.0f : dr.z;
        hr += em.x; hg += em.y; hb += em.z;
        if (hr > clampv) hr = clampv;
        if (hg > clampv) hg = clampv;
        if (hb > clampv) hb = clampv;
        float4 t = B.tot[slot];
        t.x += hr * cb.x; t.y += hg * cb.y; t.z += hb * cb.z;
        B.tot[slot] = t;
    }
};

template <int MINB, int SORT = 1, bool COUNT = false>       // MINB 6: <= 80 registers (no spills, 7 CTAs/SM at the 72 it uses)
__global__ void __launch_bounds__(TRACE_THREADS, MINB)
k_closest_bvh(DevScene S, PathBuffers B, const uint32_t* __restrict__ queue, QueueLen len, unsigned long long* work, BvhStats* stats,
              uint32_t* __restrict__ arb, uint32_t* arb_count) {
    const uint32_t count = len.get();
    BvhCount cnt{0, 0, 0};
    uint32_t mine = 0, deferred = 0;
    const ClosestIO io{B};
    trace_bvh<false, COUNT, SORT>(S, count, work, cnt, mine, deferred,
        [&](uint32_t i, BvhTraverser<false, COUNT, SORT>& T) {
            const uint32_t slot = queue ? __ldg(queue + i) : i;
            const float4 o = B.ray_o[slot], d = B.ray_d[slot];
            return T.init(S, o.x, o.y, o.z, d.x, d.y, d.z, 0.0f, 10000.0f, queue ? B.last_tri[slot] : RGK_NO_TRIANGLE);
        },
        [&](uint32_t i, bool found, const HitRec& h) { io.commit(queue ? __ldg(queue + i) : i, found, h); },
        [&](uint32_t i) { arb[atomicAdd(arb_count, 1u)] = queue ? __ldg(queue + i) : i; });
    flush_bvh_counts<COUNT>(cnt, mine, deferred, stats);
}
template <int MINB>
__global__ void __launch_bounds__(TRACE_THREADS, MINB)
k_closest_arb(DevScene S, PathBuffers B, const uint32_t* __restrict__ arb, const uint32_t* __restrict__ arb_count, unsigned long long* work, uint32_t camera_rays) {
    TravCount cnt{0, 0, 0, 0, 0, 0, 0};
    uint32_t mine = 0;
    const ClosestIO io{B};
    trace_rays<RGK_RENDER_VARIANT, false, false>(S, *arb_count, work, cnt, mine,
        [&](uint32_t i, Traverser<false, false>& T) {
            const uint32_t slot = arb[i];
            const float4 o = B.ray_o[slot], d = B.ray_d[slot];
            return T.init(S, o.x, o.y, o.z, d.x, d.y, d.z, 0.0f, 10000.0f, camera_rays ? RGK_NO_TRIANGLE : B.last_tri[slot]);
        },
        [&](uint32_t i, bool found, const HitRec& h) { io.commit(arb[i], found, h); });
}

template <int MINB, int SORT, bool COUNT = false>
__global__ void __launch_bounds__(TRACE_THREADS, MINB)
k_shadow_bvh(DevScene S, PathBuffers B, const uint32_t* __restrict__ queue, QueueLen len, float clampv, unsigned long long* work, BvhStats* stats,
             uint32_t const_light, float4 cl_pos, uint32_t* __restrict__ arb, uint32_t* arb_count) {
    const uint32_t count = len.get();
    BvhCount cnt{0, 0, 0};
    uint32_t mine = 0, deferred = 0;
    const ShadowIO io{B, clampv, const_light, cl_pos};
    trace_bvh<true, COUNT, SORT>(S, count, work, cnt, mine, deferred,
        [&](uint32_t i, BvhTraverser<true, COUNT, SORT>& T) { return io.fetch(S, __ldg(queue + i), T); },
        [&](uint32_t i, bool blocked, const HitRec&) { io.commit(__ldg(queue + i), blocked); },
        [&](uint32_t i) { arb[atomicAdd(arb_count, 1u)] = __ldg(queue + i); });
    flush_bvh_counts<COUNT>(cnt, mine, deferred, stats);
}
template <int MINB>
__global__ void __launch_bounds__(TRACE_THREADS, MINB)
k_shadow_arb(DevScene S, PathBuffers B, const uint32_t* __restrict__ arb, const uint32_t* __restrict__ arb_count, float clampv, unsigned long long* work,
             uint32_t const_light, float4 cl_pos) {
    TravCount cnt{0, 0, 0, 0, 0, 0, 0};
    uint32_t mine = 0;
    const ShadowIO io{B, clampv, const_light, cl_pos};
    trace_rays<RGK_RENDER_VARIANT, true, false>(S, *arb_count, work, cnt, mine,
        [&](uint32_t i, Traverser<true, false>& T) { return io.fetch(S, arb[i], T); },
        [&](uint32_t i, bool blocked, const HitRec&) { io.commit(arb[i], blocked); });
}

// Direction bin for the coherence reordering: octahedral map of the direction onto a 16x16 grid, cells numbered along
// a Morton curve so that consecutive bins are neighbouring directions.  255 is reserved for "no ray" (merged into 254).
__device__ __forceinline__ uint8_t dir_bin(float x, float y, float z) {
    const float inv = 1.0f / (fabsf(x) + fabsf(y) + fabsf(z));
    float u = x * inv, v = y * inv;
    if (z < 0.0f) {
        const float fu = (1.0f - fabsf(v)) * (u >= 0.0f ? 1.0f : -1.0f), fv = (1.0f - fabsf(u)) * (v >= 0.0f ? 1.0f : -1.0f);
        u = fu; v = fv;
    }
    uint32_t iu = (uint32_t)fminf(fmaxf((u * 0.5f + 0.5f) * 16.0f, 0.0f), 15.0f);
    uint32_t iv = (uint32_t)fminf(fmaxf((v * 0.5f + 0.5f) * 16.0f, 0.0f), 15.0f);
    iu = (iu | (iu << 2)) & 0x33u; iu = (iu | (iu << 1)) & 0x55u;
    iv = (iv | (iv << 2)) & 0x33u; iv = (iv | (iv << 1)) & 0x55u;
    const uint32_t m = iu | (iv << 1);
    return (uint8_t)(m > 254u ? 254u : m);
}

// Queue construction by counting sort (replaces the atomic compaction of k_shade when R.binning): one CTA owns a group of
// PG consecutive pixel positions x SG consecutive samples (path slot = sample * npix + pixel, so the group's rays leave
// one small patch of the image), counts its live keys per direction bin in shared memory, reserves a contiguous range
// of the output queue with one atomic and writes its slots there bin by bin.
constexpr int BIN_THREADS = 256;
__global__ void __launch_bounds__(BIN_THREADS)
k_bin(const uint8_t* __restrict__ keys, uint32_t npix, uint32_t ms, uint32_t PG, uint32_t SG, uint32_t n_pgroups,
      uint32_t* __restrict__ out, unsigned long long* counter, QueueLen live, uint32_t bin_thresh) {
    if (live.get() < bin_thresh) return;          // too few live paths: k_shade compacted this bounce with its atomics instead
    __shared__ uint32_t hist[256];
    __shared__ uint32_t warp_tot[BIN_THREADS / 32];
    __shared__ unsigned long long base_s;
    const uint32_t p0 = (blockIdx.x % n_pgroups) * PG, s0 = (blockIdx.x / n_pgroups) * SG;
    const uint32_t np = min(PG, npix - p0), ns = min(SG, ms - s0), n = np * ns;
    hist[threadIdx.x] = 0u;
    __syncthreads();
    // Where the rows of the group are whole aligned words (the usual case: 32 pixel positions per row) a thread reads four keys
    // at a time and keeps its words in registers for the second pass: a quarter of the loads and of the row / column divisions,
    // no second read of the keys.
    constexpr int KEEP = 4;                           // words per thread: groups of up to 4096 slots
    const uint32_t wpr = np >> 2, nwords = wpr * ns;
    const bool words = (((np | p0 | npix) & 3u) == 0u) && nwords <= (uint32_t)(KEEP * BIN_THREADS);
    uint32_t kw[KEEP], at[KEEP];
    if (words) {
#pragma unroll
        for (int i = 0; i < KEEP; i++) {
            const uint32_t w = threadIdx.x + (uint32_t)i * BIN_THREADS;
            kw[i] = 0xFFFFFFFFu; at[i] = 0u;
            if (w < nwords) {
                const uint32_t row = w / wpr, col = w - row * wpr;
                at[i] = (s0 + row) * npix + p0 + 4u * col;
                kw[i] = *reinterpret_cast<const uint32_t*>(keys + at[i]);
#pragma unroll
                for (int b = 0; b < 4; b++) { const uint32_t key = (kw[i] >> (8 * b)) & 0xFFu; if (key != 0xFFu) atomicAdd(&hist[key], 1u); }
            }
        }
    } else {
        for (uint32_t k = threadIdx.x; k < n; k += BIN_THREADS) {
            const uint32_t key = keys[(size_t)(s0 + k / np) * npix + p0 + k % np];
            if (key != 0xFFu) atomicAdd(&hist[key], 1u);
        }
    }
    __syncthreads();
    // exclusive scan of the 256 bins (one per thread)
    const uint32_t mine = hist[threadIdx.x];
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < BIN_THREADS / 32; w++) { const uint32_t t = warp_tot[w]; if (w < (int)(threadIdx.x >> 5)) before += t; total += t; }
    if (total == 0u) return;
    if (threadIdx.x == 0) base_s = atomicAdd(counter, (unsigned long long)total);
    hist[threadIdx.x] = before + incl - mine;
    __syncthreads();
    const unsigned long long base = base_s;
    if (words) {
#pragma unroll
        for (int i = 0; i < KEEP; i++) {
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint32_t key = (kw[i] >> (8 * b)) & 0xFFu;
                if (key != 0xFFu) out[base + atomicAdd(&hist[key], 1u)] = at[i] + (uint32_t)b;
            }
        }
        return;
    }
    for (uint32_t k = threadIdx.x; k < n; k += BIN_THREADS) {
        const uint32_t slot = (s0 + k / np) * npix + p0 + k % np;
        const uint32_t key = keys[slot];
        if (key != 0xFFu) out[base + atomicAdd(&hist[key], 1u)] = slot;
    }
}

__device__ __forceinline__ void push_queue(uint32_t* queue, unsigned long long* counter, bool want, uint32_t slot) {
    const unsigned mask = __ballot_sync(0xffffffffu, want);   // every lane of the warp reaches this point
    if (!want) return;
    const unsigned lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    queue[base + __popc(mask & ((1u << lane) - 1u))] = slot;
}

// One vertex of GeneratePath (src/path_tracer.cpp:122-302) plus the NEE set-up of TracePath (:405-496)
#ifndef RGK_SHADE_MINB
#define RGK_SHADE_MINB 5   // <= 102 registers: 5 CTAs of 128 threads per SM
#endif
// LAST: the launch of the last bounce (every vertex of the queue has n == depth, so no continuation is ever sampled): the same
// results from a kernel without the BxDF sampling code (4016 instead of 7352 SASS instructions; measured -3 ms per headline round)
template <bool LAST>
#ifndef RGK_SHADE_THREADS
#define RGK_SHADE_THREADS 128
#endif
#ifndef RGK_SHADE_LAST_MINB
#define RGK_SHADE_LAST_MINB RGK_SHADE_MINB
#endif
__global__ void __launch_bounds__(RGK_SHADE_THREADS, (LAST ? RGK_SHADE_LAST_MINB : RGK_SHADE_MINB) * 128 / RGK_SHADE_THREADS)
k_shade(DevScene S, RenderConst R, SamplerView smp, PathBuffers B, const uint32_t* __restrict__ queue_sorted, const uint32_t* __restrict__ queue_path_order,
        QueueLen len, QueueLen prev_len, uint32_t* __restrict__ next_queue, uint32_t* __restrict__ shadow_queue, uint32_t* __restrict__ next_unsorted,
        unsigned long long* counters) {
    // The grid covers an upper bound of the queue's length (the host launches without knowing it): CTAs past the end leave.  The
    // queue is read through its path-ordered twin when the previous bounce was binned (it was if it had at least bin_thresh live
    // paths: the same test k_bin made); this bounce is binned under the same rule.
    const uint32_t count = len.get();
    if (blockIdx.x * blockDim.x >= count) return;
    const uint32_t* __restrict__ queue = (queue_path_order && prev_len.get() >= R.bin_thresh) ? queue_path_order : queue_sorted;
    const uint32_t binmask = count >= R.bin_thresh ? R.binning : 0u;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool cont = false, shadow = false, null_shadow = false, tot_written = false;
    const bool first = R.first_bounce != 0u;
    uint32_t slot = 0;
    uint32_t cw = 0;       // counting rounds only (R.count_shade): bit 0 surface vertex, 1 sky vertex, 2 light evaluated, 4-5 LTC lobes, 8-11 texels
    if (i < count) {
        slot = queue ? __ldg(queue + i) : i;
        const float4 hit = B.hit[slot];
        const uint32_t tri = __float_as_uint(hit.w);
        const float4 ro4 = B.ray_o[slot], rd4 = B.ray_d[slot];
        const V3 ro = v3(ro4), rd = v3(rd4);
        float4 cum4 = first ? make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(0u)) : B.cum[slot];
        uint32_t n = __float_as_uint(cum4.w) + 1u;
        // everything else that is addressed by the slot alone is requested here, ahead of the dependent chain
        // triangle -> vertices / material -> textures -> LTC taps (the kernel is latency-bound): the sample's light and
        // the two sampler values this vertex may consume (continuation direction, Russian roulette)
        const float4 lp4 = R.const_light ? R.cl_pos : B.light_pos[slot];
        const float4 lc = R.const_light ? R.cl_col : B.light_col[slot];
        const uint32_t c1 = first ? 1u : B.cur1[slot];
        const uint32_t pixel = slot % R.npix, set = slot / R.npix;
        const uint32_t seed = B.pix_seed[pixel];
        const V2 sample = smp.get2d(pixel, seed, set, R.base2 + (n - 1u));
        const float roulette = smp.get1d(pixel, seed, set, c1 < R.n1d ? c1 : 0u);
        const RGB contribution = rgb(cum4.x, cum4.y, cum4.z);
        const V3 Vr = -rd;
        if (tri == RGK_NO_TRIANGLE) {
            const RGB sky = sky_radiance(S, Vr);
            float4 t = first ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : B.tot[slot];
            t.x += sky.r * contribution.r; t.y += sky.g * contribution.g; t.z += sky.b * contribution.b;
            B.tot[slot] = t; tot_written = true;
            cw = 2u;
        } else {
            const uint4 tv = __ldg(S.tri_shade + tri);
            const float ia = 1.0f - hit.y - hit.z, ib = hit.y, ic = hit.z;   // Intersection a,b,c (src/scene_intersect.cpp:280-283)
            const V3 pos = ro + hit.x * rd;
            const V3 nA = v3(__ldg(S.normals + tv.x)), nB = v3(__ldg(S.normals + tv.y)), nC = v3(__ldg(S.normals + tv.z));
            V3 faceN = ia * nA + ib * nB + ic * nC;
            bool ok = true;
            if (isnan(faceN.x)) { faceN = nA; if (isnan(faceN.x)) { faceN = nB; if (isnan(faceN.x)) { faceN = nC; if (isnan(faceN.x)) ok = false; } } }
            if (ok && length(faceN) <= 0.0f) ok = false;
            if (ok) {
                faceN = normalize(faceN);
                const DevMaterial mat = S.materials[tv.w];
                const float2 ta = __ldg(S.texcoords + tv.x), tb = __ldg(S.texcoords + tv.y), tc = __ldg(S.texcoords + tv.z);
                const V2 uv = V2{ia * ta.x + ib * tb.x + ic * tc.x, ia * ta.y + ib * tb.y + ic * tc.y};
                V3 lightN = faceN;
                float right, bottom;
                const TexPre pre = vertex_textures(S, mat, uv, right, bottom);
                if (mat.tex_bump >= 0) {
                    V3 tangent = ia * v3(__ldg(S.tangents + tv.x)) + ib * v3(__ldg(S.tangents + tv.y)) + ic * v3(__ldg(S.tangents + tv.z));
                    if (!(tangent.x * tangent.x + tangent.y * tangent.y + tangent.z * tangent.z < 0.001f)) {
                        tangent = normalize(tangent);
                        const V3 bitangent = normalize(cross(faceN, tangent));
                        const V3 tangent2 = cross(bitangent, faceN);
                        lightN = normalize(faceN + (tangent2 * right + bitangent * bottom) * R.bump_scale);
                        if (isnan(lightN.x)) lightN = faceN;
                    }
                }
                const Frame fr = system_transform_z(lightN);
                const V3 VrL = qrot(fr.g2l, Vr);
                cw = 1u | (pre.taps << 8);
                // ---- next-event estimation set-up (the visibility test runs in k_shadow)
                const uint32_t lflags = __float_as_uint(lp4.w);
                RGB emis = rgb(0.0f, 0.0f, 0.0f);
                if (dot(faceN, Vr) > 0) emis = rgb(mat.emission[0], mat.emission[1], mat.emission[2]);
                if (lflags & 1u) {
                    const V3 lpos = v3(lp4);
                    const V3 Vi = normalize(lpos - pos);
                    const V3 ViL = qrot(fr.g2l, Vi);
                    const RGB f = bxdf_value(S, tv.w, mat, ViL, VrL, uv, pre);
                    cw |= 4u;
                    if (mat.bxdf >= RGK_BXDF_LTC_BECKMANN && ViL.z > 0 && VrL.z > 0) cw += 16u;      // one LTC lobe evaluation (BxDF::value)
                    const V3 dlt = lpos - pos;
                    const float G = fabsf(dot(lightN, Vi)) / dot(dlt, dlt);
                    float df = 1.0f;
                    if (lflags & 2u) df = gmax(0.0f, dot(-Vi, v3(B.light_nrm[slot])));
                    const float k = lc.w * df;
                    const RGB inc = rgb(lc.x * k, lc.y * k, lc.z * k);
                    const RGB direct = rgb(inc.r * (G * f.r), inc.g * (G * f.g), inc.b * (G * f.b));
                    // A direct term that is exactly zero (light behind the surface: BxDF::value returns Spectrum(0) for
                    // Vi.z <= 0; hemisphere light facing away) contributes 0 whether or not the light is visible, so the
                    // Visibility query cannot change the pixel: it is not traced (R.skip_null_shadow, on by default).
                    if (!(R.skip_null_shadow && direct.r == 0.0f && direct.g == 0.0f && direct.b == 0.0f)) {
                        B.sh_pos[slot] = make_float4(pos.x, pos.y, pos.z, 0.0f);
                        if (emis.r == 0.0f && emis.g == 0.0f && emis.b == 0.0f) {
                            // no emission at this vertex (the usual case): what a visible light adds is already known,
                            // min(direct, clamp) * contribution -- the very operations k_shadow would do -- and a blocked one
                            // adds nothing, so the resolve touches one record instead of three (+ the radiance sum)
                            float hr = direct.r, hg = direct.g, hb = direct.b;
                            if (hr > R.clamp) hr = R.clamp;
                            if (hg > R.clamp) hg = R.clamp;
                            if (hb > R.clamp) hb = R.clamp;
                            B.sh_direct[slot] = make_float4(hr * contribution.r, hg * contribution.g, hb * contribution.b, __uint_as_float(0u));
                        } else {
                            B.sh_direct[slot] = make_float4(direct.r, direct.g, direct.b, __uint_as_float(1u));
                            B.sh_emis[slot] = make_float4(emis.r, emis.g, emis.b, 0.0f);
                            B.sh_contrib[slot] = make_float4(contribution.r, contribution.g, contribution.b, 0.0f);
                        }
                        shadow = true;
                    } else null_shadow = true;
                }
                if (!shadow) {
                    RGB here = emis;
                    if (here.r > R.clamp) here.r = R.clamp;
                    if (here.g > R.clamp) here.g = R.clamp;
                    if (here.b > R.clamp) here.b = R.clamp;
                    float4 t = first ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : B.tot[slot];
                    t.x += here.r * contribution.r; t.y += here.g * contribution.g; t.z += here.b * contribution.b;
                    B.tot[slot] = t; tot_written = true;
                }
                // ---- continuation
                // A path whose next vertex would exceed recursion-max ends here whatever BxDF::sample returns (the loop
                // condition `n < depth`, src/path_tracer.cpp:122): nothing of the continuation is observable, skip it.
                if (!LAST && n < R.depth) {
                    V3 dir; RGB tcf; bool may_leak;
                    bxdf_sample(S, mat, VrL, uv, sample, dir, tcf, may_leak, pre);
                    if (R.count_shade) {       // did BxDF::sample evaluate an LTC lobe?  (the specular branch of the *_diffuse kinds)
                        if (mat.bxdf == RGK_BXDF_LTC_BECKMANN || mat.bxdf == RGK_BXDF_LTC_GGX) cw += 16u;
                        else if (mat.bxdf >= RGK_BXDF_LTC_BECKMANN_DIFFUSE) {
                            const float dp = pre.diffuse.r + pre.diffuse.g + pre.diffuse.b, sp = pre.color.r + pre.color.g + pre.color.b;
                            float sx = sample.x;
                            if (!decide_and_rescale(sx, dp / (dp + sp + 0.0001f))) cw += 16u;
                        }
                    }
                    const bool inside = dir.z < 0;
                    dir = qrot(fr.l2g, dir);
                    if (!(dot(dir, faceN) * dot(Vr, faceN) > 0) && !may_leak) n += 10000u;
                    const float rcoef = (!mat.no_russian && R.russian > 0.0f && n > 1u) ? 1.0f / R.russian : 1.0f;
                    RGB cum = rgb(rcoef * contribution.r, rcoef * contribution.g, rcoef * contribution.b);
                    cum = rgb(tcf.r * cum.r, tcf.g * cum.g, tcf.b * cum.b);
                    cont = true;
                    if (gmax(gmax(cum.r, cum.g), cum.b) < 0.001f) cont = false;
                    uint32_t c1_next = c1;
                    if (cont && !mat.no_russian && R.russian >= 0.0f) {
                        c1_next = c1 + 1u;
                        if (roulette > R.russian) cont = false;
                    }
                    if (cont && n > R.depth) cont = false;
                    if (cont && !(n < R.depth)) cont = false;          // while (n < depth)
                    if (cont) {
                        const V3 no = pos + faceN * S.epsilon * 10.0f * (inside ? -1.0f : 1.0f);
                        const V3 nd = normalize(normalize(dir));
                        B.ray_o[slot] = make_float4(no.x, no.y, no.z, 0.0f);
                        B.ray_d[slot] = make_float4(nd.x, nd.y, nd.z, 0.0f);
                        B.cum[slot] = make_float4(cum.r, cum.g, cum.b, __uint_as_float(n));
                        B.last_tri[slot] = tri;
                        if (first || c1_next != c1) B.cur1[slot] = c1_next;      // (only a path that goes on looks at its cursor again)
                    }
                }
            }
        }
        // first bounce: the radiance sum starts here (k_shadow adds to it, k_finish reads it)
        if (first && !tot_written) B.tot[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    // queues: either built by k_bin from direction-bin keys (rays of one pixel block are reordered by direction so that
    // the lanes of a traversal warp follow similar paths through the tree), or by warp-aggregated atomic compaction in
    // path order.  R.binning bit 0: continuation rays, bit 1: shadow rays.  The order of a queue never changes a result.
    if (binmask & 1u) {
        if (cont) { const float4 d = B.ray_d[slot]; B.key_next[slot] = dir_bin(d.x, d.y, d.z); }
        // the same paths once more in path order: the next k_shade gathers its per-path state through this list
        // (coalesced), the traversal goes through the direction-sorted one
        push_queue(next_unsorted, counters + C_NEXT_U, cont, slot);
    } else push_queue(next_queue, counters + C_NEXT, cont, slot);
    if (binmask & 2u) {
        if (shadow) {
            const float4 a = R.const_light ? R.cl_pos : B.light_pos[slot], b = B.sh_pos[slot];
            B.key_shadow[slot] = dir_bin(b.x - a.x, b.y - a.y, b.z - a.z);
        }
    } else push_queue(shadow_queue, counters + C_SHADOW, shadow, slot);
    {   // Visibility queries of the reference that were provably irrelevant and therefore not traced
        const unsigned m = __ballot_sync(0xffffffffu, null_shadow);
        if (m && (threadIdx.x & 31) == 0) atomicAdd(counters + C_SHADOW_SKIPPED, (unsigned long long)__popc(m));
    }
    if (R.count_shade) {       // work counters of a counting round: vertices shaded, sky vertices, image texels, LTC lobes, lights, continuations
        uint32_t v[6] = {cw & 1u, (cw >> 1) & 1u, (cw >> 8) & 15u, (cw >> 4) & 3u, (cw >> 2) & 1u, cont ? 1u : 0u};
#pragma unroll
        for (int k = 0; k < 6; k++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
            if ((threadIdx.x & 31) == 0 && v[k]) atomicAdd(B.shade_counts + k, (unsigned long long)v[k]);
        }
    }
}

#include "reverse_device.cuh"

// TracePath's epilogue (clamp, NaN/negative guard, src/path_tracer.cpp:501-507), RenderPixel's in-order sum
// over the samples (:64) and EXRTexture::AddPixel (src/texture.cpp:342-348)
__global__ void k_finish(RenderConst R, PathBuffers B, float* __restrict__ fb, uint32_t* __restrict__ fb_count) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= R.npix) return;
    float sr = 0.0f, sg = 0.0f, sb = 0.0f;
    for (uint32_t s = 0; s < R.ms; s++) {
        const float4 t = B.tot[(size_t)s * R.npix + p];
        float r = t.x, g = t.y, b = t.z;
        if (r > R.clamp) r = R.clamp;
        if (g > R.clamp) g = R.clamp;
        if (b > R.clamp) b = R.clamp;
        if (isnan(r) || r < 0.0f) r = 0.0f;
        if (isnan(g) || g < 0.0f) g = 0.0f;
        if (isnan(b) || b < 0.0f) b = 0.0f;
        sr += r; sg += g; sb += b;
    }
    const uint32_t xy = B.pix_xy[p];
    const size_t px = (size_t)(xy >> 16) * R.xres + (xy & 0xffffu);
    fb[3 * px] += sr; fb[3 * px + 1] += sg; fb[3 * px + 2] += sb;
    fb_count[px] += R.ms;
}

// EXRTexture::Accumulate (src/texture.cpp:403-412) on device buffers; grid-stride, 128-bit where the pointers allow
__global__ void k_accumulate(float* __restrict__ dst, const float* __restrict__ src, uint64_t n, uint32_t* __restrict__ cnt, const uint32_t* __restrict__ ocnt, uint64_t npx) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((((uintptr_t)dst | (uintptr_t)src) & 15u) == 0) {
        const uint64_t n4 = n / 4;
        for (uint64_t i = i0; i < n4; i += stride) {
            float4 a = reinterpret_cast<float4*>(dst)[i]; const float4 b = reinterpret_cast<const float4*>(src)[i];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            reinterpret_cast<float4*>(dst)[i] = a;
        }
        for (uint64_t i = 4 * n4 + i0; i < n; i += stride) dst[i] += src[i];
    } else for (uint64_t i = i0; i < n; i += stride) dst[i] += src[i];
    if (cnt) for (uint64_t i = i0; i < npx; i += stride) cnt[i] += ocnt[i];
}

int machine_blocks(rgk_context* ctx, const void* kernel, int threads) {
    int sms = 148, per = 4;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kernel, threads, 0);
    return sms * std::max(per, 1);
}

constexpr size_t H_CHUNKS = 16;    // chunks whose counter blocks can wait in pinned memory for the end of the round
rgk_status ensure_buffers(rgk_context* ctx, size_t paths, size_t pixels, size_t t1_floats, size_t t2_float2s, size_t tiles, size_t mt_words, size_t bounces) {
    if (!ctx->paths) ctx->paths = new PathBuffers();
    PathBuffers& B = *ctx->paths;
    bool ok = true;
    if (paths > B.cap_paths) {
        RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ok = ok && alloc_dev(&B.ray_o, paths) && alloc_dev(&B.ray_d, paths) && alloc_dev(&B.hit, paths) && alloc_dev(&B.cum, paths) &&
             alloc_dev(&B.tot, paths) && alloc_dev(&B.light_pos, paths) && alloc_dev(&B.light_col, paths) && alloc_dev(&B.light_nrm, paths) &&
             alloc_dev(&B.sh_pos, paths) && alloc_dev(&B.sh_direct, paths) && alloc_dev(&B.sh_emis, paths) && alloc_dev(&B.sh_contrib, paths) &&
             alloc_dev(&B.last_tri, paths) && alloc_dev(&B.cur1, paths) && alloc_dev(&B.queue_a, paths) && alloc_dev(&B.queue_b, paths) && alloc_dev(&B.queue_ua, paths) && alloc_dev(&B.queue_ub, paths) &&
             alloc_dev(&B.queue_s, paths) && alloc_dev(&B.key_next, paths) && alloc_dev(&B.key_shadow, paths);
        B.cap_paths = ok ? paths : 0;
    }
    if (ok && pixels > B.cap_pixels) {
        RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ok = alloc_dev(&B.pix_xy, pixels) && alloc_dev(&B.pix_seed, pixels) && alloc_dev(&B.pix_src, pixels);
        B.cap_pixels = ok ? pixels : 0;
    }
    if (ok && mt_words > B.cap_mt) { RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); ok = alloc_dev(&B.mt_state, mt_words); B.cap_mt = ok ? mt_words : 0; }
    if (ok && t1_floats > B.cap_t1) { RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); ok = alloc_dev(&B.t1, t1_floats); B.cap_t1 = ok ? t1_floats : 0; }
    if (ok && t2_float2s > B.cap_t2) { RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); ok = alloc_dev(&B.t2, t2_float2s); B.cap_t2 = ok ? t2_float2s : 0; }
    if (ok && tiles > B.cap_tiles) { RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); ok = alloc_dev(&B.tiles, tiles) && alloc_dev(&B.tiles2, tiles); B.cap_tiles = ok ? tiles : 0; }
    bounces = std::max<size_t>(bounces, 1);
    if (ok && bounces > B.cap_bounces) {
        RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (B.h_counters) { cudaFreeHost(B.h_counters); B.h_counters = nullptr; }
        ok = alloc_dev(&B.counters, bounces * C_COUNT) && (B.shade_counts || alloc_dev(&B.shade_counts, (size_t)8));
        ok = ok && cudaMallocHost((void**)&B.h_counters, (H_CHUNKS * bounces * C_COUNT + bounces) * sizeof(unsigned long long)) == cudaSuccess;
        B.cap_bounces = ok ? bounces : 0;
    }
    if (!ok) { cudaGetLastError(); return rgk_fail(ctx, RGK_ERR_NOMEM, "path-state allocation failed (lower rgk_device_cfg::chunk_paths)"); }
    return RGK_OK;
}

rgk_status ensure_reverse_buffers(rgk_context* ctx, size_t paths, uint32_t depth, uint32_t reverse) {
    PathBuffers& B = *ctx->paths;
    if (!B.reverse) B.reverse = new ReverseBuffers();
    ReverseBuffers& V = *B.reverse;
    bool ok = true;
    if (paths > V.cap_paths) {
        RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ok = alloc_dev(&V.cam_o, paths) && alloc_dev(&V.lstart, paths) && alloc_dev(&V.nverts, paths) && alloc_dev(&V.nlverts, paths) && alloc_dev(&V.d2base, paths);
        V.cap_paths = ok ? paths : 0;
    }
    const size_t cam = paths * depth, light = paths * reverse;
    if (ok && cam > V.cap_cam) {
        RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ok = alloc_dev(&V.vr_pos, cam) && alloc_dev(&V.vr_nrm, cam) && alloc_dev(&V.vr_vr, cam) && alloc_dev(&V.vr_uv, cam) && alloc_dev(&V.vr_con, cam) &&
             alloc_dev(&V.vr_here, cam) && alloc_dev(&V.vr_emis, cam);
        V.cap_cam = ok ? cam : 0;
    }
    if (ok && light > V.cap_light) {
        RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ok = alloc_dev(&V.lr_pos, light) && alloc_dev(&V.lr_nrm, light) && alloc_dev(&V.lr_vr, light) && alloc_dev(&V.lr_uv, light) && alloc_dev(&V.lr_lfs, light);
        V.cap_light = ok ? light : 0;
    }
    if (!ok) { cudaGetLastError(); return rgk_fail(ctx, RGK_ERR_NOMEM, "bidirectional vertex storage allocation failed (lower rgk_device_cfg::reverse_bytes)"); }
    return RGK_OK;
}

} // namespace

void free_event_pool(struct EventPool* p);
void free_path_buffers(rgk_context* ctx) {
    if (!ctx->paths) return;
    PathBuffers& B = *ctx->paths;
    void* ptrs[] = {B.ray_o, B.ray_d, B.hit, B.cum, B.tot, B.light_pos, B.light_col, B.light_nrm, B.sh_pos, B.sh_direct, B.sh_emis,
                    B.sh_contrib, B.last_tri, B.cur1, B.queue_a, B.queue_b, B.queue_ua, B.queue_ub, B.queue_s, B.key_next, B.key_shadow, B.pix_xy, B.pix_seed, B.pix_src, B.mt_state, B.t1, B.t2,
                    B.tiles, B.tiles2, B.counters, B.shade_counts};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (B.h_counters) cudaFreeHost(B.h_counters);
    if (B.reverse) {
        ReverseBuffers& V = *B.reverse;
        void* rp[] = {V.cam_o, V.lstart, V.nverts, V.nlverts, V.d2base, V.vr_pos, V.vr_nrm, V.vr_vr, V.vr_uv, V.vr_con, V.vr_here, V.vr_emis,
                      V.lr_pos, V.lr_nrm, V.lr_vr, V.lr_uv, V.lr_lfs};
        for (void* q : rp) if (q) cudaFree(q);
        delete B.reverse;
    }
    free_event_pool(B.events);
    if (B.term_events) { for (cudaEvent_t e : *B.term_events) cudaEventDestroy(e); delete B.term_events; }
    delete ctx->paths;
    ctx->paths = nullptr;
}

// round_up_to_square, src/sampler.cpp:77-83
uint32_t host_sampler_set_size(uint32_t x) {
    const float s = (float)std::sqrt((double)x);
    float i; const float frac = std::modf(s, &i);
    if (frac < 0.0001f) return (uint32_t)(i * i);
    return (uint32_t)((i + 1) * (i + 1));
}

rgk_status launch_accumulate(rgk_context* ctx, float* d_dst, const float* d_src, uint64_t n_floats, uint32_t* d_cnt, const uint32_t* d_ocnt, cudaStream_t stream) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    k_accumulate<<<sms * 8, 256, 0, stream>>>(d_dst, d_src, n_floats, d_cnt, d_ocnt, n_floats / 3);
    ctx->launches++;
    RGK_CUDA(ctx, cudaGetLastError());
    return RGK_OK;
}

rgk_status launch_camera_rays(rgk_context* ctx, const rgk_camera* cam, uint32_t xres, uint32_t yres, const int32_t* d_xy,
                              const float* d_off, const float* d_lens, uint64_t n, rgk_ray* d_rays) {
    k_camera_rays<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(*cam, xres, yres, d_xy, d_off, d_lens, n, d_rays);
    ctx->launches++;
    RGK_CUDA(ctx, cudaGetLastError());
    return RGK_OK;
}

rgk_status launch_sampler_tables(rgk_context* ctx, const uint32_t* d_seeds, uint32_t n_seeds, uint32_t ms, uint32_t n1d, uint32_t n2d,
                                 float* d_out1, float* d_out2) {
    const uint32_t ss = host_sampler_set_size(ms);
    if ((uint64_t)ss * ss > 0xFFFFFFFFull) return rgk_fail(ctx, RGK_ERR_UNSUPPORTED, "multisample too large for the pairwise shuffle");
    const uint32_t sq = (uint32_t)(std::sqrt((double)ss) + 0.5f);
    // tables need one scratch dim each; the caller's buffers have exactly n1d / n2d dims, so run into private buffers
    float* t1 = nullptr; float2* t2 = nullptr; uint32_t* st = nullptr;
    const SamplerPlan plan = sampler_plan(ctx->device, ctx->cfg, n_seeds, ss, n1d + n2d);
    const size_t e1 = (size_t)(n1d + 1) * ss * n_seeds, e2 = (size_t)(n2d + 1) * ss * n_seeds;
    if (cudaMalloc((void**)&t1, e1 * 4) != cudaSuccess || cudaMalloc((void**)&t2, e2 * 8) != cudaSuccess ||
        cudaMalloc((void**)&st, plan.scratch_words * 4) != cudaSuccess) {
        cudaGetLastError(); if (t1) cudaFree(t1); if (t2) cudaFree(t2); if (st) cudaFree(st);
        return rgk_fail(ctx, RGK_ERR_NOMEM, "sampler table allocation failed");
    }
    launch_sampler_mt(ctx->stream, plan, ctx->cfg.sampler_smem != 0, d_seeds, n_seeds, ss, sq, n1d, n2d, t1, t2, st);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && n1d) e = cudaMemcpyAsync(d_out1, t1, (size_t)n1d * ss * n_seeds * 4, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess && n2d) e = cudaMemcpyAsync(d_out2, t2, (size_t)n2d * ss * n_seeds * 8, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(t1); cudaFree(t2); cudaFree(st);
    if (e != cudaSuccess) return rgk_fail(ctx, RGK_ERR_CUDA, cudaGetErrorString(e));
    return RGK_OK;
}

enum { T_CLOSEST = 0, T_SHADOW = 1, T_SAMPLER = 2, T_SHADE = 3, T_KINDS = 4 };
struct EventPool {
    std::vector<cudaEvent_t> ev; std::vector<int> tag; size_t used = 0;
    cudaEvent_t get() { if (used == ev.size()) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); } return ev[used++]; }
    void begin(cudaStream_t s, int kind) { tag.push_back(kind); cudaEventRecord(get(), s); }
    void end(cudaStream_t s) { cudaEventRecord(get(), s); }
    void collect(float* ms) { for (size_t i = 0; i < tag.size(); i++) { float t = 0; cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]); ms[tag[i]] += t; } used = 0; tag.clear(); }
};

void free_event_pool(EventPool* p) {
    if (!p) return;
    for (cudaEvent_t e : p->ev) cudaEventDestroy(e);
    delete p;
}

rgk_status render_round_impl(rgk_context* ctx, const rgk_camera* cam, const rgk_render_params* P, const rgk_task* tasks,
                             uint32_t n_tasks, uint32_t seedstart, uint32_t seedcount_base, float* d_rgb, uint32_t* d_count,
                             rgk_round_stats* stats) {
    const uint64_t launches0 = ctx->launches;
    const uint32_t ms = P->multisample;
    const uint32_t ss = host_sampler_set_size(ms);
    const uint32_t sq = (uint32_t)(std::sqrt((double)ss) + 0.5f);
    const uint32_t lens = cam->lens_size != 0.0f ? 1u : 0u;
    const uint32_t base2 = 4u + lens;                  // 2-D dims: jitter, [lens], areal, lightdir, choice, then one per bounce
    const uint32_t n2d = base2 + P->depth + P->reverse, n1d = 1u + P->depth;   // the light path continues the camera path's 2-D dims
    const bool user_tables = P->sampler_mode == RGK_SAMPLER_TABLES;
    if (user_tables) {
        uint64_t need = 0;
        for (uint32_t i = ctx->shard_first; i < n_tasks; i += (ctx->shard_stride ? ctx->shard_stride : 1u)) need += (uint64_t)(tasks[i].x2 - tasks[i].x1) * (tasks[i].y2 - tasks[i].y1);
        if (!ctx->d_user_t1 || !ctx->d_user_t2 || ctx->user_npix < need || ctx->user_n1d < n1d || ctx->user_n2d < n2d || ctx->user_ss != ss)
            return rgk_fail(ctx, RGK_ERR_INVALID, "RGK_SAMPLER_TABLES: rgk_render_set_tables must supply tables for every pixel of the call, "
                                                  "at least 1+depth 1-D and 4(+1 with a lens)+depth 2-D dims, set size rgk_sampler_set_size(multisample)");
    }
    if (P->sampler_mode == RGK_SAMPLER_MT19937 && (n2d > 64 || n1d > 64))
        return rgk_fail(ctx, RGK_ERR_UNSUPPORTED, "recursion-max above 59 leaves the 64 tabulated sampler dimensions (live-generator fallback, src/sampler.cpp:26-36, is not replicated)");
    if ((uint64_t)ss * ss > 0xFFFFFFFFull) return rgk_fail(ctx, RGK_ERR_UNSUPPORTED, "multisample too large");
    const bool mt = P->sampler_mode == RGK_SAMPLER_MT19937;
    const bool tables = mt || user_tables;
    size_t call_pixels = 0;
    const bool counting = ctx->counting;
    const rgk_device_cfg& cfg = ctx->cfg;              // scheduling and sizing only: none of it changes a result
    const uint32_t skip_null = cfg.skip_null_shadow ? 1u : 0u;
    const bool binning = cfg.binning && P->depth > 1;
    const bool bin_shadow0 = cfg.bin_shadow_first != 0;
    const bool path_order_shade = cfg.shade_path_order != 0;
    const double bin_min_frac = cfg.bin_min_frac;
    const size_t bin_items = cfg.bin_items;            // path slots per reordering group
    // the wide-BVH kernels' iterations are longer than the kd ones: refilling coherent warps at 24 idle lanes instead of 32
    // measured -1.9 ms per round (profiles/r1_bvh_sweep.json); the other thresholds are flat
    const bool bvh_round = ctx->dev.bvh_nodes != nullptr;
    const uint32_t refill_coherent = cfg.refill_coherent ? cfg.refill_coherent : (bvh_round ? 24u : 32u), refill_incoherent = cfg.refill_incoherent;
    const uint32_t refill_shadow = cfg.refill_shadow;  // any-hit rays end at very different times: refill sooner
    rgk_trav_stats* d_st = ctx->d_stats;               // [0] closest, [1] shadow
    if (counting) RGK_CUDA(ctx, cudaMemsetAsync(d_st, 0, 2 * sizeof(rgk_trav_stats), ctx->stream));
    bool shade_counts_cleared = false;

    // chunking: whole tiles, every multisample of a pixel in the same chunk
    size_t max_paths = (size_t)cfg.chunk_paths;        // default 128 Mi paths ~ 27 GB of path state: sized for 180 GB of HBM
    const size_t per_pixel_table = tables ? ((size_t)(n1d + 1) * 4 + (size_t)(n2d + 1) * 8) * ss : 0;
    size_t max_table_bytes = (size_t)cfg.table_bytes;
    {   // Never plan a chunk whose path state (~210 B per path, on top of what is already allocated) would not fit in 60 % of the
        // memory that is free right now (other contexts, smaller parts), nor tables beyond a quarter of it.  cudaMemGetInfo is
        // only asked when this call could need more than the context already holds: in the steady state of a frame (round
        // after round of the same size) it is skipped -- on a shared node the query was seen to take 1 - 90 ms, with the GPU idle.
        size_t call_pixels_total = 0;
        for (uint32_t i = ctx->shard_first; i < n_tasks; i += (ctx->shard_stride ? ctx->shard_stride : 1u))
            call_pixels_total += (size_t)(tasks[i].x2 - tasks[i].x1) * (tasks[i].y2 - tasks[i].y1);
        const size_t have_paths = ctx->paths ? ctx->paths->cap_paths : 0;
        const size_t have_tables = ctx->paths ? (ctx->paths->cap_t1 * 4 + ctx->paths->cap_t2 * 8) : 0;
        const bool paths_fit = have_paths >= std::min(max_paths, call_pixels_total * ms);
        const bool tables_fit = have_tables >= std::min(max_table_bytes, call_pixels_total * per_pixel_table);
        size_t free_b = 0, total_b = 0;
        if ((!paths_fit || !tables_fit) && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            if (!paths_fit) max_paths = std::min(max_paths, std::max<size_t>(have_paths + (size_t)(0.6 * (double)free_b / 210.0), (size_t)1 << 20));
            if (!tables_fit) max_table_bytes = std::min(max_table_bytes, std::max<size_t>(have_tables + (size_t)(0.25 * (double)free_b), (size_t)64 << 20));
        }
    }
    if (P->reverse) {      // bidirectional mode keeps every vertex of the camera and light paths: 112 B and 80 B per vertex
        const size_t per_path = 112 * (size_t)P->depth + 80 * (size_t)P->reverse + 64;
        max_paths = std::max<size_t>(std::min(max_paths, (size_t)cfg.reverse_bytes / per_path), 4096);
    }
    std::vector<uint4> h_tiles; std::vector<uint2> h_tiles2;
    rgk_round_stats total{};
    cudaEvent_t ev0 = ctx->ev[0], ev1 = ctx->ev[1];
    RGK_CUDA(ctx, cudaEventRecord(ev0, ctx->stream));
    if (!ctx->paths) ctx->paths = new PathBuffers();
    if (!ctx->paths->events) ctx->paths->events = new EventPool();
    EventPool& pool = *ctx->paths->events;
    pool.used = 0; pool.tag.clear();            // an earlier call that returned on an error between begin and end leaves nothing behind
    float kind_ms[T_KINDS] = {0, 0, 0, 0};
    // Scene::GetRandomLight with one point light and nothing else always returns it; with size 0 it is not jittered
    const bool const_light = !P->reverse && ctx->dev.n_point_lights == 1 && ctx->dev.n_areal_lights == 0 && ctx->first_point_light.size == 0.0f &&
                             ctx->first_point_light.intensity > 0.0f && cfg.const_light;
    // Which sampler tables does the round read?  (The generator walks through all of them either way; the warp-cooperative
    // builder does not build the others.)  2-D: the pixel jitter, the lens sample, the light's surface sample and choice
    // unless the light is the one fixed point light, and one direction per path vertex except the last, which never
    // continues.  1-D: the light's triangle sample (same condition) and the Russian-roulette cursor, which is at most
    // depth - 1 when it is last looked at.  The bidirectional mode keeps everything (its light paths continue the camera
    // path's dimensions).
    uint64_t keep1 = ~0ull, keep2 = ~0ull;
    if (!P->reverse && mt) {
        keep2 = 1ull | (lens ? 2ull : 0ull);
        keep1 = 0ull;
        if (!const_light) { keep2 |= (1ull << (base2 - 3u)) | (1ull << (base2 - 1u)); keep1 |= 1ull; }
        for (uint32_t j = 0; j + 1u < P->depth; j++) keep2 |= 1ull << (base2 + j);
        for (uint32_t c = 1; c < P->depth; c++) keep1 |= 1ull << c;
    }
    keep1 &= n1d >= 64u ? ~0ull : ((1ull << n1d) - 1ull);
    keep2 &= n2d >= 64u ? ~0ull : ((1ull << n2d) - 1ull);
    SamplerPlan splan{};
    // chunks whose counter blocks are on their way to pinned memory: (paths, bounces enqueued)
    std::vector<std::pair<uint32_t, uint32_t>> pending;
    std::vector<bool> term_recorded(P->depth + 1u, false);
    auto drain = [&]() -> rgk_status {
        RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const PathBuffers& Bd = *ctx->paths;
        for (size_t c = 0; c < pending.size(); c++) {
            const unsigned long long* h = Bd.h_counters + c * Bd.cap_bounces * C_COUNT;
            for (uint32_t b = 0; b < pending[c].second; b++) {
                total.closest_rays += b ? h[(size_t)(b - 1) * C_COUNT + C_NEXT] : pending[c].first;
                total.shadow_rays += h[(size_t)b * C_COUNT + C_SHADOW];
                total.shadow_rays_skipped += h[(size_t)b * C_COUNT + C_SHADOW_SKIPPED];
            }
        }
        pending.clear();
        return RGK_OK;
    };
    const uint32_t stride = ctx->shard_stride ? ctx->shard_stride : 1u;
    uint32_t ti = ctx->shard_first;
    while (ti < n_tasks) {
        h_tiles.clear(); h_tiles2.clear();
        size_t npix = 0;
        while (ti < n_tasks) {
            const rgk_task& t = tasks[ti];
            const size_t px = (size_t)(t.x2 - t.x1) * (t.y2 - t.y1);
            if (npix && ((npix + px) * ms > max_paths || (per_pixel_table && (npix + px) * per_pixel_table > max_table_bytes))) break;
            if (px) {
                h_tiles.push_back(make_uint4(t.x1, t.x2, t.y1, t.y2));
                h_tiles2.push_back(make_uint2((uint32_t)npix, seedstart + seedcount_base + ti));
            }
            npix += px; ti += stride;
        }
        if (npix == 0) continue;
        const size_t npaths = npix * ms;
        if (npaths > 0xFFFFFFF0ull) return rgk_fail(ctx, RGK_ERR_UNSUPPORTED, "a single chunk exceeds 2^32 paths");
        rgk_status s = ensure_buffers(ctx, npaths, npix, tables ? (size_t)(n1d + 1) * ss * npix : 0, tables ? (size_t)(n2d + 1) * ss * npix : 0,
                                      h_tiles.size(), mt ? (splan = sampler_plan(ctx->device, cfg, (uint32_t)npix, ss, (uint32_t)(__builtin_popcountll(keep1) + __builtin_popcountll(keep2)))).scratch_words : 0, P->depth);
        if (s != RGK_OK) return s;
        PathBuffers& B = *ctx->paths;
        RGK_CUDA(ctx, cudaMemcpyAsync(B.tiles, h_tiles.data(), h_tiles.size() * sizeof(uint4), cudaMemcpyHostToDevice, ctx->stream));
        RGK_CUDA(ctx, cudaMemcpyAsync(B.tiles2, h_tiles2.data(), h_tiles2.size() * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream));
        RenderConst R{};
        R.cam = *cam; R.xres = P->xres; R.yres = P->yres; R.ms = ms; R.depth = P->depth; R.clamp = P->clamp; R.russian = P->russian;
        R.bump_scale = P->bumpmap_scale; R.set_size = ss; R.n1d = n1d; R.n2d = n2d; R.base2 = base2; R.sampler_mode = P->sampler_mode;
        R.lens = lens; R.npix = (uint32_t)npix; R.skip_null_shadow = skip_null; R.binning = binning ? 1u : 0u;
        R.reverse = P->reverse; R.npaths = (uint32_t)npaths; R.count_shade = counting ? 1u : 0u;
        if (counting && !shade_counts_cleared) { RGK_CUDA(ctx, cudaMemsetAsync(B.shade_counts, 0, 8 * sizeof(unsigned long long), ctx->stream)); shade_counts_cleared = true; }
        {
            const DevPointLight& l0 = ctx->first_point_light;
            R.const_light = const_light ? 1u : 0u;
            const uint32_t flags = 1u; float fbits; std::memcpy(&fbits, &flags, 4);       // valid, FULL_SPHERE
            R.cl_pos = make_float4(l0.pos[0], l0.pos[1], l0.pos[2], fbits);
            R.cl_col = make_float4(l0.color[0], l0.color[1], l0.color[2], l0.intensity);
        }
        if (P->reverse) { s = ensure_reverse_buffers(ctx, npaths, P->depth, P->reverse); if (s != RGK_OK) return s; }
        // reordering groups: SG samples x PG pixel positions (a multiple of the 32-pixel blocks of k_pixel_setup)
        const uint32_t SG = (uint32_t)std::min<size_t>(ms, 128);
        const uint32_t PG = (uint32_t)std::max<size_t>(8, (bin_items / SG) / 8 * 8);
        const uint32_t n_pgroups = (uint32_t)((npix + PG - 1) / PG), n_sgroups = (ms + SG - 1) / SG;
        SamplerView smp{B.t1, B.t2, (uint32_t)npix, ss, sq, P->sampler_mode};
        pool.begin(ctx->stream, T_SAMPLER);
        k_pixel_setup<<<(unsigned)h_tiles.size(), 256, 0, ctx->stream>>>(B.tiles, B.tiles2, B.pix_xy, B.pix_seed, B.pix_src, (uint32_t)call_pixels);
        ctx->launches++;
        if (mt) {
            launch_sampler_mt(ctx->stream, splan, cfg.sampler_smem != 0, B.pix_seed, (uint32_t)npix, ss, sq, n1d, n2d, B.t1, B.t2, B.mt_state, keep1, keep2);
            ctx->launches++;
        } else if (user_tables) {
            k_tables_from_user<<<(unsigned)((npix + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_user_t1, ctx->d_user_t2, ctx->user_n1d, ctx->user_n2d,
                                                                                       B.pix_src, (uint32_t)npix, ss, n1d, n2d, B.t1, B.t2);
            ctx->launches++;
        }
        call_pixels += npix;
        pool.end(ctx->stream);
        pool.begin(ctx->stream, T_SHADE);
        k_raygen<<<(unsigned)((npaths + 127) / 128), 128, 0, ctx->stream>>>(ctx->dev, R, smp, B, P->reverse ? B.reverse->cam_o : nullptr);
        pool.end(ctx->stream);
        ctx->launches++;
        RGK_CUDA(ctx, cudaGetLastError());

        const int tgrid = machine_blocks(ctx, counting ? (const void*)k_closest<true, 9> : (const void*)k_closest<false, RGK_INCOH_MINB>, TRACE_THREADS);
        uint32_t count = (uint32_t)npaths;
        const uint32_t* queue = nullptr;
        uint32_t* qnext = B.queue_a;
        // wide BVH (RGK_TRAVERSAL_BVH at commit): BVH pass + kd arbiter pass per closest-hit / shadow launch.  The counting
        // instantiation stays on the kd kernels, and so do the bidirectional mode's shadow resolves and connection segments
        // (its closest-hit launches use the BVH)
        // counting rounds (rgk_render_set_counting) run the counting instantiations of the BVH kernels; the arbiter and the
        // bidirectional mode's extra segments are never counted
        const bool use_bvh = ctx->dev.bvh_nodes != nullptr;
        // the arbiter sees ~4e-4 of the rays, all of them long (grazing) traversals: spread them over many warps
        const int arb_grid = 148 * (int)std::max<uint32_t>(1u, cfg.arb_grid);
        const bool bvh_shadow_nosort = cfg.bvh_shadow_nosort != 0;        // A/B knob: any-hit children in slot order
        const bool bvh_closest_nearest = cfg.bvh_closest_nearest != 0;    // A/B knob: nearest child first, no full sort
        uint32_t* arb_list = nullptr; unsigned long long* arb_ctr = nullptr;
        if (use_bvh) {
            arb_list = (uint32_t*)rgk_scratch(ctx, 4, npaths * sizeof(uint32_t));
            unsigned long long* s3 = (unsigned long long*)rgk_scratch(ctx, 3, 256);
            if (!arb_list || !s3) return rgk_fail(ctx, RGK_ERR_NOMEM, "scratch allocation failed");
            arb_ctr = s3 + 8;               // [0] closest arbiter work, [1] its count, [2] shadow arbiter work, [3] its count
        }
        if (P->reverse) {
            // ---- bidirectional mode (reverse_device.cuh): camera paths kept vertex by vertex, then the light paths,
            // then the connections, then the per-vertex sums
            const ReverseBuffers V = *B.reverse;
            DevScene dev = ctx->dev;
            RGK_CUDA(ctx, cudaMemsetAsync(V.nverts, 0, npaths * sizeof(uint32_t), ctx->stream));
            auto counts = [&](uint32_t& next_count, uint32_t& shadow_count) -> rgk_status {
                RGK_CUDA(ctx, cudaMemcpyAsync(B.h_counters, B.counters, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
                RGK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                next_count = (uint32_t)B.h_counters[C_NEXT]; shadow_count = (uint32_t)B.h_counters[C_SHADOW];
                total.shadow_rays_skipped += B.h_counters[C_SHADOW_SKIPPED];
                return RGK_OK;
            };
            auto closest = [&](const uint32_t* q, uint32_t n, bool coherent) {
                dev.refill_threshold = coherent ? refill_coherent : refill_incoherent;
                const int g = (int)std::min<uint64_t>(tgrid, ((uint64_t)n + TRACE_THREADS - 1) / TRACE_THREADS);
                pool.begin(ctx->stream, T_CLOSEST);
                if (use_bvh) {
                    cudaMemsetAsync(arb_ctr, 0, 2 * sizeof(unsigned long long), ctx->stream);
                    k_closest_bvh<RGK_CLOSEST_BVH_MINB><<<g, TRACE_THREADS, 0, ctx->stream>>>(dev, B, q, QueueLen{nullptr, n}, B.counters + C_WORK_A, ctx->d_bvh_stats, arb_list, (uint32_t*)(arb_ctr + 1));
                    k_closest_arb<RGK_INCOH_MINB><<<std::min(g, arb_grid), TRACE_THREADS, 0, ctx->stream>>>(dev, B, arb_list, (const uint32_t*)(arb_ctr + 1), arb_ctr, 0u);
                    ctx->launches++;
                }
                else if (coherent) k_closest<false, RGK_COH_MINB><<<g, TRACE_THREADS, 0, ctx->stream>>>(dev, B, q, QueueLen{nullptr, n}, B.counters + C_WORK_A, nullptr);
                else k_closest<false, RGK_INCOH_MINB><<<g, TRACE_THREADS, 0, ctx->stream>>>(dev, B, q, QueueLen{nullptr, n}, B.counters + C_WORK_A, nullptr);
                pool.end(ctx->stream);
                ctx->launches++; total.closest_launches++; total.closest_rays += n;
            };
            for (uint32_t bounce = 0; bounce < P->depth && count > 0; bounce++) {           // camera paths
                RGK_CUDA(ctx, cudaMemsetAsync(B.counters, 0, 5 * sizeof(unsigned long long), ctx->stream));
                closest(queue, count, bounce == 0);
                pool.begin(ctx->stream, T_SHADE);
                k_shade_rev<false><<<(count + 127) / 128, 128, 0, ctx->stream>>>(ctx->dev, R, smp, B, V, queue, count, qnext, B.queue_s, B.counters);
                pool.end(ctx->stream);
                ctx->launches++;
                uint32_t next_count = 0, shadow_count = 0;
                rgk_status cs = counts(next_count, shadow_count); if (cs != RGK_OK) return cs;
                total.shadow_rays += shadow_count;
                if (shadow_count) {
                    dev.refill_threshold = bounce == 0 ? refill_coherent : refill_shadow;
                    const int g2 = (int)std::min<uint64_t>(tgrid, ((uint64_t)shadow_count + TRACE_THREADS - 1) / TRACE_THREADS);
                    pool.begin(ctx->stream, T_SHADOW);
                    k_shadow_rev<<<g2, TRACE_THREADS, 0, ctx->stream>>>(dev, R, B, V, B.queue_s, shadow_count, B.counters + C_WORK_B);
                    pool.end(ctx->stream);
                    ctx->launches++; total.shadow_launches++;
                }
                queue = qnext; qnext = (qnext == B.queue_a) ? B.queue_b : B.queue_a;
                count = next_count;
            }
            RGK_CUDA(ctx, cudaMemsetAsync(B.counters, 0, 5 * sizeof(unsigned long long), ctx->stream));   // light paths
            pool.begin(ctx->stream, T_SHADE);
            k_lightgen<<<(unsigned)((npaths + 127) / 128), 128, 0, ctx->stream>>>(ctx->dev, R, smp, B, V, B.queue_a, B.counters);
            pool.end(ctx->stream);
            ctx->launches++;
            uint32_t lcount = 0, dummy = 0;
            rgk_status cs = counts(lcount, dummy); if (cs != RGK_OK) return cs;
            const uint32_t* lq = B.queue_a; uint32_t* lnext = B.queue_b;
            for (uint32_t b = 0; b < P->reverse && lcount > 0; b++) {
                RGK_CUDA(ctx, cudaMemsetAsync(B.counters, 0, 5 * sizeof(unsigned long long), ctx->stream));
                closest(lq, lcount, false);
                pool.begin(ctx->stream, T_SHADE);
                k_shade_rev<true><<<(lcount + 127) / 128, 128, 0, ctx->stream>>>(ctx->dev, R, smp, B, V, lq, lcount, lnext, B.queue_s, B.counters);
                pool.end(ctx->stream);
                ctx->launches++;
                cs = counts(lcount, dummy); if (cs != RGK_OK) return cs;
                lq = lnext; lnext = (lnext == B.queue_a) ? B.queue_b : B.queue_a;
            }
            RGK_CUDA(ctx, cudaMemsetAsync(B.counters, 0, 5 * sizeof(unsigned long long), ctx->stream));   // connections
            pool.begin(ctx->stream, T_SHADOW);
            k_connect_camera<<<(unsigned)((npaths * P->reverse + 127) / 128), 128, 0, ctx->stream>>>(ctx->dev, R, B, V, d_rgb, B.counters);
            k_connect_vertices<<<(unsigned)((npaths * P->depth + 127) / 128), 128, 0, ctx->stream>>>(ctx->dev, R, V, B.counters);
            pool.end(ctx->stream);
            pool.begin(ctx->stream, T_SHADE);
            k_assemble<<<(unsigned)((npaths + 127) / 128), 128, 0, ctx->stream>>>(R, B, V);
            pool.end(ctx->stream);
            ctx->launches += 3;
            cs = counts(dummy, lcount); if (cs != RGK_OK) return cs;
            total.shadow_rays += lcount;                       // connection rays (Visibility calls of phases 2 and 3)
            count = 0;
        }
        // ---- unidirectional bounce loop.  Nothing in it waits for the device: queue lengths stay in the per-bounce counter
        // blocks and every kernel reads the ones it needs there (persistent grids sized for the machine, not for the queue), so the
        // whole chunk is enqueued back to back and host jitter cannot open gaps between its kernels.  The lengths travel to the
        // host once, at the end of the chunk, for the statistics.  Deep paths (recursion-max > 2): bounce b's surviving-path count
        // is also copied back right after the bounce; counts never grow, so one that has ARRIVED (no wait) bounds every later
        // bounce: it sizes their grids, and once it is zero the host stops enqueueing (the bounces enqueued in between find empty
        // queues and exit).
        if (!P->reverse) {
            const uint32_t nb = P->depth;
            RGK_CUDA(ctx, cudaMemsetAsync(B.counters, 0, (size_t)nb * C_COUNT * sizeof(unsigned long long), ctx->stream));
            if (pending.size() == H_CHUNKS) { rgk_status ds = drain(); if (ds != RGK_OK) return ds; }
            unsigned long long* h_term = B.h_counters + H_CHUNKS * B.cap_bounces * C_COUNT;
            if (!B.term_events) B.term_events = new std::vector<cudaEvent_t>();
            const uint32_t thresh = binning ? (uint32_t)std::min<double>(4294967295.0, std::ceil(bin_min_frac * (double)npaths)) : 0xFFFFFFFFu;
            uint64_t live_ub = npaths;                  // upper bound of the live paths of the bounce being enqueued
            const uint32_t* unsorted_prev = nullptr;    // the previous bounce's continuation queue in path order (if it may have been binned)
            uint32_t* unext = B.queue_ua;
            uint32_t launched = 0;
            std::fill(term_recorded.begin(), term_recorded.end(), false);
            for (uint32_t bounce = 0; bounce < nb; bounce++) {
                for (uint32_t back = 1; back <= 2 && back <= bounce; back++)       // a count that has arrived (never waits)
                    if (term_recorded[bounce - back] && cudaEventQuery((*B.term_events)[bounce - back]) == cudaSuccess)
                        live_ub = std::min<uint64_t>(live_ub, h_term[bounce - back]);
                if (live_ub == 0) break;
                unsigned long long* blk = B.counters + (size_t)bounce * C_COUNT;
                const QueueLen in{bounce ? blk - C_COUNT + C_NEXT : nullptr, (uint32_t)npaths};
                const QueueLen prev_in{bounce > 1 ? blk - 2 * C_COUNT + C_NEXT : nullptr, (uint32_t)npaths};
                const QueueLen shadow_len{blk + C_SHADOW, 0u};
                // camera rays (and the shadow rays of their hit points) are coherent: keep warps in lockstep (refill only
                // when the whole warp is done); later bounces are incoherent: refill as soon as a quarter of the warp idles
                DevScene dev = ctx->dev;
                dev.refill_threshold = bounce == 0 ? refill_coherent : refill_incoherent;
                const int g1 = (int)std::min<uint64_t>(tgrid, (live_ub + TRACE_THREADS - 1) / TRACE_THREADS);
                pool.begin(ctx->stream, T_CLOSEST);
                // two register budgets of the kd kernel: 56 registers (9 CTAs/SM) for the issue-bound coherent camera rays,
                // 48 registers (10 CTAs/SM) for later bounces, which gain from the extra warps
                if (use_bvh) {
                    uint32_t* arb_n = (uint32_t*)(blk + C_ARB_COUNT);
                    if (counting) k_closest_bvh<6, 1, true><<<g1, TRACE_THREADS, 0, ctx->stream>>>(dev, B, queue, in, blk + C_WORK_A, ctx->d_bvh_stats, arb_list, arb_n);
                    else if (bvh_closest_nearest) k_closest_bvh<RGK_CLOSEST_BVH_MINB, 2><<<g1, TRACE_THREADS, 0, ctx->stream>>>(dev, B, queue, in, blk + C_WORK_A, ctx->d_bvh_stats, arb_list, arb_n);
                    else if (bounce == 0) k_closest_bvh<RGK_CLOSEST_BVH_MINB_CAMERA><<<g1, TRACE_THREADS, 0, ctx->stream>>>(dev, B, queue, in, blk + C_WORK_A, ctx->d_bvh_stats, arb_list, arb_n);
                    else k_closest_bvh<RGK_CLOSEST_BVH_MINB><<<g1, TRACE_THREADS, 0, ctx->stream>>>(dev, B, queue, in, blk + C_WORK_A, ctx->d_bvh_stats, arb_list, arb_n);
                    k_closest_arb<RGK_INCOH_MINB><<<std::min(g1, arb_grid), TRACE_THREADS, 0, ctx->stream>>>(dev, B, arb_list, arb_n, blk + C_ARB_WORK, bounce == 0 ? 1u : 0u);
                    ctx->launches++;
                }
                else if (counting) k_closest<true, 9><<<g1, TRACE_THREADS, 0, ctx->stream>>>(dev, B, queue, in, blk + C_WORK_A, d_st);
                else if (bounce == 0) k_closest<false, RGK_COH_MINB><<<g1, TRACE_THREADS, 0, ctx->stream>>>(dev, B, queue, in, blk + C_WORK_A, nullptr);
                else k_closest<false, RGK_INCOH_MINB><<<g1, TRACE_THREADS, 0, ctx->stream>>>(dev, B, queue, in, blk + C_WORK_A, nullptr);
                pool.end(ctx->stream);
                pool.begin(ctx->stream, T_SHADE);
                const bool last_bounce = bounce + 1 >= nb;       // no continuation rays: k_shade ends every path (n == depth)
                // camera-ray hit points are already in image order: their shadow rays are binned only on request.  k_bin reads
                // every slot of the chunk, the traversal only gains on the live ones: bounces with few survivors (fewer than
                // bin_thresh, decided on the device) go back to the atomic compaction
                R.binning = (binning && !last_bounce ? 1u : 0u) | (binning && (bounce > 0 || bin_shadow0) ? 2u : 0u);
                R.bin_thresh = thresh;
                R.first_bounce = bounce == 0 ? 1u : 0u;
                if (R.binning & 1u) RGK_CUDA(ctx, cudaMemsetAsync(B.key_next, 0xFF, npaths, ctx->stream));
                if (R.binning & 2u) RGK_CUDA(ctx, cudaMemsetAsync(B.key_shadow, 0xFF, npaths, ctx->stream));
                const unsigned sg = (unsigned)((live_ub + RGK_SHADE_THREADS - 1) / RGK_SHADE_THREADS);
                if (last_bounce) k_shade<true><<<sg, RGK_SHADE_THREADS, 0, ctx->stream>>>(ctx->dev, R, smp, B, queue, unsorted_prev, in, prev_in, qnext, B.queue_s, unext, blk);
                else k_shade<false><<<sg, RGK_SHADE_THREADS, 0, ctx->stream>>>(ctx->dev, R, smp, B, queue, unsorted_prev, in, prev_in, qnext, B.queue_s, unext, blk);
                if (R.binning & 1u) {
                    k_bin<<<n_pgroups * n_sgroups, BIN_THREADS, 0, ctx->stream>>>(B.key_next, (uint32_t)npix, ms, PG, SG, n_pgroups, qnext, blk + C_NEXT, in, thresh);
                    ctx->launches++;
                }
                if (R.binning & 2u) {
                    k_bin<<<n_pgroups * n_sgroups, BIN_THREADS, 0, ctx->stream>>>(B.key_shadow, (uint32_t)npix, ms, PG, SG, n_pgroups, B.queue_s, blk + C_SHADOW, in, thresh);
                    ctx->launches++;
                }
                pool.end(ctx->stream);
                ctx->launches += 2; total.closest_launches++;
                if (nb > 2 && !last_bounce) {         // early read-back of the survivors' count (see above)
                    while (B.term_events->size() <= bounce) { cudaEvent_t e; RGK_CUDA(ctx, cudaEventCreate(&e)); B.term_events->push_back(e); }
                    RGK_CUDA(ctx, cudaMemcpyAsync(h_term + bounce, blk + C_NEXT, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
                    RGK_CUDA(ctx, cudaEventRecord((*B.term_events)[bounce], ctx->stream));
                    term_recorded[bounce] = true;
                }
                {
                    dev.refill_threshold = bounce == 0 ? refill_coherent : refill_shadow;
                    const int g2 = (int)std::min<uint64_t>(tgrid, (live_ub + TRACE_THREADS - 1) / TRACE_THREADS);
                    pool.begin(ctx->stream, T_SHADOW);
                    if (use_bvh) {
                        uint32_t* arb_n = (uint32_t*)(blk + C_ARB_SCOUNT);
                        if (counting) k_shadow_bvh<6, 1, true><<<g2, TRACE_THREADS, 0, ctx->stream>>>(dev, B, B.queue_s, shadow_len, P->clamp, blk + C_WORK_B, ctx->d_bvh_stats + 1,
                                                                                              R.const_light, R.cl_pos, arb_list, arb_n);
                        else if (bvh_shadow_nosort) k_shadow_bvh<RGK_SHADOW_BVH_MINB, 0><<<g2, TRACE_THREADS, 0, ctx->stream>>>(dev, B, B.queue_s, shadow_len, P->clamp, blk + C_WORK_B, ctx->d_bvh_stats + 1,
                                                                                                    R.const_light, R.cl_pos, arb_list, arb_n);
                        else k_shadow_bvh<RGK_SHADOW_BVH_MINB, 1><<<g2, TRACE_THREADS, 0, ctx->stream>>>(dev, B, B.queue_s, shadow_len, P->clamp, blk + C_WORK_B, ctx->d_bvh_stats + 1,
                                                                                    R.const_light, R.cl_pos, arb_list, arb_n);
                        k_shadow_arb<RGK_INCOH_MINB><<<std::min(g2, arb_grid), TRACE_THREADS, 0, ctx->stream>>>(dev, B, arb_list, arb_n, P->clamp, blk + C_ARB_SWORK,
                                                                                                         R.const_light, R.cl_pos);
                        ctx->launches++;
                    }
                    else if (counting) k_shadow<true, 9><<<g2, TRACE_THREADS, 0, ctx->stream>>>(dev, B, B.queue_s, shadow_len, P->clamp, blk + C_WORK_B, d_st + 1, R.const_light, R.cl_pos);
                    else k_shadow<false, RGK_INCOH_MINB><<<g2, TRACE_THREADS, 0, ctx->stream>>>(dev, B, B.queue_s, shadow_len, P->clamp, blk + C_WORK_B, nullptr, R.const_light, R.cl_pos);
                    pool.end(ctx->stream);
                    ctx->launches++; total.shadow_launches++;
                }
                queue = qnext; qnext = (qnext == B.queue_a) ? B.queue_b : B.queue_a;
                unsorted_prev = ((R.binning & 1u) && path_order_shade) ? unext : nullptr;
                unext = (unext == B.queue_ua) ? B.queue_ub : B.queue_ua;
                launched++;
            }
            // the chunk's counter blocks: to pinned memory now, into the statistics when the round is over
            RGK_CUDA(ctx, cudaMemcpyAsync(B.h_counters + pending.size() * B.cap_bounces * C_COUNT, B.counters, (size_t)nb * C_COUNT * sizeof(unsigned long long),
                                          cudaMemcpyDeviceToHost, ctx->stream));
            pending.push_back(std::make_pair((uint32_t)npaths, launched));
        }
        pool.begin(ctx->stream, T_SHADE);
        k_finish<<<(unsigned)((npix + 127) / 128), 128, 0, ctx->stream>>>(R, B, d_rgb, d_count);
        pool.end(ctx->stream);
        ctx->launches++;
        RGK_CUDA(ctx, cudaGetLastError());
        total.samples += npaths;
    }
    RGK_CUDA(ctx, cudaEventRecord(ev1, ctx->stream));
    RGK_CUDA(ctx, cudaEventSynchronize(ev1));
    { rgk_status ds = drain(); if (ds != RGK_OK) return ds; }
    float ms_total = 0.0f; cudaEventElapsedTime(&ms_total, ev0, ev1);
    pool.collect(kind_ms);
    total.gpu_ms = ms_total; total.closest_ms = kind_ms[T_CLOSEST]; total.shadow_ms = kind_ms[T_SHADOW];
    total.sampler_ms = kind_ms[T_SAMPLER]; total.shade_ms = kind_ms[T_SHADE];
    total.trace_ms = total.closest_ms + total.shadow_ms; total.kernel_launches = ctx->launches - launches0;
    if (counting) {
        RGK_CUDA(ctx, cudaMemcpy(&ctx->last_closest, d_st, sizeof(rgk_trav_stats), cudaMemcpyDeviceToHost));
        RGK_CUDA(ctx, cudaMemcpy(&ctx->last_shadow, d_st + 1, sizeof(rgk_trav_stats), cudaMemcpyDeviceToHost));
        if (ctx->paths && ctx->paths->shade_counts) RGK_CUDA(ctx, cudaMemcpy(ctx->last_shade, ctx->paths->shade_counts, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    }
    if (stats) *stats = total;
    return RGK_OK;
}
