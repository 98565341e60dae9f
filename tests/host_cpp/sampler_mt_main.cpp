// sampler_mt_main.cpp -- the 32-lane host build of the sampler kernels (sampler_mt.cpp) as a plain program, for the sanitizers:
// ThreadSanitizer sees every __syncwarp / __syncthreads as a pthread barrier, so a shared-memory access of one lane that is not
// ordered against another lane's by a barrier is reported as a data race; AddressSanitizer checks the bounds of the shared and
// global arrays.  tools/tsan_lanes_on_host.sh builds and runs it twice.  Prints one checksum per configuration.  (The thread-per-pixel
// kernel gets whole warps of pixels here: its lanes past the end of the chunk leave the kernel, which a pthread barrier -- unlike
// bar.warp.sync -- keeps waiting for.)
#include <cstdint>
#include <cstdio>
#include <vector>
extern "C" int doh32_sampler_tables(const uint32_t* seeds, uint32_t n, uint32_t multisample, uint32_t n1d, uint32_t n2d, float* out1, float* out2,
                                    uint32_t kernel, uint32_t slots, const char** error);
static uint32_t set_size(uint32_t x) { uint32_t i = 0; while (i * i < x) i++; return i * i; }
int main(int argc, char** argv) {
    const bool quick = argc > 1;        // any argument: the three smallest configurations (ThreadSanitizer slows every barrier a thousandfold)
    struct Cfg { uint32_t ms, n, n1d, n2d, kernel, slots; };
    const Cfg cfgs[] = {{64, 21, 3, 5, 2, 0}, {64, 9, 2, 3, 2, 2}, {36, 19, 3, 4, 2, 3}, {121, 9, 1, 2, 2, 0}, {256, 9, 1, 1, 2, 0}, {400, 9, 1, 2, 2, 2},
                        {4, 40, 5, 6, 2, 0}, {16, 75, 3, 4, 1, 0}, {1, 40, 6, 6, 2, 0}};
    int index = 0;
    for (const Cfg& c : cfgs) {
        if (quick && !(index == 1 || index == 3 || index == 4 || index == 8)) { index++; continue; }
        index++;
        std::vector<uint32_t> seeds(c.n);
        for (uint32_t i = 0; i < c.n; i++) seeds[i] = 0x42424242u * (i + 1u) + 42u;
        const uint32_t ss = set_size(c.ms);
        std::vector<float> o1((size_t)c.n1d * ss * c.n), o2((size_t)c.n2d * ss * c.n * 2);
        const char* err = nullptr;
        const int st = doh32_sampler_tables(seeds.data(), c.n, c.ms, c.n1d, c.n2d, o1.data(), o2.data(), c.kernel, c.slots, &err);
        double s = 0; for (float v : o1) s += v; for (float v : o2) s += v;
        std::printf("ms %u seeds %u kernel %u slots %u: status %d checksum %.6f %s\n", c.ms, c.n, c.kernel, c.slots, st, s, st ? err : "");
        std::fflush(stdout);
        if (st) return 1;
    }
    return 0;
}
