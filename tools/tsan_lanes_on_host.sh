#!/bin/bash
# The warp-cooperative sampler kernels on the host with 32-lane warps made of real threads (tests/host_cpp/device_shim_mt.h),
# under ThreadSanitizer (a shared-memory access not ordered by __syncwarp / __syncthreads against another lane's is a data race)
# and under AddressSanitizer + UndefinedBehaviorSanitizer (bounds of shared / global arrays, misaligned wide loads).
# compute-sanitizer is closed on the GPU pool; this is the substitute for racecheck / memcheck on that code.
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
python tests/host_cpp/gen_host_sources.py rgk_b200/csrc build/host/gen32 --lanes32
SRC="tests/host_cpp/sampler_mt.cpp tests/host_cpp/sampler_mt_main.cpp"
INC="-I/usr/local/cuda/include -Ibuild/host/gen32 -Iinclude -Itests/host_cpp"
LNK="-Lrgk_b200 -lrgk_b200 -Wl,-rpath,$PWD/rgk_b200 -pthread"
g++ -std=c++17 -O1 -g -fsanitize=thread -ffp-contract=off $INC $SRC -o build/host/sampler_mt_tsan $LNK
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off $INC $SRC -o build/host/sampler_mt_asan $LNK
TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" build/host/sampler_mt_tsan 2>&1 | tee build/host/tsan_sampler.log | grep -E "checksum|WARNING: ThreadSanitizer|SUMMARY" | head -40
ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 build/host/sampler_mt_asan 2>&1 | tee build/host/asan_sampler.log | grep -E "checksum|runtime error|AddressSanitizer" | head -40
# whole rounds of the wavefront with 32-lane warps (k_bin, queue compaction, persistent traversal, host-free chunk) under ThreadSanitizer
g++ -std=c++17 -O1 -g -fsanitize=thread -ffp-contract=off -fPIC -shared -pthread -DRGK_DOH_MT -fvisibility=hidden -fno-gnu-unique -Wl,-Bsymbolic \
    $INC tests/host_cpp/device_on_host.cpp -o build/host/libdevice_on_host_mt_tsan.so -Lrgk_b200 -lrgk_b200 -Wl,-rpath,$PWD/rgk_b200
DOH_MT_SO=$PWD/build/host/libdevice_on_host_mt_tsan.so LD_PRELOAD=$(g++ -print-file-name=libtsan.so) TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0" \
    python tools/tsan_wavefront_round.py 2>&1 | tee build/host/tsan_wavefront.log | grep -E "^round:|WARNING: ThreadSanitizer|SUMMARY" | head -40
if grep -qE "WARNING: ThreadSanitizer" build/host/tsan_wavefront.log || ! grep -q "identical to the oracle's: True" build/host/tsan_wavefront.log; then
  echo "THREAD SANITIZER FINDINGS OR MISMATCH IN THE WAVEFRONT (build/host/tsan_wavefront.log)"; exit 1; fi
if grep -qE "WARNING: ThreadSanitizer" build/host/tsan_sampler.log; then echo "THREAD SANITIZER FINDINGS (build/host/tsan_sampler.log)"; exit 1; fi
if grep -qE "runtime error|AddressSanitizer" build/host/asan_sampler.log; then echo "ADDRESS / UB SANITIZER FINDINGS (build/host/asan_sampler.log)"; exit 1; fi
echo "sampler and wavefront on the host, 32-lane warps: ThreadSanitizer clean; sampler also AddressSanitizer + UBSan clean"
