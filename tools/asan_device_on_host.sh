#!/bin/bash
# The device sources on the host (tests/host_cpp) under AddressSanitizer + UndefinedBehaviorSanitizer: every kernel of the
# wavefront, the traversal and the shading code run through tests/test_device_on_host.py with out-of-bounds and
# undefined-behaviour checks (compute-sanitizer is not available on the GPU pool; this is the substitute).
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off -fPIC -shared -Wl,-Bsymbolic \
    -I/usr/local/cuda/include -Ibuild/host/gen -Iinclude tests/host_cpp/device_on_host.cpp -o build/host/libdevice_on_host_asan.so \
    -Lrgk_b200 -lrgk_b200 -Wl,-rpath,'$ORIGIN/../../rgk_b200'
RGK_DEVICE_ON_HOST_SO=$PWD/build/host/libdevice_on_host_asan.so \
LD_PRELOAD="$(g++ -print-file-name=libasan.so) $(g++ -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
UBSAN_OPTIONS=print_stacktrace=1 python -m pytest tests/test_device_on_host.py -x -q -s 2>&1 | tee build/host/asan.log | grep -E "runtime error|AddressSanitizer|passed|failed" || true
if grep -qE "runtime error|AddressSanitizer" build/host/asan.log; then echo "SANITIZER FINDINGS (build/host/asan.log)"; exit 1; fi
echo "sanitizers: clean"
