#!/bin/bash
# The host build of the product's device headers (tests/host_device) under ASan + UBSan, driven by tests/test_device_source_on_host.py:
# closest-hit queries, BVH traversals (stack of RBRT_STACK entries, leaf decoding, 16-bit node decode), scatter, complete renders.
# compute-sanitizer is closed on the GPU pool; this is the memory-safety evidence available for that code.  Last run: round 2, 11 tests, no report.
set -e
cd "$(dirname "$0")/../.."
hd=tests/host_device
mkdir -p $hd/build
cp -f $hd/build/libhost_device.so /tmp/libhost_device_plain.so 2>/dev/null || true
g++ -O1 -g -std=c++17 -ffp-contract=off -fno-fast-math -mno-fma -DRBRT_LDG128 -fsanitize=address,undefined -fno-sanitize-recover=undefined \
    -Wno-unknown-pragmas -fPIC -shared -I $hd -I rbrt_b200/csrc -o $hd/build/libhost_device.so $hd/harness.cpp
touch $hd/build/libhost_device.so
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 \
    python -m pytest tests/test_device_source_on_host.py -x -q || rc=$?
rm -f $hd/build/libhost_device.so                     # the next test run rebuilds the plain library
exit ${rc:-0}
