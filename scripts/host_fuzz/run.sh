#!/bin/bash
# Sanitizer run of the two host-side parsers that read untrusted files (CPU only; no GPU, no librbrt_gpu.so):
#   csrc/obj_loader.cpp   under ASan + UBSan (one piece, and many 8-byte pieces on 6 threads) and under TSan
#   csrc/host/scene_yaml.hpp under ASan + UBSan
# on 1600 generated .obj and 1600 generated .yaml files (the hypothesis generators of tests/test_obj_loader.py and
# tests/test_scene_yaml.py, each file also in three byte-mutated copies).  Expected output: four "rc=0" lines and no report.
# Last run: round 2, all clean (also on the 53 MB / 1.31 M-triangle C3 .obj).
set -e
here=$(cd "$(dirname "$0")" && pwd); root=$(cd "$here/../.." && pwd)
work=${1:-/tmp/rbrt_host_fuzz}; mkdir -p "$work"; cd "$work"
san="-O1 -g -std=c++17 -ffp-contract=off"
g++ $san -fsanitize=address,undefined -fno-sanitize-recover=undefined -o obj_asan "$here/obj_main.cpp" "$root/rbrt_b200/csrc/obj_loader.cpp" -lpthread
g++ $san -fsanitize=thread -o obj_tsan "$here/obj_main.cpp" "$root/rbrt_b200/csrc/obj_loader.cpp" -lpthread
g++ $san -fsanitize=address,undefined -fno-sanitize-recover=undefined -o yaml_asan "$here/yaml_main.cpp"
python "$here/gen.py"
printf 'newmtl red\nKd 1 0 0\nnewmtl blue\n' > corpus/m.mtl; printf 'newmtl none\nNs x\n' > corpus/bad.mtl
ls corpus/*.obj | xargs -n 200 ./obj_asan; echo "obj asan rc=$?"
ls corpus/*.obj | RBRT_OBJ_PIECE_BYTES=8 RBRT_HOST_THREADS=6 xargs -n 200 ./obj_asan; echo "obj asan, many pieces rc=$?"
ls corpus/*.obj | RBRT_OBJ_PIECE_BYTES=8 RBRT_HOST_THREADS=6 xargs -n 100 ./obj_tsan; echo "obj tsan rc=$?"
ls corpus/*.yaml | xargs -n 200 ./yaml_asan; echo "yaml asan rc=$?"
