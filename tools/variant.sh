#!/bin/bash
# Builds lib/libnislam_<tag>.so from the product objects with the named translation units recompiled under extra -D flags:
#   tools/variant.sh <tag> "<-D flags>" nis_row.cu [nis_col.cu ...]      (run the product build first)
set -e
tag=$1; flags=$2; shift 2
cd "$(dirname "$0")/../ni_slam_b200"
mkdir -p build_$tag
for u in nis_col nis_row nis_misc nis_api nis_stitch; do cp -f build/$u.o build_$tag/$u.o; done
for u in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --threads 4 $flags -c csrc/$u -o build_$tag/${u%.cu}.o &
done
wait
nvcc -shared -o lib/libnislam_$tag.so build_$tag/*.o -gencode arch=compute_100a,code=sm_100a -ldl
echo built lib/libnislam_$tag.so
