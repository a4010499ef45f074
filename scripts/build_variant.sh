#!/bin/bash
# Development aid: build libhuffb200 with extra -D flags into variants/<name>/libhuffb200.so (git-ignored, travels
# to the GPU box); use with HF_LIB_PATH=variants/<name>/libhuffb200.so.   scripts/build_variant.sh name -DX=1 ...
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
dst=$root/variants/$name
mkdir -p "$dst/obj"
for f in hist codebook encode2 decode decode2 sharded programs api; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c -o "$dst/obj/$f.o" "$root/huffman_b200/csrc/$f.cu" &
done
wait; for f in hist codebook encode2 decode decode2 sharded programs api; do test -f "$dst/obj/$f.o" || { echo "compile of $f.cu failed"; exit 1; }; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$dst/libhuffb200.so" "$dst"/obj/*.o -ldl
rm -rf "$dst/obj"
echo "built $dst/libhuffb200.so"
