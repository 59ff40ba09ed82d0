#!/bin/bash
# Builds a side-by-side variant of the library from another source tree (development only):
#   tools/build_variant.sh NAME /path/to/tree   ->  py-tdgl_b200/_variants/NAME.so
# run it with TDGL_B200_LIB=py-tdgl_b200/_variants/NAME.so (see py-tdgl_b200/_lib.py).
set -e
name="$1"; tree="${2:-.}"
mkdir -p py-tdgl_b200/_variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
  -o "py-tdgl_b200/_variants/$name.so" "$tree/py-tdgl_b200/csrc/tdgl_b200.cu" -ldl
echo "built py-tdgl_b200/_variants/$name.so"
