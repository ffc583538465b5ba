#!/bin/bash
# Builds liborcdemux.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
     -shared -Xcompiler -fPIC ${ORC_NVCC_EXTRA:-} \
     -o "$here/orcdemux/liborcdemux.so" "$here/csrc/orc_api.cu" "$here/csrc/orc_edit.cu" "$here/csrc/orc_io.cpp" -lz -lpthread
