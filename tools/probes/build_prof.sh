#!/bin/bash
# MM_PROF build of the library (per-warp time stamps) next to the product build: 3d-magic-mirror_b200/libmagicmirror_prof.so
cd "$(dirname "$0")/../../3d-magic-mirror_b200/csrc" && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -shared -DMM_PROF -o ../libmagicmirror_prof.so mm_abi.cu mm_vertex.cu mm_raster.cu mm_fused.cu mm_loss.cu mm_meshreg.cu mm_template.cu mm_texflow.cu
