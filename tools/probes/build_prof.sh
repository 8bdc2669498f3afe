#!/bin/bash
# MM_PROF build of the library (per-warp time stamps) next to the product build: 3d-magic-mirror_b200/libmagicmirror_prof.so
# (the stamps cost registers: the shading kernel is built at 4 CTAs / SM here -- at the product's 96-register cap they would
# spill inside its loops and distort what is being measured; the hard pass runs at 37 instead of 28 registers, i.e. 6 instead of
# 8 CTAs / SM, which shows as ~5 % late starters in its timeline)
cd "$(dirname "$0")/../../3d-magic-mirror_b200/csrc" && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -shared -DMM_PROF -DFUSED_MINB=4 -o ../libmagicmirror_prof.so mm_abi.cu mm_vertex.cu mm_raster.cu mm_fused.cu mm_loss.cu mm_meshreg.cu mm_template.cu mm_texflow.cu
