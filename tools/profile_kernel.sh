#!/bin/bash
# usage (GPU box): tools/profile_kernel.sh <tag> <kernel regex> [skip] [count]  -- ncu --set full of selected kernels (with source)
TAG=$1; KRE=$2; SKIP=${3:-8}; CNT=${4:-1}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c $CNT -f -o gpurun_out/${TAG} \
    python bench.py --steps 6 --warmup 3 --profile > gpurun_out/${TAG}.log 2>&1
