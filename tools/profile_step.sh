#!/bin/bash
# usage (GPU box): tools/profile_step.sh <tag> [skip]  -- ncu launch list + one --set full capture of every kernel of one step
# skip = k_* launches before the captured step (4 gt renders x 5 kernels + (warmup 3 + 1 timed) steps x NK kernels)
TAG=${1:-r1}; NK=${NK:-11}; SKIP=${2:-$((20 + 4 * NK))}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 160 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 6 --warmup 3 --profile > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_ -s $SKIP -c $NK -f -o gpurun_out/${TAG}_full \
    python bench.py --steps 6 --warmup 3 --profile > gpurun_out/${TAG}_full.log 2>&1
ls -la gpurun_out | tail -5
