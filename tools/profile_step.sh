#!/bin/bash
# usage (GPU box): tools/profile_step.sh <tag>
#   1. ncu launch list (gpu__time_duration.sum) of bench.py --profile: fused steps, then 4 eager steps of the API path
#   2. ncu --set full of ONE fused step (6 kernels) and ONE API step (8 kernels), raw pages exported as CSV
# kernels before the captured fused step: 4 GT renders x 4 + (3 warm-up + 1 timed) x 6 = 40; the API steps follow the 9 fused steps
TAG=${1:-r2}
mkdir -p gpurun_out
MM_PROFILE_API=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 6 --warmup 3 --profile > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_ -s 40 -c 6 -f -o gpurun_out/${TAG}_fused_full \
    python bench.py --steps 6 --warmup 3 --profile > gpurun_out/${TAG}_fused_full.log 2>&1
MM_PROFILE_API=1 ncu --set full --clock-control none --import-source on -k regex:^k_ -s $((16 + 9 * 6 + 8)) -c 8 -f -o gpurun_out/${TAG}_api_full \
    python bench.py --steps 6 --warmup 3 --profile > gpurun_out/${TAG}_api_full.log 2>&1
ncu -i gpurun_out/${TAG}_fused_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_fused_raw.csv
ncu -i gpurun_out/${TAG}_api_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_api_raw.csv
ls -la gpurun_out | grep ${TAG}
