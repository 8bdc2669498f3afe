#!/bin/bash
# usage: tools/sweep_env.sh  -- prints ms_per_step + per-kernel times for tile-shape settings
for rows in 1 2 4 8; do for rs in 1 0; do
  echo -n "MM_ST_ROWS=$rows MM_REC_SMEM=$rs : "
  MM_ST_ROWS=$rows MM_REC_SMEM=$rs timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['kernel_ms'].items()})"
done; done
