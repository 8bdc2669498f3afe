#!/bin/bash
# usage: multi_gpu_bench.sh N TAG  (on the GPU box) -- the default bench under torchrun at N ranks
N=$1; TAG=${2:-r2j}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 1000 --warmup 20 2>gpurun_out/bench$N.err | tail -1 > gpurun_out/${TAG}_bench_n$N.json
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_n$N.json")); print("bench", d["n_gpus"], d["value"], d["ms_per_step"], d["value_api"], d["e2e"]["value"], d["e2e"]["h2d_gbs_per_rank"], d["e2e"]["pcie_h2d_ceiling_gbs"])
PY
