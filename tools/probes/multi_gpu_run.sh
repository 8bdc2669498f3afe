#!/bin/bash
# usage: multi_gpu_run.sh N   (on the GPU box) -- cfg4-ddp trainer step + the default bench under torchrun at N ranks
N=$1
mkdir -p gpurun_out
(nvidia-smi topo -m; lscpu | grep -E "^CPU\(s\)|NUMA|Socket|Model name") > gpurun_out/r2_topo_n$N.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --workload cfg4-ddp --steps 60 --warmup 10 2>gpurun_out/ddp$N.err | tail -1 > gpurun_out/r2_cfg4_ddp_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 500 --warmup 10 2>gpurun_out/bench$N.err | tail -1 > gpurun_out/r2_bench_n$N.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2_cfg4_ddp_n$N.json")); print("cfg4", d["n_gpus"], d["value"], d["ms_per_step"], d["detail"].get("allreduce"), d["detail"].get("step_nosync_ms"))
d=json.load(open("gpurun_out/r2_bench_n$N.json")); print("bench", d["n_gpus"], d["value"], d["ms_per_step"], d["value_api"], d["e2e"]["value"], d["e2e"]["h2d_gbs_per_rank"], d["e2e"]["pcie_h2d_ceiling_gbs"])
PY
cat gpurun_out/r2_topo_n$N.txt | head -16
