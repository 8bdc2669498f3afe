# the round's records on one GPU box: bench (default arguments), the reference arm, the ncu evidence of the same command, the other
# BASELINE configs and the per-warp timeline.  usage: bash tools/probes/round_records.sh <tag>
TAG=${1:-r2l}
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null
timeout 900 bash tools/profile_step.sh ${TAG} > /dev/null
timeout 600 python tools/config_sweep.py 200 > gpurun_out/${TAG}_config_sweep.log 2>&1; cp gpurun_out/config_sweep.json gpurun_out/${TAG}_config_sweep.json
timeout 300 python tools/probes/timeline.py 2 2>&1 | grep -v Warning > gpurun_out/${TAG}_timeline.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/${TAG}_pytest_gpu.log; tail -2 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
