# bench (default arguments), the reference arm, and the ncu evidence of the same command -- one GPU box
TAG=${1:-r2g}
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null
timeout 900 bash tools/profile_step.sh ${TAG}
