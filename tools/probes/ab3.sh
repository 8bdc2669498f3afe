run() { env $1 timeout 300 python tools/probes/lib_ab.py 1500 2>&1 | grep ms_per_step | cut -c1-230; }
{ run "MM_X=0"; run "MM_X=0"; } | tee gpurun_out/ab3.txt
timeout 300 python tools/probes/timeline.py 1 2>&1 | grep -v Warning > gpurun_out/timeline.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
