run() { env $1 timeout 300 python tools/probes/lib_ab.py 2>&1 | grep ms_per_step; }
{ run "MM_LIB=libmagicmirror_head.so"
  run "MM_X=0"
  run "MM_X=0"
} | tee gpurun_out/ab3.txt
timeout 300 python tools/probes/timeline.py 2 2>&1 | grep -v Warning > gpurun_out/timeline.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
