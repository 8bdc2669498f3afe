timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/ab_pytest.log
tail -3 gpurun_out/ab_pytest.log
run() { env $1 timeout 300 python tools/probes/lib_ab.py 2>&1 | grep ms_per_step; }
{ run "MM_LIB=libmagicmirror_head.so"
  run "MM_X=0"
} | tee gpurun_out/ab3.txt
