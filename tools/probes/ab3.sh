run() { env $1 timeout 300 python tools/probes/lib_ab.py 1500 2>&1 | grep -E "ms_per_step|rror" | cut -c1-230; }
{ run "MM_LIB=libmagicmirror_var_base.so"; run "MM_LIB=libmagicmirror_var_clrfirst.so"; run "MM_X=0"; run "MM_LIB=libmagicmirror_var_base.so"; run "MM_LIB=libmagicmirror_var_clrfirst.so"; run "MM_X=0"; } | tee gpurun_out/ab3.txt
timeout 300 python tools/quick_bench.py 50 2>&1 | grep "^parity"
