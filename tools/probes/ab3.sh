run() { env $1 timeout 300 python tools/probes/lib_ab.py 1000 2>&1 | grep -E "ms_per_step|rror" | sed -E 's/ +ms_per_step/ ms/; s/groups_us.*vertex_bwd/ vbwd/' | cut -c1-120; }
{ run "MM_X=0"
  for v in vb4_640 vb4_768 vb4_1024 vb2_1024 vb1_1024 vb2_768; do run "MM_LIB=libmagicmirror_var_$v.so"; done
  run "MM_X=0"; } | tee gpurun_out/ab3.txt
