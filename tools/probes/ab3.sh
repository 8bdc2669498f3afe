run() { env $1 timeout 300 python tools/probes/lib_ab.py 800 2>&1 | grep ms_per_step | cut -c1-200; }
{ run "MM_X=0"
  for v in sb8_8 sb8_2 sb16_2 sb4_2 sb6_1; do run "MM_LIB=libmagicmirror_var_$v.so"; done
  for l in 14 13 7 31 0 23 11; do run "MM_PDL_LATE=$l"; done
  run "MM_X=0"
} | tee gpurun_out/ab3.txt
