for m in 0 8 16 24 1 2 3 31 27; do echo -n "MM_PDL_LATE=$m  "; MM_PDL_LATE=$m python bench.py --steps 1500 --warmup 30 --profile 2>/dev/null | tail -1; done
