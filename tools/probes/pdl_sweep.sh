for m in 31 15 31 15 7; do echo -n "MM_PDL_LATE=$m  "; MM_PDL_LATE=$m python bench.py --steps 2000 --warmup 30 --profile 2>/dev/null | tail -1; done
