#!/bin/bash
# usage (GPU box): tools/sweep_build.sh "<-DFOO=1>" "<-DFOO=2>" ...   -- rebuild with each flag set, run the quick bench
for flags in "$@"; do
  MM_NVCC_EXTRA="$flags" python -c "import __graft_entry__ as g; g.build_cuda(force=True)" || exit 1
  echo "== $flags"; python tools/quick_bench.py 300 2>&1 | tail -1
done
