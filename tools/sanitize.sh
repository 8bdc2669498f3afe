#!/bin/bash
# usage (GPU box): tools/sanitize.sh  -- compute-sanitizer memcheck + racecheck over two small fused/unfused parity cases
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import __graft_entry__ as g
mm = g.load_package()
import parity_utils as pu
for c in (dict(mesh="icosphere", B=2, image_size=32, no_mask=True, contour=0.1, seed=3),
          dict(mesh="sphere", B=2, image_size=22, ratio=1.5, no_mask=True, contour=0.1, seed=19),
          dict(mesh="sphere", B=1, image_size=48, no_mask=True, contour=0.0, seed=13, dist_range=(6.5, 7.0))):
    r = pu.run_parity_case(mm, **c)
    print("case ok", c["mesh"], r["face_idx_mismatch_staged"], r["loss_rel_err"], flush=True)
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -c 'case ok' gpurun_out/sanitize_$tool.log) cases, $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
done
