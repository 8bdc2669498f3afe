"""Scratch driver for the first GPU runs: prints the parity dict of a few cases."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as g
g.build()
mm = g.load_package()
import parity_utils as pu
cases = [dict(mesh="icosphere", B=2, image_size=32, no_mask=True, contour=0.1, seed=3),
         dict(mesh="icosphere", B=3, image_size=64, no_mask=False, contour=0.0, seed=5),
         dict(mesh="icosphere", B=4, image_size=128, no_mask=True, contour=0.1, seed=7),
         dict(mesh="icosphere", B=2, image_size=64, ratio=2, init_ellipsoid=2, no_mask=True, contour=0.1, seed=9)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for c in cases:
    t = time.time()
    r = pu.run_parity_case(mm, **c)
    print(json.dumps({"case": c, "sec": round(time.time() - t, 2), **r}), flush=True)
