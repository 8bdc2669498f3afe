"""A/B of library builds / switches on the GPU box: the fused step's time and the event-timed launch groups.
usage: [MM_LIB=libmagicmirror_head.so] [MM_PDL_LATE=..] python tools/probes/lib_ab.py [steps]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as g
import bench
mm = g.load_package()
from magic_mirror_b200 import _lib
if os.environ.get("MM_LIB"): _lib.LIB_PATH = os.path.join(g.PKG_DIR, os.environ["MM_LIB"])
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
dev = "cuda:0"
dr, sets = bench.build_workload(mm, dev, 0)
fr = bench.FusedRunner(mm, dr, sets, dev)
for i in range(20): fr.step(i)
ms = min(bench.timed(torch, 1, fr.step, steps) / steps for _ in range(3))
L = mm.lib(); h = fr.h.handle
L.mm_ctx_set_timing(h, 1)
acc = [0.0] * 7; buf = (ctypes.c_float * 8)()
n = 200
for i in range(n):
    fr.step(i); L.mm_ctx_get_timing(h, buf, 8)
    for j in range(7): acc[j] += buf[j]
L.mm_ctx_set_timing(h, 0)
tag = " ".join("%s=%s" % (k, os.environ[k]) for k in ("MM_LIB", "MM_PDL_LATE") if k in os.environ)
print("%-60s ms_per_step %.4f  groups_us %s" % (tag or "(defaults)", ms, {k: round(1e3 * a / n, 1) for k, a in zip(bench.KERNELS, acc)}), flush=True)
