"""Iteration tool (GPU box): two parity cases + per-kernel timing of the fused step, one process.
usage: python tools/quick_bench.py [steps]"""
import ctypes, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
t00 = time.time()
import torch
print('import torch %.1fs' % (time.time() - t00), flush=True)
import __graft_entry__ as g
import bench
mm = g.load_package()
import parity_utils as pu
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for c in (dict(mesh="ellipsoid", B=4, image_size=128, no_mask=True, contour=0.1, seed=7),
          dict(mesh="sphere", B=2, image_size=64, no_mask=True, contour=0.0, seed=13, dist_range=(6.5, 7.0))):
    r = pu.run_parity_case(mm, **c)
    keys = ["face_idx_mismatch_staged", "face_idx_mismatch_e2e", "soft_staged_max_abs_err", "rgba_staged_max_abs_err",
            "loss_rel_err", "grad_vertices_rel_err", "grad_textures_rel_err", "grad_azimuths_rel_err", "grad_lights_rel_err"]
    print("parity", c["mesh"], {k: (r[k] if isinstance(r[k], int) else float("%.3g" % r[k])) for k in keys}, flush=True)
print('parity done %.1fs' % (time.time() - t00), flush=True)
dev = "cuda:0"
dr, sets = bench.build_workload(mm, dev, 0)
fr = bench.FusedRunner(mm, dr, sets, dev)
print('workload built %.1fs' % (time.time() - t00), flush=True)
for i in range(10): fr.step(i)
ms = bench.timed(torch, 1, fr.step, steps) / steps
L = mm.lib(); h = fr.h.handle
L.mm_ctx_set_timing(h, 1)
acc = [0.0] * 7; buf = (ctypes.c_float * 8)()
n = min(steps, 200)
for i in range(n):
    fr.step(i); L.mm_ctx_get_timing(h, buf, 8)
    for j in range(7): acc[j] += buf[j]
L.mm_ctx_set_timing(h, 0)
print("ms_per_step %.4f  img/s %.0f  kernels_us %s" % (ms, 48 / ms * 1e3, {k: round(1e3 * a / n, 1) for k, a in zip(bench.KERNELS, acc)}), flush=True)

# SURVEY 8(f)-3 variant: mirrored texture (the upper half of the atlas only); informational, not the headline workload
fm = bench.FusedRunner(mm, dr, sets, dev, tex_mirror=True)
for i in range(10): fm.step(i)
msm = bench.timed(torch, 1, fm.step, steps) / steps
print("mirrored-texture variant: ms_per_step %.4f  img/s %.0f" % (msm, 48 / msm * 1e3), flush=True)
