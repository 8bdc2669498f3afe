"""GPU box: phase breakdown of the band rasteriser at the bench workload (MM_BAND_PROF=1)."""
import ctypes, os, sys
os.environ["MM_BAND_PROF"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import __graft_entry__ as g
import bench
mm = g.load_package()
dr, sets = bench.build_workload(mm, "cuda:0", 0)
fr = bench.FusedRunner(mm, dr, sets, "cuda:0")
for i in range(8):
    fr.step(i)
torch.cuda.synchronize()
L = mm.lib()
nb = L.mm_ctx_get_int(fr.h.handle, b"band_count")
n = nb * fr.B
buf = np.zeros((n, 8), dtype=np.int64)
assert L.mm_debug_band_profile(fr.h.handle, buf.ctypes.data_as(ctypes.c_void_p), n) == 0
d = np.diff(buf[:, :7], axis=1) / 1965.0          # us at 1965 MHz
names = ["vertex", "select", "hard", "soft", "ovf", "out"]
print("bands/img", nb, "CTAs", n)
print("phase      mean    p50    p90    max (us)")
for j, nm in enumerate(names):
    print("%-8s %6.2f %6.2f %6.2f %6.2f" % (nm, d[:, j].mean(), np.median(d[:, j]), np.percentile(d[:, j], 90), d[:, j].max()))
tot = (buf[:, 6] - buf[:, 0]) / 1965.0
print("total    %6.2f %6.2f %6.2f %6.2f" % (tot.mean(), np.median(tot), np.percentile(tot, 90), tot.max()))
span = (buf[:, 6].max() - buf[:, 0].min()) / 1965.0
print("kernel span (first start -> last end) %.1f us; start spread %.1f us" % (span, (buf[:, 0].max() - buf[:, 0].min()) / 1965.0))
per_img = tot.reshape(fr.B, nb).max(1)
dist = sets[(8 - 1) % len(sets)][0]['distances'].numpy()
order = np.argsort(-per_img)[:6]
print("slowest images (max band us, distance, relevant faces of slowest band):")
for i in order:
    print("   img %2d  %6.1f us  dist %.2f  nL %s" % (i, per_img[i], dist[i], buf[i * nb:(i + 1) * nb, 7].tolist()))
