"""Per-warp timeline of one fused step (GPU box; needs the MM_PROF build: tools/probes/build_prof.sh).
Every warp of the six kernels stamps %globaltimer at entry (0), once its dependencies are satisfied (1, where the kernel has
such a point) and at exit (2).  Prints, per kernel: when its warps started / became ready / ended relative to the step's first
stamp, how long they ran, and how many warps of each kernel were alive at every 5 us tick.
usage: [MM_FLOW=0|1] [MM_MIXED=0|1] python tools/probes/timeline.py [reps]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import __graft_entry__ as g
import bench
mm = g.load_package()
from magic_mirror_b200 import _lib
_lib.LIB_PATH = os.path.join(g.PKG_DIR, "libmagicmirror_prof.so")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = "cuda:0"
dr, sets = bench.build_workload(mm, dev, 0)
fr = bench.FusedRunner(mm, dr, sets, dev)
L = mm.lib()
L.mm_debug_profile.restype = ctypes.c_int
L.mm_debug_profile.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
NK, NW = 6, 16384
prof = torch.zeros(NK * NW * 4, dtype=torch.int64, device=dev)
for i in range(20): fr.step(i)
torch.cuda.synchronize()
ms = bench.timed(torch, 1, fr.step, 500) / 500
print("flow=%s mixed=%s  ms_per_step (no stamps) %.4f" % (os.environ.get("MM_FLOW", "1"), os.environ.get("MM_MIXED", "1"), ms))
L.mm_debug_profile(fr.h.handle, ctypes.c_void_p(prof.data_ptr()))
def ws_counters(ws, h, B):
    """{truncated pixels, candidate pairs, -, -} and the shading schedule's class sizes of the last step"""
    L.mm_debug_workspace_offset.restype = ctypes.c_size_t
    L.mm_debug_workspace_offset.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p]
    rd = lambda name, n: ws[L.mm_debug_workspace_offset(h, B, name):][:4 * n].view(torch.int32).cpu().numpy()
    return np.concatenate([rd(b"ovf_count", 4), rd(b"sched_n", 5)])
c = ws_counters(fr.sets[0]['out']['ws'], fr.h.handle, fr.B)
print("set 0: truncated pixels %d, candidate pairs %d, strips per class (0..4 rounds) %s" % (c[0], c[1], c[4:9].tolist()))
names = ["vertex_fwd", "hard", "soft_fwd", "shade", "soft_bwd", "vertex_bwd"]
def q(a, ps=(0, 50, 90, 100)): return " ".join("%6.1f" % np.percentile(a, p) for p in ps)
for r in range(reps):
    for i in range(4): fr.step(r * 5 + i)          # a few back-to-back steps: the stamped one runs as inside a loop
    torch.cuda.synchronize()
    prof.zero_(); torch.cuda.synchronize()
    fr.step(r * 5 + 4); fr.step(r * 5 + 5)          # (the second step overwrites the first one's stamps: steady state)
    torch.cuda.synchronize()
    t = prof.cpu().numpy().reshape(NK, NW, 4).astype(np.float64)
    live = t[:, :, 0] > 0
    t0 = t[:, :, 0][live].min()
    print("rep %d: step span %.1f us  (percentiles 0/50/90/100, us from the step's first stamp)" % (r, (t[:, :, 2].max() - t0) / 1e3))
    ticks = np.arange(0, (t[:, :, 2].max() - t0) / 1e3 + 5, 5.0)
    for k in range(NK):
        m = live[k] & (t[k, :, 2] > 0)
        if not m.any(): continue
        s, e = (t[k, m, 0] - t0) / 1e3, (t[k, m, 2] - t0) / 1e3
        line = "  %-10s n=%5d  start %s | end %s | run %s" % (names[k], m.sum(), q(s), q(e), q(e - s))
        rd = t[k, m, 1]
        if (rd > 0).any():
            w = (rd[rd > 0] - t0) / 1e3 - s[rd > 0]
            line += " | wait %s" % q(w)
        print(line)
        if k == 3:                                     # shading: what would a work-ordered dispatch buy?  (greedy list scheduling
            import heapq                               # of the measured per-CTA run times on the kernel's 5 x SMs slots)
            st = t[k, :, 0].reshape(-1, 4); en = t[k, :, 2].reshape(-1, 4)
            ok = (st > 0).all(axis=1) & (en > 0).all(axis=1)
            run_cta = ((en.max(axis=1) - st.min(axis=1)) / 1e3)[ok]
            def makespan(order, slots=148 * 5):
                h = [0.0] * slots
                for r in order: heapq.heapreplace(h, h[0] + r)
                return max(h)
            two = np.concatenate([run_cta[run_cta > 5.0], run_cta[run_cta <= 5.0]])
            print("    shade CTAs %d: sum of runs / slots %.1f us; list-scheduled makespan: grid order %.1f, longest first %.1f, two classes (> 5 us first) %.1f, shortest first %.1f" % (
                len(run_cta), run_cta.sum() / (148 * 5), makespan(run_cta), makespan(np.sort(run_cta)[::-1]), makespan(two), makespan(np.sort(run_cta))))
        if k == 4:                                     # soft_bwd: the overflow-role CTAs (the first 8 * SMs CTAs) separately
            nl = 148 * 16 * 4
            mm_ = m.copy(); mm_[148 * 8 * 4:] = False
            if mm_.any():
                s2, e2 = (t[k, mm_, 0] - t0) / 1e3, (t[k, mm_, 2] - t0) / 1e3
                r2 = (t[k, mm_, 1] - t0) / 1e3
                print("    overflow-role CTAs n=%d start %s | listed %s | end %s | run %s ; busy (run > 2 us): %d" % (mm_.sum(), q(s2), q(r2), q(e2), q(e2 - s2), int(((e2 - s2) > 2.0).sum())))
        if k == 1:                                     # hard pass: run time against the warp's (face, pixel) pairs
            pairs = t[k, m, 3]
            run = e - s
            print("    start percentiles 50/90/95/99/100: %s" % q(s, (50, 90, 95, 99, 100)))
            for lo, hi in ((0, 3), (3, 5), (5, 7), (7, 9), (9, 99)):
                sel = (run >= lo) & (run < hi)
                if sel.any():
                    print("    run %2d-%2d us: %5d warps, pairs p50/max %5d/%5d, start p50 %.1f" % (lo, hi, sel.sum(), np.median(pairs[sel]), pairs[sel].max(), np.median(s[sel])))
            print("    all warps: pairs sum %d, p50 %d, p90 %d, max %d; corr(run, pairs) %.2f" % (pairs.sum(), np.median(pairs), np.percentile(pairs, 90), pairs.max(), np.corrcoef(run, pairs)[0, 1]))
        if k == 2:                                     # soft forward: run time against the warp's row-walk length and candidates
            extra = t[k, m, 3].astype(np.int64)
            mseg, ncand = extra >> 32, extra & 0xffffffff
            run = e - s
            for lo, hi in ((0, 4), (4, 6), (6, 8), (8, 10), (10, 13), (13, 99)):
                sel = (run >= lo) & (run < hi)
                if sel.any():
                    print("    run %2d-%2d us: %5d warps, iterations p50/max %3d/%3d, candidates p50/max %4d/%4d, start p50 %.1f" % (
                        lo, hi, sel.sum(), np.median(mseg[sel]), mseg[sel].max(), np.median(ncand[sel]), ncand[sel].max(), np.median(s[sel])))
            print("    all warps: iterations sum %d, candidates sum %d; corr(run, iterations) %.2f, corr(run, candidates) %.2f" % (
                mseg.sum(), ncand.sum(), np.corrcoef(run, mseg)[0, 1], np.corrcoef(run, ncand)[0, 1]))
        alive = [(int(((s <= x) & (e > x)).sum())) for x in ticks]
        print("             alive@5us: " + " ".join("%5d" % a for a in alive))
