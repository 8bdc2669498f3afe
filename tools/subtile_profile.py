"""Per-sub-tile cycle profile of the fused step (GPU box).  Prints the heaviest sub-tiles."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as g
import bench
mm = g.load_package()
dev = "cuda:0"
dr, sets = bench.build_workload(mm, dev, 0, nsets=1)
fr = bench.FusedRunner(mm, dr, sets, dev)
for i in range(3): fr.step(i)
nst = ((dr.image_size + 7) // 8) * ((dr.height + 3) // 4)
buf = torch.zeros(48, nst, 8, dtype=torch.int64, device=dev)
mm.lib().mm_debug_set_profile_buffer(fr.h.handle, ctypes.c_void_p(buf.data_ptr()))
fr.step(0); torch.cuda.synchronize()
mm.lib().mm_debug_set_profile_buffer(fr.h.handle, ctypes.c_void_p(0))
p = buf.cpu()
for name, col in (("fwd", 0), ("bwd", 1)):
    cyc = p[..., col].flatten().float()
    print(name, "sum Mcyc %.1f  mean %.0f  median %.0f  p99 %.0f  max %.0f" % (cyc.sum() / 1e6, cyc.mean(), cyc.median(), cyc.quantile(0.99), cyc.max()))
    top = torch.argsort(cyc, descending=True)[:12]
    for t in top:
        b, st = divmod(int(t), nst)
        print("   b=%2d st=%3d cycles=%8d  S=%4d H=%4d  dist=%.2f" % (b, st, int(cyc[t]), int(p[b, st, 2]), int(p[b, st, 3]), float(sets[0][0]['distances'][b])))
    if col == 0:
        for nm, cc in (("hard", 4), ("soft-mark(ph1+2)", 5), ("soft-pairs(ph3)", 6)):
            print("   phase %-18s total Mcyc %.1f" % (nm, p[..., cc].sum().item() / 1e6))
        print("   pairs total %d  (%.2f per pixel)" % (p[..., 7].sum().item(), p[..., 7].sum().item() / (48 * 128 * 128)))
    S = p[..., 2].flatten().float()
    for lo, hi in ((0, 0), (1, 31), (32, 127), (128, 511), (512, 99999)):
        m = (S >= lo) & (S <= hi)
        if m.any(): print("   S in [%d,%d]: n=%d  mean cycles %.0f  total Mcyc %.1f" % (lo, hi, int(m.sum()), cyc[m].mean(), cyc[m].sum() / 1e6))
