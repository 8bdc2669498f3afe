"""Host-side profile of the eager three-call API (GPU box): where do the ~0.29 ms of Python / autograd time per step go?
usage: python tools/probes/eager_profile.py [steps]  -> cProfile top entries (cumulative and own time)"""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as g
import bench
mm = g.load_package()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = "cuda:0"
dr, sets = bench.build_workload(mm, dev, 0)
fr = bench.FusedRunner(mm, dr, sets, dev)
api = bench.ApiRunner(mm, dr, fr, graph=False)
for i in range(30): api.step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(steps): api.step(i)
t1 = time.perf_counter()            # host time to ISSUE the steps (the GPU runs behind)
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host issue time %.1f us / step; with the GPU drained %.1f us / step" % ((t1 - t0) / steps * 1e6, (t2 - t0) / steps * 1e6))
pr = cProfile.Profile()
pr.enable()
for i in range(steps): api.step(i)
pr.disable()
torch.cuda.synchronize()
for key in ("cumulative", "tottime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).strip_dirs().sort_stats(key).print_stats(22)
    print("\n".join(l for l in s.getvalue().splitlines() if l.strip())[:6000])
