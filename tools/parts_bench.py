"""Fused-step time for MM_PARTS = 1..4 (sub-batches on concurrent streams), one process.  usage: python tools/parts_bench.py [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as g
import bench
mm = g.load_package()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dev = "cuda:0"
dr, sets = bench.build_workload(mm, dev, 0)
fr = bench.FusedRunner(mm, dr, sets, dev)
L = mm.lib()
for rep in range(2):
    for n in (1, 2, 3, 4):
        L.mm_ctx_set_parts(fr.h.handle, n)
        for i in range(20): fr.step(i)
        ms = bench.timed(torch, 1, fr.step, steps) / steps
        print("parts %d: ms_per_step %.4f  img/s %.0f" % (n, ms, 48 / ms * 1e3), flush=True)

# same steps replayed as CUDA graphs (one graph per rotating input set): removes the host's launch cost from the picture
for n in (1, 2, 3, 4):
    L.mm_ctx_set_parts(fr.h.handle, n)
    for i in range(8): fr.step(i)
    torch.cuda.synchronize()
    graphs = []
    cap = torch.cuda.Stream()
    for i in range(len(fr.sets)):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=cap):
            fr.stream = torch.cuda.current_stream()
            fr.step(i)
        graphs.append(gr)
    fr.stream = torch.cuda.current_stream()
    for i in range(20): graphs[i % len(graphs)].replay()
    ms = bench.timed(torch, 1, lambda i: graphs[i % len(graphs)].replay(), steps) / steps
    print("graph, parts %d: ms_per_step %.4f  img/s %.0f" % (n, ms, 48 / ms * 1e3), flush=True)
