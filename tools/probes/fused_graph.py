"""The fused C-ABI step launched from the stream (what bench.py's `value` times) against the same call captured once per input
set into a CUDA graph and replayed: does graph replay shorten the ~1 us between the step's six dependent kernels?
usage (GPU box): python tools/probes/fused_graph.py [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as g
import bench
mm = g.load_package()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
dev = "cuda:0"
dr, sets = bench.build_workload(mm, dev, 0)
fr = bench.FusedRunner(mm, dr, sets, dev)
for i in range(20): fr.step(i)
torch.cuda.synchronize()
ms_stream = min(bench.timed(torch, 1, fr.step, steps) / steps for _ in range(3))
graphs = []
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    fr.stream = side
    for i in range(len(fr.sets)): fr.step(i)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
for i in range(len(fr.sets)):
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fr.stream = torch.cuda.current_stream()
        fr.step(i)
    graphs.append(gr)
fr.stream = torch.cuda.current_stream()
ref = fr.step(0); torch.cuda.synchronize()
loss_stream = ref['loss'].clone(); gv = ref['g_v'].clone()
graphs[0].replay(); torch.cuda.synchronize()
o = fr.sets[0]['out']
# (the loss is a fixed-point sum: bit-identical; the vertex gradient goes through float reductions whose order is free)
same = "loss %s, g_vertices max rel diff %.1e" % ("bit-identical" if bool((o['loss'] == loss_stream).all()) else "DIFFERS",
                                                   float((o['g_v'] - gv).abs().max() / gv.abs().max()))
replay = lambda i: graphs[i % len(graphs)].replay()      # noqa: E731
for i in range(20): replay(i)
ms_graph = min(bench.timed(torch, 1, replay, steps) / steps for _ in range(3))
print("fused step: stream launches %.4f ms, graph replay %.4f ms per step; stream vs graph: %s" % (ms_stream, ms_graph, same), flush=True)
