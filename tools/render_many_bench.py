"""SURVEY 8(f)-2 timing: three separate render + backward calls of B=48 (trainer.py:276,345,347) vs render_many of the three
sets (one pass over 144 images).  usage (GPU box): python tools/render_many_bench.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as g
import parity_utils as pu
mm = g.load_package()
dev = "cuda:0"
dr = mm.DiffRender(pu.get_mesh(mm, "ellipsoid"), 128, image_weight=1.0)
keys = ('vertices', 'azimuths', 'elevations', 'distances', 'biases', 'textures', 'lights', 'bg')
sets = [{k: (v.requires_grad_(k in keys) if torch.is_tensor(v) else v)
         for k, v in pu.to_device(pu.make_attributes(dr.vertices_init, 48, 128, 128, 70 + i), dev).items()} for i in range(3)]
w = torch.randn(48, 4, 128, 128, device=dev)

def separate():
    outs = [dr.render(no_mask=True, **dict(A)) for A in sets]
    sum((img * w).sum() for img, _ in outs).backward()

def many():
    outs = dr.render_many([dict(A) for A in sets], no_mask=True)
    sum((img * w).sum() for img, _ in outs).backward()

for name, fn in (("3 x render(B=48) + backward", separate), ("render_many(3 x 48) + backward", many)):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-32s %.3f ms" % (name, e0.elapsed_time(e1) / 50), flush=True)
