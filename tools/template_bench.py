"""SURVEY 8(f)-3 timing: DiffRender.template_features vs the reference's torch ops (model_res.py:317-325) at the trainer's
shape (B=48, C=288, 8x4 feature map, V=642), forward + backward.  usage (GPU box): python tools/template_bench.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
import __graft_entry__ as g
import parity_utils as pu
mm = g.load_package()
dev = "cuda:0"
B, C, h, w = 48, 288, 8, 4
dr = mm.DiffRender(pu.get_mesh(mm, "sphere"), 128)
V = dr.num_vertices
x = torch.randn(B, C, h, w, device=dev, requires_grad=True)
template = dr.vertices_init[None].to(dev)
lpl = dr.vertices_laplacian_matrix.to(dev)
wl = torch.randn(B, C, V, 1, device=dev); wn = torch.randn(B, C, V, 1, device=dev)

def ours():
    l, n = dr.template_features(x, template)
    torch.autograd.grad([l, n], [x], [wl, wn])

def ref():
    uv = template.repeat(B, 1, 1).view(B, V, 1, 3)[:, :, :, 0:2].detach()
    l = F.grid_sample(x, uv, mode='bilinear', align_corners=True, padding_mode="zeros")
    n = torch.mm(l.view(-1, V), lpl).view(B, -1, V, 1)
    torch.autograd.grad([l, n], [x], [wl, wn])

def ours_fwd():
    with torch.no_grad():
        dr.template_features(x, template)

def ref_fwd():
    with torch.no_grad():
        uv = template.repeat(B, 1, 1).view(B, V, 1, 3)[:, :, :, 0:2]
        l = F.grid_sample(x, uv, mode='bilinear', align_corners=True, padding_mode="zeros")
        torch.mm(l.view(-1, V), lpl)

for name, fn, tf32 in (("ours", ours, False), ("torch fp32", ref, False), ("torch tf32 matmul", ref, True),
                       ("ours, forward only", ours_fwd, False), ("torch fp32, forward only", ref_fwd, False)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print("%-18s fwd+bwd %.3f ms  (outputs 2 x %.1f MB)" % (name, ms, B * C * V * 4 / 1e6), flush=True)
