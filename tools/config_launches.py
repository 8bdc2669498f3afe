"""A few fused steps of one BASELINE config, for an ncu launch list; prints the overflow / pair counters of the last step.
usage (GPU box): ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file out.csv \
                 python tools/config_launches.py cfg-4|cfg-5|cfg-5s2|cfg-2"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as g
import bench
import parity_utils as pu
mm = g.load_package()
CFG = {
    "cfg-2": dict(mesh="ellipsoid", B=48, size=128, ratio=1, ell=1, Ht=256, Wt=128, kw={}),
    "cfg-4": dict(mesh="smpl_uv_642", B=48, size=128, ratio=2, ell=2, Ht=512, Wt=128,
                  kw=dict(elev_range=(-15.0, 15.0), dist_range=(2.0, 6.0), bias_range=0.5)),
    "cfg-5": dict(mesh="sphere", B=64, size=256, ratio=1, ell=1, Ht=512, Wt=512, kw={}),
    "cfg-5s2": dict(mesh="sphere2", B=64, size=256, ratio=1, ell=1, Ht=512, Wt=512, kw={}),
}
c = CFG[sys.argv[1]]
dev = "cuda:0"
dr = mm.DiffRender(pu.get_mesh(mm, c["mesh"]), c["size"], ratio=c["ratio"], init_ellipsoid=c["ell"], image_weight=1.0)
H, W = dr.height, dr.image_size
sets = [(pu.make_attributes(dr.vertices_init, c["B"], H, W, 900, Ht=c["Ht"], Wt=c["Wt"], **c["kw"]),
         pu.make_attributes(dr.vertices_init, c["B"], H, W, 950, Ht=c["Ht"], Wt=c["Wt"], **c["kw"]))]
fr = bench.FusedRunner(mm, dr, sets, dev)
for i in range(3): fr.step(i)
torch.cuda.synchronize()
al = lambda x: (x + 255) // 256 * 256
B, F = c["B"], dr.num_faces
off = al(B * F * 12 * 4); off += al(B * H * W * 8) * 2; off += al(B * H * ((W + 31) // 32) * 4)
cnt = fr.sets[0]['out']['ws'][off:off + 16].view(torch.int32).cpu().tolist()
print("%s: overflow pixels %d of %d (%.2f %%), candidate pairs %d (%.2f per pixel)" %
      (sys.argv[1], cnt[0], B * H * W, 100.0 * cnt[0] / (B * H * W), cnt[1], cnt[1] / (B * H * W)), flush=True)
