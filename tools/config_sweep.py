"""Fused-step time and whole-step HBM-roofline fraction on the other BASELINE.json configs (informational; bench.py's headline
stays configs[1]).  usage (GPU box): python tools/config_sweep.py [steps]  -> gpurun_out/config_sweep.json"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as g
import bench
import parity_utils as pu
mm = g.load_package()
if os.environ.get("MM_LIB"):          # A/B of another build of the library (tools/probes/lib_ab.py)
    from magic_mirror_b200 import _lib
    _lib.LIB_PATH = os.path.join(g.PKG_DIR, os.environ["MM_LIB"])
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = "cuda:0"
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
CFG = [
    dict(name="cfg-2", mesh="ellipsoid", B=48, size=128, ratio=1, ell=1, Ht=256, Wt=128, kw={}),
    dict(name="cfg-4", mesh="smpl_uv_642", B=48, size=128, ratio=2, ell=2, Ht=512, Wt=128,
         kw=dict(elev_range=(-15.0, 15.0), dist_range=(2.0, 6.0), bias_range=0.5)),
    dict(name="cfg-5", mesh="sphere", B=64, size=256, ratio=1, ell=1, Ht=512, Wt=512, kw={}),
    dict(name="cfg-5/sphere2", mesh="sphere2", B=64, size=256, ratio=1, ell=1, Ht=512, Wt=512, kw={}),
]
rows = []
for c in CFG:
    dr = mm.DiffRender(pu.get_mesh(mm, c["mesh"]), c["size"], ratio=c["ratio"], init_ellipsoid=c["ell"], image_weight=1.0)
    H, W = dr.height, dr.image_size
    nsets = 4 if c["size"] <= 128 else 2           # cfg-5: 2 x 1 GB sets, each far larger than the 126 MB L2
    sets = [(pu.make_attributes(dr.vertices_init, c["B"], H, W, 900 + i, Ht=c["Ht"], Wt=c["Wt"], **c["kw"]),
             pu.make_attributes(dr.vertices_init, c["B"], H, W, 950 + i, Ht=c["Ht"], Wt=c["Wt"], **c["kw"])) for i in range(nsets)]
    fr = bench.FusedRunner(mm, dr, sets, dev)
    for i in range(5): fr.step(i)
    ms = bench.timed(torch, 1, fr.step, steps) / steps
    L = mm.lib(); h = fr.h.handle
    L.mm_ctx_set_timing(h, 1)
    acc = [0.0] * 7; buf = (ctypes.c_float * 8)()
    n = min(steps, 100)
    for i in range(n):
        fr.step(i); L.mm_ctx_get_timing(h, buf, 8)
        for j in range(7): acc[j] += buf[j]
    L.mm_ctx_set_timing(h, 0)
    _, _, nbytes = bench.algorithmic_bytes(c["B"], dr.num_vertices, dr.num_faces, H, W, c["Ht"], c["Wt"])
    ach = nbytes / (ms * 1e-3) / 1e9
    t_sh = acc[2] / n
    nb_sh = nbytes - c["B"] * 4 * (3 * dr.num_vertices * 3 + 3 * dr.num_faces + 14 * 3) - dr.num_faces * 36
    row = {"config": c["name"], "mesh": c["mesh"], "B": c["B"], "H": H, "W": W, "F": dr.num_faces, "tex": [c["Ht"], c["Wt"]],
           "ms_per_step": ms, "images_per_s": c["B"] / ms * 1e3, "algorithmic_MB_per_step": nbytes / 1e6,
           "step_achieved_GBs": ach, "step_frac_of_hbm_peak": ach / peak,
           "shade_fused_ms": t_sh, "shade_fused_frac_of_hbm_peak": nb_sh / (t_sh * 1e-3) / 1e9 / peak,
           "kernel_us": {k: round(1e3 * a / n, 1) for k, a in zip(bench.KERNELS, acc)}}
    rows.append(row)
    print(json.dumps(row), flush=True)
    del fr, sets
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"peak_hbm_GBs": peak, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "config_sweep.json"), "w"), indent=1)
