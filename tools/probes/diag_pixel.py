"""One-pixel upstream gradient through product and oracle (sphere2, 256^2, B=2, seed 56, pixel (95,143) of image 0)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
import parity_utils as pu
mm = pu.load_mm()
dr = mm.DiffRender(pu.get_mesh(mm, "sphere2"), 256, image_weight=1.0)
A = pu.make_attributes(dr.vertices_init, 2, 256, 256, 56, Ht=512, Wt=512)
o32 = pu.oracle_for(dr)
for (iy, ix) in ((95, 143), (96, 143), (95, 142)):
    up = torch.zeros(2, 4, 256, 256)
    up[0, :3, iy, ix] = torch.tensor([1.0, 0.5, -0.7])
    Ag = {k: v.detach().clone().requires_grad_(k != 'delta_vertices') for k, v in A.items()}
    rgb, fn, _, fidx = o32.render(no_mask=True, **Ag)
    rgb.backward(up)
    Ac = pu.to_device({k: v.detach() for k, v in A.items()}, "cuda:0", requires_grad=True)
    Ac['_want_face_idx'] = True
    rc, out = dr.render(no_mask=True, **Ac)
    rc.backward(up.cuda())
    f = int(fidx[0, iy, ix])
    vs = dr.faces[f].tolist() if f >= 0 else []
    print("pixel", iy, ix, "face", f, "cuda face", int(out['face_idx'][0, iy, ix]), "rgb", rgb[0, :3, iy, ix].tolist(), rc[0, :3, iy, ix].tolist())
    for v in vs:
        print("   v", v, "oracle", Ag['vertices'].grad[0, v].tolist(), "cuda", Ac['vertices'].grad[0, v].cpu().tolist())
    gt_o, gt_c = Ag['textures'].grad, Ac['textures'].grad.cpu()
    print("   g_tex nonzero oracle", int((gt_o != 0).sum()), "cuda", int((gt_c != 0).sum()), "max diff", float((gt_o - gt_c).abs().max()),
          "g_vert max diff", float((Ag['vertices'].grad - Ac['vertices'].grad.cpu()).abs().max()), "scale", float(Ag['vertices'].grad.abs().max()))
