"""Diagnostic (GPU box): where do the product's geometry gradients differ from the fp32 / fp64 oracle on one parity case?"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
import parity_utils as pu

case = json.loads(sys.argv[1]) if len(sys.argv) > 1 else dict(mesh="sphere2", B=2, image_size=256, seed=56, tex=(512, 512))
mm = pu.load_mm()
dr = mm.DiffRender(pu.get_mesh(mm, case["mesh"]), case["image_size"], ratio=case.get("ratio", 1),
                   init_ellipsoid=case.get("init_ellipsoid", 1), image_weight=1.0)
H, W, B = dr.height, dr.image_size, case["B"]
Ht, Wt = case.get("tex", (None, None))
kw = dict(Ht=Ht, Wt=Wt, dist_range=tuple(case.get("dist_range", (2.0, 7.0))))
A = pu.make_attributes(dr.vertices_init, B, H, W, case["seed"], **kw)
G = pu.make_attributes(dr.vertices_init, B, H, W, case["seed"] + 1000, **kw)
o32, o64 = pu.oracle_for(dr), pu.oracle_for(dr, torch.float64)
with torch.no_grad():
    gt = o32.render(no_mask=True, **G)[0]
res = {}
for name, orc, dt in (("o32", o32, torch.float32), ("o64", o64, torch.float64)):
    Ag = {k: v.detach().clone().to(dt).requires_grad_(k != 'delta_vertices') for k, v in A.items()}
    rgb, fn, _, fidx = orc.render(no_mask=True, **Ag)
    parts = {}
    for term in ("image", "mask", "all"):
        for v in Ag.values():
            if v.grad is not None:
                v.grad = None
        pred_img, pred_mask = rgb[:, :3], rgb[:, 3]
        gtt = gt.to(dt)
        gm = gtt[:, 3:4]
        l_img = ((pred_img * gm + (1 - gm)) - (gtt[:, :3] * gm + (1 - gm))).abs().mean()
        import kaolin_shim as kal
        l_mask = kal.mask_iou(pred_mask, gtt[:, 3])
        loss = {"image": l_img, "mask": l_mask, "all": l_img + l_mask}[term]
        loss.backward(retain_graph=True)
        parts[term] = {k: Ag[k].grad.detach().clone().double() for k in ('vertices', 'elevations', 'distances')}
    res[name] = (parts, fidx, rgb.detach())
# product: image-only and mask-only through g_rgba
if not torch.cuda.is_available():
    print("no GPU: oracle part ran"); sys.exit(0)
dev = "cuda:0"
prod = {}
for term in ("image", "mask", "all"):
    Ac = pu.to_device({k: v.detach() for k, v in A.items()}, dev, requires_grad=True)
    rgb, _ = dr.render(no_mask=True, **Ac)
    gtd = gt.to(dev)
    gm = gtd[:, 3:4]
    l_img = ((rgb[:, :3] * gm + (1 - gm)) - (gtd[:, :3] * gm + (1 - gm))).abs().mean()
    m = rgb[:, 3]
    g = gtd[:, 3]
    l_mask = 1 - ((m * g).flatten(1).sum(1) / ((m + g - m * g).flatten(1).sum(1) + 1e-10)).mean()
    {"image": l_img, "mask": l_mask, "all": l_img + l_mask}[term].backward()
    prod[term] = {k: Ac[k].grad.detach().cpu().double() for k in ('vertices', 'elevations', 'distances')}
for term in ("image", "mask", "all"):
    for k in ('vertices', 'elevations', 'distances'):
        c, a32, a64 = prod[term][k], res["o32"][0][term][k], res["o64"][0][term][k]
        sc = a64.abs().max().item()
        print("%-6s %-11s scale %.3e  |c-o32| %.3e  |c-o64| %.3e  |o32-o64| %.3e" % (
            term, k, sc, (c - a32).abs().max().item() / sc, (c - a64).abs().max().item() / sc, (a32 - a64).abs().max().item() / sc))
# worst vertices for the 'all' term
c, a64, a32 = prod["all"]["vertices"], res["o64"][0]["all"]["vertices"], res["o32"][0]["all"]["vertices"]
err = (c - a64).abs().sum(-1)
top = err.flatten().topk(6).indices
V = c.shape[1]
faces = dr.faces
for i in top.tolist():
    b, v = i // V, i % V
    print("b %d v %d  prod %s  o32 %s  o64 %s" % (b, v, c[b, v].tolist(), a32[b, v].tolist(), a64[b, v].tolist()))
    adj = (faces == v).any(1).nonzero().flatten().tolist()
    fidx = res["o64"][1]
    print("   adjacent faces", adj, "winner pixels", [int((fidx[b] == f).sum()) for f in adj])
print("elev", prod["all"]["elevations"].tolist(), res["o32"][0]["all"]["elevations"].tolist(), res["o64"][0]["all"]["elevations"].tolist())
