"""Shared helpers for the parity tests, smoke() and bench.py's checker leg.

`make_attributes` draws the seeded synthetic inputs of SURVEY.md section 8(d);
`run_parity_case` runs the CUDA product (through the Python DiffRender -> C ABI)
and the CPU oracle (oracle/ref_pipeline.py) on the same inputs and returns error
figures.  The oracle is only ever the checker here.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "oracle") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))


def load_mm():
    import __graft_entry__ as g
    return g.load_package()


def get_mesh(mm, name):
    """'icosphere' / 'icosphere2' (procedural) or a golden template name (tests/golden/templates/<name>.npz)."""
    if name == "icosphere":
        return mm.icosphere(3)
    if name == "icosphere2":
        return mm.icosphere(4)
    path = os.path.join(GOLDEN, "templates", name + ".npz")
    z = np.load(path)
    return mm.TemplateMesh(torch.from_numpy(z["vertices"]), torch.from_numpy(z["faces"]).long(),
                           torch.from_numpy(z["uvs"]), torch.from_numpy(z["face_uvs_idx"]).long())


def make_attributes(vertices_init, B, H, W, seed, Ht=None, Wt=None, elev_range=(0.0, 30.0), dist_range=(2.0, 7.0),
                    bias_range=0.3, deform=0.05):
    """SURVEY 8(d) cfg-2 recipe (train.py:123-127 ranges), CPU tensors, seeded."""
    g = torch.Generator().manual_seed(seed)
    V = vertices_init.shape[0]
    Ht = 2 * H if Ht is None else Ht
    Wt = W if Wt is None else Wt
    u = lambda *s: torch.rand(*s, generator=g)          # noqa: E731
    n = lambda *s: torch.randn(*s, generator=g)         # noqa: E731
    delta = deform * torch.tanh(n(B, V, 3))
    delta = delta - delta.mean(dim=1, keepdim=True)
    lights = torch.tensor([3.0] + [0.0] * 8) + torch.tensor([0.5] + [0.1] * 8) * torch.tanh(n(B, 9))
    A = {
        'azimuths': u(B) * 360.0 - 180.0,
        'elevations': elev_range[0] + u(B) * (elev_range[1] - elev_range[0]),
        'distances': dist_range[0] + u(B) * (dist_range[1] - dist_range[0]),
        'biases': (u(B, 2) * 2 - 1) * bias_range,
        'vertices': vertices_init[None] + delta,
        'delta_vertices': delta,
        'textures': u(B, 3, Ht, Wt),
        'lights': lights,
        'bg': u(B, 3, H, W),
    }
    return A


def to_device(A, device, requires_grad=False):
    out = {}
    for k, v in A.items():
        t = v.to(device).clone()
        if requires_grad and k != 'delta_vertices':
            t.requires_grad_(True)
        out[k] = t
    return out


GRAD_KEYS = ['vertices', 'azimuths', 'elevations', 'distances', 'biases', 'textures', 'lights', 'bg']


def rel_err(a, b):
    """max |a-b| / max(|b|) -- the 'relative to the tensor's scale' figure used for every gradient."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    scale = b.abs().max().item()
    return (a - b).abs().max().item() / (scale if scale > 0 else 1.0)


def oracle_for(dr, dtype=torch.float32, **variant):
    import ref_pipeline
    return ref_pipeline.OracleRender(dr.faces.cpu().numpy(), dr.face_uvs.cpu().numpy(), dr.image_size, dr.ratio,
                                     dr.image_weight, dtype=dtype)


def boundary_check(orc64, A64, fidx_a, fidx_b, H, W, tol=2e-4):
    """For every pixel where the two fp32 pipelines picked different faces: how close is that pixel, IN THE FP64 ORACLE, to
    flipping?  margin = the smallest |barycentric weight| of the two faces at the pixel (an edge passes through the pixel
    centre), or their relative depth gap when both contain it.  Returns (#pixels with margin > tol, worst margin).
    A flip with a large margin is a real disagreement; a flip inside fp32 rounding of a boundary is the reference algorithm's
    own ill-conditioning (two fp32 vertex stages differ by ~2e-7)."""
    diff = (fidx_a != fidx_b).nonzero()
    if diff.numel() == 0:
        return 0, 0.0
    with torch.no_grad():
        fvc, fvi, _ = orc64.vertex_stage({k: v.detach() for k, v in A64.items()})
    worst, bad = 0.0, 0
    for b, iy, ix in diff.tolist():
        x0 = (2 * ix + 1 - W) / W
        y0 = (H - 2 * iy - 1) / H
        margins, depths = [], []
        for f in (int(fidx_a[b, iy, ix]), int(fidx_b[b, iy, ix])):
            if f < 0:
                continue
            (ax, ay), (bx, by), (cx, cy) = fvi[b, f].tolist()
            m, p, n, q, s_, t = bx - ax, by - ay, cx - ax, cy - ay, x0 - ax, y0 - ay
            k3 = m * q - n * p
            w1, w2 = (s_ * q - n * t) / (k3 + 1e-14), (m * t - s_ * p) / (k3 + 1e-14)
            w0 = 1 - w1 - w2
            margins.append(min(abs(w0), abs(w1), abs(w2)))
            z = fvc[b, f, :, 2].tolist()
            depths.append((min(w0, w1, w2), w0 * z[0] + w1 * z[1] + w2 * z[2]))
        margin = min(margins) if margins else 1.0
        if len(depths) == 2 and depths[0][0] >= 0 and depths[1][0] >= 0:          # both contain the pixel: a depth tie?
            margin = min(margin, abs(depths[0][1] - depths[1][1]) / max(abs(depths[0][1]), 1e-12))
        worst = max(worst, margin)
        bad += margin > tol
    return bad, worst


def run_parity_case(mm, mesh="icosphere", B=2, image_size=32, ratio=1, no_mask=True, contour=0.1, seed=0,
                    device="cuda:0", init_ellipsoid=1, dist_range=(2.0, 7.0), image_weight=1.0, fused=True, tex=None,
                    elev_range=(0.0, 30.0), bias_range=0.3):
    """CUDA product vs CPU oracle on one seeded case.  Returns a dict of error figures."""
    import ctypes
    import kaolin_shim as kal
    tm = get_mesh(mm, mesh)
    dr = mm.DiffRender(tm, image_size, ratio=ratio, init_ellipsoid=init_ellipsoid, image_weight=image_weight)
    H, W = dr.height, dr.image_size
    Ht, Wt = tex if tex is not None else (None, None)
    kw = dict(dist_range=dist_range, elev_range=elev_range, bias_range=bias_range, Ht=Ht, Wt=Wt)
    A_cpu = make_attributes(dr.vertices_init, B, H, W, seed, **kw)
    gt_src = make_attributes(dr.vertices_init, B, H, W, seed + 1000, **kw)
    orc = oracle_for(dr)

    # ---------------- oracle (CPU): GT image = render of an independent sample
    with torch.no_grad():
        gt_cpu, _, _, _ = orc.render(no_mask=no_mask, **gt_src)
        # The masked-L1 term has a KINK at pred == gt: d|x|/dx = sign(x).  Where |pred - gt| is below the fp32 noise of pred
        # (up to ~3e-3 at sliver / grazing faces) two equally valid fp32 renderings take opposite signs, and at a pixel whose
        # colour is very sensitive to the geometry (a grazing face less than a pixel high: measured 1.7e4 per unit of colour)
        # that ONE sign moves the camera gradients by tens of per cent (found on sphere2 @ 256^2, seed 56: gt_R - pred_R =
        # 3.8e-4 with the fp32 oracle, -6.4e-5 with the product).  A gradient comparison is only meaningful away from the
        # kink, so the synthetic GT is moved to at least 5e-3 from the (fp32 oracle's) prediction; the count is reported.
        pred0 = orc.render(no_mask=no_mask, **A_cpu)[0]
        d = gt_cpu[:, :3] - pred0[:, :3]
        near = d.abs() < 5e-3
        tgt = pred0[:, :3] + torch.where(d >= 0, 5e-3, -5e-3)
        tgt = torch.where((tgt < 0) | (tgt > 1), 2 * pred0[:, :3] - tgt, tgt)          # stay inside [0,1]: step the other way
        gt_cpu[:, :3] = torch.where(near, tgt, gt_cpu[:, :3])
    Ao = to_device(A_cpu, "cpu", requires_grad=True)
    rgb_o, fn_o, imn_o, fidx_o = orc.render(no_mask=no_mask, **Ao)
    loss_o, parts_o = orc.recon_data(rgb_o, gt_cpu, no_mask=no_mask, contour=contour, return_parts=True)
    # an extra consumer of face_normals so its gradient path is exercised too
    wfn = torch.randn(fn_o.shape, generator=torch.Generator().manual_seed(seed + 7)) * 1e-3
    (loss_o + (fn_o * wfn).sum()).backward()

    # ---------------- product (CUDA) through the reference-shaped Python API
    Ac = to_device(A_cpu, device, requires_grad=True)
    Ac['_want_face_idx'] = True
    gt_dev = gt_cpu.to(device)
    rgb_c, Aout = dr.render(no_mask=no_mask, **Ac)
    loss_c = dr.recon_data(rgb_c, gt_dev, no_mask=no_mask, contour=contour)
    (loss_c + (Aout['face_normals'] * wfn.to(device)).sum()).backward()
    torch.cuda.synchronize()

    res = {"l1_kink_pixels_moved": int(near.sum())}
    fidx_c = Aout['face_idx'].cpu().long()
    res["face_idx_mismatch_e2e"] = int((fidx_c != fidx_o).sum())
    res["covered_frac"] = float((fidx_o >= 0).float().mean())

    # ---------------- staged: oracle rasteriser on the product's own vertex-stage output -> bit-exact face_idx
    h = dr._ctx(torch.device(device))
    F = dr.num_faces
    fvi = torch.empty(B, F, 3, 2, device=device)
    fvz = torch.empty(B, F, 3, device=device)
    fnz = torch.empty(B, F, device=device)
    # re-run the forward through the fused entry point to get a workspace we own
    with torch.no_grad():
        out = dr.render_compare(gt_dev, no_mask=no_mask, contour=contour, **{k: v.detach() for k, v in Ac.items()
                                                                             if k != '_want_face_idx'})
    ws = out['_workspace']
    rc = mm.lib().mm_debug_export_faces(h.handle, B, ctypes.c_void_p(ws.data_ptr()), ws.numel(), ctypes.c_void_p(fvi.data_ptr()),
                                        ctypes.c_void_p(fvz.data_ptr()), ctypes.c_void_p(fnz.data_ptr()),
                                        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    feats = torch.ones(B, F, 3, 1)
    _, fidx_s = kal.rasterize(H, W, fvz.cpu(), fvi.cpu(), feats, fnz.cpu() >= 0)
    soft_s = kal.dibr_soft_mask(fvi.cpu(), fidx_s)
    res["face_idx_mismatch_staged"] = int((fidx_c != fidx_s).sum())
    res["soft_staged_max_abs_err"] = float((rgb_c[:, 3].detach().cpu() - soft_s).abs().max())
    # full staged image: oracle rasteriser + shading on the product's (fvi, fvz, unit normals)
    fn_c = Aout['face_normals'].detach().cpu()
    attrs = [torch.ones(B, F, 3, 1), orc.face_uvs.repeat(B, 1, 1, 1), fn_c.unsqueeze(-2).repeat(1, 1, 3, 1)]
    with torch.no_grad():
        (tm_s, tc_s, imn_s), soft_s2, _ = kal.dibr_rasterization(H, W, fvz.cpu(), fvi.cpu(), attrs, fnz.cpu())
        rgb_s = orc.shade(A_cpu, no_mask, tm_s, tc_s, imn_s, soft_s2)
    res["rgba_staged_max_abs_err"] = float((rgb_c.detach().cpu() - rgb_s).abs().max())
    res["imnormal_staged_max_abs_err"] = float((Aout['imnormal'].detach().cpu() - imn_s).abs().max())
    res["vertex_stage_fvi_rel_err"] = rel_err(fvi, orc.vertex_stage(A_cpu)[1])

    # ---------------- end-to-end numbers (pixels whose winner differs between the two vertex stages are excluded
    # from the max-abs figure and counted separately)
    same = (fidx_c == fidx_o)[:, None].expand(-1, 4, -1, -1)
    diff = (rgb_c.detach().cpu() - rgb_o.detach()).abs()
    res["rgba_max_abs_err"] = float(diff[same].max())
    res["rgba_mean_abs_err"] = float(diff.mean())
    res["loss_cuda"] = float(loss_c)
    res["loss_oracle"] = float(loss_o)
    res["loss_rel_err"] = abs(float(loss_c) - float(loss_o)) / max(abs(float(loss_o)), 1e-12)
    res["face_normals_rel_err"] = rel_err(Aout['face_normals'], fn_o)
    # fp32 noise floor of the reference algorithm itself: the same oracle in fp64 is the arbiter.  Sliver faces at
    # the silhouette (k3 -> 0) and small faces far from the camera amplify the ~2e-7 difference between two fp32
    # vertex stages; whatever the fp32 oracle loses against fp64, the product may lose too (x4), never more.
    orc64 = oracle_for(dr, torch.float64)
    A64 = {k: v.double().requires_grad_(k != 'delta_vertices') for k, v in A_cpu.items()}
    rgb64, fn64, _, fidx64 = orc64.render(no_mask=no_mask, **A64)
    loss64 = orc64.recon_data(rgb64, gt_cpu.double(), no_mask=no_mask, contour=contour)
    (loss64 + (fn64 * wfn.double()).sum()).backward()
    rgb64, fn64 = rgb64.detach(), fn64.detach()
    agree = ((fidx_c == fidx_o) & (fidx_o == fidx64))[:, None].expand(-1, 4, -1, -1)
    res["rgba_noise_f32_oracle_vs_f64"] = float((rgb_o.detach().double() - rgb64).abs()[agree].max())
    res["rgba_err_vs_f64"] = float((rgb_c.detach().cpu().double() - rgb64).abs()[agree].max())
    # distribution-level figure: which fraction of the (agreeing) pixels is further than 1e-4 from the fp64 oracle -- for the
    # product and for the fp32 oracle itself (random U[0,1] texels on a 256..512-row atlas amplify a 1e-7 error of u,v to 1e-4)
    npx = max(int(agree.sum()), 1)
    res["rgba_frac_gt_1e-4_cuda_vs_f64"] = float(((rgb_c.detach().cpu().double() - rgb64).abs()[agree] > 1e-4).sum()) / npx
    res["rgba_frac_gt_1e-4_f32_oracle_vs_f64"] = float(((rgb_o.detach().double() - rgb64).abs()[agree] > 1e-4).sum()) / npx
    res["face_normals_noise_f32_oracle_vs_f64"] = rel_err(fn_o, fn64)
    res["face_normals_err_vs_f64"] = rel_err(Aout['face_normals'], fn64)
    res["loss_noise_f32_oracle_vs_f64"] = abs(float(loss_o) - float(loss64)) / max(abs(float(loss64)), 1e-12)
    # every pixel whose winner differs end to end must sit on a decision boundary of the fp64 oracle (an edge of one of the two
    # faces within fp32 rounding of the pixel centre, or a depth tie): `face_idx_unexcused` counts the ones that do not
    res["face_idx_unexcused"], res["face_idx_worst_margin"] = boundary_check(orc64, A64, fidx_c, fidx_o, H, W)
    only32 = (fidx_c == fidx_o) & (fidx_o != fidx64)
    res["face_idx_f32_oracle_vs_f64"] = int((fidx_o != fidx64).sum())
    res["face_idx_both_f32_differ_from_f64"] = int(only32.sum())
    for k in GRAD_KEYS:
        if k == 'bg' and not no_mask:
            continue
        res["grad_" + k + "_rel_err"] = rel_err(Ac[k].grad, Ao[k].grad)
        res["gnoise_" + k] = rel_err(Ao[k].grad, A64[k].grad)              # fp32 oracle vs fp64 oracle
        res["gerr64_" + k] = rel_err(Ac[k].grad, A64[k].grad)              # product vs fp64 oracle
    # fused entry point vs the two-call path
    if fused:
        res["fused_loss_rel_err"] = abs(float(out['loss'][0]) - float(loss_o)) / max(abs(float(loss_o)), 1e-12)
        res["fused_rgba_max_abs_vs_unfused"] = float((out['rgba'] - rgb_c.detach()).abs().max())
    # the same step with lazy fusion off (recon_data materialises its gradient, the render backward takes it as g_rgba)
    dr.lazy_fusion = False
    An = to_device(A_cpu, device, requires_grad=True)
    rgb_n, Aout_n = dr.render(no_mask=no_mask, **An)
    loss_n = dr.recon_data(rgb_n, gt_dev, no_mask=no_mask, contour=contour)
    (loss_n + (Aout_n['face_normals'] * wfn.to(device)).sum()).backward()
    dr.lazy_fusion = True
    res["lazy_vs_materialised_rgba"] = float((rgb_n.detach() - rgb_c.detach()).abs().max())
    res["lazy_vs_materialised_loss"] = abs(float(loss_n) - float(loss_c))
    res["lazy_vs_materialised_grad"] = max(rel_err(Ac[k].grad, An[k].grad) for k in GRAD_KEYS if not (k == 'bg' and not no_mask))
    return res


# ------------------------------------------------------------------ mesh regularisers fixtures
REG_WEIGHTS = [1.0, 0.7, 1.3, 0.9, 1.1, 0.5, 2.0]                # calc_reg_loss, edge, depth, depthR, depthC, deform, flip


def reg_inputs(V, F, seed=5, B=3):
    g = torch.Generator().manual_seed(seed)
    delta = 0.05 * torch.randn(B, V, 3, generator=g)
    delta[0, :7] = 0.0
    fn = torch.nn.functional.normalize(torch.randn(B, F, 3, generator=g), dim=2)
    return delta, fn


def reg_values(dr, delta, fn, temp=1.5):
    """The seven reference calls of trainer.py:54-68 on one attribute set -> (values[7], d/d delta, d/d face_normals).
    `dr`: the product's DiffRender (CUDA tensors: the fused kernel) or tests/reg_torch.TorchRegularisers (the torch checker)."""
    vinit = (dr.dr if hasattr(dr, 'dr') else dr).vertices_init
    d = delta.clone().requires_grad_(True)
    n = fn.clone().requires_grad_(True)
    att = {'delta_vertices': d, 'face_normals': n, 'vertices': vinit.to(d.device)[None] + d}
    vals = torch.stack([dr.calc_reg_loss(att), dr.calc_reg_edge(att['vertices']), dr.calc_reg_depth(att['vertices']),
                        dr.calc_reg_depthR(att['vertices'], temp=temp), dr.calc_reg_depthC(att['vertices']),
                        dr.calc_reg_deform(att['delta_vertices']), dr.recon_flip(att, False)])
    (vals * torch.tensor(REG_WEIGHTS, device=vals.device)).sum().backward()
    return vals.detach(), d.grad, n.grad
