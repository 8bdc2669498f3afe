"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Every compute call goes through the
C ABI of libmagicmirror.so via the reference-shaped Python DiffRender; the CPU oracle is the checker.

PARITY IS UNPINNED for the Kaolin part: the oracle restates Kaolin's DIB-R kernels from the published algorithm (Kaolin is
not installable offline); docs/DIBR_SPEC.md lists the assumptions, docs/DIBR_SENSITIVITY.md what each alternative would change.

Bars (fp32 path; BASELINE.json north_star: "bit-exact face indices / visibility, RGBA and loss / grad within 1e-4 rel"):
  * face_idx: BIT-EXACT against the oracle rasteriser fed the product's own vertex-stage output ("staged").  End to end the
    two pipelines have different fp32 vertex stages (torch-CPU vs our kernel, ~2e-7 rel apart): a pixel may only flip if it
    sits on a decision boundary of the SAME oracle run in fp64 (an edge through the pixel centre within 2e-4 barycentric
    units, or a depth tie) -- `face_idx_unexcused` must be 0.
  * RGBA: staged (identical face records) <= 5e-5 abs (values in [0,1]; bilinear texel coordinates reach 511, 1 ulp there =
    3e-5).  End to end the bar is 1e-4, with the NOISE RULE: the same oracle run in fp64 is the arbiter, and where the fp32
    ORACLE ITSELF is further than 1e-4 / NOISE_X from it (the reference algorithm's own conditioning: k/k3 barycentrics of
    sliver faces, U[0,1] texels on a 256..512-row atlas turn a 1e-7 error of u,v into 1e-4 of colour) the bar is NOISE_X x
    that noise, for the product against the fp32 oracle AND against the fp64 oracle.  Measured on the B200 (round 2): the
    product's error is of the size of the fp32 oracle's own (0.3x - 4x).  Plus a distribution figure: the fraction of pixels
    further than 1e-4 from the fp64 oracle may not exceed 4x the fp32 oracle's own fraction (+ 2e-4).
  * The synthetic GT is kept 5e-3 away from the prediction (parity_utils.run_parity_case): the masked-L1 term has a kink at
    pred == gt, and at a pixel whose colour is very sensitive to the geometry ONE sign decided by fp32 noise moves the camera
    gradients by tens of per cent -- for the fp32 oracle against the fp64 oracle just as for the product (this, not a sliver
    face, is what made round 1 skip seed 23: there the fp32 ORACLE was the outlier, 3 % off the fp64 oracle).
  * loss: 1e-5 rel.  Gradients: max|a-b| / max|b| <= 1e-4 under the same noise rule (`gnoise_*` = fp32 oracle vs fp64 oracle,
    `gerr64_*` = product vs fp64 oracle, reported for every tensor).  NOISE_X = 8: the noise is estimated from ONE sample of a
    heavy-tailed max statistic.
  * lazy fusion (recon_data's gradient formed inside the render backward) == the materialised path: image and loss bit-equal,
    gradients to float-atomics order (2e-5).
"""
import glob
import os

import numpy as np
import pytest
import torch

import parity_utils as pu

pytestmark = pytest.mark.gpu

TOL_STAGED_RGBA = 5e-5
TOL_E2E_RGBA = 1e-4
TOL_LOSS = 1e-5
TOL_GRAD = 1e-4
NOISE_X = 8.0
DEV = "cuda:0"


def _fmt(res):
    return {k: (float("%.3g" % v) if isinstance(v, float) else v) for k, v in res.items()}


def _check(res, B, H, W):
    print("PARITY", B, H, W, _fmt(res))
    assert res["face_idx_mismatch_staged"] == 0, res
    assert res["face_idx_unexcused"] == 0, res
    assert res["face_idx_mismatch_e2e"] <= max(2, int(2e-4 * B * H * W)), res
    assert res["soft_staged_max_abs_err"] <= 2e-6, res
    assert res["rgba_staged_max_abs_err"] <= TOL_STAGED_RGBA, res
    assert res["imnormal_staged_max_abs_err"] <= 2e-6, res
    noise = res["rgba_noise_f32_oracle_vs_f64"]
    assert res["rgba_err_vs_f64"] <= max(TOL_E2E_RGBA, NOISE_X * noise), res
    assert res["rgba_max_abs_err"] <= max(TOL_E2E_RGBA, NOISE_X * noise), res
    # (a sanity bound on small counts, not a precise one: measured 0.3x - 4x of the fp32 oracle's own fraction)
    assert res["rgba_frac_gt_1e-4_cuda_vs_f64"] <= 4.0 * res["rgba_frac_gt_1e-4_f32_oracle_vs_f64"] + 2e-4, res
    assert res["rgba_mean_abs_err"] <= 2e-6 + 2.0 * res["face_idx_mismatch_e2e"] / (B * H * W), res
    assert res["loss_rel_err"] <= TOL_LOSS and res["fused_loss_rel_err"] <= TOL_LOSS, res
    assert res["fused_rgba_max_abs_vs_unfused"] == 0.0, res
    assert res["face_normals_err_vs_f64"] <= max(1e-4, NOISE_X * res["face_normals_noise_f32_oracle_vs_f64"]), res
    assert res["lazy_vs_materialised_rgba"] == 0.0 and res["lazy_vs_materialised_loss"] == 0.0, res
    assert res["lazy_vs_materialised_grad"] <= 5e-5, res                       # same kernels, float-atomics order only
    for k, v in res.items():
        if k.startswith("grad_"):
            name = k[5:-8]
            bar = max(TOL_GRAD, NOISE_X * res["gnoise_" + name])
            assert v <= bar, (k, v, bar, res)
            assert res["gerr64_" + name] <= bar, (k, res)


CASES = [
    dict(mesh="icosphere", B=2, image_size=32, no_mask=True, contour=0.1, seed=3),
    dict(mesh="icosphere", B=3, image_size=64, no_mask=False, contour=0.0, seed=5),
    dict(mesh="ellipsoid", B=4, image_size=128, no_mask=True, contour=0.1, seed=7),            # cfg-2 shape, small B
    dict(mesh="smpl_uv_642", B=2, image_size=64, ratio=2, init_ellipsoid=2, no_mask=True, contour=0.1, seed=9,
         dist_range=(2.0, 6.0)),                                                                # cfg-4 shape (Market)
    dict(mesh="sphere", B=1, image_size=20, ratio=1.8, no_mask=False, contour=0.1, seed=11),    # ragged sub-tiles
    dict(mesh="sphere", B=2, image_size=64, no_mask=True, contour=0.0, seed=13, dist_range=(6.5, 7.0)),  # knum truncation
    dict(mesh="sphere", B=2, image_size=48, no_mask=True, contour=0.1, seed=15, dist_range=(1.6, 2.0)),  # fills the frame
    dict(mesh="sphere2", B=2, image_size=64, no_mask=True, contour=0.1, seed=17),               # F=5120: records not in smem
    dict(mesh="sphere", B=2, image_size=22, ratio=1.5, no_mask=True, contour=0.1, seed=19),     # 33x22: scalar (non-float4) rows,
                                                                                                 # contour through index tables
    dict(mesh="icosphere", B=2, image_size=30, ratio=1.2, no_mask=False, contour=0.1, seed=21), # 36x30: H % 4 == 0, W % 4 != 0
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%d-s%d" % (c["mesh"], c["image_size"], c["seed"]))
def test_parity_against_oracle(mm, case):
    res = pu.run_parity_case(mm, device=DEV, **case)
    H = int(round(case.get("ratio", 1) * case["image_size"]))
    _check(res, case["B"], H, case["image_size"])


# BASELINE.json configs at their REAL sizes, against the oracle (not fused-vs-unfused): the bench workload itself (cfg-2, B=48,
# seed 1234), the Market shape (cfg-4: 256x128, smpl_uv_642, ratio 2, train_market.py:126-130 camera ranges) and the high-res
# shape (cfg-5: 256x256, 512x512 atlas; sphere and the 5120-face sphere2).  The C/OpenMP oracle needs a few seconds per case.
FULL_SIZE = [
    dict(mesh="ellipsoid", B=48, image_size=128, no_mask=True, contour=0.1, seed=1234),
    dict(mesh="smpl_uv_642", B=8, image_size=128, ratio=2, init_ellipsoid=2, no_mask=True, contour=0.1, seed=4321,
         dist_range=(2.0, 6.0), elev_range=(-15.0, 15.0), bias_range=0.5),
    dict(mesh="sphere", B=2, image_size=256, no_mask=True, contour=0.1, seed=55, tex=(512, 512)),
    dict(mesh="sphere2", B=2, image_size=256, no_mask=True, contour=0.1, seed=56, tex=(512, 512)),
]


@pytest.mark.parametrize("case", FULL_SIZE, ids=["cfg2-B48-128", "cfg4-B8-256x128", "cfg5-sphere-256-tex512", "cfg5-sphere2-256-tex512"])
def test_parity_against_oracle_full_size(mm, case):
    res = pu.run_parity_case(mm, device=DEV, **case)
    H = int(round(case.get("ratio", 1) * case["image_size"]))
    print("parity", case["mesh"], {k: (float("%.3g" % v) if isinstance(v, float) else v) for k, v in res.items()})
    _check(res, case["B"], H, case["image_size"])


def test_seed23_sliver_winner_is_the_reference_algorithms_conditioning(mm):
    """Round 1 skipped this input ('sliver winner: flat grad tolerance fails').  It is re-admitted under the noise rule: the
    test prints how far the fp32 ORACLE's own gradients are from the fp64 oracle's on it, and the product must stay within
    4 x that (or 1e-4, whichever is larger) of BOTH -- i.e. it may be as ill-conditioned as the reference algorithm, not more."""
    case = dict(mesh="ellipsoid", B=3, image_size=64, no_mask=True, contour=0.1, seed=23)
    res = pu.run_parity_case(mm, device=DEV, **case)
    print("seed23", {k: float("%.3g" % v) for k, v in res.items() if k.startswith(("grad_", "gnoise_", "gerr64_"))})
    _check(res, 3, 64, 64)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(pu.GOLDEN, "render_*.npz"))),
                         ids=lambda p: os.path.basename(p)[7:-4])
def test_against_committed_golden(mm, path):
    """CUDA product vs fixtures produced by the unmodified reference Python (over the kaolin shim)."""
    z = np.load(path)
    S, ratio, ell, B, no_mask, contour, seed = z["meta"]
    S, B, no_mask = int(S), int(B), bool(no_mask)
    ratio = int(ratio) if float(ratio).is_integer() else float(ratio)
    mesh = {"sphere": "sphere", "ellips": "ellipsoid", "smpl_6": "smpl_uv_642"}[os.path.basename(path)[7:13]]
    dr = mm.DiffRender(pu.get_mesh(mm, mesh), S, ratio=ratio, init_ellipsoid=int(ell), image_weight=1.0)
    A = {k[3:]: torch.from_numpy(z[k]).to(DEV).requires_grad_(k != "in_delta_vertices") for k in z.files if k.startswith("in_")}
    rgbs, Aout = dr.render(no_mask=no_mask, **A)
    loss = dr.recon_data(rgbs, torch.from_numpy(z["gt"]).to(DEV), no_mask=no_mask, contour=float(contour))
    loss.backward()
    want = torch.from_numpy(z["rgbs"])
    diff = (rgbs.detach().cpu() - want).abs()
    assert float(diff.mean()) <= 5e-6 and float((diff > 1e-3).float().mean()) <= 2e-4
    assert abs(float(loss) - float(z["loss"])) <= 1e-4 * abs(float(z["loss"]))
    assert pu.rel_err(Aout['face_normals'], torch.from_numpy(z["face_normals"])) <= 1e-4
    for k in pu.GRAD_KEYS:
        if "grad_" + k in z.files:
            # the fixtures hold fp32 results only (no fp64 arbiter here): tensors 1e-4; the per-image camera scalars are sums
            # of ~1e3 cancelling per-vertex terms (B values, the max-normalisation has nothing to average over): 5e-4
            bar = 5e-4 if k in ('azimuths', 'elevations', 'distances', 'biases') else TOL_GRAD
            assert pu.rel_err(A[k].grad, torch.from_numpy(z["grad_" + k])) <= bar, k


def _cfg2(mm, B=48, seed=1234):
    dr = mm.DiffRender(pu.get_mesh(mm, "ellipsoid"), 128, image_weight=1.0)
    A = pu.to_device(pu.make_attributes(dr.vertices_init, B, 128, 128, seed), DEV)
    return dr, A


def test_cfg2_full_size_properties(mm):
    """BASELINE configs[1] at full size (B=48, 128^2) through size-independent properties."""
    dr, A = _cfg2(mm)
    B = 48
    A['_want_face_idx'] = True
    with torch.no_grad():
        rgb, out = dr.render(no_mask=True, **A)
        rgb2, out2 = dr.render(no_mask=True, **A)
    fidx = out['face_idx']
    soft = rgb[:, 3]
    assert torch.equal(rgb, rgb2) and torch.equal(fidx, out2['face_idx'])                      # deterministic forward
    assert bool(((soft >= 0) & (soft <= 1)).all()) and bool((soft[fidx >= 0] == 1).all())
    assert bool(((rgb[:, :3] >= 0) & (rgb[:, :3] <= 1)).all())
    assert int(fidx.max()) < dr.num_faces and int(fidx.min()) >= -1
    cov = (fidx >= 0).float().mean().item()
    assert 0.02 < cov < 0.95
    # uncovered pixels carry no normal; covered ones a unit normal scaled by w0+w1+w2 ~ 1
    nrm = out['imnormal'].norm(dim=-1)
    assert bool((nrm[fidx < 0] == 0).all()) and bool(((nrm[fidx >= 0] - 1).abs() < 1e-3).all())
    # visible faces are front-facing
    fn = out['face_normals']
    bidx = torch.arange(B, device=DEV)[:, None, None].expand_as(fidx)
    assert bool((fn[bidx[fidx >= 0], fidx[fidx >= 0].long(), 2] >= 0).all())
    # batch-permutation equivariance (bitwise) and azimuth + 360 invariance (fp32 trig: tolerance)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0)).to(DEV)
    Ap = {k: (v[perm] if torch.is_tensor(v) else v) for k, v in A.items()}
    with torch.no_grad():
        rgbp, _ = dr.render(no_mask=True, **Ap)
        A360 = dict(A); A360['azimuths'] = A['azimuths'] + 360.0
        rgb360, _ = dr.render(no_mask=True, **A360)
    assert torch.equal(rgbp, rgb[perm])
    assert float(((rgb360 - rgb).abs() > 1e-3).float().mean()) < 1e-3
    # recon_data(x, x): image term 0, mask term 1 - sum(m^2)/sum(2m-m^2), contour term 0
    got = float(dr.recon_data(rgb, rgb, no_mask=True, contour=0.1))
    m = soft.reshape(B, -1).double()
    want = float(1 - ((m * m).sum(1) / ((2 * m - m * m).sum(1) + 1e-10)).mean())
    assert abs(got - want) < 1e-5


def test_cfg2_fused_equals_unfused_and_linearity(mm):
    dr, A = _cfg2(mm, B=48, seed=99)
    gt, _ = dr.render(no_mask=True, **pu.to_device(pu.make_attributes(dr.vertices_init, 48, 128, 128, 7), DEV))
    Ag = {k: v.clone().requires_grad_(k != 'delta_vertices') for k, v in A.items()}
    rgb, _ = dr.render(no_mask=True, **Ag)
    loss = dr.recon_data(rgb, gt, no_mask=True, contour=0.1)
    loss.backward(retain_graph=True)
    out = dr.render_compare(gt, no_mask=True, contour=0.1, **A)
    assert torch.equal(out['rgba'], rgb.detach())
    assert abs(float(out['loss'][0]) - float(loss)) <= 1e-6 * abs(float(loss))
    names = dict(vertices='g_vertices', azimuths='g_azimuths', elevations='g_elevations', distances='g_distances',
                 biases='g_biases', textures='g_textures', lights='g_lights', bg='g_bg')
    for k, gk in names.items():
        assert pu.rel_err(out[gk], Ag[k].grad) <= 2e-5, k            # same kernels, atomics order only
    # linearity in loss_scale and in the extra upstream gradient
    out2 = dr.render_compare(gt, no_mask=True, contour=0.1, loss_scale=2.0, **A)
    assert pu.rel_err(out2['g_textures'], 2 * out['g_textures']) <= 2e-5
    assert pu.rel_err(out2['g_vertices'], 2 * out['g_vertices']) <= 2e-5
    gx = torch.randn(48, 4, 128, 128, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1)) * 1e-4
    out3 = dr.render_compare(gt, no_mask=True, contour=0.1, g_rgba_extra=gx, **A)
    rgb.backward(gx)            # accumulates into Ag grads: loss grad + extra
    for k, gk in names.items():
        assert pu.rel_err(out3[gk], Ag[k].grad) <= 5e-5, k


@pytest.mark.parametrize("cfg", [
    dict(mesh="smpl_uv_642", B=8, image_size=128, ratio=2, init_ellipsoid=2, dist_range=(2.0, 6.0)),     # cfg-4 shape: 256x128, tex 512x128
    dict(mesh="sphere2", B=4, image_size=256, ratio=1, init_ellipsoid=1, dist_range=(2.0, 7.0), Ht=512, Wt=512),  # cfg-5 shape: F=5120, 256^2, atlas 512^2
], ids=["cfg4-market-256x128", "cfg5-sphere2-256x256-tex512"])
def test_large_configs_fused_equals_unfused(mm, cfg):
    """BASELINE configs[3] / configs[4] shapes: the fused step must reproduce render -> recon_data -> backward."""
    B, S = cfg["B"], cfg["image_size"]
    dr = mm.DiffRender(pu.get_mesh(mm, cfg["mesh"]), S, ratio=cfg["ratio"], init_ellipsoid=cfg["init_ellipsoid"], image_weight=1.0)
    H, W = dr.height, dr.image_size
    mk = lambda seed: pu.to_device(pu.make_attributes(dr.vertices_init, B, H, W, seed, Ht=cfg.get("Ht"), Wt=cfg.get("Wt"),  # noqa: E731
                                                      dist_range=cfg["dist_range"]), DEV)
    A = mk(31)
    with torch.no_grad():
        gt, _ = dr.render(no_mask=True, **mk(32))
    Ag = {k: v.clone().requires_grad_(k != 'delta_vertices') for k, v in A.items()}
    rgb, _ = dr.render(no_mask=True, **Ag)
    loss = dr.recon_data(rgb, gt, no_mask=True, contour=0.1)
    loss.backward()
    out = dr.render_compare(gt, no_mask=True, contour=0.1, **A)
    assert torch.equal(out['rgba'], rgb.detach())
    assert bool(torch.isfinite(out['rgba']).all())
    assert abs(float(out['loss'][0]) - float(loss)) <= 1e-6 * abs(float(loss))
    names = dict(vertices='g_vertices', azimuths='g_azimuths', elevations='g_elevations', distances='g_distances',
                 biases='g_biases', textures='g_textures', lights='g_lights', bg='g_bg')
    for k, gk in names.items():
        assert bool(torch.isfinite(out[gk]).all()), k
        assert pu.rel_err(out[gk], Ag[k].grad) <= 5e-5, k            # same arithmetic, atomics order only
    cov = (rgb[:, 3] == 1).float().mean().item()
    assert 0.005 < cov < 0.98


@pytest.mark.parametrize("shape", [(2, 32, 32), (3, 30, 22), (2, 160, 96), (1, 7, 5)])
@pytest.mark.parametrize("contour", [0.0, 0.1])
def test_recon_data_vs_oracle_any_size(mm, shape, contour):
    """Stand-alone recon_data (value + gradient) incl. the nearest-down/up contour tables on sizes not divisible by 4."""
    B, H, W = shape
    dr = mm.DiffRender(mm.icosphere(1), W, ratio=H / W, image_weight=0.7)
    assert dr.height == H
    g = torch.Generator().manual_seed(H * W)
    pred = torch.rand(B, 4, H, W, generator=g)
    gt = torch.rand(B, 4, H, W, generator=g)
    gt[:, 3] = (gt[:, 3] > 0.5).float()
    orc = pu.oracle_for(dr)
    po = pred.clone().requires_grad_(True)
    lo = orc.recon_data(po, gt, contour=contour)
    lo.backward()
    pc = pred.to(DEV).requires_grad_(True)
    lc = dr.recon_data(pc, gt.to(DEV), contour=contour)
    (3.0 * lc).backward()
    assert abs(float(lc) - float(lo)) <= TOL_LOSS * abs(float(lo))
    assert pu.rel_err(pc.grad, 3.0 * po.grad) <= 2e-5


def test_error_behaviour(mm):
    dr = mm.DiffRender(mm.icosphere(2), 32)
    A = pu.to_device(pu.make_attributes(dr.vertices_init, 2, 32, 32, 0), DEV)
    with pytest.raises(KeyError):
        dr.render(no_mask=False, **{k: v for k, v in A.items() if k != 'lights'})
    A2 = dict(A); A2['bg'] = None
    with pytest.raises(TypeError):
        dr.render(no_mask=True, **A2)
    dr.render(no_mask=False, **A2)           # bg=None is fine when it is not used (networks.py:263,312-313)
    A3 = dict(A); A3['vertices'] = A['vertices'][:, :100]
    with pytest.raises(ValueError):
        dr.render(no_mask=False, **A3)
    with pytest.raises(ValueError):
        dr.recon_data(torch.rand(2, 4, 16, 16, device=DEV), torch.rand(2, 4, 16, 16, device=DEV))
    # raw ABI: NULL pointers are rejected with an error code and a message, not a crash
    import ctypes
    L = mm.lib()
    h = dr._ctx(torch.device(DEV))
    ws = h.workspace(2)
    nul = ctypes.c_void_p(0)
    rc = L.mm_render_forward(h.handle, 2, *([nul] * 6), 64, 32, 0, nul, nul, 0, *([nul] * 4), ctypes.c_void_p(ws.data_ptr()),
                             ws.numel(), nul)
    assert rc == -1 and b"invalid argument" in L.mm_last_error()
    # an undersized / missing / misaligned workspace is detected, not overrun
    good = [ctypes.c_void_p(A[k].data_ptr()) for k in ('vertices', 'azimuths', 'elevations', 'distances', 'biases', 'textures')]
    rgba = torch.empty(2, 4, 32, 32, device=DEV)
    tail = [64, 32, 0, ctypes.c_void_p(A['lights'].data_ptr()), nul, 0, ctypes.c_void_p(rgba.data_ptr()), nul, nul, nul]
    rc = L.mm_render_forward(h.handle, 2, *good, *tail, ctypes.c_void_p(ws.data_ptr()), ws.numel() - 1, nul)
    assert rc == -1 and b"workspace holds" in L.mm_last_error()
    rc = L.mm_render_forward(h.handle, 2, *good, *tail, nul, ws.numel(), nul)
    assert rc == -1 and b"workspace is NULL" in L.mm_last_error()
    rc = L.mm_render_forward(h.handle, 2, *good, *tail, ctypes.c_void_p(ws.data_ptr() + 4), ws.numel(), nul)
    assert rc == -1 and b"256-byte aligned" in L.mm_last_error()
    rc = L.mm_render_forward(h.handle, 70000, *good, *tail, ctypes.c_void_p(ws.data_ptr()), ws.numel(), nul)
    assert rc == -1
    # render_compare validates like render does (shapes, devices) instead of handing bad pointers to the kernels
    gt = torch.rand(2, 4, 32, 32, device=DEV)
    with pytest.raises(ValueError):
        dr.render_compare(gt[:, :3], no_mask=True, **A)
    with pytest.raises(mm.MagicMirrorError):
        dr.render_compare(gt.cpu(), no_mask=True, **A)
    with pytest.raises(ValueError):
        dr.render_compare(gt, no_mask=True, g_rgba_extra=torch.zeros(2, 4, 16, 16, device=DEV), **A)


def test_soft_backward_fallback_when_pair_list_overflows(mm, monkeypatch):
    """The backward replays the forward's (face, pixel) candidate list; if that list outgrows its buffer the backward
    re-walks the bboxes instead.  MM_PLIST_CAP (read at ctx creation) shrinks the buffer so that path runs."""
    monkeypatch.setenv("MM_PLIST_CAP", "257")
    case = dict(mesh="sphere", B=2, image_size=64, no_mask=True, contour=0.1, seed=13, dist_range=(3.0, 7.0))
    res = pu.run_parity_case(mm, **case)
    _check(res, case["B"], 64, 64)


def test_empty_scene_and_offscreen(mm):
    """Object entirely outside the frame: nothing covered, silhouette 0, gradients finite (zeros for geometry)."""
    dr = mm.DiffRender(mm.icosphere(3), 64)
    A = pu.to_device(pu.make_attributes(dr.vertices_init, 2, 64, 64, 4), DEV)
    A['vertices'] = A['vertices'] + torch.tensor([[[0.0, 50.0, 0.0]], [[-40.0, 0.0, 0.0]]], device=DEV)
    A['_want_face_idx'] = True
    Ag = {k: (v.clone().requires_grad_(True) if torch.is_tensor(v) and k != 'delta_vertices' else v) for k, v in A.items()}
    rgb, out = dr.render(no_mask=True, **Ag)
    assert bool((out['face_idx'] == -1).all()) and float(rgb[:, 3].abs().max()) == 0.0
    coef = 0.28209479177 * A['lights'][:, 0] - 0.31539156525 * A['lights'][:, 6]
    want = torch.clamp(A['bg'] * coef[:, None, None, None], 0, 1)
    assert float((rgb[:, :3] - want).abs().max()) < 1e-6
    rgb.sum().backward()
    for k in ('vertices', 'azimuths', 'textures'):
        assert bool(torch.isfinite(Ag[k].grad).all()) and float(Ag[k].grad.abs().max()) == 0.0
    assert float(Ag['bg'].grad.abs().max()) > 0


@pytest.mark.parametrize("size,ratio,mesh,kw", [(128, 1, "ellipsoid", {}), (64, 2, "smpl_uv_642", {}), (30, 1.2, "icosphere", {}),
                                                 (64, 1, "sphere", dict(dist_range=(6.5, 7.0)))])
def test_shading_schedule_and_truncation_counters(mm, size, ratio, mesh, kw):
    """The soft pass classes every 4-tile strip of the shading kernel by the rounds its dense pass needs (the shading CTAs take
    the strips longest class first) and lists the pixels with more than knum candidates: every strip must be listed exactly once,
    in the class its covered pixels (from the image's own face_idx) put it in, and the far-camera case must have truncated pixels
    (so the parity cases at that distance do exercise the in-kernel re-scan)."""
    import ctypes
    dr = mm.DiffRender(pu.get_mesh(mm, mesh), size, ratio=ratio)
    H, W = dr.height, dr.image_size
    B = 5
    A = pu.to_device(pu.make_attributes(dr.vertices_init, B, H, W, 31, **kw), DEV)
    A['_want_face_idx'] = True
    h = dr._ctx(torch.device(DEV))
    ws = h.workspace(B)
    L = mm.lib()
    rgba = torch.empty(B, 4, H, W, device=DEV); fn = torch.empty(B, dr.num_faces, 3, device=DEV)
    imn = torch.empty(B, H, W, 3, device=DEV); fidx = torch.empty(B, H, W, device=DEV, dtype=torch.int32)
    P = lambda t: ctypes.c_void_p(t.data_ptr())     # noqa: E731
    c = lambda k: A[k].contiguous()                 # noqa: E731
    tex = c('textures')
    rc = L.mm_render_forward(h.handle, B, P(c('vertices')), P(c('azimuths')), P(c('elevations')), P(c('distances')), P(c('biases')),
                             P(tex), tex.shape[2], tex.shape[3], 0, P(c('lights')), P(c('bg')), 1, P(rgba), P(fn), P(imn), P(fidx),
                             P(ws), ws.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, L.mm_last_error()
    torch.cuda.synchronize()
    off = lambda name: L.mm_debug_workspace_offset(h.handle, B, name.encode())      # noqa: E731
    assert off("no such block") == ctypes.c_size_t(-1).value
    word = lambda name, n: ws[off(name):off(name) + 4 * n].view(torch.int32).cpu().numpy()      # noqa: E731
    ntx, nty = (W + 15) // 16, (H + 7) // 8
    nstrips = (ntx * nty + 3) // 4
    n = word("sched_n", 5)
    assert int(n.sum()) == B * nstrips
    ent = word("sched_list", 5 * B * nstrips * 16).reshape(5, B * nstrips, 16)       # entry: strip id, the image's 9 lights, padding
    lists = ent[:, :, 0]
    covered = (fidx >= 0).cpu().numpy()
    want = np.zeros(B * nstrips, dtype=np.int64)
    for b in range(B):
        for t in range(ntx * nty):
            ty, tx = divmod(t, ntx)
            want[b * nstrips + t // 4] += covered[b, ty * 8:ty * 8 + 8, tx * 16:tx * 16 + 16].sum()
    seen = []
    for k in range(5):
        ids = lists[k, :n[k]]
        assert (np.minimum(4, (want[ids] + 127) // 128) == k).all(), k
        lights = ent[k, :n[k], 1:10].copy().view(np.float32)
        assert (lights == A['lights'].cpu().numpy()[ids // nstrips]).all()
        seen.append(ids)
    assert sorted(np.concatenate(seen).tolist()) == list(range(B * nstrips))
    ntrunc = int(word("ovf_count", 1)[0])
    if kw:
        assert ntrunc > 0
        pix = word("ovf_list", ntrunc)
        assert len(set(pix.tolist())) == ntrunc and bool((covered.reshape(-1)[pix] == 0).all())     # uncovered, listed once


# ------------------------------------------------------------------ fused mesh regularisers
@pytest.mark.parametrize("mesh,ratio,ell", [("sphere", 1, 1), ("smpl_uv_642", 2, 2), ("sphere2", 1, 1)])
def test_mesh_regularisers_fused_kernel_vs_torch_statement(mm, mesh, ratio, ell):
    """mm_mesh_reg_forward/backward (the product) against the torch statement of networks.py:392-491 in tests/reg_torch.py
    (itself checked against the unmodified reference in tests/test_host_setup.py).  fp32: values 2e-5 rel, grads 2e-4."""
    dr = mm.DiffRender(pu.get_mesh(mm, mesh), 64, ratio=ratio, init_ellipsoid=ell)
    g = torch.Generator().manual_seed(5)
    B, V, F = 5, dr.num_vertices, dr.num_faces
    delta = 0.05 * torch.randn(B, V, 3, generator=g)
    delta[0, :7] = 0.0                                   # zero displacements: norm sub-gradient, sign(0)
    fn = torch.nn.functional.normalize(torch.randn(B, F, 3, generator=g), dim=2)

    import reg_torch
    chk = reg_torch.TorchRegularisers(dr)

    def run(dev):
        d = delta.clone().to(dev).requires_grad_(True)
        n = fn.clone().to(dev).requires_grad_(True)
        att = {'delta_vertices': d, 'face_normals': n, 'vertices': dr.vertices_init.to(dev)[None] + d}
        R = chk if dev == "cpu" else dr              # CPU: the torch checker; CUDA: the product's fused kernel
        vals = [R.calc_reg_loss(att), R.calc_reg_edge(att['vertices']), R.calc_reg_depth(att['vertices']),
                R.calc_reg_depthR(att['vertices'], temp=1.5), R.calc_reg_depthC(att['vertices']),
                R.calc_reg_deform(att['delta_vertices']), R.recon_flip(att, False), R.recon_flip(att, True)]
        w = torch.tensor([1.0, 0.7, 1.3, 0.9, 1.1, 0.5, 2.0, 0.3], device=dev)
        (torch.stack(vals) * w).sum().backward()
        return torch.stack(vals).detach().cpu(), d.grad.cpu(), n.grad.cpu()

    v_ref, gd_ref, gn_ref = run("cpu")
    v_cu, gd_cu, gn_cu = run(DEV)
    assert torch.allclose(v_cu, v_ref, rtol=2e-5, atol=1e-8), (v_cu, v_ref)
    assert pu.rel_err(gd_cu, gd_ref) <= 2e-4 and pu.rel_err(gn_cu, gn_ref) <= 2e-4
    # all eight terms in one launch == the individual calls; deterministic
    att = {'delta_vertices': delta.to(DEV), 'face_normals': fn.to(DEV), 'vertices': dr.vertices_init.to(DEV)[None] + delta.to(DEV)}
    t1, t2 = dr.regularizer_terms(att, temp=1.5), dr.regularizer_terms(att, temp=1.5)
    assert all(torch.equal(t1[k], t2[k]) for k in t1)
    lam = dr.lambda_lpl * t1['laplacian'] + dr.lambda_flat * t1['flat']
    assert torch.allclose(lam.cpu(), v_ref[0], rtol=2e-5)
    for k, i in (('edge', 1), ('depth', 2), ('depthR', 3), ('depthC', 4), ('deform', 5), ('flip', 6)):
        assert torch.allclose(t1[k].cpu(), v_ref[i], rtol=2e-5, atol=1e-8), k


def test_mesh_regularisers_gradient_reaches_render_inputs(mm):
    """calc_reg_loss consumes the `face_normals` the render returns (trainer.py:56): its gradient must flow back through
    mm_render_backward's g_face_normals input into the vertices."""
    dr, A = _cfg2(mm, B=4, seed=3)
    Ag = {k: v.clone().requires_grad_(k == 'vertices') for k, v in A.items()}
    _, out = dr.render(no_mask=True, **Ag)
    att = {'delta_vertices': Ag['vertices'] - dr.vertices_init.to(DEV)[None], 'face_normals': out['face_normals']}
    dr.calc_reg_loss(att).backward()
    g = Ag['vertices'].grad
    assert g is not None and bool(torch.isfinite(g).all()) and float(g.abs().max()) > 0


@pytest.mark.parametrize("mesh", ["sphere", "smpl_uv_642"])
def test_mesh_regularisers_fused_kernel_vs_reference_golden(mm, mesh):
    """The fused kernel against values + gradients produced by the UNMODIFIED reference (tests/golden/reg_<mesh>.npz)."""
    z = np.load(os.path.join(pu.GOLDEN, "reg_%s.npz" % mesh))
    dr = mm.DiffRender(pu.get_mesh(mm, mesh), 64, ratio=int(z["ratio"]), init_ellipsoid=int(z["init_ellipsoid"]))
    delta, fn = pu.reg_inputs(dr.num_vertices, dr.num_faces)
    vals, gd, gn = pu.reg_values(dr, delta.to(DEV), fn.to(DEV))
    assert np.allclose(vals.cpu().numpy(), z["values"], rtol=2e-5, atol=1e-8)
    assert pu.rel_err(gd.cpu(), torch.from_numpy(z["grad_delta"])) <= 2e-4
    assert pu.rel_err(gn.cpu(), torch.from_numpy(z["grad_face_normals"])) <= 2e-4


def test_render_without_image_matches_full_render(mm):
    """render(_need_image=False) = vertex stage only; face_normals and their gradient path must equal the
    full render's (bitwise forward; backward through g_face_normals only)."""
    dr, A = _cfg2(mm, B=6, seed=11)
    keys = ('vertices', 'azimuths', 'elevations', 'distances', 'biases')
    w = torch.randn(6, dr.num_faces, 3, device=DEV, generator=torch.Generator(device=DEV).manual_seed(2))
    grads = []
    for need in (True, False):
        Ag = {k: (v.clone().requires_grad_(k in keys) if torch.is_tensor(v) else v) for k, v in A.items()}
        img, out = dr.render(no_mask=True, _need_image=need, **Ag)
        assert (img is None) == (not need)
        (out['face_normals'] * w).sum().backward()
        grads.append((out['face_normals'].detach(), {k: Ag[k].grad.clone() for k in keys}))
    assert torch.equal(grads[0][0], grads[1][0])
    for k in keys:
        # shared-memory float atomics order; the camera scalars are sums of ~1e3 cancelling per-vertex terms
        assert pu.rel_err(grads[1][1][k], grads[0][1][k]) <= (TOL_GRAD if k == 'vertices' else 5e-4), k


@pytest.mark.parametrize("size,ratio,mesh", [(128, 1, "ellipsoid"), (64, 2, "smpl_uv_642"), (30, 1.2, "icosphere")])
def test_mirrored_texture_equals_concatenated_atlas(mm, size, ratio, mesh):
    """TextureEncoder hands the renderer cat([t, t.flip(2)], 2) (model_res.py:609-610).  render(_tex_mirror=True)
    on t alone must give the SAME image bit for bit (same texels, same weights), and d/dt must equal autograd's sum over the two
    halves of the atlas gradient (float atomics order only); same through the unfused and the fused entry points."""
    B = 4
    dr = mm.DiffRender(pu.get_mesh(mm, mesh), size, ratio=ratio, init_ellipsoid=2 if ratio == 2 else 1, image_weight=1.0)
    H, W = dr.height, dr.image_size
    A = pu.to_device(pu.make_attributes(dr.vertices_init, B, H, W, 77, Ht=H, Wt=W), DEV)          # 'textures' = the half t
    G = pu.to_device(pu.make_attributes(dr.vertices_init, B, H, W, 78), DEV)
    with torch.no_grad():
        gt, _ = dr.render(no_mask=True, **G)
    half = A['textures']
    keys = ('vertices', 'azimuths', 'elevations', 'distances', 'biases', 'lights', 'bg')
    res = []
    for mirror in (False, True):
        Ag = {k: (v.clone().requires_grad_(k in keys) if torch.is_tensor(v) else v) for k, v in A.items()}
        t = half.clone().requires_grad_(True)
        Ag['textures'] = t if mirror else torch.cat([t, t.flip([2])], dim=2)
        rgb, _ = dr.render(no_mask=True, _tex_mirror=mirror, **Ag)
        loss = dr.recon_data(rgb, gt, no_mask=True, contour=0.1)
        loss.backward()
        res.append((rgb.detach(), loss.detach(), t.grad.clone(), {k: Ag[k].grad.clone() for k in keys}))
    assert torch.equal(res[0][0], res[1][0])
    assert torch.equal(res[0][1], res[1][1])
    assert float(res[0][2].abs().max()) > 0
    assert pu.rel_err(res[1][2], res[0][2]) <= 1e-5
    for k in keys:
        assert pu.rel_err(res[1][3][k], res[0][3][k]) <= TOL_GRAD, k
    # fused entry point: mirrored half vs concatenated atlas
    full = torch.cat([half, half.flip([2])], dim=2)
    Af = {k: v for k, v in A.items() if k != 'textures'}
    o0 = dr.render_compare(gt, no_mask=True, contour=0.1, textures=full, **Af)
    o1 = dr.render_compare(gt, no_mask=True, contour=0.1, tex_mirror=True, textures=half, **Af)
    assert torch.equal(o0['rgba'], o1['rgba']) and torch.equal(o0['rgba'], res[0][0])
    assert torch.equal(o0['loss'], o1['loss'])
    g_full = o0['g_textures']
    assert o1['g_textures'].shape == half.shape
    assert pu.rel_err(o1['g_textures'], g_full[:, :, :H] + g_full[:, :, H:].flip([2])) <= 1e-5
    assert pu.rel_err(o1['g_textures'], res[0][2]) <= 1e-5
    for k in ('vertices', 'azimuths', 'lights', 'bg'):
        assert pu.rel_err(o1['g_' + k], o0['g_' + k]) <= TOL_GRAD, k
    # the switch does not leak into the next call on the same ctx
    with torch.no_grad():
        again, _ = dr.render(no_mask=True, **{**Af, 'textures': full})
    assert torch.equal(again, res[0][0])


@pytest.mark.parametrize("mesh,shape", [("sphere", (3, 16, 8, 4)), ("smpl_uv_642", (2, 8, 4, 4)), ("icosphere", (2, 5, 7, 9)),
                                         ("sphere", (48, 288, 8, 4))])
def test_template_features_vs_reference_torch_ops(mm, mesh, shape):
    """SURVEY 8(f)-3, encoder side: DiffRender.template_features == the reference's own statement, network/model_res.py:317-325
    (F.grid_sample(align_corners=True, zeros) at the template's (x, y), then torch.mm with the dense V x V Laplacian),
    forward and d/dx, fp32 tolerance 1e-5 of the tensor's scale (sparse 7-term rows vs a 642-term dense dot product).
    A stretched template puts some vertices outside [-1,1]: zero padding."""
    import torch.nn.functional as F
    B, C, h, w = shape
    dr = mm.DiffRender(pu.get_mesh(mm, mesh), 64, image_weight=1.0)
    V = dr.num_vertices
    gen = torch.Generator().manual_seed(B * 1000 + C)
    x = torch.randn(B, C, h, w, generator=gen).to(DEV).requires_grad_(True)
    template = (dr.vertices_init * torch.tensor([1.3, 1.1, 1.0]))[None].to(DEV)
    wl = torch.randn(B, C, V, 1, generator=gen).to(DEV)
    wn = torch.randn(B, C, V, 1, generator=gen).to(DEV)
    local, ndiff = dr.template_features(x, template)
    ((local * wl).sum() + (ndiff * wn).sum()).backward()
    gx = x.grad.clone()
    # the reference's lines, verbatim in form
    x2 = x.detach().clone().requires_grad_(True)
    current_position = template.repeat(B, 1, 1).view(B, V, 1, 3).detach()
    uv_sampler = current_position[:, :, :, 0:2].detach()
    local_r = F.grid_sample(x2, uv_sampler, mode='bilinear', align_corners=True, padding_mode="zeros")
    lpl = dr.vertices_laplacian_matrix.to(DEV)
    torch.backends.cuda.matmul.allow_tf32 = False
    nd_r = torch.mm(local_r.view(-1, V), lpl).view(B, -1, V, 1)
    ((local_r * wl).sum() + (nd_r * wn).sum()).backward()
    assert local.shape == local_r.shape and ndiff.shape == nd_r.shape
    assert pu.rel_err(local, local_r) <= 1e-5
    assert pu.rel_err(ndiff, nd_r) <= 1e-5
    assert pu.rel_err(gx, x2.grad) <= 1e-5
    # only one of the two outputs used downstream
    x3 = x.detach().clone().requires_grad_(True)
    l3, n3 = dr.template_features(x3, template)
    (n3 * wn).sum().backward()
    x4 = x.detach().clone().requires_grad_(True)
    l4 = F.grid_sample(x4, uv_sampler, mode='bilinear', align_corners=True, padding_mode="zeros")
    (torch.mm(l4.view(-1, V), lpl).view(B, -1, V, 1) * wn).sum().backward()
    assert pu.rel_err(x3.grad, x4.grad) <= 1e-5


def test_template_features_errors(mm):
    dr = mm.DiffRender(pu.get_mesh(mm, "icosphere"), 32)
    with pytest.raises(mm.MagicMirrorError):
        dr.template_features(torch.zeros(1, 2, 4, 4), dr.vertices_init)                     # CPU tensor: no fallback
    with pytest.raises(ValueError):
        dr.template_features(torch.zeros(1, 2, 4, 4, device=DEV), dr.vertices_init[:10])    # wrong vertex count
    x = torch.zeros(1, 1, 64, 64, device=DEV, requires_grad=True)                           # plane too large for the backward
    l, n = dr.template_features(x, dr.vertices_init)
    with pytest.raises(mm.MagicMirrorError):
        (l.sum() + n.sum()).backward()


def test_render_many_equals_separate_renders(mm):
    """render_many([A1, A2, A3]) == three render calls (trainer.py:276,345,347), images bit-identical, gradients
    of a loss over all three equal up to float-atomics order; uneven batch sizes."""
    dr = mm.DiffRender(pu.get_mesh(mm, "ellipsoid"), 128, image_weight=1.0)
    sizes = (6, 3, 5)
    sets = [pu.to_device(pu.make_attributes(dr.vertices_init, b, 128, 128, 60 + i), DEV) for i, b in enumerate(sizes)]
    keys = ('vertices', 'azimuths', 'distances', 'textures', 'lights', 'bg')
    w = [torch.randn(b, 4, 128, 128, device=DEV, generator=torch.Generator(device=DEV).manual_seed(i)) for i, b in enumerate(sizes)]

    def run(many):
        As = [{k: (v.clone().requires_grad_(k in keys) if torch.is_tensor(v) else v) for k, v in A.items()} for A in sets]
        outs = dr.render_many(As, no_mask=True) if many else [dr.render(no_mask=True, **A) for A in As]
        loss = sum((img * wi).sum() + out['face_normals'].sum() for (img, out), wi in zip(outs, w))
        loss.backward()
        return [img.detach() for img, _ in outs], [{k: A[k].grad.clone() for k in keys} for A in As], outs

    img_s, g_s, _ = run(False)
    img_m, g_m, outs = run(True)
    for i, b in enumerate(sizes):
        assert img_m[i].shape == (b, 4, 128, 128) and torch.equal(img_m[i], img_s[i])
        assert outs[i][1]['face_normals'].shape == (b, dr.num_faces, 3) and outs[i][1]['imnormal'].shape == (b, 128, 128, 3)
        for k in keys:
            assert pu.rel_err(g_m[i][k], g_s[i][k]) <= TOL_GRAD, (i, k)


@pytest.mark.parametrize("shape,concat", [((4, 3, 128, 128, 128, 128), True), ((2, 3, 40, 24, 33, 17), False),
                                          ((48, 3, 128, 128, 128, 128), True)])
def test_texture_flow_vs_reference_torch_ops(mm, shape, concat):
    """SURVEY 8(f)-3, texture side: DiffRender.texture_flow == the reference's own lines, network/model_res.py:598-599,609-610
    (bicubic F.grid_sample with align_corners=True at the predicted flow, then cat([t, t.flip([2])], 2)), forward and the
    gradients w.r.t. the image and the flow.  A flow reaching outside [-1,1] exercises the zero padding of the 4x4 taps."""
    import torch.nn.functional as F
    B, C, Hi, Wi, Ho, Wo = shape
    dr = mm.DiffRender(mm.icosphere(1), 32)
    gen = torch.Generator().manual_seed(B * 7 + Ho)
    img = torch.rand(B, C, Hi, Wi, generator=gen).to(DEV).requires_grad_(True)
    flow = (torch.rand(B, 2, Ho, Wo, generator=gen) * 2.3 - 1.15).to(DEV).requires_grad_(True)
    w = torch.randn(B, C, Ho * (2 if concat else 1), Wo, generator=gen).to(DEV)
    out = dr.texture_flow(img, flow, concat=concat)
    (out * w).sum().backward()
    gi, gf = img.grad.clone(), flow.grad.clone()
    # the reference's lines, verbatim in form
    img2, flow2 = img.detach().clone().requires_grad_(True), flow.detach().clone().requires_grad_(True)
    uv_sampler = flow2.permute(0, 2, 3, 1)
    textures = F.grid_sample(img2, uv_sampler, mode='bicubic', align_corners=True)
    if concat:
        textures_flip = textures.flip([2])
        textures = torch.cat([textures, textures_flip], dim=2)
    (textures * w).sum().backward()
    assert out.shape == textures.shape
    assert pu.rel_err(out, textures) <= 2e-6
    assert pu.rel_err(gi, img2.grad) <= 2e-5          # float atomics order
    assert pu.rel_err(gf, flow2.grad) <= 2e-5
    # the un-concatenated half feeds render(_tex_mirror=True): same image as rendering the concatenated atlas
    if concat and Ho == 128:
        dr2 = mm.DiffRender(pu.get_mesh(mm, "ellipsoid"), 128, image_weight=1.0)
        A = pu.to_device(pu.make_attributes(dr2.vertices_init, B, 128, 128, 5), DEV)
        with torch.no_grad():
            half = dr2.texture_flow(img, flow, concat=False)
            full, _ = dr2.render(no_mask=True, **{**A, 'textures': out.detach()})
            mirr, _ = dr2.render(no_mask=True, _tex_mirror=True, **{**A, 'textures': half})
        assert torch.equal(full, mirr)
    with pytest.raises(mm.MagicMirrorError):
        dr.texture_flow(img.detach().cpu(), flow.detach().cpu())


def test_wgan_gp_consumer_gradient_reaches_render_inputs(mm):
    """the WGAN-GP critic is the second consumer of the rendered RGBA (trainer.py:391-438).  After a D step
    (gradient penalty: a real double backward through the critic) the generator term -D(render) sends a MATERIALISED gradient
    into the render backward while recon_data's gradient arrives lazily; autograd sums the two.  Checked three ways:
      (a) the autograd path (render -> {recon_data, critic} -> backward) against the same step through the CPU oracle,
      (b) the fused entry point fed d(-D)/d(rgba) as g_rgba_extra against (a),
      (c) lazy fusion on against off."""
    import wgan_consumer as wg
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B, S = 4, 64
    dr = mm.DiffRender(pu.get_mesh(mm, "sphere"), S, image_weight=1.0)
    A_cpu = pu.make_attributes(dr.vertices_init, B, S, S, 91)
    G_cpu = pu.make_attributes(dr.vertices_init, B, S, S, 92)
    orc = pu.oracle_for(dr)
    with torch.no_grad():
        real = orc.render(no_mask=True, **G_cpu)[0]
    alpha = torch.rand(B, 1, 1, 1, generator=torch.Generator().manual_seed(3))
    lam = 1e-2                                                   # adversarial weight large enough to matter next to the data term

    def run(device, render, recon, lazy=True):
        D = wg.make_critic(seed=7).to(device)
        optD = torch.optim.Adam(D.parameters(), lr=1e-3)
        A = pu.to_device(A_cpu, device, requires_grad=True)
        dr.lazy_fusion = lazy
        X, fn = render(A)
        wg.d_step(D, optD, real.to(device), X, alpha.to(device), lambda_gan=lam)      # double backward inside
        loss = recon(X, real.to(device)) + wg.g_loss(D, X, lambda_gan=lam)
        loss.backward()
        dr.lazy_fusion = True
        return A, X.detach(), float(loss), D

    Ac, Xc, lc, Dc = run(DEV, lambda A: (dr.render(no_mask=True, **A)[0], None),
                         lambda p, g: dr.recon_data(p, g, no_mask=True, contour=0.1))
    Ao, Xo, lo, _ = run("cpu", lambda A: (orc.render(no_mask=True, **A)[0], None),
                        lambda p, g: orc.recon_data(p, g, no_mask=True, contour=0.1))
    An, Xn, ln, _ = run(DEV, lambda A: (dr.render(no_mask=True, **A)[0], None),
                        lambda p, g: dr.recon_data(p, g, no_mask=True, contour=0.1), lazy=False)
    assert abs(lc - lo) <= 1e-4 * abs(lo)
    assert torch.equal(Xc, Xn) and lc == ln
    for k in pu.GRAD_KEYS:
        assert float(Ac[k].grad.abs().max()) > 0, k
        assert pu.rel_err(Ac[k].grad, Ao[k].grad) <= 5e-4, (k, pu.rel_err(Ac[k].grad, Ao[k].grad))      # conv stacks on two devices
        assert pu.rel_err(Ac[k].grad, An[k].grad) <= 5e-5, k                                            # lazy vs materialised
    # (b) the fused entry point with the critic's gradient as g_rgba_extra
    x = Xc.clone().requires_grad_(True)
    gx, = torch.autograd.grad(wg.g_loss(Dc, x, lambda_gan=lam), x)
    out = dr.render_compare(real.to(DEV), no_mask=True, contour=0.1, g_rgba_extra=gx,
                            **{k: v.detach() for k, v in Ac.items()})
    assert torch.equal(out['rgba'], Xc)
    names = dict(vertices='g_vertices', azimuths='g_azimuths', elevations='g_elevations', distances='g_distances',
                 biases='g_biases', textures='g_textures', lights='g_lights', bg='g_bg')
    for k, gk in names.items():
        assert pu.rel_err(out[gk], Ac[k].grad) <= 5e-5, k


def test_lazy_fusion_edge_cases(mm):
    """recon_data's gradient is handed to the render backward lazily; these are the autograd situations in which
    that hand-over must still give what plain autograd would: a retained graph walked twice, a second recon_data on the same
    image, another consumer of the image, an image that is not (any more) the untouched render output, and the whole step
    captured into a CUDA graph."""
    dr = mm.DiffRender(pu.get_mesh(mm, "sphere"), 64, image_weight=1.0)
    A0 = pu.to_device(pu.make_attributes(dr.vertices_init, 3, 64, 64, 70), DEV)
    with torch.no_grad():
        gt1, _ = dr.render(no_mask=True, **pu.to_device(pu.make_attributes(dr.vertices_init, 3, 64, 64, 71), DEV))
        gt2, _ = dr.render(no_mask=True, **pu.to_device(pu.make_attributes(dr.vertices_init, 3, 64, 64, 72), DEV))
    keys = ('vertices', 'azimuths', 'textures', 'lights', 'bg')
    w = torch.randn(3, 4, 64, 64, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1)) * 1e-4

    def leaf():
        return {k: (v.clone().requires_grad_(k in keys) if torch.is_tensor(v) else v) for k, v in A0.items()}

    def grads(A):
        return {k: A[k].grad.clone() for k in keys}

    def close(a, b, tol=3e-5):
        for k in keys:
            assert pu.rel_err(a[k], b[k]) <= tol, k

    def reference(fn):                       # the same expression with lazy fusion off
        dr.lazy_fusion = False
        try:
            A = leaf()
            fn(A).backward()
            return grads(A)
        finally:
            dr.lazy_fusion = True

    one = lambda A: dr.recon_data(dr.render(no_mask=True, **A)[0], gt1, no_mask=True, contour=0.1)      # noqa: E731
    g_one = reference(one)
    # (a) retained graph, two backward walks: gradients accumulate to twice the single walk
    A = leaf()
    loss = one(A)
    loss.backward(retain_graph=True)
    loss.backward()
    close(grads(A), {k: 2 * v for k, v in g_one.items()})
    # (b) two recon_data calls on one image (the second cannot share the workspace's IoU sums: it takes the materialised path)
    two = lambda A: (lambda X: dr.recon_data(X, gt1, no_mask=True, contour=0.1) + 0.5 * dr.recon_data(X, gt2, no_mask=True, contour=0.0))(  # noqa: E731
        dr.render(no_mask=True, **A)[0])
    A = leaf()
    two(A).backward()
    close(grads(A), reference(two))
    # (c) another consumer of the image next to the lazily fused loss
    mix = lambda A: (lambda X: 3.0 * dr.recon_data(X, gt1, no_mask=True, contour=0.1) + (X * w).sum())(dr.render(no_mask=True, **A)[0])  # noqa: E731
    A = leaf()
    mix(A).backward()
    close(grads(A), reference(mix))
    # (d) not the untouched render output: a clone, and an in-place edit -> materialised path, same numbers
    for edit in (lambda X: X.clone(), lambda X: X.mul_(1.0)):
        A = leaf()
        X = edit(dr.render(no_mask=True, **A)[0])
        dr.recon_data(X, gt1, no_mask=True, contour=0.1).backward()
        close(grads(A), g_one)
    # (e) the gradient with respect to the image ITSELF is the documented placeholder in lazy mode, the real one with it off
    A = leaf()
    X = dr.render(no_mask=True, **A)[0]
    gX, = torch.autograd.grad(dr.recon_data(X, gt1, no_mask=True, contour=0.1), X)
    assert float(gX.abs().max()) == 0.0
    dr.lazy_fusion = False
    X = dr.render(no_mask=True, **leaf())[0]
    gX, = torch.autograd.grad(dr.recon_data(X, gt1, no_mask=True, contour=0.1), X)
    dr.lazy_fusion = True
    assert float(gX.abs().max()) > 0.0
    # (f) the three calls captured into a CUDA graph and replayed: same loss, same gradients as eager
    static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in A0.items()}
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            Aw = {k: (v.detach().requires_grad_(k in keys) if torch.is_tensor(v) else v) for k, v in static.items()}
            one(Aw).backward()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        Ag = {k: (v.detach().requires_grad_(k in keys) if torch.is_tensor(v) else v) for k, v in static.items()}
        lg = one(Ag)
        lg.backward()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    A = leaf()
    le = one(A)
    le.backward()
    assert abs(float(lg) - float(le)) <= 1e-6 * abs(float(le))
    close({k: Ag[k].grad for k in keys}, grads(A))
