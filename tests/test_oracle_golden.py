"""CPU tests of the ORACLE: the restated pipeline against the committed golden fixtures (generated
by the unmodified reference Python running over the kaolin shim), the camera helpers against the
real reference outputs, closed-form answers and finite differences for the DIB-R C kernels."""
import glob
import math
import os

import numpy as np
import pytest
import torch

import parity_utils as pu
import kaolin_shim as kal
import ref_pipeline

GOLD = pu.GOLDEN


def _case_files():
    return sorted(glob.glob(os.path.join(GOLD, "render_*.npz")))


def test_camera_helpers_match_reference_outputs(mm):
    z = np.load(os.path.join(GOLD, "camera.npz"))
    t = lambda k: torch.from_numpy(z[k])   # noqa: E731
    for fns in ((ref_pipeline.camera_position_from_spherical_angles, ref_pipeline.generate_transformation_matrix),
                (mm.camera_position_from_spherical_angles, mm.generate_transformation_matrix)):
        pos = fns[0](t("dist"), t("elev"), t("azim"), degrees=True)
        T = fns[1](pos, t("look_at"), t("up"))
        assert torch.equal(pos, t("camera_position"))
        assert torch.allclose(T, t("transform"), rtol=0, atol=1e-6)


@pytest.mark.parametrize("path", _case_files(), ids=lambda p: os.path.basename(p)[7:-4])
def test_oracle_reproduces_reference_python(path):
    """OracleRender (our restatement of networks.py:258-324,364-390) == the reference's own code over the shim."""
    z = np.load(path)
    S, ratio, ell, B, no_mask, contour, seed = z["meta"]
    S, B, no_mask = int(S), int(B), bool(no_mask)
    ratio = int(ratio) if float(ratio).is_integer() else float(ratio)
    mesh = {"sphere": "sphere", "ellips": "ellipsoid", "smpl_6": "smpl_uv_642"}[os.path.basename(path)[7:13]]
    tz = np.load(os.path.join(GOLD, "templates", mesh + ".npz"))
    face_uvs = tz["uvs"][tz["face_uvs_idx"]]
    orc = ref_pipeline.OracleRender(tz["faces"], face_uvs, S, ratio, image_weight=1.0)
    A = {k[3:]: torch.from_numpy(z[k]).clone().requires_grad_(k != "in_delta_vertices") for k in z.files if k.startswith("in_")}
    rgbs, fn, imn, fidx = orc.render(no_mask=no_mask, **A)
    loss = orc.recon_data(rgbs, torch.from_numpy(z["gt"]), no_mask=no_mask, contour=float(contour))
    loss.backward()
    assert torch.equal(rgbs.detach(), torch.from_numpy(z["rgbs"]))
    assert torch.equal(fn.detach(), torch.from_numpy(z["face_normals"]))
    assert abs(float(loss) - float(z["loss"])) <= 1e-7
    for k in pu.GRAD_KEYS:
        if "grad_" + k in z.files:
            assert pu.rel_err(A[k].grad, torch.from_numpy(z["grad_" + k])) < 1e-6, k


def test_soft_mask_closed_form_single_triangle():
    """One triangle, pixel at perpendicular distance d from an edge: soft = 1 - (1 - p) = p = exp(-7000 d^2)
    (the "1 - exp" in SURVEY 8c(4) contradicts A.4 and is not what DIB-R computes)."""
    H = W = 16
    fvi = torch.tensor([[[[-0.5, -0.5], [0.5, -0.5], [0.0, 0.6]]]])          # CCW
    fvz = -torch.ones(1, 1, 3)
    feats = torch.ones(1, 1, 3, 1)
    _, fidx = kal.rasterize(H, W, fvz, fvi, feats, torch.ones(1, 1, dtype=torch.bool))
    soft = kal.dibr_soft_mask(fvi, fidx)
    assert (soft[fidx >= 0] == 1).all()
    # pixel centres: x = (2ix+1-W)/W, y = (H-2iy-1)/H ; take the one just below the bottom edge y=-0.5, x inside
    ix, iy = 8, 12          # x = 0.0625, y = -0.5625 -> d = 0.0625 > boxlen -> outside enlarged bbox -> 0
    assert fidx[0, iy, ix] == -1 and soft[0, iy, ix] == 0
    fvi2 = fvi.clone(); fvi2[..., 1] -= 0.05                                   # edge at y=-0.55 -> d = 0.0125
    _, fidx2 = kal.rasterize(H, W, fvz, fvi2, feats, torch.ones(1, 1, dtype=torch.bool))
    soft2 = kal.dibr_soft_mask(fvi2, fidx2)
    want = math.exp(-7000 * 0.0125 ** 2)
    assert fidx2[0, iy, ix] == -1
    assert abs(float(soft2[0, iy, ix]) - want) < 2e-5


def test_raster_invariants_and_knum_order():
    torch.manual_seed(0)
    B, F, H, W = 2, 60, 24, 20
    fvi = torch.rand(B, F, 3, 2) * 0.3 - 0.15                                  # many overlapping faces -> > knum hits
    fvz = -1 - torch.rand(B, F, 3)
    feats = torch.rand(B, F, 3, 3)
    valid = torch.rand(B, F) > 0.3
    interp, fidx = kal.rasterize(H, W, fvz, fvi, feats, valid)
    soft = kal.dibr_soft_mask(fvi, fidx, knum=30)
    assert ((soft >= 0) & (soft <= 1)).all()
    assert (soft[fidx >= 0] == 1).all()
    assert (interp[fidx < 0] == 0).all()
    assert valid[torch.arange(B)[:, None, None].expand_as(fidx)[fidx >= 0], fidx[fidx >= 0]].all()
    # truncation really triggers and is order dependent: reversing the face order changes the result
    soft5 = kal.dibr_soft_mask(fvi, fidx, knum=5)
    soft5r = kal.dibr_soft_mask(fvi.flip(1), torch.where(fidx >= 0, F - 1 - fidx, fidx), knum=5)
    assert (soft5 != soft).any() and (soft5 != soft5r).any()
    # batch-permutation equivariance
    p = torch.tensor([1, 0])
    _, fidx_p = kal.rasterize(H, W, fvz[p], fvi[p], feats[p], valid[p])
    assert torch.equal(fidx_p, fidx[p])


def test_dibr_backward_matches_finite_differences_f64():
    torch.manual_seed(1)
    B, F, H, W, D = 1, 5, 12, 12, 3
    dt = torch.float64
    fvi = (torch.rand(B, F, 3, 2, dtype=dt) * 1.6 - 0.8)
    fvz = -torch.rand(B, F, 3, dtype=dt) - 1
    feat = torch.rand(B, F, 3, D, dtype=dt)
    valid = torch.ones(B, F, dtype=torch.bool)
    _, idx = kal.rasterize(H, W, fvz, fvi, feat, valid)
    g = torch.randn(B, H, W, D, dtype=dt)
    g2 = torch.randn(B, H, W, dtype=dt)

    def total(x):
        return float((kal.rasterize(H, W, fvz, x, feat, valid)[0] * g).sum() + (kal.dibr_soft_mask(x, idx) * g2).sum())
    x = fvi.clone().requires_grad_(True)
    ((kal.rasterize(H, W, fvz, x, feat, valid)[0] * g).sum() + (kal.dibr_soft_mask(x, idx) * g2).sum()).backward()
    num = torch.zeros_like(fvi)
    eps = 1e-7
    for i in range(fvi.numel()):
        xp = fvi.clone(); xp.view(-1)[i] += eps
        xm = fvi.clone(); xm.view(-1)[i] -= eps
        num.view(-1)[i] = (total(xp) - total(xm)) / (2 * eps)
    assert pu.rel_err(x.grad, num) < 1e-4          # 1e-6 guard in the soft backward is the only approximation


def test_recon_data_identity():
    """recon_data(x, x) = 0 + (1 - sum(m^2)/sum(2m - m^2)) (SURVEY 8c invariant 3)."""
    torch.manual_seed(3)
    x = torch.rand(3, 4, 16, 12)
    orc = ref_pipeline.OracleRender(np.zeros((1, 3), dtype=np.int64), np.zeros((1, 3, 2), dtype=np.float32), 12, 4 / 3, 1.0)
    got = orc.recon_data(x, x, contour=0.1)
    m = x[:, 3].reshape(3, -1)
    want = 1 - ((m * m).sum(1) / ((2 * m - m * m).sum(1) + 1e-10)).mean()
    assert abs(float(got) - float(want)) < 1e-6


def test_full_pipeline_gradients_match_f64_central_differences(mm):
    """the oracle's analytic backward (autograd glue + the C DIB-R backward kernels) against fp64 central
    differences of loss = recon_data(render(A), gt) on EVERY input of the path: vertices, the five camera scalars, texture,
    lights, background.  The loss is piecewise smooth (visibility is discrete), so a probe that straddles a visibility change
    is allowed to disagree: at most one of the probes per input may."""
    dt = torch.float64
    dr = mm.DiffRender(mm.icosphere(1), 16, image_weight=1.0)
    orc = pu.oracle_for(dr, dtype=dt)
    A = {k: v.to(dt) for k, v in pu.make_attributes(dr.vertices_init, 1, 16, 16, 5, dist_range=(2.5, 3.0)).items()}
    G = {k: v.to(dt) for k, v in pu.make_attributes(dr.vertices_init, 1, 16, 16, 6, dist_range=(2.5, 3.0)).items()}
    with torch.no_grad():
        gt = orc.render(no_mask=True, **G)[0]

    def loss_of(Ax):
        return orc.recon_data(orc.render(no_mask=True, **Ax)[0], gt, no_mask=True, contour=0.1)

    keys = ['vertices', 'azimuths', 'elevations', 'distances', 'biases', 'textures', 'lights', 'bg']
    Ag = {k: (v.clone().requires_grad_(k in keys)) for k, v in A.items()}
    loss_of(Ag).backward()
    gen = torch.Generator().manual_seed(0)
    for k in keys:
        g = Ag[k].grad.reshape(-1)
        n = g.numel()
        # probe where the analytic gradient is largest (informative) plus random entries
        cand = torch.unique(torch.cat([g.abs().topk(min(4, n)).indices, torch.randint(0, n, (4,), generator=gen)]))
        eps = 1e-6 if k not in ('azimuths', 'elevations') else 1e-5          # degrees
        bad = 0
        for i in cand.tolist():
            vals = []
            for sgn in (1.0, -1.0):
                Ax = {kk: vv.clone() for kk, vv in A.items()}
                Ax[k].view(-1)[i] += sgn * eps
                with torch.no_grad():
                    vals.append(float(loss_of(Ax)))
            fd = (vals[0] - vals[1]) / (2 * eps)
            if abs(fd - float(g[i])) > 2e-5 * max(1e-3, abs(fd), abs(float(g[i]))) + 1e-9:      # 2e-5 relative
                bad += 1
        assert bad <= 1 and float(g.abs().max()) > 0, (k, bad, len(cand))
