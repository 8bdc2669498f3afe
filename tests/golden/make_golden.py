"""Generates the committed golden fixtures under tests/golden/ IN THE BUILD CONTAINER
(needs /root/reference; the GPU box does not have it, it only reads the .npz files).

What is real reference code here and what is not:
  * camera.npz        -- outputs of the UNMODIFIED reference smr_utils.py
                         (camera_position_from_spherical_angles, generate_transformation_matrix).
  * setup_<mesh>.npz  -- attributes computed by the UNMODIFIED reference networks.DiffRender.__init__
                         (vertices_init, face_uvs, flip_index, edges, edge2faces, Laplacian) with the
                         kaolin calls it makes served by oracle/kaolin_shim.py.
  * render_<case>.npz -- inputs/outputs/gradients of the UNMODIFIED reference DiffRender.render +
                         recon_data Python code, again over the kaolin shim.  The DIB-R rasteriser
                         inside is our C restatement (parity unpinned, docs/DIBR_SPEC.md); everything
                         around it is the reference's own code executing.
  * templates/<mesh>.npz -- the reference template meshes (data, not code) parsed into arrays.
  * reg_<mesh>.npz    -- the UNMODIFIED reference's mesh regularisers (networks.py:392-491: calc_reg_loss, calc_reg_edge,
                         calc_reg_depth/R/C, calc_reg_deform, recon_flip) on seeded inputs: values and the gradients of a
                         fixed weighted sum.  Pure torch in the reference: no shim arithmetic involved.

Usage:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_import          # noqa: E402
import parity_utils as pu  # noqa: E402

REF = ref_import.REFERENCE_ROOT
TEMPLATES = {"sphere": 1, "ellipsoid": 1, "smpl_uv_642": 2, "sphere2": 1}     # name -> init_ellipsoid used in fixtures

RENDER_CASES = {
    # name: (mesh, image_size, ratio, init_ellipsoid, B, no_mask, contour, seed, dist_range)
    "sphere_64_masked": ("sphere", 64, 1, 1, 2, False, 0.0, 11, (2.0, 7.0)),
    "ellipsoid_64_bg_contour": ("ellipsoid", 64, 1, 1, 2, True, 0.1, 12, (2.0, 7.0)),
    "smpl_64x32_bg_contour": ("smpl_uv_642", 32, 2, 2, 2, True, 0.1, 13, (2.0, 6.0)),
    "sphere_cfg1": ("sphere", 64, 1, 1, 1, False, 0.0, 14, None),                  # BASELINE.json configs[0]
}


def cfg1_attributes(vertices_init, H, W, seed):
    """azim 30, elev 15, dist 4.5, bias 0; texture U[0,1]; undeformed template."""
    g = torch.Generator().manual_seed(seed)
    return {
        'azimuths': torch.tensor([30.0]), 'elevations': torch.tensor([15.0]), 'distances': torch.tensor([4.5]),
        'biases': torch.zeros(1, 2), 'vertices': vertices_init[None].clone(),
        'delta_vertices': torch.zeros(1, vertices_init.shape[0], 3),
        'textures': torch.rand(1, 3, 2 * H, W, generator=g), 'lights': torch.tensor([[3.0] + [0.0] * 8]),
        'bg': torch.rand(1, 3, H, W, generator=g),
    }


def cfg1_gt(H, W, seed):
    """GT mask = centred disc of radius 0.45*W; GT rgb U[0,1]."""
    g = torch.Generator().manual_seed(seed + 1)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    disc = (((xx + 0.5 - W / 2) ** 2 + (yy + 0.5 - H / 2) ** 2) <= (0.45 * W) ** 2).float()
    return torch.cat([torch.rand(1, 3, H, W, generator=g), disc[None, None]], dim=1)


REG_CASES = {"sphere": (1, 1), "smpl_uv_642": (2, 2)}          # mesh -> (ratio, init_ellipsoid)


def make_reg(net):
    for mesh, (ratio, ell) in REG_CASES.items():
        dr = net.DiffRender(os.path.join(REF, "template", mesh + ".obj"), 64, ratio=ratio, init_ellipsoid=ell)
        dr.sign_init = dr.sign_init.cpu()
        delta, fn = pu.reg_inputs(dr.num_vertices, dr.num_faces)
        vals, gd, gn = pu.reg_values(dr, delta, fn)
        np.savez_compressed(os.path.join(HERE, "reg_%s.npz" % mesh), ratio=ratio, init_ellipsoid=ell, values=vals.numpy(),
                            grad_delta=gd.numpy(), grad_face_normals=gn.numpy())
        print("reg", mesh, vals.numpy())


def main():
    net, smr = ref_import.import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "reg":
        return make_reg(net)
    mm = pu.load_mm()
    os.makedirs(os.path.join(HERE, "templates"), exist_ok=True)
    make_reg(net)

    # ---- templates
    for name in TEMPLATES:
        tm = mm.load_obj(os.path.join(REF, "template", name + ".obj"))
        np.savez_compressed(os.path.join(HERE, "templates", name + ".npz"), vertices=tm.vertices.numpy(),
                            faces=tm.faces.numpy().astype(np.int32), uvs=tm.uvs.numpy(),
                            face_uvs_idx=tm.face_uvs_idx.numpy().astype(np.int32))

    # ---- camera helpers (real reference code, no shim involved)
    g = torch.Generator().manual_seed(2024)
    n = 64
    dist = 2 + 5 * torch.rand(n, generator=g)
    elev = -30 + 90 * torch.rand(n, generator=g)
    azim = -180 + 360 * torch.rand(n, generator=g)
    look = torch.cat([torch.rand(n, 2, generator=g) - 0.5, torch.zeros(n, 1)], dim=1)
    up = torch.tensor([[0., 1., 0.]]).repeat(n, 1)
    pos = smr.camera_position_from_spherical_angles(dist, elev, azim, degrees=True)
    T = smr.generate_transformation_matrix(pos, look, up)
    np.savez_compressed(os.path.join(HERE, "camera.npz"), dist=dist.numpy(), elev=elev.numpy(), azim=azim.numpy(),
                        look_at=look.numpy(), up=up.numpy(), camera_position=pos.numpy(), transform=T.numpy())

    # ---- DiffRender.__init__ products (reference code over the shim)
    for name, ell in TEMPLATES.items():
        dr = net.DiffRender(os.path.join(REF, "template", name + ".obj"), 64, ratio=1, init_ellipsoid=ell)
        lap = dr.vertices_laplacian_matrix
        nz = lap.nonzero()
        np.savez_compressed(os.path.join(HERE, "setup_%s.npz" % name), init_ellipsoid=ell,
                            vertices_init=dr.vertices_init.numpy(), faces=dr.faces.numpy().astype(np.int32),
                            face_uvs=dr.face_uvs.numpy(), flip_index=dr.flip_index.numpy().astype(np.int32),
                            edges=dr.edges.numpy().astype(np.int32), edge2faces=dr.edge2faces.numpy().astype(np.int32),
                            lap_idx=nz.numpy().astype(np.int32), lap_val=lap[nz[:, 0], nz[:, 1]].numpy(),
                            cam_proj=dr.cam_proj.numpy(), sign_init=dr.sign_init.numpy())

    # ---- render + recon_data (reference Python over the shim)
    for case, (mesh, S, ratio, ell, B, no_mask, contour, seed, dist_range) in RENDER_CASES.items():
        dr = net.DiffRender(os.path.join(REF, "template", mesh + ".obj"), S, ratio=ratio, init_ellipsoid=ell,
                            image_weight=1.0)
        H, W = round(ratio * S), S
        if dist_range is None:
            A = cfg1_attributes(dr.vertices_init, H, W, seed)
            gt = cfg1_gt(H, W, seed)
        else:
            A = pu.make_attributes(dr.vertices_init, B, H, W, seed, dist_range=dist_range)
            with torch.no_grad():
                gt, _ = dr.render(no_mask=no_mask, **pu.make_attributes(dr.vertices_init, B, H, W, seed + 1000,
                                                                        dist_range=dist_range))
        Ag = pu.to_device(A, "cpu", requires_grad=True)
        rgbs, Aout = dr.render(no_mask=no_mask, **Ag)
        loss = dr.recon_data(rgbs, gt, no_mask=no_mask, contour=contour)
        loss.backward()
        out = {"in_" + k: v.numpy() for k, v in A.items()}
        out.update({"gt": gt.numpy(), "rgbs": rgbs.detach().numpy(), "loss": loss.detach().numpy(),
                    "face_normals": Aout['face_normals'].detach().numpy(),
                    "imnormal": Aout['imnormal'].detach().numpy()})
        for k in pu.GRAD_KEYS:
            if Ag[k].grad is not None:
                out["grad_" + k] = Ag[k].grad.numpy()
        out["meta"] = np.array([S, ratio, ell, B, int(no_mask), contour, seed], dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, "render_%s.npz" % case), **out)
        print(case, "loss", float(loss), "soft mean", float(rgbs[:, 3].mean()))


if __name__ == "__main__":
    main()
