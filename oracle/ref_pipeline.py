"""ref_pipeline -- TEST INFRASTRUCTURE ONLY.  **parity unpinned** for the Kaolin part.

CPU restatement of the reference's render-and-compare path on top of
oracle/kaolin_shim.py:
    OracleRender.render      follows /root/reference/networks.py:258-324
    OracleRender.recon_data  follows /root/reference/networks.py:364-390
    camera helpers           follow /root/reference/smr_utils.py:257-311
Plain torch ops on CPU (autograd provides every backward except the four DIB-R
kernels, which are the C restatement in dibr_oracle_impl.h).  The restated Python
glue is pinned against the UNMODIFIED reference `networks.DiffRender` run through
the same shim (tests/golden/make_golden.py, tests/test_oracle_golden.py); what
stays unpinned is Kaolin itself (docs/DIBR_SPEC.md).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this file.
"""
import math
import os
import sys

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
import kaolin_shim as kal   # noqa: E402


def camera_position_from_spherical_angles(dist, elev, azim, degrees=True):
    """smr_utils.py:257-281."""
    if degrees:
        elev = math.pi / 180.0 * elev
        azim = math.pi / 180.0 * azim
    x = dist * torch.cos(elev) * torch.sin(azim)
    y = dist * torch.sin(elev)
    z = dist * torch.cos(elev) * torch.cos(azim)
    return torch.stack([x, y, z], dim=1).reshape(-1, 3)


def generate_transformation_matrix(camera_position, look_at, camera_up_direction):
    """smr_utils.py:284-311."""
    z_axis = camera_position - look_at
    z_axis = z_axis / z_axis.norm(dim=1, keepdim=True)
    x_axis = torch.cross(camera_up_direction, z_axis, dim=1)
    x_axis = x_axis / x_axis.norm(dim=1, keepdim=True)
    y_axis = torch.cross(z_axis, x_axis, dim=1)
    rot_part = torch.stack([x_axis, y_axis, z_axis], dim=2)
    trans_part = -camera_position.unsqueeze(1) @ rot_part
    return torch.cat([rot_part, trans_part], dim=1)


class OracleRender(object):
    """The hot path of networks.DiffRender restated for CPU; takes raw topology arrays so that it
    shares no code with the product package."""

    def __init__(self, faces, face_uvs, image_size, ratio=1, image_weight=0.1, dtype=torch.float32):
        self.faces = torch.as_tensor(faces).long()
        self.face_uvs = torch.as_tensor(face_uvs).to(dtype).reshape(1, -1, 3, 2)
        self.image_size = image_size
        self.ratio = ratio
        self.image_weight = image_weight
        self.dtype = dtype
        fovy = np.arctan(1.0 / 2.5) * 2                                  # networks.py:172
        self.cam_proj = kal.generate_perspective_projection(fovy, ratio=1 / ratio, dtype=dtype)   # :174

    # ---- staged pieces (also used one by one by the parity tests)
    def vertex_stage(self, A):
        B = A['azimuths'].shape[0]
        dt = self.dtype
        object_pos = torch.cat((A['biases'], torch.zeros(B, 1, dtype=dt)), dim=1)                  # :278
        camera_up = torch.tensor([[0., 1., 0.]], dtype=dt).repeat(B, 1)                            # :279
        camera_pos = camera_position_from_spherical_angles(A['distances'], A['elevations'], A['azimuths'])  # :281
        cam_transform = generate_transformation_matrix(camera_pos, object_pos, camera_up)          # :282
        fvc, fvi, fn = kal.prepare_vertices(vertices=A['vertices'], faces=self.faces, camera_proj=self.cam_proj,
                                            camera_transform=cam_transform)                        # :284-287
        return fvc, fvi, fn

    def shade(self, A, no_mask, texmask, texcoord, imnormal, soft_mask):
        texcolor = kal.texture_mapping(texcoord, A['textures'], mode='bilinear')                   # :305
        coef = kal.spherical_harmonic_lighting(imnormal, A['lights'])                              # :306
        if no_mask:                                                                                # :307-313
            bg = A['bg'].permute(0, 2, 3, 1)
            image = texcolor * texmask + bg * (1 - texmask)
            image = image * coef.unsqueeze(-1)
        else:
            image = texcolor * texmask * coef.unsqueeze(-1) + torch.ones_like(texcolor) * (1 - texmask)
        render_img = torch.clamp(image, 0, 1)                                                      # :314
        return torch.cat([render_img, soft_mask[..., None]], dim=-1).permute(0, 3, 1, 2)           # :316-317

    def render(self, no_mask=False, **A):
        """networks.py:258-324 -> (rgbs, face_normals, imnormal, face_idx)."""
        B = A['azimuths'].shape[0]
        H, W = round(self.ratio * self.image_size), self.image_size
        fvc, fvi, fn = self.vertex_stage(A)
        fn_unit = kal.face_normals(fvc, unit=True).unsqueeze(-2).repeat(1, 1, 3, 1)               # :289-290
        F = self.faces.shape[0]
        face_attributes = [torch.ones((B, F, 3, 1), dtype=self.dtype), self.face_uvs.repeat(B, 1, 1, 1), fn_unit]
        (texmask, texcoord, imnormal), soft_mask, face_idx = kal.dibr_rasterization(
            H, W, fvc[:, :, :, -1], fvi, face_attributes, fn[:, :, -1])                            # :297-299
        rgbs = self.shade(A, no_mask, texmask, texcoord, imnormal, soft_mask)
        return rgbs, fn, imnormal, face_idx

    def recon_data(self, pred_data, gt_data, no_mask=False, contour=0, return_parts=False):
        """networks.py:364-390."""
        F_ = torch.nn.functional
        pred_img, pred_mask = pred_data[:, :3], pred_data[:, 3]
        gt_img, gt_mask = gt_data[:, :3], gt_data[:, 3]
        gm = gt_mask.unsqueeze(1)
        gt_img = gt_img * gm + torch.ones_like(gt_img) * (1 - gm)                                  # :374
        pred_img = pred_img * gm + torch.ones_like(pred_img) * (1 - gm)                            # :375
        loss_image = torch.mean(torch.abs(pred_img - gt_img))                                      # :376
        loss_iou = kal.mask_iou(pred_mask, gt_mask)                                                # :377
        loss_mask = loss_iou
        loss_contour = torch.zeros((), dtype=pred_data.dtype)
        if contour > 0:                                                                            # :379-386
            n, h, w = gt_mask.shape
            up = lambda m: F_.interpolate(F_.interpolate(m, size=(h // 4, w // 4)), size=(h, w))   # noqa: E731
            gt_c = torch.abs(gm - up(gm))
            pr_c = torch.abs(pred_mask.unsqueeze(1) - up(pred_mask.unsqueeze(1)))
            loss_contour = torch.mean((pr_c - gt_c) ** 2)
            loss_mask = loss_mask + loss_contour * contour
        loss_data = self.image_weight * loss_image + 1.0 * loss_mask                               # :389
        if return_parts:
            return loss_data, (loss_image, loss_iou, loss_contour)
        return loss_data
