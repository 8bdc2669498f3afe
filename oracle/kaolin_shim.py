"""kaolin_shim -- TEST INFRASTRUCTURE ONLY.  **parity unpinned.**

A minimal CPU stand-in for the slice of NVIDIA Kaolin that the reference's hot
path touches (imports at /root/reference/networks.py:6-11; call sites
networks.py:174,176,201,249,285,289,297,305,306,377).  Kaolin itself is a
third-party dependency pinned to v0.12.0 / 0.18.0 (/root/reference/INSTALL.md:31-34,45)
that is neither vendored in /root/reference nor installable offline, so every
function here is a restatement of Kaolin's published algorithm (docs/DIBR_SPEC.md
lists each assumption).  The Python-level helpers are restated with plain torch
ops (autograd gives their backward); the four DIB-R CUDA kernels are restated in
C (oracle/dibr_oracle_impl.h) and wrapped in torch.autograd.Function.

`install()` registers this module tree as `kaolin` in sys.modules so that the
UNMODIFIED reference `networks.py` can be imported and its own
`DiffRender.render` / `recon_data` Python code run on CPU (used by
tests/golden/make_golden.py to generate golden fixtures in this container).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this file.
"""
import ctypes
import math
import os
import subprocess
import sys
import types

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libdibr_oracle.so")
_lib = None


def build_oracle(force=False):
    """Compile oracle/dibr_oracle.c -> oracle/_build/libdibr_oracle.so (gcc, OpenMP)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []),
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


# ---- variant switches (ORACLE ONLY; tools/dibr_sensitivity.py): bits of mmo_set_variant in dibr_oracle.c, plus the one
# assumption that lives at the Python level of Kaolin (valid faces: normal_z >= 0 vs > 0)
VARIANTS = {"bbox_closed": 1, "soft_bbox_closed": 2, "depth_ge": 4, "pixel_order": 8, "edge_bary": 16, "eps_zero": 32,
            "inside_strict": 64}
NZ_STRICT = False


def set_variant(name=None):
    """name: None (the spec), a key of VARIANTS, or 'nz_strict'."""
    global NZ_STRICT
    NZ_STRICT = (name == "nz_strict")
    lib().mmo_set_variant(VARIANTS.get(name, 0) if name else 0)


def _real(dtype):
    if dtype == torch.float32:
        return "f32", ctypes.c_float
    if dtype == torch.float64:
        return "f64", ctypes.c_double
    raise TypeError(dtype)


# ----------------------------------------------------------------------------
# kaolin.render.camera
# ----------------------------------------------------------------------------
def generate_perspective_projection(fovyangle, ratio=1.0, dtype=torch.float):
    """[K-recall] kaolin.render.camera.generate_perspective_projection -> (3,1)."""
    tanfov = math.tan(fovyangle / 2.0)
    return torch.tensor([[1.0 / (ratio * tanfov)], [1.0 / tanfov], [-1]], dtype=dtype)


def perspective_camera(points, camera_proj):
    """[K-recall] kaolin.render.camera.perspective_camera."""
    projected_points = points * camera_proj.transpose(0, 1)
    return projected_points[..., :2] / projected_points[..., 2:3]


# ----------------------------------------------------------------------------
# kaolin.ops.mesh
# ----------------------------------------------------------------------------
def index_vertices_by_faces(vertices_features, faces):
    """[K-recall] (B,V,C),(F,3) -> (B,F,3,C)."""
    return vertices_features[:, faces]


def face_normals(face_vertices, unit=False):
    """[K-recall] kaolin.ops.mesh.face_normals: cross(v1-v0, v2-v0), / (|n| + 1e-10)."""
    e0 = face_vertices[:, :, 1] - face_vertices[:, :, 0]
    e1 = face_vertices[:, :, 2] - face_vertices[:, :, 0]
    n = torch.cross(e0, e1, dim=2)
    if unit:
        n = n / (n.norm(dim=2, keepdim=True) + 1e-10)
    return n


def adjacency_matrix(num_vertices, faces, sparse=False):
    """[K-recall] symmetric binary vertex adjacency (dense)."""
    f = faces.long()
    i = torch.cat([f[:, 0], f[:, 1], f[:, 2], f[:, 1], f[:, 2], f[:, 0]])
    j = torch.cat([f[:, 1], f[:, 2], f[:, 0], f[:, 0], f[:, 1], f[:, 2]])
    adj = torch.zeros(num_vertices, num_vertices, dtype=torch.float32)
    adj[i, j] = 1.0
    return adj


def uniform_laplacian(num_vertices, faces):
    """[K-recall] kaolin.ops.mesh.uniform_laplacian: A / deg with -1 on the diagonal."""
    adj = adjacency_matrix(num_vertices, faces)
    deg = adj.sum(dim=1).view(-1, 1)
    L = adj / deg
    torch.diagonal(L)[:] = -1
    L[torch.isnan(L)] = 0
    return L


# ----------------------------------------------------------------------------
# kaolin.io.obj
# ----------------------------------------------------------------------------
class _Mesh(object):
    def __init__(self, vertices, faces, uvs, face_uvs_idx):
        self.vertices, self.faces, self.uvs, self.face_uvs_idx = vertices, faces, uvs, face_uvs_idx
        self.materials = None


def import_mesh(path, with_materials=False, with_normals=False, **_):
    """[K-recall] kaolin.io.obj.import_mesh: v / vt / f (a, a/b, a/b/c), triangles only."""
    v, vt, fi, fti = [], [], [], []
    with open(path, "r") as fh:
        for line in fh:
            d = line.split()
            if not d:
                continue
            if d[0] == "v":
                v.append([float(x) for x in d[1:4]])
            elif d[0] == "vt":
                vt.append([float(x) for x in d[1:3]])
            elif d[0] == "f":
                tri = [c.split("/") for c in d[1:]]
                assert len(tri) == 3, "oracle OBJ reader: triangles only"
                fi.append([int(c[0]) for c in tri])
                if len(tri[0]) > 1 and tri[0][1] != "":
                    fti.append([int(c[1]) for c in tri])
    vertices = torch.tensor(v, dtype=torch.float32).view(-1, 3)
    faces = torch.tensor(fi, dtype=torch.long).view(-1, 3) - 1
    uvs = torch.tensor(vt, dtype=torch.float32).view(-1, 2)
    face_uvs_idx = torch.tensor(fti, dtype=torch.long).view(-1, 3) - 1
    return _Mesh(vertices, faces, uvs, face_uvs_idx)


# ----------------------------------------------------------------------------
# kaolin.render.mesh
# ----------------------------------------------------------------------------
def prepare_vertices(vertices, faces, camera_proj, camera_rot=None, camera_trans=None,
                     camera_transform=None):
    """[K-recall] kaolin.render.mesh.prepare_vertices (camera_transform form only)."""
    assert camera_transform is not None and camera_transform.shape[1:] == (4, 3)
    padded = torch.nn.functional.pad(vertices, (0, 1), mode="constant", value=1.0)
    vertices_camera = padded @ camera_transform
    vertices_image = perspective_camera(vertices_camera, camera_proj)
    fvc = index_vertices_by_faces(vertices_camera, faces)
    fvi = index_vertices_by_faces(vertices_image, faces)
    fn = face_normals(fvc, unit=True)
    return fvc, fvi, fn


class _RasterizeOracle(torch.autograd.Function):
    """[K-recall] kaolin RasterizeCuda (packed_rasterize_forward / rasterize_backward)."""

    @staticmethod
    def forward(ctx, height, width, fvz, fvi, feat, valid, multiplier, eps):
        B, F = fvz.shape[:2]
        D = feat.shape[-1]
        fvz, fvi, feat = fvz.contiguous(), fvi.contiguous(), feat.contiguous()
        valid_u8 = valid.to(torch.uint8).contiguous()
        sfx, creal = _real(fvi.dtype)
        face_idx = torch.empty(B, height, width, dtype=torch.long)
        weights = torch.empty(B, height, width, 3, dtype=fvi.dtype)
        interp = torch.empty(B, height, width, D, dtype=fvi.dtype)
        getattr(lib(), "mmo_rasterize_forward_" + sfx)(
            B, height, width, F, D, _p(fvz), _p(fvi), _p(feat), _p(valid_u8),
            creal(multiplier), creal(eps), _p(face_idx), _p(weights), _p(interp))
        ctx.save_for_backward(face_idx, weights, fvi, feat)
        ctx.cfg = (height, width, multiplier, eps)
        ctx.mark_non_differentiable(face_idx)
        return interp, face_idx

    @staticmethod
    def backward(ctx, g_interp, _g_idx):
        face_idx, weights, fvi, feat = ctx.saved_tensors
        height, width, multiplier, eps = ctx.cfg
        B, F = fvi.shape[:2]
        D = feat.shape[-1]
        sfx, creal = _real(fvi.dtype)
        g_fvi = torch.zeros_like(fvi)
        g_feat = torch.zeros_like(feat)
        getattr(lib(), "mmo_rasterize_backward_" + sfx)(
            B, height, width, F, D, _p(g_interp.contiguous()), _p(face_idx), _p(weights),
            _p(fvi), _p(feat), creal(multiplier), creal(eps), _p(g_fvi), _p(g_feat))
        return None, None, None, g_fvi, g_feat, None, None, None


class _SoftMaskOracle(torch.autograd.Function):
    """[K-recall] kaolin DibrSoftMaskCuda (dibr_soft_mask_forward / backward)."""

    @staticmethod
    def forward(ctx, fvi, face_idx, sigmainv, boxlen, knum, multiplier):
        B, F = fvi.shape[:2]
        H, W = face_idx.shape[1:]
        fvi = fvi.contiguous()
        sfx, creal = _real(fvi.dtype)
        soft = torch.empty(B, H, W, dtype=fvi.dtype)
        cprob = torch.empty(B, H, W, knum, dtype=fvi.dtype)
        cidx = torch.empty(B, H, W, knum, dtype=torch.long)
        ctype = torch.empty(B, H, W, knum, dtype=torch.uint8)
        getattr(lib(), "mmo_soft_mask_forward_" + sfx)(
            B, H, W, F, knum, _p(fvi), _p(face_idx), creal(sigmainv), creal(boxlen),
            creal(multiplier), _p(soft), _p(cprob), _p(cidx), _p(ctype))
        ctx.save_for_backward(soft, face_idx, cprob, cidx, ctype, fvi)
        ctx.cfg = (sigmainv, knum, multiplier)
        return soft

    @staticmethod
    def backward(ctx, g_soft):
        soft, face_idx, cprob, cidx, ctype, fvi = ctx.saved_tensors
        sigmainv, knum, multiplier = ctx.cfg
        B, F = fvi.shape[:2]
        H, W = face_idx.shape[1:]
        sfx, creal = _real(fvi.dtype)
        g_fvi = torch.zeros_like(fvi)
        getattr(lib(), "mmo_soft_mask_backward_" + sfx)(
            B, H, W, F, knum, _p(g_soft.contiguous()), _p(soft), _p(face_idx), _p(cprob),
            _p(cidx), _p(ctype), _p(fvi), creal(sigmainv), creal(multiplier), _p(g_fvi))
        return g_fvi, None, None, None, None, None


def rasterize(height, width, face_vertices_z, face_vertices_image, face_features,
              valid_faces=None, multiplier=None, eps=None, backend="cuda"):
    multiplier = 1000 if multiplier is None else multiplier
    eps = 1e-8 if eps is None else eps
    if valid_faces is None:
        valid_faces = torch.ones(face_vertices_z.shape[:2], dtype=torch.bool)
    return _RasterizeOracle.apply(height, width, face_vertices_z, face_vertices_image,
                                  face_features, valid_faces, float(multiplier), float(eps))


def dibr_soft_mask(face_vertices_image, selected_face_idx, sigmainv=7000, boxlen=0.02,
                   knum=30, multiplier=1000.):
    return _SoftMaskOracle.apply(face_vertices_image, selected_face_idx, float(sigmainv),
                                 float(boxlen), int(knum), float(multiplier))


def dibr_rasterization(height, width, face_vertices_z, face_vertices_image, face_features,
                       face_normals_z, sigmainv=7000, boxlen=0.02, knum=30, multiplier=None,
                       eps=None, rast_backend="cuda"):
    """[K-recall] kaolin.render.mesh.dibr_rasterization (call site networks.py:297-299)."""
    multiplier = 1000 if multiplier is None else multiplier
    eps = 1e-8 if eps is None else eps
    is_list = isinstance(face_features, (list, tuple))
    if is_list:
        dims = [f.shape[-1] for f in face_features]
        feats = torch.cat(list(face_features), dim=-1)
    else:
        feats = face_features
    valid_faces = (face_normals_z > 0) if NZ_STRICT else (face_normals_z >= 0)
    interp, face_idx = rasterize(height, width, face_vertices_z, face_vertices_image, feats,
                                 valid_faces, multiplier, eps)
    if is_list:
        interp = tuple(torch.split(interp, dims, dim=-1))
    soft_mask = dibr_soft_mask(face_vertices_image, face_idx, sigmainv, boxlen, knum, multiplier)
    return interp, soft_mask, face_idx


def texture_mapping(texture_coordinates, texture_maps, mode="nearest"):
    """[K-recall] kaolin.render.mesh.texture_mapping: grid_sample with v flipped,
    align_corners=False, padding_mode='border'."""
    batch_size = texture_coordinates.shape[0]
    num_channels = texture_maps.shape[1]
    tc = texture_coordinates.reshape(batch_size, -1, 1, 2)
    tc = tc * 2.0 - 1.0
    tc = torch.stack([tc[..., 0], -tc[..., 1]], dim=-1)
    out = torch.nn.functional.grid_sample(texture_maps, tc, mode=mode, align_corners=False,
                                          padding_mode="border")
    out = out.permute(0, 2, 3, 1)
    return out.reshape(*texture_coordinates.shape[:-1], num_channels)


def spherical_harmonic_lighting(imnormal, lights):
    """[K-recall] kaolin.render.mesh.spherical_harmonic_lighting (9 bands)."""
    x, y, z = imnormal[..., 0], imnormal[..., 1], imnormal[..., 2]
    band0 = 0.28209479177 * torch.ones_like(x)
    band1_m1 = 0.4886025119 * x
    band1_0 = 0.4886025119 * z
    band1_p1 = 0.4886025119 * y
    band2_m2 = 1.09254843059 * (x * y)
    band2_m1 = 1.09254843059 * (y * z)
    band2_0 = 0.94617469575 * (z * z) - 0.31539156525
    band2_p1 = 0.77254840404 * (x * z)
    band2_p2 = 0.38627420202 * (x * x - y * y)
    bands = torch.stack([band0, band1_m1, band1_0, band1_p1, band2_m2, band2_m1, band2_0,
                         band2_p1, band2_p2], dim=-1)
    shape = [lights.shape[0]] + [1] * (imnormal.dim() - 2) + [9]
    return torch.sum(bands * lights.view(*shape), dim=-1)


# ----------------------------------------------------------------------------
# kaolin.metrics.render
# ----------------------------------------------------------------------------
def mask_iou(lhs_mask, rhs_mask):
    """[K-recall] kaolin.metrics.render.mask_iou = 1 - mean_b(sum(l*r) / (sum(l+r-l*r) + 1e-10))."""
    batch_size = lhs_mask.shape[0]
    sil_mul = lhs_mask * rhs_mask
    sil_add = lhs_mask + rhs_mask
    iou_up = torch.sum(sil_mul.reshape(batch_size, -1), dim=1)
    iou_down = torch.sum((sil_add - sil_mul).reshape(batch_size, -1), dim=1)
    iou_neg = iou_up / (iou_down + 1e-10)
    return 1.0 - torch.mean(iou_neg)


# ----------------------------------------------------------------------------
# sys.modules registration so the unmodified reference can `import kaolin`
# ----------------------------------------------------------------------------
def install():
    """Register `kaolin` (+ the submodules networks.py names) backed by this shim."""
    if "kaolin" in sys.modules and getattr(sys.modules["kaolin"], "__mm_shim__", False):
        return sys.modules["kaolin"]

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    camera = mod("kaolin.render.camera", generate_perspective_projection=generate_perspective_projection,
                 perspective_camera=perspective_camera)
    rmesh = mod("kaolin.render.mesh", dibr_rasterization=dibr_rasterization, texture_mapping=texture_mapping,
                spherical_harmonic_lighting=spherical_harmonic_lighting, prepare_vertices=prepare_vertices,
                rasterize=rasterize, dibr_soft_mask=dibr_soft_mask)
    render = mod("kaolin.render", camera=camera, mesh=rmesh)
    omesh = mod("kaolin.ops.mesh", index_vertices_by_faces=index_vertices_by_faces, face_normals=face_normals,
                uniform_laplacian=uniform_laplacian, adjacency_matrix=adjacency_matrix)
    ops = mod("kaolin.ops", mesh=omesh)
    obj = mod("kaolin.io.obj", import_mesh=import_mesh)
    io = mod("kaolin.io", obj=obj)
    mrender = mod("kaolin.metrics.render", mask_iou=mask_iou)
    metrics = mod("kaolin.metrics", render=mrender)
    return mod("kaolin", render=render, ops=ops, io=io, metrics=metrics, __mm_shim__=True,
               __version__="0.0-mm-oracle-shim")
