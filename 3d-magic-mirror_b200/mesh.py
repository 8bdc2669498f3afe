"""Template-mesh setup: everything DiffRender.__init__ derives once from the OBJ
template (reference: networks.py:165-256 and the kaolin helpers it calls --
io.obj.import_mesh :176, ops.mesh.index_vertices_by_faces :201,
ops.mesh.uniform_laplacian :249).  Host-side, runs once per process; numpy.
"""
import numpy as np
import torch


class TemplateMesh(object):
    """Raw OBJ content: vertices (V,3) f32, faces (F,3) i64, uvs (VT,2) f32, face_uvs_idx (F,3) i64."""

    def __init__(self, vertices, faces, uvs, face_uvs_idx):
        self.vertices, self.faces, self.uvs, self.face_uvs_idx = vertices, faces, uvs, face_uvs_idx


def load_obj(path):
    """Minimal OBJ reader for the template meshes: `v`, `vt`, triangular `f a/b` or
    `f a/b/c`; `vn`, `mtllib`, `usemtl`, comments are ignored."""
    verts, uvs, fv, ft = [], [], [], []
    with open(path, "r") as fh:
        for raw in fh:
            tok = raw.strip().split()
            if not tok or tok[0].startswith("#"):
                continue
            key = tok[0]
            if key == "v":
                verts.append((float(tok[1]), float(tok[2]), float(tok[3])))
            elif key == "vt":
                uvs.append((float(tok[1]), float(tok[2])))
            elif key == "f":
                if len(tok) != 4:
                    raise ValueError("%s: only triangular faces are supported (got %d corners)" % (path, len(tok) - 1))
                corner = [t.split("/") for t in tok[1:]]
                fv.append([int(c[0]) - 1 for c in corner])
                if len(corner[0]) >= 2 and corner[0][1]:
                    ft.append([int(c[1]) - 1 for c in corner])
    if not verts or not fv:
        raise ValueError("%s: no geometry found" % path)
    if len(ft) != len(fv):
        raise ValueError("%s: every face needs texture-coordinate indices (f v/vt ...)" % path)
    return TemplateMesh(torch.tensor(verts, dtype=torch.float32), torch.tensor(fv, dtype=torch.long),
                        torch.tensor(uvs, dtype=torch.float32), torch.tensor(ft, dtype=torch.long))


def save_obj(path, vertices, faces, uvs=None, face_uvs_idx=None):
    """Writes a template back out (used by tests to materialise golden meshes as .obj)."""
    with open(path, "w") as fh:
        for v in np.asarray(vertices, dtype=np.float64):
            fh.write("v %.9g %.9g %.9g\n" % (v[0], v[1], v[2]))
        if uvs is not None:
            for t in np.asarray(uvs, dtype=np.float64):
                fh.write("vt %.9g %.9g\n" % (t[0], t[1]))
        fa = np.asarray(faces)
        if uvs is not None and face_uvs_idx is not None:
            ta = np.asarray(face_uvs_idx)
            for f, t in zip(fa, ta):
                fh.write("f %d/%d %d/%d %d/%d\n" % (f[0] + 1, t[0] + 1, f[1] + 1, t[1] + 1, f[2] + 1, t[2] + 1))
        else:
            for f in fa:
                fh.write("f %d %d %d\n" % (f[0] + 1, f[1] + 1, f[2] + 1))


def normalise_template(vertices, init_ellipsoid=1):
    """networks.py:181-194: per-axis min-max to [-1,1]; z/2 (unless init_ellipsoid == -1);
    x and z divided again by init_ellipsoid when != 1; everything * 0.9."""
    v = vertices.clone().to(torch.float32)
    vmax = v.max(0, True)[0]
    vmin = v.min(0, True)[0]
    v = (v - vmin) / (vmax - vmin)
    v = v * 2.0 - 1.0
    if not init_ellipsoid == -1:
        v[:, 2] = v[:, 2] / 2
        if init_ellipsoid != 1:
            v[:, 0] = v[:, 0] / init_ellipsoid
            v[:, 2] = v[:, 2] / init_ellipsoid
    v *= 0.9
    return v


def mirror_index(vertices_init):
    """networks.py:215-217: for each vertex, index of the vertex nearest to its z-mirror image."""
    a = vertices_init.to(torch.float64)
    m = a.clone()
    m[:, 2] *= -1
    d2 = ((a[:, None, :] - m[None, :, :]) ** 2).sum(-1)
    return d2.argmin(dim=1)


def edge_tables(faces):
    """networks.py:220-246: unique undirected edges (lexicographically sorted) and, per edge,
    the first two faces that own it, in the order (corner-pair index, face index)."""
    f = faces.cpu().numpy().astype(np.int64)
    F = f.shape[0]
    pairs = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
    pairs.sort(axis=1)
    owner = np.tile(np.arange(F, dtype=np.int64), 3)
    edges, inverse = np.unique(pairs, axis=0, return_inverse=True)
    inverse = inverse.reshape(-1)
    edge2faces = np.zeros((edges.shape[0], 2), dtype=np.int64)
    seen = np.zeros(edges.shape[0], dtype=np.int64)
    for pos in range(pairs.shape[0]):
        e = inverse[pos]
        k = seen[e]
        if k < 2:
            edge2faces[e, k] = owner[pos]
        elif k >= 2:
            raise ValueError("non-manifold edge shared by more than two faces")
        seen[e] = k + 1
    return torch.from_numpy(edges), torch.from_numpy(edge2faces)


def uniform_laplacian(num_vertices, faces):
    """kaolin.ops.mesh.uniform_laplacian (networks.py:249): L = A/deg, diag = -1, NaN -> 0 (dense V x V)."""
    f = faces.cpu().numpy().astype(np.int64)
    adj = np.zeros((num_vertices, num_vertices), dtype=np.float32)
    for a, b in ((0, 1), (1, 2), (2, 0)):
        adj[f[:, a], f[:, b]] = 1.0
        adj[f[:, b], f[:, a]] = 1.0
    deg = adj.sum(axis=1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        L = adj / deg
    np.fill_diagonal(L, -1.0)
    L[np.isnan(L)] = 0.0
    return torch.from_numpy(L)


def icosphere(level=3):
    """Procedural unit icosphere with per-corner spherical UVs; level 3 gives V=642, F=1280
    (the size of template/sphere.obj).  Lets the package run without any OBJ asset."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    verts = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2),
             (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5),
             (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(level):
        cache, nf = {}, []

        def mid(a, b):
            key = (a, b) if a < b else (b, a)
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]
        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = nf
    V = np.array(verts, dtype=np.float32) * 0.5
    Fa = np.array(faces, dtype=np.int64)
    # per-corner UVs (3 unique per face) so that seams need no vertex duplication
    uv = np.zeros((Fa.shape[0] * 3, 2), dtype=np.float32)
    for i, tri in enumerate(Fa):
        p = V[tri].astype(np.float64)
        u = np.arctan2(p[:, 0], p[:, 2]) / (2 * np.pi) + 0.5
        w = np.arcsin(np.clip(p[:, 1] / 0.5, -1, 1)) / np.pi + 0.5
        if u.max() - u.min() > 0.5:          # face straddles the seam
            u = np.where(u < 0.5, u + 1.0, u)
            if u.min() >= 1.0:
                u -= 1.0
            u = np.clip(u, 0.0, 1.0)
        uv[i * 3:(i + 1) * 3, 0] = u
        uv[i * 3:(i + 1) * 3, 1] = w
    ft = np.arange(Fa.shape[0] * 3, dtype=np.int64).reshape(-1, 3)
    return TemplateMesh(torch.from_numpy(V), torch.from_numpy(Fa), torch.from_numpy(uv), torch.from_numpy(ft))
