// mm_camera.cuh -- camera_position_from_spherical_angles (smr_utils.py:257-281) + generate_transformation_matrix
// (smr_utils.py:284-311) + the vertex transform of kaolin prepare_vertices: ONE statement of the arithmetic for every kernel
// that needs it (mm_vertex.cu forward and backward), so the face records are bit-identical whoever builds them.
#pragma once
#include "mm_common.cuh"

namespace {

struct Cam {
    float T[12];        // 4x3 row-major: rows 0-2 rotation (columns = x,y,z axes), row 3 translation
    float cam[3];       // camera position
    float zr[3], zl;    // cam - look_at, its norm
    float za[3];
    float xr[3], xl;    // cross(up, za), its norm
    float xa[3];
    float ya[3];
    float ce, se, ca, sa, d;
};

// DIBR_SPEC V.1-V.2.  Angles in degrees (smr_utils.py:274-276).
__device__ inline void camera_setup(float az_deg, float el_deg, float d, float bx, float by, Cam& c) {
    const float k = 3.14159265358979323846f / 180.0f;
    const float e = k * el_deg, a = k * az_deg;
    c.ce = cosf(e); c.se = sinf(e); c.ca = cosf(a); c.sa = sinf(a); c.d = d;
    c.cam[0] = d * c.ce * c.sa;
    c.cam[1] = d * c.se;
    c.cam[2] = d * c.ce * c.ca;
    c.zr[0] = c.cam[0] - bx; c.zr[1] = c.cam[1] - by; c.zr[2] = c.cam[2] - 0.0f;
    c.zl = sqrtf(c.zr[0] * c.zr[0] + c.zr[1] * c.zr[1] + c.zr[2] * c.zr[2]);
    for (int i = 0; i < 3; ++i) c.za[i] = c.zr[i] / c.zl;
    // cross((0,1,0), za) = (za_z, 0, -za_x)
    c.xr[0] = c.za[2]; c.xr[1] = 0.0f; c.xr[2] = -c.za[0];
    c.xl = sqrtf(c.xr[0] * c.xr[0] + c.xr[1] * c.xr[1] + c.xr[2] * c.xr[2]);
    for (int i = 0; i < 3; ++i) c.xa[i] = c.xr[i] / c.xl;
    // ya = cross(za, xa)
    c.ya[0] = c.za[1] * c.xa[2] - c.za[2] * c.xa[1];
    c.ya[1] = c.za[2] * c.xa[0] - c.za[0] * c.xa[2];
    c.ya[2] = c.za[0] * c.xa[1] - c.za[1] * c.xa[0];
    for (int i = 0; i < 3; ++i) { c.T[i * 3 + 0] = c.xa[i]; c.T[i * 3 + 1] = c.ya[i]; c.T[i * 3 + 2] = c.za[i]; }
    for (int j = 0; j < 3; ++j)
        c.T[9 + j] = (-c.cam[0]) * c.T[0 + j] + (-c.cam[1]) * c.T[3 + j] + (-c.cam[2]) * c.T[6 + j];
}

__device__ inline void transform_vertex(const float* T, float x, float y, float z, float& cx, float& cy, float& cz) {
    cx = x * T[0] + y * T[3] + z * T[6] + T[9];
    cy = x * T[1] + y * T[4] + z * T[7] + T[10];
    cz = x * T[2] + y * T[5] + z * T[8] + T[11];
}

// one vertex: camera space + kaolin perspective_camera ((x*px, y*py) / (z * -1)); returns the unscaled image-plane xy
__device__ inline void project_vertex(const float* T, float proj_x, float proj_y, float x, float y, float z,
                                      float& cx, float& cy, float& cz, float& xi, float& yi) {
    transform_vertex(T, x, y, z, cx, cy, cz);
    const float den = cz * -1.0f;
    xi = (cx * proj_x) / den;
    yi = (cy * proj_y) / den;
}

// one face from its three camera-space corners: unit normal, / (|n| + 1e-10) (kaolin face_normals(unit=True))
__device__ inline void face_normal(float ax, float ay, float az, float bx, float by, float bz, float cx, float cy, float cz,
                                   float& nx, float& ny, float& nz) {
    const float e0x = bx - ax, e0y = by - ay, e0z = bz - az;
    const float e1x = cx - ax, e1y = cy - ay, e1z = cz - az;
    nx = e0y * e1z - e0z * e1y;
    ny = e0z * e1x - e0x * e1z;
    nz = e0x * e1y - e0y * e1x;
    const float len = sqrtf(nx * nx + ny * ny + nz * nz);
    const float inv = len + 1e-10f;
    nx /= inv; ny /= inv; nz /= inv;
}

}  // namespace
