// mm_vertex.cu -- vertex stage of the render path, forward and backward.
//
// Forward fuses, per image: camera_position_from_spherical_angles
// (smr_utils.py:257-281), generate_transformation_matrix (smr_utils.py:284-311),
// kaolin prepare_vertices (networks.py:284-287: [v,1]*T, perspective divide, gather
// by faces, unit face normals) and the second face_normals call (networks.py:289),
// and emits the 48-byte face records the geometry kernels read.
// The reference runs ~25 tiny kernels + a cuBLAS matmul for this.
//
// Backward consumes the per-face accumulators (d/d fvi, d/d unit normal) produced by
// the raster backward plus the optional upstream gradient of `face_normals`, and
// chains through normals, projection, transform and the look-at construction down
// to vertices, azimuth, elevation, distance and bias.
#include "mm_device.cuh"
#include "mm_camera.cuh"
#include <math_constants.h>
#include <cooperative_groups.h>
#include <mutex>

namespace {

// ------------------------------------------------------------------ forward
// grid = (nchunks, B): every CTA of an image recomputes the (tiny) vertex transform into shared memory and emits the
// face records / normals of its contiguous 1/nchunks share of the faces.
struct VertexFwdParams {
    int V, F, nchunks;
    float proj_x, proj_y, multiplier;
    // buffers this kernel clears for the rest of the step (16-byte units): visibility buffer + silhouette accumulators +
    // coverage bitmap + counters, and (fused step) the texture-gradient output.  Folding the clears in here removes two
    // memset nodes (~31 MB at cfg-2) from the head of the dependency chain: the stores drain while the CTAs do the
    // latency-bound camera / vertex work.
    uint4* clr0; size_t n0;
    uint4* clr1; size_t n1;
    unsigned long long* prof;
    // the faces' exact pixel rectangles for the hard and the soft pass (face_rects; NULL: vertex stage alone)
    uint4* frect;
    int H, W;
    float sx, sy, blen;
};

__global__ void __launch_bounds__(MM_VTHREADS)
k_vertex_fwd(const VertexFwdParams q,
             const int32_t* __restrict__ faces, const float* __restrict__ vertices,
             const float* __restrict__ azim, const float* __restrict__ elev, const float* __restrict__ dist,
             const float* __restrict__ bias,
             float* __restrict__ frec, float* __restrict__ vimg, float* __restrict__ face_normals,
             float* __restrict__ gfacc_zero, long long* __restrict__ img_fwd, long long* __restrict__ img_bwd)
{
    // First kernel of a step: WAIT FIRST, then release the dependents.  Nothing of this library can then run before the work
    // that precedes the step in the stream (the producer of vertices / camera scalars / textures) is complete and visible --
    // which is what lets k_vertex_bwd read the caller's inputs ahead of its own wait.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ float sm[];
    const int V = q.V, F = q.F;
    float* sT = sm;                    // 12
    float* svc = sm + 16;              // V*3 camera-space
    float* svi = svc + (size_t)V * 3;  // V*2 image-plane (unscaled)
    const int chunk = blockIdx.x, b = blockIdx.y;
    MM_PROF_MARK(q.prof, 0, (b * gridDim.x + chunk) * (MM_VTHREADS / 32) + (threadIdx.x >> 5), 0);
    {
        const size_t nthreads = (size_t)gridDim.x * gridDim.y * blockDim.x;
        const size_t t = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (size_t i = t; i < q.n0; i += nthreads) q.clr0[i] = z;
        for (size_t i = t; i < q.n1; i += nthreads) q.clr1[i] = z;
    }
    // this thread's first face: its vertex indices are static data of the ctx -- fetched here, under the camera chain and the
    // vertex transform, instead of as one more round trip behind them
    const int per_chunk = (F + q.nchunks - 1) / q.nchunks;
    const int f_end = min(F, (chunk + 1) * per_chunk);
    const int f_first = chunk * per_chunk + threadIdx.x;
    int pi0 = 0, pi1 = 0, pi2 = 0;
    if (f_first < f_end) { pi0 = faces[f_first * 3]; pi1 = faces[f_first * 3 + 1]; pi2 = faces[f_first * 3 + 2]; }
    if (threadIdx.x == 0) {
        Cam c;
        camera_setup(azim[b], elev[b], dist[b], bias[b * 2], bias[b * 2 + 1], c);
        for (int i = 0; i < 12; ++i) sT[i] = c.T[i];
    }
    if (chunk == 0 && threadIdx.x < 16) {          // the per-image fixed-point accumulators start at zero
        if (threadIdx.x < 4) img_fwd[b * 4 + threadIdx.x] = 0;
        if (threadIdx.x < 12) img_bwd[b * 12 + threadIdx.x] = 0;
    }
    __syncthreads();
    const float* vb = vertices + (size_t)b * V * 3;
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        float cx, cy, cz, xi, yi;
        project_vertex(sT, q.proj_x, q.proj_y, vb[v * 3], vb[v * 3 + 1], vb[v * 3 + 2], cx, cy, cz, xi, yi);
        svc[v * 3] = cx; svc[v * 3 + 1] = cy; svc[v * 3 + 2] = cz;
        svi[v * 2] = xi; svi[v * 2 + 1] = yi;
        if (chunk == 0) {
            vimg[((size_t)b * V + v) * 2] = xi;
            vimg[((size_t)b * V + v) * 2 + 1] = yi;
        }
    }
    __syncthreads();
    float4* rec = reinterpret_cast<float4*>(frec + (size_t)b * F * MM_REC_FLOATS);
    for (int f = f_first; f < f_end; f += blockDim.x) {
        int i0 = pi0, i1 = pi1, i2 = pi2;            // (the first face's indices were fetched ahead of the two barriers)
        if (f != f_first) { i0 = faces[f * 3]; i1 = faces[f * 3 + 1]; i2 = faces[f * 3 + 2]; }
        const float ax = svc[i0 * 3], ay = svc[i0 * 3 + 1], az = svc[i0 * 3 + 2];
        const float bx = svc[i1 * 3], by = svc[i1 * 3 + 1], bz = svc[i1 * 3 + 2];
        const float cx = svc[i2 * 3], cy = svc[i2 * 3 + 1], cz = svc[i2 * 3 + 2];
        float nx, ny, nz;
        face_normal(ax, ay, az, bx, by, bz, cx, cy, cz, nx, ny, nz);
        // image-plane corners scaled by `multiplier` (DIBR_SPEC A.1), one rounding each
        rec[f * 3 + 0] = make_float4(__fmul_rn(svi[i0 * 2], q.multiplier), __fmul_rn(svi[i0 * 2 + 1], q.multiplier),
                                     __fmul_rn(svi[i1 * 2], q.multiplier), __fmul_rn(svi[i1 * 2 + 1], q.multiplier));
        rec[f * 3 + 1] = make_float4(__fmul_rn(svi[i2 * 2], q.multiplier), __fmul_rn(svi[i2 * 2 + 1], q.multiplier), az, bz);
        rec[f * 3 + 2] = make_float4(cz, nx, ny, nz);
        if (q.frect) {
            FaceRec r;
            r.ax = __fmul_rn(svi[i0 * 2], q.multiplier); r.ay = __fmul_rn(svi[i0 * 2 + 1], q.multiplier);
            r.bx = __fmul_rn(svi[i1 * 2], q.multiplier); r.by = __fmul_rn(svi[i1 * 2 + 1], q.multiplier);
            r.cx = __fmul_rn(svi[i2 * 2], q.multiplier); r.cy = __fmul_rn(svi[i2 * 2 + 1], q.multiplier);
            r.az = az; r.bz = bz; r.cz = cz; r.nx = nx; r.ny = ny; r.nz = nz;
            q.frect[(size_t)b * F + f] = face_rects(q, r);
        }
        if (face_normals) {
            float* fn = face_normals + ((size_t)b * F + f) * 3;
            fn[0] = nx; fn[1] = ny; fn[2] = nz;
        }
        if (gfacc_zero) {
            float4* g = reinterpret_cast<float4*>(gfacc_zero + ((size_t)b * F + f) * MM_GF);
            g[0] = g[1] = g[2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
    }
    MM_PROF_MARK(q.prof, 0, (b * gridDim.x + chunk) * (MM_VTHREADS / 32) + (threadIdx.x >> 5), 2);
}

// ------------------------------------------------------------------ backward
// One thread-block CLUSTER of VB_CLUSTER CTAs per image.  Every CTA scatters the per-face gradients of its quarter of the
// faces into its own shared-memory copy of the per-vertex accumulator; after a cluster barrier each CTA reduces its quarter of
// the VERTICES across the four copies through distributed shared memory, emits g_vertices and its partial camera sums; rank 0
// finishes the camera chain.  (The first version was one 512-thread CTA per image: 48 busy SMs and ~26 block barriers, 20 us.)
// The last CTA of image 0 also finalises the loss scalars of the fused step (was a separate single-CTA launch).
#ifndef VB_CLUSTER
#define VB_CLUSTER 4
#endif
#ifndef VB_THREADS
#define VB_THREADS 512       // (swept in round 2, cluster x threads: 4 x 384 0.0864 ms, 4 x 512 0.0856, 2 x 768 0.0856, 2 x 1024 0.0861,
#endif                       //  4 x 256 0.0873, 4 x 768 0.0917, 1 x 1024 0.0880, 8 x 256 0.0888)
#define VB_WARPS (VB_THREADS / 32)

struct VertexBwdParams {
    int B, V, F, H, W;
    float proj_x, proj_y;
    // loss finalisation (fused step only; loss == NULL otherwise)
    float* loss;
    const long long* img_fwd;
    float image_weight, contour;
    unsigned long long* prof;
};

__device__ __forceinline__ float fxv(const long long* a, double scale) { return (float)((double)(*a) / scale); }

__global__ void __cluster_dims__(VB_CLUSTER, 1, 1) __launch_bounds__(VB_THREADS)
k_vertex_bwd(const VertexBwdParams q,
             const int32_t* __restrict__ faces, const float* __restrict__ vertices,
             const float* __restrict__ azim, const float* __restrict__ elev, const float* __restrict__ dist,
             const float* __restrict__ bias,
             float* __restrict__ gfacc, const float* __restrict__ g_face_normals,
             long long* __restrict__ img_bwd,
             float* __restrict__ g_vertices, float* __restrict__ g_azim, float* __restrict__ g_elev,
             float* __restrict__ g_dist, float* __restrict__ g_bias, float* __restrict__ g_lights)
{
    // Programmatic dependent launch: everything up to the face loop reads only the CALLER's inputs (vertices, camera scalars),
    // which were complete before the step's first kernel ran (that kernel waited for them) -- so this prologue (two cold loads,
    // the camera chain, 642 transforms) runs while the previous kernel (k_soft_bwd) is still draining; the wait sits right in
    // front of the first read of what that kernel produced (gfacc, img_bwd).
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    MM_PROF_MARK(q.prof, 5, (blockIdx.y * gridDim.x + blockIdx.x) * VB_WARPS + (threadIdx.x >> 5), 0);
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ float sm[];
    __shared__ Cam sc;
    __shared__ float s_part[VB_WARPS][12];
    __shared__ float s_cl[VB_CLUSTER][12];          // rank 0's copy collects the partial camera sums of the cluster
    const int V = q.V, F = q.F;
    float* svc = sm;                      // V*3 camera-space positions
    float* sgv = sm + (size_t)V * 3;      // V*3 gradient w.r.t. camera-space positions (this CTA's faces only)
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* vb = vertices + (size_t)b * V * 3;
    // raw vertices -> shared while thread 0 builds the camera (the two latencies overlap), then transform in place
    for (int i = threadIdx.x; i < V * 3; i += blockDim.x) { svc[i] = vb[i]; sgv[i] = 0.0f; }
    if (threadIdx.x == 0) camera_setup(azim[b], elev[b], dist[b], bias[b * 2], bias[b * 2 + 1], sc);
    __syncthreads();
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        float cx, cy, cz;
        transform_vertex(sc.T, svc[v * 3], svc[v * 3 + 1], svc[v * 3 + 2], cx, cy, cz);
        svc[v * 3] = cx; svc[v * 3 + 1] = cy; svc[v * 3 + 2] = cz;
    }
    __syncthreads();
    // (this thread's first face: vertex indices -- static ctx data -- fetched ahead of the wait too)
    const int fper = (F + VB_CLUSTER - 1) / VB_CLUSTER;
    const int f_end = min(F, (rank + 1) * fper);
    const int f_first = rank * fper + threadIdx.x;
    int pidx[3] = {0, 0, 0};
    if (f_first < f_end) { pidx[0] = faces[f_first * 3]; pidx[1] = faces[f_first * 3 + 1]; pidx[2] = faces[f_first * 3 + 2]; }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    MM_PROF_MARK(q.prof, 5, (blockIdx.y * gridDim.x + blockIdx.x) * VB_WARPS + (threadIdx.x >> 5), 1);
    for (int f = f_first; f < f_end; f += blockDim.x) {
        int idx[3] = {pidx[0], pidx[1], pidx[2]};
        if (f != f_first) { idx[0] = faces[f * 3]; idx[1] = faces[f * 3 + 1]; idx[2] = faces[f * 3 + 2]; }
        float P[3][3];
        #pragma unroll
        for (int i = 0; i < 3; ++i) { P[i][0] = svc[idx[i] * 3]; P[i][1] = svc[idx[i] * 3 + 1]; P[i][2] = svc[idx[i] * 3 + 2]; }
        float ga[9];
        {
            // consume and clear: the accumulators are left zeroed for the next backward on this workspace (the vertex forward
            // zeroes them for the first one), so no call needs a memset node for them
            float4* gp = reinterpret_cast<float4*>(gfacc + ((size_t)b * F + f) * MM_GF);
            const float4 g0 = gp[0], g1 = gp[1], g2 = gp[2];
            {
                const float4 z4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (g0.x != 0.0f || g0.y != 0.0f || g0.z != 0.0f || g0.w != 0.0f) gp[0] = z4;
                if (g1.x != 0.0f || g1.y != 0.0f) gp[1] = z4;
                if (g2.x != 0.0f || g2.y != 0.0f || g2.z != 0.0f) gp[2] = z4;
            }
            ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g2.x; ga[7] = g2.y; ga[8] = g2.z;
            bool any = g_face_normals != nullptr;
            #pragma unroll
            for (int i = 0; i < 9; ++i) any = any || (ga[i] != 0.0f);
            if (!any) continue;          // most faces (back-facing, interior, off-screen) received no gradient at all
        }
        float G[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        // (1) image-plane gradient -> camera space: xi = -px*x/z, yi = -py*y/z.  One IEEE reciprocal per corner: the
        // 15 divisions of the direct form mostly had zero numerators, which take the ~60-instruction slow path of the
        // IEEE division (ncu: half of the kernel's instructions)
        #pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float gx = ga[i * 2], gy = ga[i * 2 + 1];
            const float rz = 1.0f / P[i][2];
            const float xi = -(P[i][0] * q.proj_x) * rz, yi = -(P[i][1] * q.proj_y) * rz;
            G[i][0] += -(gx * q.proj_x) * rz;
            G[i][1] += -(gy * q.proj_y) * rz;
            G[i][2] += -(gx * xi + gy * yi) * rz;
        }
        // (2) unit normal gradient (raster path + upstream face_normals gradient)
        float gu[3] = {ga[6], ga[7], ga[8]};
        if (g_face_normals) {
            const float* ge = g_face_normals + ((size_t)b * F + f) * 3;
            gu[0] += ge[0]; gu[1] += ge[1]; gu[2] += ge[2];
        }
        if (gu[0] != 0.0f || gu[1] != 0.0f || gu[2] != 0.0f) {
            const float e0[3] = {P[1][0] - P[0][0], P[1][1] - P[0][1], P[1][2] - P[0][2]};
            const float e1[3] = {P[2][0] - P[0][0], P[2][1] - P[0][1], P[2][2] - P[0][2]};
            const float n[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
            const float L = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            const float Le = L + 1e-10f;
            float gn[3] = {0, 0, 0};
            if (L > 0.0f) {
                const float dot = n[0] * gu[0] + n[1] * gu[1] + n[2] * gu[2];
                const float k = dot / (L * Le * Le);
                gn[0] = gu[0] / Le - n[0] * k; gn[1] = gu[1] / Le - n[1] * k; gn[2] = gu[2] / Le - n[2] * k;
            }
            // n = e0 x e1 :  g_e0 = e1 x gn ; g_e1 = gn x e0
            const float ge0[3] = {e1[1] * gn[2] - e1[2] * gn[1], e1[2] * gn[0] - e1[0] * gn[2], e1[0] * gn[1] - e1[1] * gn[0]};
            const float ge1[3] = {gn[1] * e0[2] - gn[2] * e0[1], gn[2] * e0[0] - gn[0] * e0[2], gn[0] * e0[1] - gn[1] * e0[0]};
            #pragma unroll
            for (int k2 = 0; k2 < 3; ++k2) { G[1][k2] += ge0[k2]; G[2][k2] += ge1[k2]; G[0][k2] -= ge0[k2] + ge1[k2]; }
        }
        #pragma unroll
        for (int i = 0; i < 3; ++i)
            #pragma unroll
            for (int k2 = 0; k2 < 3; ++k2)
                if (G[i][k2] != 0.0f) atomicAdd(&sgv[idx[i] * 3 + k2], G[i][k2]);
    }
    cluster.sync();
    // (3) vcam = v*R + t : g_v = g_vcam * R^T ; g_R[i][j] = sum v_i g_j ; g_t[j] = sum g_j.   This CTA's quarter of the
    // vertices, summed over the cluster's four accumulator copies (distributed shared memory, fixed order)
    const float* rsgv[VB_CLUSTER];
    #pragma unroll
    for (int r = 0; r < VB_CLUSTER; ++r) rsgv[r] = cluster.map_shared_rank(sgv, r);
    float acc[12];
    #pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = 0.0f;
    float* gvb = g_vertices + (size_t)b * V * 3;
    const int vper = (V + VB_CLUSTER - 1) / VB_CLUSTER;
    const int v_end = min(V, (rank + 1) * vper);
    for (int v = rank * vper + threadIdx.x; v < v_end; v += blockDim.x) {
        float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
        #pragma unroll
        for (int r = 0; r < VB_CLUSTER; ++r) { g0 += rsgv[r][v * 3]; g1 += rsgv[r][v * 3 + 1]; g2 += rsgv[r][v * 3 + 2]; }
        const float x = vb[v * 3], y = vb[v * 3 + 1], z = vb[v * 3 + 2];
        gvb[v * 3 + 0] = g0 * sc.T[0] + g1 * sc.T[1] + g2 * sc.T[2];
        gvb[v * 3 + 1] = g0 * sc.T[3] + g1 * sc.T[4] + g2 * sc.T[5];
        gvb[v * 3 + 2] = g0 * sc.T[6] + g1 * sc.T[7] + g2 * sc.T[8];
        acc[0] += x * g0; acc[1] += x * g1; acc[2] += x * g2;
        acc[3] += y * g0; acc[4] += y * g1; acc[5] += y * g2;
        acc[6] += z * g0; acc[7] += z * g1; acc[8] += z * g2;
        acc[9] += g0; acc[10] += g1; acc[11] += g2;
    }
    #pragma unroll
    for (int i = 0; i < 12; ++i) {
        float v = acc[i];
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float v = 0.0f;
        #pragma unroll
        for (int w = 0; w < VB_WARPS; ++w) v += s_part[w][threadIdx.x];
        float* dst = cluster.map_shared_rank(&s_cl[0][0], 0);
        dst[rank * 12 + threadIdx.x] = v;
    }
    cluster.sync();                        // partial sums visible to rank 0; nobody reads a remote sgv after this point
    if (rank == 0 && threadIdx.x == 0) {
        float gR[9], gt[3], gcam[3];
        for (int i = 0; i < 9; ++i) { gR[i] = 0.0f; for (int r = 0; r < VB_CLUSTER; ++r) gR[i] += s_cl[r][i]; }
        for (int j = 0; j < 3; ++j) { gt[j] = 0.0f; for (int r = 0; r < VB_CLUSTER; ++r) gt[j] += s_cl[r][9 + j]; }
        // t_j = -sum_i cam_i R_ij
        for (int i = 0; i < 3; ++i) {
            gcam[i] = -(gt[0] * sc.T[i * 3 + 0] + gt[1] * sc.T[i * 3 + 1] + gt[2] * sc.T[i * 3 + 2]);
            for (int j = 0; j < 3; ++j) gR[i * 3 + j] += -sc.cam[i] * gt[j];
        }
        float gxa[3], gya[3], gza[3];
        for (int i = 0; i < 3; ++i) { gxa[i] = gR[i * 3 + 0]; gya[i] = gR[i * 3 + 1]; gza[i] = gR[i * 3 + 2]; }
        // ya = za x xa : g_za += xa x g_ya ; g_xa += g_ya x za
        gza[0] += sc.xa[1] * gya[2] - sc.xa[2] * gya[1];
        gza[1] += sc.xa[2] * gya[0] - sc.xa[0] * gya[2];
        gza[2] += sc.xa[0] * gya[1] - sc.xa[1] * gya[0];
        gxa[0] += gya[1] * sc.za[2] - gya[2] * sc.za[1];
        gxa[1] += gya[2] * sc.za[0] - gya[0] * sc.za[2];
        gxa[2] += gya[0] * sc.za[1] - gya[1] * sc.za[0];
        // xa = xr/|xr|
        const float dx = sc.xa[0] * gxa[0] + sc.xa[1] * gxa[1] + sc.xa[2] * gxa[2];
        float gxr[3];
        for (int i = 0; i < 3; ++i) gxr[i] = (gxa[i] - sc.xa[i] * dx) / sc.xl;
        // xr = (za_z, 0, -za_x)
        gza[2] += gxr[0];
        gza[0] += -gxr[2];
        // za = zr/|zr|
        const float dz = sc.za[0] * gza[0] + sc.za[1] * gza[1] + sc.za[2] * gza[2];
        float gzr[3];
        for (int i = 0; i < 3; ++i) gzr[i] = (gza[i] - sc.za[i] * dz) / sc.zl;
        for (int i = 0; i < 3; ++i) gcam[i] += gzr[i];
        g_bias[b * 2] = -gzr[0];
        g_bias[b * 2 + 1] = -gzr[1];
        const float k = 3.14159265358979323846f / 180.0f;
        g_dist[b] = gcam[0] * sc.ce * sc.sa + gcam[1] * sc.se + gcam[2] * sc.ce * sc.ca;
        g_elev[b] = k * (gcam[0] * (-sc.d * sc.se * sc.sa) + gcam[1] * (sc.d * sc.ce) + gcam[2] * (-sc.d * sc.se * sc.ca));
        g_azim[b] = k * (gcam[0] * (sc.d * sc.ce * sc.ca) + gcam[2] * (-sc.d * sc.ce * sc.sa));
    }
    // (4) light gradient: per-image fixed-point sums accumulated by the shading backward
    if (rank == 1 && threadIdx.x < 9 && g_lights) {
        g_lights[b * 9 + threadIdx.x] = (float)((double)img_bwd[b * 12 + 1 + threadIdx.x] / 17592186044416.0);   // MM_FX_GRAD
        img_bwd[b * 12 + 1 + threadIdx.x] = 0;                  // consume and clear (entry 0, the contour sum, belongs to the
    }                                                           // loss finalisation below and to the next vertex forward)
    // (5) fused step: loss[0..3] = data, image, mask (1 - mean IoU), contour term (networks.py:364-390), one warp, fixed order
    if (q.loss && b == 0 && rank == VB_CLUSTER - 1 && warp == 0) {
        const double FXL = 1099511627776.0;                    // MM_FX_LOSS
        float a_l1 = 0.0f, a_c = 0.0f, a_iou = 0.0f;
        for (int i = lane; i < q.B; i += 32) {
            a_l1 += fxv(q.img_fwd + i * 4 + 0, FXL);
            a_c += fxv(q.img_fwd + i * 4 + 3, FXL) + fxv(img_bwd + i * 12 + 0, FXL);
            const float n = fxv(q.img_fwd + i * 4 + 1, FXL), d = fxv(q.img_fwd + i * 4 + 2, FXL);
            a_iou += n / (d + 1e-10f);
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a_l1 += __shfl_xor_sync(0xffffffffu, a_l1, o);
            a_c += __shfl_xor_sync(0xffffffffu, a_c, o);
            a_iou += __shfl_xor_sync(0xffffffffu, a_iou, o);
        }
        if (lane == 0) {
            const float npx = (float)q.B * (float)q.H * (float)q.W;
            const float l_img = a_l1 / (npx * 3.0f);
            const float l_iou = 1.0f - a_iou / (float)q.B;
            const float l_cont = (q.contour > 0.0f) ? a_c / npx : 0.0f;
            const float l_mask = l_iou + ((q.contour > 0.0f) ? l_cont * q.contour : 0.0f);
            q.loss[0] = q.image_weight * l_img + l_mask;
            q.loss[1] = l_img;
            q.loss[2] = l_iou;
            q.loss[3] = l_cont;
        }
    }
    MM_PROF_MARK(q.prof, 5, (blockIdx.y * gridDim.x + blockIdx.x) * VB_WARPS + (threadIdx.x >> 5), 2);
}

__global__ void k_export_faces(int V, int F, float multiplier, const int32_t* __restrict__ faces,
                               const float* __restrict__ frec, const float* __restrict__ vimg,
                               float* fvi, float* fvz, float* fnz, int total)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int b = i / F, f = i - b * F;
    const float* r = frec + (size_t)i * MM_REC_FLOATS;
    if (fvi) {
        for (int c = 0; c < 3; ++c) {
            const int v = faces[f * 3 + c];
            fvi[(size_t)i * 6 + c * 2] = vimg[((size_t)b * V + v) * 2];
            fvi[(size_t)i * 6 + c * 2 + 1] = vimg[((size_t)b * V + v) * 2 + 1];
        }
    }
    if (fvz) { fvz[(size_t)i * 3] = r[6]; fvz[(size_t)i * 3 + 1] = r[7]; fvz[(size_t)i * 3 + 2] = r[8]; }
    if (fnz) fnz[i] = r[11];
}

}  // namespace

cudaError_t mm_launch_vertex_fwd(const mm_ctx* c, int B, const float* vertices, const float* azim, const float* elev,
                          const float* dist, const float* bias, float* frec,
                          float* vimg, float* face_normals, float* gfacc_zero, long long* img_fwd, long long* img_bwd,
                          void* clr0, size_t bytes0, void* clr1, size_t bytes1, uint4* frect, cudaStream_t s)
{
    VertexFwdParams q;
    q.frect = frect; q.H = c->H; q.W = c->W; q.sx = c->sx; q.sy = c->sy; q.blen = c->blen;
    q.clr0 = (uint4*)clr0; q.n0 = bytes0 / 16; q.clr1 = (uint4*)clr1; q.n1 = bytes1 / 16; q.prof = c->prof;
    q.V = c->V; q.F = c->F; q.nchunks = c->nchunks;
    q.proj_x = c->proj_x; q.proj_y = c->proj_y; q.multiplier = c->multiplier;
    const dim3 grid(c->nchunks, B);
    // programmatic launch here too: in a loop of steps the launch latency of this first kernel hides under the tail of whatever
    // ran before on the stream (the prologue still waits for that work to complete before touching memory)
    return mm_launch(k_vertex_fwd, grid, dim3(MM_VTHREADS), c->smem_vertex_fwd, s, c->pdl != 0, q, (const int32_t*)c->d_faces, vertices,
              azim, elev, dist, bias, frec, vimg, face_normals, gfacc_zero, img_fwd, img_bwd);
}

cudaError_t mm_launch_vertex_bwd(const mm_ctx* c, int B, const float* vertices, const float* azim, const float* elev,
                          const float* dist, const float* bias, float* gfacc, const float* g_face_normals,
                          long long* img_bwd, float* g_vertices, float* g_azim, float* g_elev, float* g_dist,
                          float* g_bias, float* g_lights, float* loss, const long long* img_fwd, float image_weight,
                          float contour, cudaStream_t s)
{
    const size_t smem = ((size_t)c->V * 6) * sizeof(float);
    VertexBwdParams q;
    q.B = B; q.V = c->V; q.F = c->F; q.H = c->H; q.W = c->W; q.proj_x = c->proj_x; q.proj_y = c->proj_y;
    q.loss = loss; q.img_fwd = img_fwd; q.image_weight = image_weight; q.contour = contour; q.prof = c->prof;
    return mm_launch(k_vertex_bwd, dim3(VB_CLUSTER, B), dim3(VB_THREADS), smem, s, c->pdl != 0, q,
              (const int32_t*)c->d_faces, vertices, azim, elev, dist, bias, gfacc, g_face_normals, img_bwd, g_vertices,
              g_azim, g_elev, g_dist, g_bias, g_lights);
}

cudaError_t mm_launch_export_faces(const mm_ctx* c, int B, const float* frec, const float* vimg, float* fvi, float* fvz,
                            float* fnz, cudaStream_t s)
{
    const int total = B * c->F;
    k_export_faces<<<(total + 255) / 256, 256, 0, s>>>(c->V, c->F, c->multiplier, c->d_faces, frec, vimg, fvi, fvz,
                                                       fnz, total);
    return cudaPeekAtLastError();
}

size_t mm_vertex_smem_fwd(const mm_ctx* c) { return (16 + (size_t)c->V * 5) * sizeof(float); }
size_t mm_vertex_smem_bwd(int V) { return ((size_t)V * 6) * sizeof(float); }
// The opt-in dynamic shared memory limit is a per-device attribute of the KERNEL, shared by every ctx of the process: keep the
// largest requirement seen per device and only ever raise it (a second ctx with a smaller mesh must not lower the limit
// under a first one with a larger mesh).
cudaError_t mm_vertex_set_smem(int device, size_t fwd, size_t bwd) {
    static std::mutex mu;
    static size_t cur_fwd[64] = {0}, cur_bwd[64] = {0};
    std::lock_guard<std::mutex> lock(mu);
    const int d = (device >= 0 && device < 64) ? device : 0;
    if (fwd > cur_fwd[d]) {
        cudaError_t e = cudaFuncSetAttribute(k_vertex_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd);
        if (e != cudaSuccess) return e;
        cur_fwd[d] = fwd;
    }
    if (bwd > cur_bwd[d]) {
        cudaError_t e = cudaFuncSetAttribute(k_vertex_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd);
        if (e != cudaSuccess) return e;
        cur_bwd[d] = bwd;
    }
    return cudaSuccess;
}
