// mm_template.cu -- SURVEY 8(f)-3: the encoder-side glue that touches the template mesh.
//
// ShapeEncoder.forward (network/model_res.py:314-328) conditions the per-vertex shape head on the template:
//     uv      = template[:, :, 0:2]                                                       (model_res.py:317-318, detached)
//     local   = F.grid_sample(x, uv, 'bilinear', align_corners=True, padding_mode='zeros')  (:321)   [B,C,V,1]
//     ndiff   = torch.mm(local.view(-1, V), lpl)                                          (:324)   [B*C,V] x [V,V] dense
// with lpl = DiffRender.vertices_laplacian_matrix (trainer.py:91), 6-7 non-zeros per row.  The reference runs a gather
// kernel, a .repeat and a dense (B*C) x V x V GEMM (11 GFLOP at B=48, C=288); here ONE kernel per direction does both with
// the Laplacian in sparse form: a warp owns a (b, c) row, the V sampled features stay in shared memory between the gather
// and the Laplacian, and both outputs leave as coalesced row stores.  HBM traffic = the two [B*C, V] outputs (written once).
// The sampling positions depend on the template only, so every CTA builds the table of 4 texel indices + 4 weights per
// vertex once and reuses it for all of its rows.
//
// Backward (x only: the reference detaches the sampling positions): g_tot = g_local + lpl @ g_ndiff (row form of the same
// sparse matrix), then every plane element gathers its (vertex, weight) contributions from a per-texel list.
#include "mm_common.cuh"

namespace {

#define TF_THREADS 128
#define TF_WARPS (TF_THREADS / 32)
#define TF_MAX_PLANE 2048          // h*w bound of the backward's per-texel lists in shared memory

struct TFParams {
    int N, V, h, w;                // N = B*C rows
    const float* tmpl;             // [V,3] template positions (x, y used), in [-1,1]
    const int32_t* off;            // forward: CSR of lpl^T (column j -> rows i);  backward: CSR of lpl (row i -> columns j)
    const int32_t* idx;
    const float* val;
};

// grid_sample(align_corners=True): pixel = (g + 1) / 2 * (size - 1); corners outside the plane contribute zero
__device__ __forceinline__ void tf_build_table(const TFParams& q, int* s_idx, float* s_w)
{
    for (int v = threadIdx.x; v < q.V; v += blockDim.x) {
        const float gx = q.tmpl[v * 3], gy = q.tmpl[v * 3 + 1];
        const float ix = ((gx + 1.0f) / 2.0f) * (float)(q.w - 1);
        const float iy = ((gy + 1.0f) / 2.0f) * (float)(q.h - 1);
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
        const float wx1 = ix - fx, wx0 = (float)x1 - ix, wy1 = iy - fy, wy0 = (float)y1 - iy;
        const bool vx0 = x0 >= 0 && x0 < q.w, vx1 = x1 >= 0 && x1 < q.w, vy0 = y0 >= 0 && y0 < q.h, vy1 = y1 >= 0 && y1 < q.h;
        // a NaN / inf position makes every validity test false: the vertex samples zeros, as in torch
        // [corner][vertex] layout: consecutive lanes read consecutive words
        s_idx[0 * q.V + v] = (vx0 && vy0) ? y0 * q.w + x0 : -1; s_w[0 * q.V + v] = wx0 * wy0;     // nw
        s_idx[1 * q.V + v] = (vx1 && vy0) ? y0 * q.w + x1 : -1; s_w[1 * q.V + v] = wx1 * wy0;     // ne
        s_idx[2 * q.V + v] = (vx0 && vy1) ? y1 * q.w + x0 : -1; s_w[2 * q.V + v] = wx0 * wy1;     // sw
        s_idx[3 * q.V + v] = (vx1 && vy1) ? y1 * q.w + x1 : -1; s_w[3 * q.V + v] = wx1 * wy1;     // se
    }
}

// dynamic smem: [V*4] int table | [V*4] float weights | TF_WARPS x [V] floats (the row's sampled features)
__global__ void __launch_bounds__(TF_THREADS)
k_template_features_fwd(const TFParams q, const float* __restrict__ x, float* __restrict__ local, float* __restrict__ ndiff)
{
    mm_pdl_prologue();
    extern __shared__ float tf_sm[];
    int* s_idx = reinterpret_cast<int*>(tf_sm);
    float* s_w = tf_sm + (size_t)q.V * 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* s_loc = s_w + (size_t)q.V * 4 + (size_t)warp * q.V;
    tf_build_table(q, s_idx, s_w);
    __syncthreads();
    const int hw = q.h * q.w;
    for (int row = blockIdx.x * TF_WARPS + warp; row < q.N; row += gridDim.x * TF_WARPS) {
        const float* plane = x + (size_t)row * hw;
        for (int v = lane; v < q.V; v += 32) {
            float a = 0.0f;
            #pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int t = s_idx[k * q.V + v];
                if (t >= 0) a += __ldg(plane + t) * s_w[k * q.V + v];
            }
            s_loc[v] = a;
            local[(size_t)row * q.V + v] = a;
        }
        __syncwarp();
        if (ndiff) {
            for (int j = lane; j < q.V; j += 32) {
                float a = 0.0f;
                for (int k = __ldg(q.off + j), ke = __ldg(q.off + j + 1); k < ke; ++k) a += s_loc[__ldg(q.idx + k)] * __ldg(q.val + k);
                ndiff[(size_t)row * q.V + j] = a;
            }
        }
        __syncwarp();
    }
}

// dynamic smem: table [4V] | weights [4V] | texel lists: begins [hw+1], cursors [hw], entry vertex [4V], entry weight [4V]
// | TF_WARPS x 2 x [V] floats (g_ndiff row, g_tot row).
// The scatter through the bilinear weights is inverted ONCE per CTA into a list of (vertex, weight) contributions per texel
// (the sampling positions depend on the template only); per row, lane t then GATHERS plane element t from the g_tot row in
// shared memory.  (The first version scattered with shared-memory float atomics: ~80-fold address conflicts per texel.)
__global__ void __launch_bounds__(TF_THREADS)
k_template_features_bwd(const TFParams q, const float* __restrict__ g_local, const float* __restrict__ g_ndiff,
                        float* __restrict__ g_x)
{
    mm_pdl_prologue();
    extern __shared__ float tf_sm[];
    const int V = q.V, hw = q.h * q.w;
    int* s_idx = reinterpret_cast<int*>(tf_sm);
    float* s_w = tf_sm + (size_t)V * 4;
    int* s_beg = reinterpret_cast<int*>(s_w + (size_t)V * 4);             // [hw + 1]
    int* s_cur = s_beg + hw + 1;                                          // [hw]
    int* s_tv = s_cur + hw;                                               // [4V]
    float* s_tw = reinterpret_cast<float*>(s_tv + (size_t)V * 4);         // [4V]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* s_g = s_tw + (size_t)V * 4 + (size_t)warp * 2 * V;             // this warp's g_ndiff row
    float* s_t = s_g + V;                                                 // this warp's g_tot row
    tf_build_table(q, s_idx, s_w);
    for (int t = threadIdx.x; t <= hw; t += blockDim.x) s_beg[t] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < 4 * V; e += blockDim.x) { const int t = s_idx[e]; if (t >= 0) atomicAdd(&s_beg[t + 1], 1); }
    __syncthreads();
    if (warp == 0) {                                                      // inclusive scan of the counts: s_beg[t] = first entry of texel t
        int carry = 0;
        for (int t0 = 1; t0 <= hw; t0 += 32) {
            const int t = t0 + lane;
            int v = (t <= hw) ? s_beg[t] : 0;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
            if (t <= hw) s_beg[t] = v + carry;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < hw; t += blockDim.x) s_cur[t] = s_beg[t];
    __syncthreads();
    for (int e = threadIdx.x; e < 4 * V; e += blockDim.x) {
        const int t = s_idx[e];
        if (t >= 0) {
            const int d = atomicAdd(&s_cur[t], 1);
            s_tv[d] = e % V;                                              // table layout is [corner][vertex]
            s_tw[d] = s_w[e];
        }
    }
    __syncthreads();
    for (int row = blockIdx.x * TF_WARPS + warp; row < q.N; row += gridDim.x * TF_WARPS) {
        if (g_ndiff) for (int j = lane; j < V; j += 32) s_g[j] = g_ndiff[(size_t)row * V + j];
        __syncwarp();
        for (int v = lane; v < V; v += 32) {
            float g = g_local ? g_local[(size_t)row * V + v] : 0.0f;
            if (g_ndiff)                                                  // (lpl @ g_ndiff)[v]
                for (int k = __ldg(q.off + v), ke = __ldg(q.off + v + 1); k < ke; ++k) g += __ldg(q.val + k) * s_g[__ldg(q.idx + k)];
            s_t[v] = g;
        }
        __syncwarp();
        for (int t = lane; t < hw; t += 32) {
            float a = 0.0f;
            for (int d = s_beg[t], de = s_beg[t + 1]; d < de; ++d) a += s_tw[d] * s_t[s_tv[d]];
            g_x[(size_t)row * hw + t] = a;
        }
        __syncwarp();
    }
}

}  // namespace

static size_t tf_smem_fwd(int V) { return ((size_t)V * 8 + (size_t)TF_WARPS * V) * sizeof(float); }
static size_t tf_smem_bwd(int V, int hw) { return ((size_t)V * 16 + 2 * (size_t)hw + 1 + (size_t)TF_WARPS * 2 * V) * sizeof(float); }

int mm_template_max_plane(void) { return TF_MAX_PLANE; }

cudaError_t mm_launch_template_fwd(const mm_ctx* c, int N, int h, int w, const float* x, const float* tmpl, float* local,
                                   float* ndiff, cudaStream_t s)
{
    TFParams q;
    q.N = N; q.V = c->V; q.h = h; q.w = w; q.tmpl = tmpl;
    q.off = c->d_lapT_off; q.idx = c->d_lapT_row; q.val = c->d_lapT_val;
    const size_t smem = tf_smem_fwd(c->V);
    cudaError_t e = cudaFuncSetAttribute(k_template_features_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int grid = (N + TF_WARPS - 1) / TF_WARPS;
    const int cap = c->num_sms * 16;                      // persistent over rows: the per-CTA table is built once
    if (grid > cap) grid = cap;
    return mm_launch(k_template_features_fwd, dim3(grid), dim3(TF_THREADS), smem, s, c->pdl != 0, q, x, local, ndiff);
}

cudaError_t mm_launch_template_bwd(const mm_ctx* c, int N, int h, int w, const float* tmpl, const float* g_local,
                                   const float* g_ndiff, float* g_x, cudaStream_t s)
{
    TFParams q;
    q.N = N; q.V = c->V; q.h = h; q.w = w; q.tmpl = tmpl;
    q.off = c->d_lap_off; q.idx = c->d_lap_col; q.val = c->d_lap_val;
    const size_t smem = tf_smem_bwd(c->V, h * w);
    cudaError_t e = cudaFuncSetAttribute(k_template_features_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int grid = (N + TF_WARPS - 1) / TF_WARPS;
    const int cap = c->num_sms * 8;
    if (grid > cap) grid = cap;
    return mm_launch(k_template_features_bwd, dim3(grid), dim3(TF_THREADS), smem, s, c->pdl != 0, q, g_local, g_ndiff, g_x);
}
