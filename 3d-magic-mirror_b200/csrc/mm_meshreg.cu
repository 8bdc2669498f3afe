// mm_meshreg.cu -- the mesh regularisers of DiffRender (networks.py:392-491), SURVEY 8(f)-1: the rows next to the render path.
//
// The reference evaluates them with ~80 small torch kernels per training iteration (index_select / matmul with a dense
// V x V Laplacian / norm / mean ...), all on B x V or B x E sized data.  Here ONE launch per direction computes every term:
//   0 laplacian  mean((Lap @ delta)^2) * V * 3                                calc_reg_loss   networks.py:425-426
//   1 flat       mean((<n_f0, n_f1> - 1)^2) * E over the edge -> face pairs   calc_reg_loss   networks.py:428-431
//   2 edge       0.1 * mean_b || len_e - mean_e(len_e) ||_2                   calc_reg_edge   networks.py:453-461
//   3 depth      mean(z^2)                                                    calc_reg_depth  networks.py:463-466
//   4 depthR     mean((z -+ eps)^2 * exp(temp (x^2 + (y/ratio)^2)))           calc_reg_depthR networks.py:468-475
//   5 depthC     mean((z -+ eps)^2 * (x^2 + (y/ratio)^2))                     calc_reg_depthC networks.py:477-485
//   6 deform     mean ||delta_v||_2                                           calc_reg_deform networks.py:487-491
//   7 flip       mean(|delta_v - M delta_flip(v)| * mask)                     recon_flip      networks.py:392-410
// One CTA per image; per-image partials are reduced over the batch by the last CTA in a fixed order (deterministic).
// The backward takes the upstream gradient of each term (g_terms[8]) and emits d/d delta_vertices, d/d vertices and
// d/d face_normals; outputs are overwritten.  The dense Laplacian is used in CSR form (6-7 non-zeros per row).
#include "mm_device.cuh"

namespace {

#define REG_THREADS 256
#define REG_TERMS 8

struct RegParams {
    int B, V, F, E;
    const int32_t* edges;        // [E,2]
    const int32_t* edge2faces;   // [E,2]
    const int32_t* flip;         // [V]
    const float* sign_init;      // [V]
    const int32_t* lap_off;      // [V+1]
    const int32_t* lap_col;      // [nnz]
    const float* lap_val;        // [nnz]
    float ratio, temp, eps;
    int flip_l1;
    unsigned mask;               // which terms to evaluate (bit k = term k)
    const float* delta;          // [B,V,3] or NULL
    const float* vertices;       // [B,V,3] or NULL
    const float* fn;             // [B,F,3] or NULL
};

__device__ __forceinline__ float reg_block_sum(float v, float* red) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.0f;
    #pragma unroll
    for (int i = 0; i < REG_THREADS / 32; ++i) r += red[i];
    return r;
}

__device__ __forceinline__ float flip_mask(const RegParams& q, const float* d, int j) {
    // relu(sign(Na_z) * sign_init) of vertex j (networks.py:404)
    const float z = d[j * 3 + 2];
    const float s = (z > 0.0f) ? 1.0f : ((z < 0.0f) ? -1.0f : 0.0f);
    return fmaxf(s * q.sign_init[j], 0.0f);
}

// per-image sums of every term (un-normalised), then the batch reduction by the last CTA
__global__ void __launch_bounds__(REG_THREADS)
k_meshreg_fwd(const RegParams q, float* __restrict__ partials /*[B,8]*/, unsigned* __restrict__ ticket, float* __restrict__ terms /*[8]*/)
{
    mm_pdl_prologue();
    __shared__ float red[REG_THREADS / 32];
    __shared__ bool s_last;
    const int b = blockIdx.x;
    const float* d = q.delta ? q.delta + (size_t)b * q.V * 3 : nullptr;
    const float* v = q.vertices ? q.vertices + (size_t)b * q.V * 3 : nullptr;
    const float* n = q.fn ? q.fn + (size_t)b * q.F * 3 : nullptr;
    float acc[REG_TERMS];
    #pragma unroll
    for (int k = 0; k < REG_TERMS; ++k) acc[k] = 0.0f;
    float edge_sum = 0.0f;
    for (int i = threadIdx.x; i < q.V; i += REG_THREADS) {
        if ((q.mask & 1u) && d) {                                  // Laplacian row i
            float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f;
            for (int k = q.lap_off[i]; k < q.lap_off[i + 1]; ++k) {
                const int j = q.lap_col[k];
                const float w = q.lap_val[k];
                l0 += w * d[j * 3]; l1 += w * d[j * 3 + 1]; l2 += w * d[j * 3 + 2];
            }
            acc[0] += l0 * l0 + l1 * l1 + l2 * l2;
        }
        if (v) {
            const float x = v[i * 3], y = v[i * 3 + 1], z = v[i * 3 + 2];
            if (q.mask & 8u) acc[3] += z * z;
            if (q.mask & 48u) {
                const float yr = y / q.ratio;
                const float r2 = x * x + yr * yr;
                const float zz = (q.sign_init[i] >= 0.0f) ? (z - q.eps) : (z + q.eps);
                if (q.mask & 16u) acc[4] += zz * zz * expf(q.temp * r2);
                if (q.mask & 32u) acc[5] += zz * zz * r2;
            }
        }
        if (d) {
            if (q.mask & 64u) acc[6] += sqrtf(d[i * 3] * d[i * 3] + d[i * 3 + 1] * d[i * 3 + 1] + d[i * 3 + 2] * d[i * 3 + 2]);
            if (q.mask & 128u) {
                const int f = q.flip[i];
                const float dx = d[i * 3] - d[f * 3], dy = d[i * 3 + 1] - d[f * 3 + 1], dz = d[i * 3 + 2] + d[f * 3 + 2];
                const float m = flip_mask(q, d, f);
                acc[7] += (q.flip_l1 ? (fabsf(dx) + fabsf(dy) + fabsf(dz)) : sqrtf(dx * dx + dy * dy + dz * dz)) * m;
            }
        }
    }
    for (int e = threadIdx.x; e < q.E; e += REG_THREADS) {
        if ((q.mask & 2u) && n) {
            const int f0 = q.edge2faces[e * 2], f1 = q.edge2faces[e * 2 + 1];
            const float c = n[f0 * 3] * n[f1 * 3] + n[f0 * 3 + 1] * n[f1 * 3 + 1] + n[f0 * 3 + 2] * n[f1 * 3 + 2];
            acc[1] += (c - 1.0f) * (c - 1.0f);
        }
        if ((q.mask & 4u) && v) {
            const int a = q.edges[e * 2], c2 = q.edges[e * 2 + 1];
            const float dx = v[a * 3] - v[c2 * 3], dy = v[a * 3 + 1] - v[c2 * 3 + 1], dz = v[a * 3 + 2] - v[c2 * 3 + 2];
            edge_sum += sqrtf(dx * dx + dy * dy + dz * dz);
        }
    }
    if ((q.mask & 4u) && v) {                                      // edge term: second pass around the per-image mean length
        const float mean = reg_block_sum(edge_sum, red) / (float)q.E;
        float dev = 0.0f;
        for (int e = threadIdx.x; e < q.E; e += REG_THREADS) {
            const int a = q.edges[e * 2], c2 = q.edges[e * 2 + 1];
            const float dx = v[a * 3] - v[c2 * 3], dy = v[a * 3 + 1] - v[c2 * 3 + 1], dz = v[a * 3 + 2] - v[c2 * 3 + 2];
            const float bias = sqrtf(dx * dx + dy * dy + dz * dz) - mean;
            dev += bias * bias;
        }
        const float tot = reg_block_sum(dev, red);
        acc[2] = (threadIdx.x == 0) ? sqrtf(tot) : 0.0f;
    }
    #pragma unroll
    for (int k = 0; k < REG_TERMS; ++k) {
        const float s = reg_block_sum(acc[k], red);
        if (threadIdx.x == 0) partials[b * REG_TERMS + k] = s;
    }
    // ---- last CTA: batch reduction in image order + the reference's normalisations
    __threadfence();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == (unsigned)(q.B - 1));
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < REG_TERMS) {
        float s = 0.0f;
        for (int i = 0; i < q.B; ++i) s += __ldcg(partials + i * REG_TERMS + threadIdx.x);
        const float fB = (float)q.B, fBV = (float)q.B * (float)q.V;
        float out;
        switch (threadIdx.x) {
            case 0: out = s / fB; break;                           // mean over B*V*3, times V*3
            case 1: out = s / fB; break;                           // mean over B*E, times E
            case 2: out = 0.1f * s / fB; break;
            case 3: case 4: case 5: case 6: out = s / fBV; break;
            default: out = s / (q.flip_l1 ? fBV * 3.0f : fBV); break;
        }
        terms[threadIdx.x] = out;
    }
    if (threadIdx.x == 0) *ticket = 0u;                            // self-resetting
}

// gradients of sum_k g_terms[k] * term_k; one CTA per image, outputs zero-filled then accumulated with atomics
__global__ void __launch_bounds__(REG_THREADS)
k_meshreg_bwd(const RegParams q, const float* __restrict__ g_terms, float* __restrict__ g_delta, float* __restrict__ g_vertices,
              float* __restrict__ g_fn)
{
    mm_pdl_prologue();
    __shared__ float red[REG_THREADS / 32];
    const int b = blockIdx.x;
    const float* d = q.delta ? q.delta + (size_t)b * q.V * 3 : nullptr;
    const float* v = q.vertices ? q.vertices + (size_t)b * q.V * 3 : nullptr;
    const float* n = q.fn ? q.fn + (size_t)b * q.F * 3 : nullptr;
    float* gd = g_delta ? g_delta + (size_t)b * q.V * 3 : nullptr;
    float* gv = g_vertices ? g_vertices + (size_t)b * q.V * 3 : nullptr;
    float* gn = g_fn ? g_fn + (size_t)b * q.F * 3 : nullptr;
    for (int i = threadIdx.x; i < q.V * 3; i += REG_THREADS) { if (gd) gd[i] = 0.0f; if (gv) gv[i] = 0.0f; }
    for (int i = threadIdx.x; i < q.F * 3; i += REG_THREADS) if (gn) gn[i] = 0.0f;
    __syncthreads();
    float g[REG_TERMS];
    #pragma unroll
    for (int k = 0; k < REG_TERMS; ++k) g[k] = ((q.mask >> k) & 1u) ? g_terms[k] : 0.0f;
    const float fB = (float)q.B, fBV = (float)q.B * (float)q.V;
    float edge_sum = 0.0f;
    for (int i = threadIdx.x; i < q.V; i += REG_THREADS) {
        if (g[0] != 0.0f && d && gd) {                              // d/d delta of sum (Lap delta)^2 / B : 2/B Lap^T (Lap delta)
            float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f;
            for (int k = q.lap_off[i]; k < q.lap_off[i + 1]; ++k) {
                const int j = q.lap_col[k];
                const float w = q.lap_val[k];
                l0 += w * d[j * 3]; l1 += w * d[j * 3 + 1]; l2 += w * d[j * 3 + 2];
            }
            const float c = 2.0f * g[0] / fB;
            for (int k = q.lap_off[i]; k < q.lap_off[i + 1]; ++k) {
                const int j = q.lap_col[k];
                const float w = q.lap_val[k] * c;
                atomicAdd(gd + j * 3, w * l0); atomicAdd(gd + j * 3 + 1, w * l1); atomicAdd(gd + j * 3 + 2, w * l2);
            }
        }
        if (v && gv) {
            const float x = v[i * 3], y = v[i * 3 + 1], z = v[i * 3 + 2];
            float gz = 0.0f;
            if (g[3] != 0.0f) gz += g[3] * 2.0f * z / fBV;
            if (g[4] != 0.0f || g[5] != 0.0f) {                    // x, y enter detached (networks.py:470-471, 479-480)
                const float yr = y / q.ratio;
                const float r2 = x * x + yr * yr;
                const float zz = (q.sign_init[i] >= 0.0f) ? (z - q.eps) : (z + q.eps);
                if (g[4] != 0.0f) gz += g[4] * 2.0f * zz * expf(q.temp * r2) / fBV;
                if (g[5] != 0.0f) gz += g[5] * 2.0f * zz * r2 / fBV;
            }
            if (gz != 0.0f) atomicAdd(gv + i * 3 + 2, gz);
        }
        if (d && gd) {
            if (g[6] != 0.0f) {
                const float nr = sqrtf(d[i * 3] * d[i * 3] + d[i * 3 + 1] * d[i * 3 + 1] + d[i * 3 + 2] * d[i * 3 + 2]);
                if (nr > 0.0f) {
                    const float c = g[6] / (fBV * nr);
                    atomicAdd(gd + i * 3, c * d[i * 3]); atomicAdd(gd + i * 3 + 1, c * d[i * 3 + 1]); atomicAdd(gd + i * 3 + 2, c * d[i * 3 + 2]);
                }
            }
            if (g[7] != 0.0f) {
                const int f = q.flip[i];
                const float m = flip_mask(q, d, f);                 // the mask carries no gradient (sign / relu of a sign)
                if (m != 0.0f) {
                    const float dx = d[i * 3] - d[f * 3], dy = d[i * 3 + 1] - d[f * 3 + 1], dz = d[i * 3 + 2] + d[f * 3 + 2];
                    float ux, uy, uz;
                    if (q.flip_l1) {
                        const float c = g[7] * m / (fBV * 3.0f);
                        ux = c * sgnf(dx); uy = c * sgnf(dy); uz = c * sgnf(dz);
                    } else {
                        const float nr = sqrtf(dx * dx + dy * dy + dz * dz);
                        const float c = nr > 0.0f ? g[7] * m / (fBV * nr) : 0.0f;
                        ux = c * dx; uy = c * dy; uz = c * dz;
                    }
                    atomicAdd(gd + i * 3, ux); atomicAdd(gd + i * 3 + 1, uy); atomicAdd(gd + i * 3 + 2, uz);
                    atomicAdd(gd + f * 3, -ux); atomicAdd(gd + f * 3 + 1, -uy); atomicAdd(gd + f * 3 + 2, uz);
                }
            }
        }
    }
    for (int e = threadIdx.x; e < q.E; e += REG_THREADS) {
        if (g[1] != 0.0f && n && gn) {
            const int f0 = q.edge2faces[e * 2], f1 = q.edge2faces[e * 2 + 1];
            const float c = n[f0 * 3] * n[f1 * 3] + n[f0 * 3 + 1] * n[f1 * 3 + 1] + n[f0 * 3 + 2] * n[f1 * 3 + 2];
            const float k = 2.0f * g[1] * (c - 1.0f) / fB;
            #pragma unroll
            for (int a = 0; a < 3; ++a) { atomicAdd(gn + f0 * 3 + a, k * n[f1 * 3 + a]); atomicAdd(gn + f1 * 3 + a, k * n[f0 * 3 + a]); }
        }
        if (g[2] != 0.0f && v) {
            const int a = q.edges[e * 2], c2 = q.edges[e * 2 + 1];
            const float dx = v[a * 3] - v[c2 * 3], dy = v[a * 3 + 1] - v[c2 * 3 + 1], dz = v[a * 3 + 2] - v[c2 * 3 + 2];
            edge_sum += sqrtf(dx * dx + dy * dy + dz * dz);
        }
    }
    if (g[2] != 0.0f && v && gv) {                                  // d/d v of 0.1/B * sqrt(sum_e (len_e - mean)^2): bias_e / norm * dlen_e
        const float mean = reg_block_sum(edge_sum, red) / (float)q.E;
        float dev = 0.0f;
        for (int e = threadIdx.x; e < q.E; e += REG_THREADS) {
            const int a = q.edges[e * 2], c2 = q.edges[e * 2 + 1];
            const float dx = v[a * 3] - v[c2 * 3], dy = v[a * 3 + 1] - v[c2 * 3 + 1], dz = v[a * 3 + 2] - v[c2 * 3 + 2];
            const float bias = sqrtf(dx * dx + dy * dy + dz * dz) - mean;
            dev += bias * bias;
        }
        const float nrm = sqrtf(reg_block_sum(dev, red));
        if (nrm > 0.0f) {
            const float c = 0.1f * g[2] / (fB * nrm);
            for (int e = threadIdx.x; e < q.E; e += REG_THREADS) {
                const int a = q.edges[e * 2], c2 = q.edges[e * 2 + 1];
                const float dx = v[a * 3] - v[c2 * 3], dy = v[a * 3 + 1] - v[c2 * 3 + 1], dz = v[a * 3 + 2] - v[c2 * 3 + 2];
                const float len = sqrtf(dx * dx + dy * dy + dz * dz);
                if (len > 0.0f) {
                    const float k = c * (len - mean) / len;          // sum_e bias_e = 0, so the mean's own gradient cancels
                    atomicAdd(gv + a * 3, k * dx); atomicAdd(gv + a * 3 + 1, k * dy); atomicAdd(gv + a * 3 + 2, k * dz);
                    atomicAdd(gv + c2 * 3, -k * dx); atomicAdd(gv + c2 * 3 + 1, -k * dy); atomicAdd(gv + c2 * 3 + 2, -k * dz);
                }
            }
        }
    }
}

}  // namespace

static void fill_reg(const mm_ctx* c, int B, const float* delta, const float* vertices, const float* fn, float temp, float eps,
                     int flip_l1, unsigned mask, RegParams& q)
{
    q.B = B; q.V = c->V; q.F = c->F; q.E = c->reg_E;
    q.edges = c->d_edges; q.edge2faces = c->d_edge2faces; q.flip = c->d_flip; q.sign_init = c->d_sign_init;
    q.lap_off = c->d_lap_off; q.lap_col = c->d_lap_col; q.lap_val = c->d_lap_val;
    q.ratio = c->reg_ratio; q.temp = temp; q.eps = eps; q.flip_l1 = flip_l1; q.mask = mask;
    q.delta = delta; q.vertices = vertices; q.fn = fn;
}

cudaError_t mm_launch_meshreg_fwd(const mm_ctx* c, int B, const float* delta, const float* vertices, const float* fn, float temp,
                           float eps, int flip_l1, unsigned mask, float* partials, unsigned* ticket, float* terms, cudaStream_t s)
{
    RegParams q;
    fill_reg(c, B, delta, vertices, fn, temp, eps, flip_l1, mask, q);
    return mm_launch(k_meshreg_fwd, dim3(B), dim3(REG_THREADS), 0, s, false, q, partials, ticket, terms);
}

cudaError_t mm_launch_meshreg_bwd(const mm_ctx* c, int B, const float* delta, const float* vertices, const float* fn, float temp,
                           float eps, int flip_l1, unsigned mask, const float* g_terms, float* g_delta, float* g_vertices,
                           float* g_fn, cudaStream_t s)
{
    RegParams q;
    fill_reg(c, B, delta, vertices, fn, temp, eps, flip_l1, mask, q);
    return mm_launch(k_meshreg_bwd, dim3(B), dim3(REG_THREADS), 0, s, false, q, g_terms, g_delta, g_vertices, g_fn);
}
