// mm_texflow.cu -- SURVEY 8(f)-3, texture side: the tail of TextureEncoder.forward (network/model_res.py:598-611), the
// PRODUCER of the atlas the renderer reads:
//     uv_sampler = texture_flow.permute(0, 2, 3, 1)
//     textures   = F.grid_sample(img, uv_sampler, mode='bicubic', align_corners=True)        (padding_mode='zeros')
//     textures   = torch.cat([textures, textures.flip([2])], dim=2)                           (when `concat`)
// One kernel per direction; the permute, the flip and the cat never materialise (forward: each sample is stored twice;
// backward: the two halves of g_out are summed on the fly).  Bicubic = cubic convolution with A = -0.75 on the 4x4 taps
// around the unnormalised coordinate, out-of-range taps read as zero: exactly ATen's grid_sampler_2d (checked against
// torch's own op in tests/test_gpu_parity.py).
#include "mm_common.cuh"

namespace {

#define TXF_A (-0.75f)

__device__ __forceinline__ float cc1(float x) { return ((TXF_A + 2.0f) * x - (TXF_A + 3.0f)) * x * x + 1.0f; }
__device__ __forceinline__ float cc2(float x) { return ((TXF_A * x - 5.0f * TXF_A) * x + 8.0f * TXF_A) * x - 4.0f * TXF_A; }

__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
    c[0] = cc2(t + 1.0f); c[1] = cc1(t); c[2] = cc1(1.0f - t); c[3] = cc2(2.0f - t);
}
__device__ __forceinline__ void cubic_coeffs_grad(float t, float (&c)[4]) {
    float x = -1.0f - t; c[0] = (-3.0f * TXF_A * x - 10.0f * TXF_A) * x - 8.0f * TXF_A;
    x = -t;              c[1] = (-3.0f * (TXF_A + 2.0f) * x - 2.0f * (TXF_A + 3.0f)) * x;
    x = 1.0f - t;        c[2] = (3.0f * (TXF_A + 2.0f) * x - 2.0f * (TXF_A + 3.0f)) * x;
    x = 2.0f - t;        c[3] = (3.0f * TXF_A * x - 10.0f * TXF_A) * x + 8.0f * TXF_A;
}

struct TxfParams {
    int B, C, Hi, Wi, Ho, Wo, concat;
    const float* img;      // [B,C,Hi,Wi]
    const float* flow;     // [B,2,Ho,Wo]
    float* out;            // [B,C,Ho*(1+concat),Wo]
    const float* g_out;
    float* g_img;
    float* g_flow;
};

template <bool BWD>
__global__ void __launch_bounds__(256)
k_texflow(const TxfParams q)
{
    mm_pdl_prologue();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z;
    if (x >= q.Wo) return;
    const size_t HWo = (size_t)q.Ho * q.Wo, HWi = (size_t)q.Hi * q.Wi;
    const size_t po = (size_t)y * q.Wo + x;
    const float gx = __ldg(q.flow + ((size_t)b * 2 + 0) * HWo + po), gy = __ldg(q.flow + ((size_t)b * 2 + 1) * HWo + po);
    // align_corners=True: [-1,1] -> [0, size-1]
    const float ix = (gx + 1.0f) * 0.5f * (float)(q.Wi - 1), iy = (gy + 1.0f) * 0.5f * (float)(q.Hi - 1);
    const float fx = floorf(ix), fy = floorf(iy);
    const float tx = ix - fx, ty = iy - fy;
    // non-finite coordinates: every tap is out of range (ATen's float -> int conversion of such values is undefined)
    const bool finite = (fabsf(ix) < 1e9f) && (fabsf(iy) < 1e9f);
    const int x0 = finite ? (int)fx - 1 : -8, y0 = finite ? (int)fy - 1 : -8;
    float cx[4], cy[4];
    cubic_coeffs(tx, cx); cubic_coeffs(ty, cy);
    const int Hout = q.Ho * (q.concat ? 2 : 1);
    if (!BWD) {
        for (int c = 0; c < q.C; ++c) {
            const float* ip = q.img + ((size_t)b * q.C + c) * HWi;
            float acc = 0.0f;
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int yy = y0 + j;
                float row = 0.0f;
                if (yy >= 0 && yy < q.Hi) {
                    #pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int xx = x0 + i;
                        const float v = (xx >= 0 && xx < q.Wi) ? __ldg(ip + (size_t)yy * q.Wi + xx) : 0.0f;
                        row += v * cx[i];
                    }
                }
                acc += row * cy[j];
            }
            float* op = q.out + ((size_t)b * q.C + c) * Hout * q.Wo;
            op[po] = acc;
            if (q.concat) op[(size_t)(2 * q.Ho - 1 - y) * q.Wo + x] = acc;       // cat([t, t.flip(2)], 2)
        }
    } else {
        float gcx[4], gcy[4];
        cubic_coeffs_grad(tx, gcx); cubic_coeffs_grad(ty, gcy);
        float gix = 0.0f, giy = 0.0f;
        for (int c = 0; c < q.C; ++c) {
            const float* gp = q.g_out + ((size_t)b * q.C + c) * Hout * q.Wo;
            float g = __ldg(gp + po);
            if (q.concat) g += __ldg(gp + (size_t)(2 * q.Ho - 1 - y) * q.Wo + x);
            const float* ip = q.img + ((size_t)b * q.C + c) * HWi;
            float* gi = q.g_img + ((size_t)b * q.C + c) * HWi;
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int yy = y0 + j;
                if (yy < 0 || yy >= q.Hi) continue;
                #pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int xx = x0 + i;
                    if (xx < 0 || xx >= q.Wi) continue;
                    const size_t t = (size_t)yy * q.Wi + xx;
                    const float w = g * cx[i] * cy[j];
                    if (w != 0.0f) atomicAdd(gi + t, w);
                    const float v = __ldg(ip + t);
                    gix -= v * gcx[i] * cy[j] * g;
                    giy -= v * gcy[j] * cx[i] * g;
                }
            }
        }
        q.g_flow[((size_t)b * 2 + 0) * HWo + po] = gix * (0.5f * (float)(q.Wi - 1));
        q.g_flow[((size_t)b * 2 + 1) * HWo + po] = giy * (0.5f * (float)(q.Hi - 1));
    }
}

}  // namespace

static cudaError_t launch(const mm_ctx* c, bool bwd, const TxfParams& q, cudaStream_t s)
{
    if (q.Ho > 65535 || q.B > 65535) return cudaErrorInvalidValue;
    const dim3 grid((q.Wo + 255) / 256, q.Ho, q.B);
    // g_img is cleared by a memset in front of the backward: no programmatic launch across it
    return bwd ? mm_launch(k_texflow<true>, grid, dim3(256), 0, s, false, q)
               : mm_launch(k_texflow<false>, grid, dim3(256), 0, s, c->pdl != 0, q);
}

cudaError_t mm_launch_texflow_fwd(const mm_ctx* c, int B, int C, int Hi, int Wi, int Ho, int Wo, int concat, const float* img,
                                  const float* flow, float* out, cudaStream_t s)
{
    TxfParams q = {};
    q.B = B; q.C = C; q.Hi = Hi; q.Wi = Wi; q.Ho = Ho; q.Wo = Wo; q.concat = concat ? 1 : 0;
    q.img = img; q.flow = flow; q.out = out;
    return launch(c, false, q, s);
}

cudaError_t mm_launch_texflow_bwd(const mm_ctx* c, int B, int C, int Hi, int Wi, int Ho, int Wo, int concat, const float* img,
                                  const float* flow, const float* g_out, float* g_img, float* g_flow, cudaStream_t s)
{
    TxfParams q = {};
    q.B = B; q.C = C; q.Hi = Hi; q.Wi = Wi; q.Ho = Ho; q.Wo = Wo; q.concat = concat ? 1 : 0;
    q.img = img; q.flow = flow; q.g_out = g_out; q.g_img = g_img; q.g_flow = g_flow;
    return launch(c, true, q, s);
}
