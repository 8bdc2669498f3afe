// mm_loss.cu -- DiffRender.recon_data (networks.py:364-390) as stand-alone kernels
// (masked L1 + kaolin mask_iou + contour MSE) and the loss finalisation used by both
// the stand-alone and the fused render-compare path.
#include "mm_device.cuh"

namespace {

// One CTA per (contiguous pixel range, image): partial sums (L1, N, D, contour) -> part_fwd[b][band][4]
__global__ void __launch_bounds__(MM_THREADS)
k_recon_fwd(int H, int W, int nparts, float contour,
            const float* __restrict__ pred, const float* __restrict__ gt, const int32_t* __restrict__ tab,
            float* __restrict__ part_fwd)
{
    mm_pdl_prologue();
    __shared__ float red[MM_WARPS];
    const int b = blockIdx.y, band = blockIdx.x;
    const size_t HW = (size_t)H * W;
    const float* pb = pred + (size_t)b * 4 * HW;
    const float* gb = gt + (size_t)b * 4 * HW;
    const int32_t* refrow = tab;
    const int32_t* refcol = tab + 3 * H;
    const int per = (H * W + nparts - 1) / nparts;               // contiguous pixel range of this CTA
    const int i0 = band * per, i1 = min(H * W, i0 + per);
    float a_l1 = 0.0f, a_n = 0.0f, a_d = 0.0f, a_c = 0.0f;
    for (int i = i0 + threadIdx.x; i < i1; i += MM_THREADS) {
        const float gm = gb[3 * HW + i], m = pb[3 * HW + i];
        #pragma unroll
        for (int c = 0; c < 3; ++c) a_l1 += fabsf(l1_term(pb[c * HW + i], gb[c * HW + i], gm));
        const float mul = m * gm;
        a_n += mul;
        a_d += (m + gm) - mul;
        if (contour > 0.0f) {
            const int iy = i / W, ix = i - iy * W;
            const size_t rp = (size_t)refrow[iy] * W + refcol[ix];
            const float d = fabsf(m - pb[3 * HW + rp]) - fabsf(gm - gb[3 * HW + rp]);
            a_c += d * d;
        }
    }
    const float s0 = block_sum(a_l1, red), s1 = block_sum(a_n, red), s2 = block_sum(a_d, red), s3 = block_sum(a_c, red);
    if (threadIdx.x == 0) {
        float* pf = part_fwd + ((size_t)b * nparts + band) * 4;
        pf[0] = s0; pf[1] = s1; pf[2] = s2; pf[3] = s3;
    }
}

__global__ void __launch_bounds__(MM_THREADS)
k_recon_bwd(int B, int H, int W, int nparts, float image_weight, float contour, float loss_scale,
            const float* __restrict__ pred, const float* __restrict__ gt, const int32_t* __restrict__ tab,
            const long long* __restrict__ img_fwd, float* __restrict__ g_pred)
{
    mm_pdl_prologue();
    const int b = blockIdx.y, band = blockIdx.x;
    const size_t HW = (size_t)H * W;
    const float* pb = pred + (size_t)b * 4 * HW;
    const float* gb = gt + (size_t)b * 4 * HW;
    float* go = g_pred + (size_t)b * 4 * HW;
    const int32_t* refrow = tab;
    const int32_t* rowlo = tab + H;
    const int32_t* rowhi = tab + 2 * H;
    const int32_t* refcol = tab + 3 * H;
    const int32_t* collo = tab + 3 * H + W;
    const int32_t* colhi = tab + 3 * H + 2 * W;
    const float Nb = fx_get(img_fwd + b * 4 + 1, MM_FX_LOSS);
    const float De = fx_get(img_fwd + b * 4 + 2, MM_FX_LOSS) + 1e-10f;
    const float k_img = loss_scale * image_weight / ((float)B * 3.0f * (float)HW);
    const float k_iou = loss_scale / (float)B;
    const float k_cont = loss_scale * contour / ((float)B * (float)HW);
    const int per = (H * W + nparts - 1) / nparts;
    const int i0 = band * per, i1 = min(H * W, i0 + per);
    for (int i = i0 + threadIdx.x; i < i1; i += MM_THREADS) {
        const float gm = gb[3 * HW + i], m = pb[3 * HW + i];
        #pragma unroll
        for (int c = 0; c < 3; ++c)
            go[c * HW + i] = k_img * sgnf(l1_term(pb[c * HW + i], gb[c * HW + i], gm)) * gm;
        float g = -k_iou * (gm * De - Nb * (1.0f - gm)) / (De * De);
        if (contour > 0.0f) {
            const int iy = i / W, ix = i - iy * W;
            const size_t rp = (size_t)refrow[iy] * W + refcol[ix];
            const float mref = pb[3 * HW + rp], gref = gb[3 * HW + rp];
            const float dlt = fabsf(m - mref) - fabsf(gm - gref);
            float gc = 2.0f * dlt * sgnf(m - mref);
            for (int yy = rowlo[iy]; yy < rowhi[iy]; ++yy)
                for (int xx = collo[ix]; xx < colhi[ix]; ++xx) {
                    const size_t q = (size_t)yy * W + xx;
                    const float mq = pb[3 * HW + q], gq = gb[3 * HW + q];
                    const float dq = fabsf(mq - m) - fabsf(gq - gm);
                    gc -= 2.0f * dq * sgnf(mq - m);
                }
            g += k_cont * gc;
        }
        go[3 * HW + i] = g;
    }
}

// part_fwd [B][np][4] -> img_fwd [B][4] (fixed point); one warp per image, fixed order
__global__ void k_image_reduce(int np, const float* __restrict__ part_fwd, long long* __restrict__ img_fwd)
{
    mm_pdl_prologue();
    const int b = blockIdx.x, lane = threadIdx.x;
    float a[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int k = lane; k < np; k += 32)
        #pragma unroll
        for (int i = 0; i < 4; ++i) a[i] += part_fwd[((size_t)b * np + k) * 4 + i];
    #pragma unroll
    for (int i = 0; i < 4; ++i) {
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
        if (lane == 0) img_fwd[b * 4 + i] = __double2ll_rn((double)a[i] * MM_FX_LOSS);
    }
}

// loss[0..3] = data, image, mask (1 - mean IoU), contour term, from the per-image sums.  Single CTA, fixed order.
__global__ void __launch_bounds__(MM_THREADS)
k_loss_finalize(int B, int H, int W, float image_weight, float contour,
                const long long* __restrict__ img_fwd, long long* __restrict__ img_bwd,
                float* __restrict__ loss, float* __restrict__ iou_out)
{
    mm_pdl_prologue();
    __shared__ float red[MM_WARPS];
    float a_l1 = 0.0f, a_c = 0.0f, a_iou = 0.0f;
    for (int b = threadIdx.x; b < B; b += MM_THREADS) {
        a_l1 += fx_get(img_fwd + b * 4 + 0, MM_FX_LOSS);
        a_c += fx_get(img_fwd + b * 4 + 3, MM_FX_LOSS);
        if (img_bwd) {          // fused path: the contour sum came from the shading backward; leave the workspace reusable
            a_c += fx_get(img_bwd + b * 12 + 0, MM_FX_LOSS);
            for (int i = 0; i < 12; ++i) img_bwd[b * 12 + i] = 0;
        }
        const float n = fx_get(img_fwd + b * 4 + 1, MM_FX_LOSS), d = fx_get(img_fwd + b * 4 + 2, MM_FX_LOSS);
        a_iou += n / (d + 1e-10f);
        if (iou_out) { iou_out[b * 2] = n; iou_out[b * 2 + 1] = d; }
    }
    const float s_l1 = block_sum(a_l1, red), s_c = block_sum(a_c, red), s_iou = block_sum(a_iou, red);
    if (threadIdx.x == 0) {
        const float l_img = s_l1 / ((float)B * 3.0f * (float)H * (float)W);
        const float l_iou = 1.0f - s_iou / (float)B;
        const float l_cont = (contour > 0.0f) ? s_c / ((float)B * (float)H * (float)W) : 0.0f;
        const float l_mask = l_iou + ((contour > 0.0f) ? l_cont * contour : 0.0f);
        loss[0] = image_weight * l_img + l_mask;
        loss[1] = l_img;
        loss[2] = l_iou;
        loss[3] = l_cont;
    }
}

}  // namespace

void mm_launch_recon_fwd(const mm_ctx* c, int B, const float* pred, const float* gt, float contour, float* part_fwd,
                         cudaStream_t s)
{
    const dim3 grid(c->nparts_recon, B);
    mm_launch(k_recon_fwd, grid, dim3(MM_THREADS), 0, s, false, c->H, c->W, c->nparts_recon, contour, pred, gt, (const int32_t*)c->d_tab, part_fwd);
}

void mm_launch_recon_bwd(const mm_ctx* c, int B, const float* pred, const float* gt, const long long* img_fwd,
                         float image_weight, float contour, float loss_scale, float* g_pred, cudaStream_t s)
{
    const dim3 grid(c->nparts_recon, B);
    mm_launch(k_recon_bwd, grid, dim3(MM_THREADS), 0, s, g_mm_pdl != 0, B, c->H, c->W, c->nparts_recon, image_weight, contour, loss_scale,
              pred, gt, (const int32_t*)c->d_tab, img_fwd, g_pred);
}

void mm_launch_loss_finalize(const mm_ctx* c, int B, const long long* img_fwd, long long* img_bwd,
                             float image_weight, float contour, float* loss, float* iou_out, cudaStream_t s)
{
    mm_launch(k_loss_finalize, dim3(1), dim3(MM_THREADS), 0, s, g_mm_pdl != 0, B, c->H, c->W, image_weight, contour, img_fwd, img_bwd, loss, iou_out);
}

void mm_launch_image_reduce(const mm_ctx* c, int B, int np, const float* part_fwd, long long* img_fwd, cudaStream_t s)
{
    (void)c;
    mm_launch(k_image_reduce, dim3(B), dim3(32), 0, s, g_mm_pdl != 0, np, part_fwd, img_fwd);
}
