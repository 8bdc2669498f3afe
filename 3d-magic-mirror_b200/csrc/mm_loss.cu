// mm_loss.cu -- DiffRender.recon_data (networks.py:364-390) as stand-alone kernels: masked L1 + kaolin mask_iou + contour MSE.
//
//   k_recon_fwd   ONE launch: per-CTA partial sums (L1, N, D, contour) -> per-image fixed-point integer atomics (order
//                 independent: the loss is deterministic) -> the last CTA to finish (ticket) reduces the images in index
//                 order and writes the four loss scalars.  The per-image IoU sums stay in the workspace (`img_fwd`): when
//                 `pred` is the output of mm_render_forward and the call is made on THAT render's workspace, the render
//                 backward forms the loss gradient in-kernel from them (lazy fusion, see mm_render_backward).
//   k_recon_bwd   the materialised gradient d(loss_scale * loss)/d(pred), for a `pred` that did not come from this library.
#include "mm_device.cuh"

namespace {

struct ReconParams {
    int B, H, W, nparts;
    float image_weight, contour, loss_scale;
    const float* loss_scale_dev;
    const float* pred;
    const float* gt;
    const int32_t* tab;
    long long* img_fwd;      // [B,4] fixed point
    unsigned* ticket;
    float* loss;             // [4]
    float* iou_out;          // [B,2] or NULL
    float* g_pred;           // backward only
};

template <bool VEC>
__global__ void __launch_bounds__(MM_THREADS)
k_recon_fwd(const ReconParams q)
{
    mm_pdl_prologue();
    __shared__ float red[MM_WARPS];
    __shared__ int s_last;
    const int b = blockIdx.y, band = blockIdx.x;
    const int H = q.H, W = q.W;
    const size_t HW = (size_t)H * W;
    const float* pb = q.pred + (size_t)b * 4 * HW;
    const float* gb = q.gt + (size_t)b * 4 * HW;
    const int32_t* refrow = q.tab;
    const int32_t* refcol = q.tab + 3 * H;
    float a_l1 = 0.0f, a_n = 0.0f, a_d = 0.0f, a_c = 0.0f;
    if (VEC) {
        // 4 consecutive pixels of a row per thread, 16-byte loads (W % 4 == 0: a quad never straddles a row)
        const int nq = (int)(HW >> 2);
        const int per = (nq + q.nparts - 1) / q.nparts;
        const int i0 = band * per, i1 = min(nq, i0 + per);
        for (int i = i0 + threadIdx.x; i < i1; i += MM_THREADS) {
            const size_t pix = (size_t)i * 4;
            const float4 gm4 = __ldg(reinterpret_cast<const float4*>(gb + 3 * HW + pix));
            const float4 m4 = __ldg(reinterpret_cast<const float4*>(pb + 3 * HW + pix));
            const float gm[4] = {gm4.x, gm4.y, gm4.z, gm4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w};
            #pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 p4 = __ldg(reinterpret_cast<const float4*>(pb + c * HW + pix));
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(gb + c * HW + pix));
                a_l1 += fabsf(l1_term(p4.x, g4.x, gm[0])) + fabsf(l1_term(p4.y, g4.y, gm[1])) +
                        fabsf(l1_term(p4.z, g4.z, gm[2])) + fabsf(l1_term(p4.w, g4.w, gm[3]));
            }
            float mref = 0.0f, gref = 0.0f;
            if (q.contour > 0.0f) {
                const int iy = (int)(pix / W), ix = (int)(pix - (size_t)iy * W);
                // W % 4 == 0: the quad is one 4-pixel column block, whose nearest-down/up reference column is ix (refcol[ix..ix+3]
                // are equal only when W % 4 == 0 AND the table maps blocks of 4, which holds for W a multiple of 4)
                const size_t rp = (size_t)refrow[iy] * W + refcol[ix];
                mref = __ldg(pb + 3 * HW + rp); gref = __ldg(gb + 3 * HW + rp);
            }
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float mul = m[j] * gm[j];
                a_n += mul;
                a_d += (m[j] + gm[j]) - mul;
                if (q.contour > 0.0f) {
                    const float d = fabsf(m[j] - mref) - fabsf(gm[j] - gref);
                    a_c += d * d;
                }
            }
        }
    } else {
        const int per = (H * W + q.nparts - 1) / q.nparts;               // contiguous pixel range of this CTA
        const int i0 = band * per, i1 = min(H * W, i0 + per);
        for (int i = i0 + threadIdx.x; i < i1; i += MM_THREADS) {
            const float gm = gb[3 * HW + i], m = pb[3 * HW + i];
            #pragma unroll
            for (int c = 0; c < 3; ++c) a_l1 += fabsf(l1_term(pb[c * HW + i], gb[c * HW + i], gm));
            const float mul = m * gm;
            a_n += mul;
            a_d += (m + gm) - mul;
            if (q.contour > 0.0f) {
                const int iy = i / W, ix = i - iy * W;
                const size_t rp = (size_t)refrow[iy] * W + refcol[ix];
                const float d = fabsf(m - pb[3 * HW + rp]) - fabsf(gm - gb[3 * HW + rp]);
                a_c += d * d;
            }
        }
    }
    const float s0 = block_sum(a_l1, red), s1 = block_sum(a_n, red), s2 = block_sum(a_d, red), s3 = block_sum(a_c, red);
    if (threadIdx.x == 0) {
        if (s0 != 0.0f) fx_add(q.img_fwd + b * 4 + 0, s0, MM_FX_LOSS);
        if (s1 != 0.0f) fx_add(q.img_fwd + b * 4 + 1, s1, MM_FX_LOSS);
        if (s2 != 0.0f) fx_add(q.img_fwd + b * 4 + 2, s2, MM_FX_LOSS);
        if (s3 != 0.0f) fx_add(q.img_fwd + b * 4 + 3, s3, MM_FX_LOSS);
        __threadfence();
        s_last = (atomicAdd(q.ticket, 1u) == gridDim.x * gridDim.y - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    // ---- last CTA: loss[0..3] = data, image, mask (1 - mean IoU), contour term; images in index order (deterministic)
    __threadfence();
    float a1 = 0.0f, ac = 0.0f, aiou = 0.0f;
    for (int i = threadIdx.x; i < q.B; i += MM_THREADS) {
        const volatile long long* f = q.img_fwd + i * 4;
        const float l1 = (float)((double)f[0] / MM_FX_LOSS), n = (float)((double)f[1] / MM_FX_LOSS);
        const float d = (float)((double)f[2] / MM_FX_LOSS), cc = (float)((double)f[3] / MM_FX_LOSS);
        a1 += l1; ac += cc;
        aiou += n / (d + 1e-10f);
        if (q.iou_out) { q.iou_out[i * 2] = n; q.iou_out[i * 2 + 1] = d; }
    }
    const float t1 = block_sum(a1, red), tc = block_sum(ac, red), ti = block_sum(aiou, red);
    if (threadIdx.x == 0) {
        const float npx = (float)q.B * (float)H * (float)W;
        const float l_img = t1 / (npx * 3.0f);
        const float l_iou = 1.0f - ti / (float)q.B;
        const float l_cont = (q.contour > 0.0f) ? tc / npx : 0.0f;
        q.loss[0] = q.image_weight * l_img + l_iou + ((q.contour > 0.0f) ? l_cont * q.contour : 0.0f);
        q.loss[1] = l_img;
        q.loss[2] = l_iou;
        q.loss[3] = l_cont;
        *q.ticket = 0u;
    }
}

__global__ void __launch_bounds__(MM_THREADS)
k_recon_bwd(const ReconParams q)
{
    mm_pdl_prologue();
    const int b = blockIdx.y, band = blockIdx.x;
    const int B = q.B, H = q.H, W = q.W;
    const size_t HW = (size_t)H * W;
    const float* pb = q.pred + (size_t)b * 4 * HW;
    const float* gb = q.gt + (size_t)b * 4 * HW;
    float* go = q.g_pred + (size_t)b * 4 * HW;
    const int32_t* refrow = q.tab;
    const int32_t* rowlo = q.tab + H;
    const int32_t* rowhi = q.tab + 2 * H;
    const int32_t* refcol = q.tab + 3 * H;
    const int32_t* collo = q.tab + 3 * H + W;
    const int32_t* colhi = q.tab + 3 * H + 2 * W;
    const float Nb = fx_get(q.img_fwd + b * 4 + 1, MM_FX_LOSS);
    const float De = fx_get(q.img_fwd + b * 4 + 2, MM_FX_LOSS) + 1e-10f;
    const float ls = q.loss_scale_dev ? q.loss_scale * __ldg(q.loss_scale_dev) : q.loss_scale;
    const float k_img = ls * q.image_weight / ((float)B * 3.0f * (float)HW);
    const float k_iou = ls / (float)B;
    const float k_cont = ls * q.contour / ((float)B * (float)HW);
    const int per = (H * W + q.nparts - 1) / q.nparts;
    const int i0 = band * per, i1 = min(H * W, i0 + per);
    for (int i = i0 + threadIdx.x; i < i1; i += MM_THREADS) {
        const float gm = gb[3 * HW + i], m = pb[3 * HW + i];
        #pragma unroll
        for (int c = 0; c < 3; ++c)
            go[c * HW + i] = k_img * sgnf(l1_term(pb[c * HW + i], gb[c * HW + i], gm)) * gm;
        float g = -k_iou * (gm * De - Nb * (1.0f - gm)) / (De * De);
        if (q.contour > 0.0f) {
            const int iy = i / W, ix = i - iy * W;
            const size_t rp = (size_t)refrow[iy] * W + refcol[ix];
            const float mref = pb[3 * HW + rp], gref = gb[3 * HW + rp];
            const float dlt = fabsf(m - mref) - fabsf(gm - gref);
            float gc = 2.0f * dlt * sgnf(m - mref);
            for (int yy = rowlo[iy]; yy < rowhi[iy]; ++yy)
                for (int xx = collo[ix]; xx < colhi[ix]; ++xx) {
                    const size_t r = (size_t)yy * W + xx;
                    const float mq = pb[3 * HW + r], gq = gb[3 * HW + r];
                    const float dq = fabsf(mq - m) - fabsf(gq - gm);
                    gc -= 2.0f * dq * sgnf(mq - m);
                }
            g += k_cont * gc;
        }
        go[3 * HW + i] = g;
    }
}

}  // namespace

// `img_fwd` and `ticket` must be zero when the kernel starts (the caller clears them with one memset)
cudaError_t mm_launch_recon_fwd(const mm_ctx* c, int B, const float* pred, const float* gt, float image_weight, float contour,
                                long long* img_fwd, unsigned* ticket, float* loss, float* iou_out, cudaStream_t s)
{
    ReconParams q = {};
    q.B = B; q.H = c->H; q.W = c->W; q.nparts = c->nparts_recon;
    q.image_weight = image_weight; q.contour = contour;
    q.pred = pred; q.gt = gt; q.tab = c->d_tab; q.img_fwd = img_fwd; q.ticket = ticket; q.loss = loss; q.iou_out = iou_out;
    const bool vec = (c->W & 3) == 0 && (c->H & 3) == 0 && (((uintptr_t)pred | (uintptr_t)gt) & 15) == 0;
    const dim3 grid(c->nparts_recon, B);
    // first kernel after the caller's memset: no programmatic launch across a memset node
    return mm_launch(vec ? k_recon_fwd<true> : k_recon_fwd<false>, grid, dim3(MM_THREADS), 0, s, false, q);
}

cudaError_t mm_launch_recon_bwd(const mm_ctx* c, int B, const float* pred, const float* gt, const long long* img_fwd,
                                float image_weight, float contour, float loss_scale, const float* loss_scale_dev,
                                float* g_pred, cudaStream_t s)
{
    ReconParams q = {};
    q.B = B; q.H = c->H; q.W = c->W; q.nparts = c->nparts_recon;
    q.image_weight = image_weight; q.contour = contour; q.loss_scale = loss_scale; q.loss_scale_dev = loss_scale_dev;
    q.pred = pred; q.gt = gt; q.tab = c->d_tab; q.img_fwd = const_cast<long long*>(img_fwd); q.g_pred = g_pred;
    const dim3 grid(c->nparts_recon, B);
    return mm_launch(k_recon_bwd, grid, dim3(MM_THREADS), 0, s, c->pdl != 0, q);
}
