// mm_raster.cu -- per-pixel stage: DIB-R hard visibility + soft silhouette + UV
// texture sampling + SH lighting + composite (+ loss partial sums), forward and
// backward.  Replaces kaolin dibr_rasterization / texture_mapping /
// spherical_harmonic_lighting and the ~40 elementwise torch kernels of
// networks.py:297-317, and their autograd.
//
// Work decomposition (B200: 148 SMs):
//   grid = (nparts, B); one CTA = 8 warps = 8 consecutive sub-tiles of one image, one
//   warp per sub-tile (8x4 pixels, one lane per pixel).
//     1. The CTA stages the bitmask rows ("tile face lists") of its 8 sub-tiles -- one
//        contiguous block per mask -- into shared memory with TMA bulk copies
//        (cp.async.bulk + mbarrier).  The masks were produced by the vertex stage.
//     2. Hard pass: the warp walks the set bits of its H mask in index order; all 32
//        lanes visit the same face at the same time (the 48-byte face record is one
//        broadcast load served by L1) and test their own pixel.  Bit-exact DIB-R
//        arithmetic (see mm_device.cuh).
//     3. Soft pass, only if some lane is uncovered: phase A walks the S mask and lets
//        every uncovered lane append the faces whose enlarged bbox contains ITS pixel
//        to a private list in shared memory (stops at knum, so DIB-R's order-dependent
//        truncation falls out for free); phase B lets every lane evaluate only its own
//        list -- the expensive distance/exp code runs on (pixel, face) pairs that
//        matter instead of on every face of the sub-tile for every lane.
//   Backward re-derives the same per-pixel state from `face_idx` (saved) and the same
//   masks instead of storing Kaolin's knum-deep side buffers
//   (B*H*W*30*(4+8+1) B = 307 MB at B=48,128^2).
#include "mm_device.cuh"

namespace {

#define FULL 0xffffffffu

struct CtaCtx {
    const uint32_t* mS;     // this warp's S mask row (shared memory)
    const uint32_t* mH;     // this warp's H mask row (shared memory)
    uint16_t* slist;        // [knum][MM_THREADS] per-lane soft candidate lists
    float* lights;          // 9
    float* red;             // MM_WARPS
    const float* rec;       // face records of this image (global, read through L1)
    int st, ix, iy;
    bool st_valid, active;
};

// dynamic smem: | mbarrier 16 B | maskS 8*nwords*4 | maskH 8*nwords*4 | slist knum*256*2 |
__host__ __device__ inline size_t raster_smem(int nwords, int knum) {
    return 16 + 2 * (size_t)MM_WARPS * nwords * 4 + (size_t)knum * MM_THREADS * 2;
}

__device__ __forceinline__ void cta_prologue(const mm_raster_params& p, unsigned char* smem, float* s_lights, float* s_red,
                                             CtaCtx& c)
{
    const int b = blockIdx.y, g = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint32_t* smS = reinterpret_cast<uint32_t*>(smem + 16);
    uint32_t* smH = smS + MM_WARPS * p.nwords;
    c.slist = reinterpret_cast<uint16_t*>(smH + MM_WARPS * p.nwords);
    c.lights = s_lights; c.red = s_red;
    c.rec = p.frec + (size_t)b * p.F * MM_REC_FLOATS;
    const int st0 = g * MM_WARPS;
    const int nsub = min(MM_WARPS, p.nst - st0);
    if (threadIdx.x == 0) mbar_init(bar, 1);
    if (threadIdx.x < 9) s_lights[threadIdx.x] = p.lights[b * 9 + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)nsub * p.nwords * 4;
        const size_t off = ((size_t)b * p.nst + st0) * p.nwords;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(2 * bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smS)), "l"(p.maskS + off), "r"(bytes), "r"(smem_u32(bar)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smH)), "l"(p.maskH + off), "r"(bytes), "r"(smem_u32(bar)) : "memory");
    }
    c.st = st0 + warp;
    c.st_valid = c.st < p.nst;
    const int sty = c.st / p.nstx, stx = c.st - sty * p.nstx;
    c.ix = stx * MM_ST_W + (lane & 7);
    c.iy = sty * MM_ST_H + (lane >> 3);
    c.active = c.st_valid && (c.ix < p.W) && (c.iy < p.H);
    c.mS = smS + warp * p.nwords;
    c.mH = smH + warp * p.nwords;
    mbar_wait(bar, 0);
}

// Walks the set bits of one sub-tile mask row in face-index order; fn(f) is warp-uniform.
template <typename Fn>
__device__ __forceinline__ void for_each_face(const uint32_t* row, int nwords, int lane, Fn fn)
{
    for (int wd0 = 0; wd0 < nwords; wd0 += 32) {
        const uint32_t w = (wd0 + lane < nwords) ? row[wd0 + lane] : 0u;
        uint32_t nz = __ballot_sync(FULL, w != 0u);
        while (nz) {
            const int src = __ffs(nz) - 1;
            nz &= nz - 1;
            uint32_t m = __shfl_sync(FULL, w, src);
            const int base = (wd0 + src) << 5;
            while (m) {
                const int f = base + __ffs(m) - 1;
                m &= m - 1;
                fn(f);
            }
        }
    }
}

__device__ __forceinline__ bool mask_empty(const uint32_t* row, int nwords, int lane)
{
    uint32_t any = 0u;
    for (int wd0 = 0; wd0 < nwords; wd0 += 32) any |= (wd0 + lane < nwords) ? row[wd0 + lane] : 0u;
    return __ballot_sync(FULL, any != 0u) == 0u;
}

// Hard pass (DIBR_SPEC A.2)
__device__ __forceinline__ void hard_pass(const mm_raster_params& p, const CtaCtx& c, int lane, float x0, float y0,
                                          int& best_f, float& bw0, float& bw1, float& bw2)
{
    float best_z = -INFINITY;
    best_f = -1; bw0 = bw1 = bw2 = 0.0f;
    for_each_face(c.mH, p.nwords, lane, [&](int f) {
        const FaceRec r = load_rec(c.rec, f);
        float w0, w1, w2, zz;
        if (hard_test(r, x0, y0, p.eps, w0, w1, w2, zz)) {
            if (!(zz <= best_z)) { best_z = zz; best_f = f; bw0 = w0; bw1 = w1; bw2 = w2; }
        }
    });
}

// Soft pass phase A: per-lane candidate list (first knum faces, in index order, whose enlarged bbox holds the pixel)
__device__ __forceinline__ int soft_collect(const mm_raster_params& p, const CtaCtx& c, int lane, bool need, float x0, float y0)
{
    int cnt = 0;
    uint16_t* mine = c.slist + threadIdx.x;
    for_each_face(c.mS, p.nwords, lane, [&](int f) {
        const FaceRec r = load_rec(c.rec, f);
        if (need && cnt < p.knum && soft_bbox_test(r, x0, y0, p.blen)) {
            mine[cnt * MM_THREADS] = (uint16_t)f;
            ++cnt;
        }
    });
    return cnt;
}

// ---------------------------------------------------------------------------------------------- forward
template <bool WITH_LOSS>
__global__ void __launch_bounds__(MM_THREADS)
k_raster_fwd(const mm_raster_params p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ float s_lights[16];
    __shared__ float s_red[MM_WARPS];
    CtaCtx c;
    cta_prologue(p, smem, s_lights, s_red, c);
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const size_t HW = (size_t)p.H * p.W;
    float acc_l1 = 0.0f, acc_n = 0.0f, acc_d = 0.0f;

    if (c.st_valid) {
        const float x0 = pix_x(c.ix, p.W, p.sx), y0 = pix_y(c.iy, p.H, p.sy);
        int best_f = -1;
        float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f, soft = 0.0f;
        if (!mask_empty(c.mS, p.nwords, lane)) {
            hard_pass(p, c, lane, x0, y0, best_f, w0, w1, w2);
            const bool need_soft = c.active && (best_f < 0);
            if (__any_sync(FULL, need_soft)) {
                const int cnt = soft_collect(p, c, lane, need_soft, x0, y0);
                float allprob = 1.0f;
                const uint16_t* mine = c.slist + threadIdx.x;
                for (int k = 0; k < cnt; ++k) {
                    const FaceRec r = load_rec(c.rec, (int)mine[k * MM_THREADS]);
                    int type;
                    const float d2 = soft_d2(r, x0, y0, p.multiplier, type);
                    allprob = allprob * (1.0f - soft_prob(d2, p.sigmainv, p.multiplier));
                }
                soft = 1.0f - allprob;
            }
            if (best_f >= 0) soft = 1.0f;
        }

        if (c.active) {
            const size_t pix = (size_t)c.iy * p.W + c.ix;
            // ---- shading (networks.py:303-314)
            float tm = 0.0f, nrm[3] = {0.0f, 0.0f, 0.0f}, tcol[3] = {0.0f, 0.0f, 0.0f};
            if (best_f >= 0) {
                const float* uvp = p.face_uvs + best_f * 6;
                // interpolation in the rasteriser's operation order (w0*c0 + w1*c1) + w2*c2, uncontracted
                const float u = interp3(w0, w1, w2, __ldg(uvp + 0), __ldg(uvp + 2), __ldg(uvp + 4));
                const float v = interp3(w0, w1, w2, __ldg(uvp + 1), __ldg(uvp + 3), __ldg(uvp + 5));
                const FaceRec r = load_rec(c.rec, best_f);
                tm = ADD(ADD(w0, w1), w2);
                nrm[0] = interp3(w0, w1, w2, r.nx, r.nx, r.nx);
                nrm[1] = interp3(w0, w1, w2, r.ny, r.ny, r.ny);
                nrm[2] = interp3(w0, w1, w2, r.nz, r.nz, r.nz);
                Bilin bl;
                bilin_setup(u, v, p.Ht, p.Wt, bl);
                const float* tb = p.tex + (size_t)b * 3 * p.Ht * p.Wt;
                #pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const TexFetch t = tex_fetch(tb + (size_t)ch * p.Ht * p.Wt, bl, p.Ht, p.Wt);
                    tcol[ch] = t.nw * bl.nw + t.ne * bl.ne + t.sw * bl.sw + t.se * bl.se;
                }
            }
            float bnd[9];
            sh_bands(nrm[0], nrm[1], nrm[2], bnd);
            const float coef = sh_coef(bnd, c.lights);
            float img[3];
            #pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float v;
                if (p.no_mask) {
                    const float bgc = __ldg(p.bg + ((size_t)b * 3 + ch) * HW + pix);
                    v = (tcol[ch] * tm + bgc * (1.0f - tm)) * coef;
                } else {
                    v = tcol[ch] * tm * coef + (1.0f - tm);
                }
                img[ch] = clamp01(v);
            }
            float* out = p.rgba + (size_t)b * 4 * HW + pix;
            out[0] = img[0]; out[HW] = img[1]; out[2 * HW] = img[2]; out[3 * HW] = soft;
            p.face_idx_ws[(size_t)b * HW + pix] = best_f;
            if (p.face_idx_out) p.face_idx_out[(size_t)b * HW + pix] = best_f;
            if (p.imnormal) {
                float* no = p.imnormal + ((size_t)b * HW + pix) * 3;
                no[0] = nrm[0]; no[1] = nrm[1]; no[2] = nrm[2];
            }
            if (WITH_LOSS) {
                const float* g = p.gt + (size_t)b * 4 * HW + pix;
                const float gm = __ldg(g + 3 * HW);
                #pragma unroll
                for (int ch = 0; ch < 3; ++ch) acc_l1 += fabsf(l1_term(img[ch], __ldg(g + ch * HW), gm));
                const float mul = soft * gm;
                acc_n += mul;
                acc_d += (soft + gm) - mul;
            }
        }
    }
    if (WITH_LOSS) {
        const float s0 = block_sum(acc_l1, c.red);
        const float s1 = block_sum(acc_n, c.red);
        const float s2 = block_sum(acc_d, c.red);
        if (threadIdx.x == 0) {
            float* pf = p.part_fwd + ((size_t)b * p.nparts + blockIdx.x) * 4;
            pf[0] = s0; pf[1] = s1; pf[2] = s2; pf[3] = 0.0f;
        }
    }
}

// ---------------------------------------------------------------------------------------------- backward
__device__ __forceinline__ float contour_c(float m, float mref) { return fabsf(m - mref); }

__global__ void __launch_bounds__(MM_THREADS)
k_raster_bwd(const mm_raster_params p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ float s_lights[16];
    __shared__ float s_red[MM_WARPS];
    __shared__ float s_iou[2];
    CtaCtx c;
    cta_prologue(p, smem, s_lights, s_red, c);
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const size_t HW = (size_t)p.H * p.W;
    const int H = p.H, W = p.W;
    const int32_t* refrow = p.tab;
    const int32_t* rowlo = p.tab + H;
    const int32_t* rowhi = p.tab + 2 * H;
    const int32_t* refcol = p.tab + 3 * H;
    const int32_t* collo = p.tab + 3 * H + W;
    const int32_t* colhi = p.tab + 3 * H + 2 * W;

    float acc_contour = 0.0f;
    float acc_l[9];
    #pragma unroll
    for (int i = 0; i < 9; ++i) acc_l[i] = 0.0f;

    // loss-gradient constants; the per-image IoU sums are re-derived from the forward partials in a fixed
    // order by warp 0 (every CTA of the image gets bit-identical sums)
    float k_img = 0.0f, k_iou = 0.0f, k_cont = 0.0f, Nb = 0.0f, De = 1.0f;
    if (p.analytic_loss) {
        if (threadIdx.x < 32) {
            float n = 0.0f, d = 0.0f;
            for (int k = lane; k < p.nparts; k += 32) {
                n += p.part_fwd_in[((size_t)b * p.nparts + k) * 4 + 1];
                d += p.part_fwd_in[((size_t)b * p.nparts + k) * 4 + 2];
            }
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) { n += __shfl_xor_sync(FULL, n, o); d += __shfl_xor_sync(FULL, d, o); }
            if (lane == 0) { s_iou[0] = n; s_iou[1] = d; }
        }
        __syncthreads();
        k_img = p.loss_scale * p.image_weight / ((float)p.B * 3.0f * (float)HW);
        k_iou = p.loss_scale / (float)p.B;
        k_cont = p.loss_scale * p.contour / ((float)p.B * (float)HW);
        Nb = s_iou[0];
        De = s_iou[1] + 1e-10f;
    }
    const float* rg = p.rgba + (size_t)b * 4 * HW;         // forward output (silhouette re-read)
    const float* gtb = p.gt ? p.gt + (size_t)b * 4 * HW : nullptr;
    const float* gup = p.g_rgba ? p.g_rgba + (size_t)b * 4 * HW : nullptr;
    float* gacc = p.gfacc + (size_t)b * p.F * 9;
    float* gtex = p.g_tex + (size_t)b * 3 * p.Ht * p.Wt;

    if (c.st_valid) {
        const int ix = c.ix, iy = c.iy;
        const bool active = c.active;
        const float x0 = pix_x(ix, W, p.sx), y0 = pix_y(iy, H, p.sy);
        const size_t pix = active ? (size_t)iy * W + ix : 0;

        const int best_f = active ? p.face_idx_ws[(size_t)b * HW + pix] : -2;   // -2: inactive lane
        // ---- upstream gradient of the 4 output channels
        float g_img[3] = {0.0f, 0.0f, 0.0f}, g_soft = 0.0f;
        float soft = 0.0f;
        if (active) {
            soft = rg[3 * HW + pix];
            if (gup) { g_img[0] = gup[pix]; g_img[1] = gup[HW + pix]; g_img[2] = gup[2 * HW + pix]; g_soft = gup[3 * HW + pix]; }
            if (p.analytic_loss) {
                const float gm = __ldg(gtb + 3 * HW + pix);
                #pragma unroll
                for (int ch = 0; ch < 3; ++ch)
                    g_img[ch] += k_img * sgnf(l1_term(rg[ch * HW + pix], __ldg(gtb + ch * HW + pix), gm)) * gm;
                // soft IoU: -(1/B) * (gm*De - Nb*(1-gm)) / De^2
                g_soft += -k_iou * (gm * De - Nb * (1.0f - gm)) / (De * De);
                if (p.contour > 0.0f) {
                    const int ry = refrow[iy], rx = refcol[ix];
                    const size_t rp = (size_t)ry * W + rx;
                    const float mref = rg[3 * HW + rp], gref = __ldg(gtb + 3 * HW + rp);
                    const float dlt = contour_c(soft, mref) - contour_c(gm, gref);
                    acc_contour += dlt * dlt;
                    float gc = 2.0f * dlt * sgnf(soft - mref);
                    // this pixel may itself be the reference of a block of pixels
                    const int y_lo = rowlo[iy], y_hi = rowhi[iy], x_lo = collo[ix], x_hi = colhi[ix];
                    for (int yy = y_lo; yy < y_hi; ++yy)
                        for (int xx = x_lo; xx < x_hi; ++xx) {
                            const size_t q = (size_t)yy * W + xx;
                            const float mq = rg[3 * HW + q], gq = __ldg(gtb + 3 * HW + q);
                            const float dq = contour_c(mq, soft) - contour_c(gq, gm);
                            gc -= 2.0f * dq * sgnf(mq - soft);
                        }
                    g_soft += k_cont * gc;
                }
            }
        }

        // ---- soft silhouette backward (DIBR_SPEC A.5): uncovered pixels only
        const bool need_soft = active && (best_f == -1) && (g_soft != 0.0f) && (soft > 0.0f);
        if (__any_sync(FULL, need_soft)) {
            const int cnt = soft_collect(p, c, lane, need_soft, x0, y0);
            const float one_m_all = 1.0f - soft;
            const uint16_t* mine = c.slist + threadIdx.x;
            for (int k = 0; k < cnt; ++k) {
                const int f = (int)mine[k * MM_THREADS];
                const FaceRec r = load_rec(c.rec, f);
                int type;
                const float d2s = soft_d2(r, x0, y0, p.multiplier, type);
                const float prob = soft_prob(d2s, p.sigmainv, p.multiplier);
                // dLdz = -sigmainv * dLdp * (1-allprob) / (1-prob+1e-6) * prob
                const float dLdz = MUL(DIV(MUL(MUL(MUL(-1.0f, p.sigmainv), g_soft), one_m_all),
                                           ADD(SUB(1.0f, prob), 1e-6f)), prob);
                float* g = gacc + (size_t)f * 9;
                if (type >= 3) {
                    const int i = type - 3;
                    const float x1 = (i == 0) ? r.ax : ((i == 1) ? r.bx : r.cx);
                    const float y1 = (i == 0) ? r.ay : ((i == 1) ? r.by : r.cy);
                    atomicAdd(g + 2 * i,     DIV(MUL(MUL(dLdz, 2.0f), SUB(x1, x0)), p.multiplier));
                    atomicAdd(g + 2 * i + 1, DIV(MUL(MUL(dLdz, 2.0f), SUB(y1, y0)), p.multiplier));
                } else {
                    const int i = type, j = (type + 1) % 3;
                    const float x1 = (i == 0) ? r.ax : ((i == 1) ? r.bx : r.cx);
                    const float y1 = (i == 0) ? r.ay : ((i == 1) ? r.by : r.cy);
                    const float x2 = (j == 0) ? r.ax : ((j == 1) ? r.bx : r.cx);
                    const float y2 = (j == 0) ? r.ay : ((j == 1) ? r.by : r.cy);
                    const float A = SUB(y2, y1), Bc = SUB(x1, x2), C = SUB(MUL(x2, y1), MUL(x1, y2));
                    const float up = ADD(ADD(MUL(A, x0), MUL(Bc, y0)), C);
                    const float dn = ADD(ADD(MUL(A, A), MUL(Bc, Bc)), 1e-10f);
                    const float d2 = DIV(MUL(up, up), dn);
                    const float dzdA = DIV(MUL(2.0f, SUB(MUL(x0, up), MUL(d2, A))), dn);
                    const float dzdB = DIV(MUL(2.0f, SUB(MUL(y0, up), MUL(d2, Bc))), dn);
                    const float dzdC = DIV(MUL(2.0f, up), dn);
                    atomicAdd(g + 2 * i,     DIV(MUL(dLdz, SUB(dzdB, MUL(y2, dzdC))), p.multiplier));
                    atomicAdd(g + 2 * i + 1, DIV(MUL(dLdz, SUB(MUL(x2, dzdC), dzdA)), p.multiplier));
                    atomicAdd(g + 2 * j,     DIV(MUL(dLdz, SUB(MUL(y1, dzdC), dzdB)), p.multiplier));
                    atomicAdd(g + 2 * j + 1, DIV(MUL(dLdz, SUB(dzdA, MUL(x1, dzdC))), p.multiplier));
                }
            }
        }

        if (active) {
            // ---- shading backward
            float tm = 0.0f, nrm[3] = {0.0f, 0.0f, 0.0f}, tcol[3] = {0.0f, 0.0f, 0.0f};
            FaceRec r;
            Bary bar;
            Bilin bl;
            TexFetch tf[3];
            float uv[6];
            if (best_f >= 0) {
                r = load_rec(c.rec, best_f);
                bary_eval(r, x0, y0, p.eps, bar);
                const float* uvp = p.face_uvs + best_f * 6;
                #pragma unroll
                for (int i = 0; i < 6; ++i) uv[i] = __ldg(uvp + i);
                const float u = interp3(bar.w0, bar.w1, bar.w2, uv[0], uv[2], uv[4]);
                const float v = interp3(bar.w0, bar.w1, bar.w2, uv[1], uv[3], uv[5]);
                tm = ADD(ADD(bar.w0, bar.w1), bar.w2);
                nrm[0] = interp3(bar.w0, bar.w1, bar.w2, r.nx, r.nx, r.nx);
                nrm[1] = interp3(bar.w0, bar.w1, bar.w2, r.ny, r.ny, r.ny);
                nrm[2] = interp3(bar.w0, bar.w1, bar.w2, r.nz, r.nz, r.nz);
                bilin_setup(u, v, p.Ht, p.Wt, bl);
                const float* tb = p.tex + (size_t)b * 3 * p.Ht * p.Wt;
                #pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    tf[ch] = tex_fetch(tb + (size_t)ch * p.Ht * p.Wt, bl, p.Ht, p.Wt);
                    tcol[ch] = tf[ch].nw * bl.nw + tf[ch].ne * bl.ne + tf[ch].sw * bl.sw + tf[ch].se * bl.se;
                }
            }
            float bnd[9];
            sh_bands(nrm[0], nrm[1], nrm[2], bnd);
            const float coef = sh_coef(bnd, c.lights);
            float g_coef = 0.0f, g_tcol[3];
            #pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float pre, bgc = 0.0f;
                if (p.no_mask) {
                    bgc = __ldg(p.bg + ((size_t)b * 3 + ch) * HW + pix);
                    pre = (tcol[ch] * tm + bgc * (1.0f - tm)) * coef;
                } else {
                    pre = tcol[ch] * tm * coef + (1.0f - tm);
                }
                const float g = (pre >= 0.0f && pre <= 1.0f) ? g_img[ch] : 0.0f;     // torch.clamp backward
                g_tcol[ch] = g * tm * coef;
                if (p.no_mask) {
                    g_coef += g * (tcol[ch] * tm + bgc * (1.0f - tm));
                    if (p.g_bg) p.g_bg[((size_t)b * 3 + ch) * HW + pix] = g * (1.0f - tm) * coef;
                } else {
                    g_coef += g * (tcol[ch] * tm);
                }
            }
            #pragma unroll
            for (int i = 0; i < 9; ++i) acc_l[i] += g_coef * bnd[i];

            if (best_f >= 0) {
                // texture gradient + d/d(u,v)
                float gix = 0.0f, giy = 0.0f;
                const bool xe = (bl.ix + 1) < p.Wt, ys = (bl.iy + 1) < p.Ht;
                const float tx = bl.x - (float)bl.ix, ty = bl.y - (float)bl.iy;
                #pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const float g = g_tcol[ch];
                    if (g != 0.0f) {
                        float* gp = gtex + ((size_t)ch * p.Ht + bl.iy) * p.Wt + bl.ix;
                        atomicAdd(gp, g * bl.nw);
                        if (xe) atomicAdd(gp + 1, g * bl.ne);
                        if (ys) atomicAdd(gp + p.Wt, g * bl.sw);
                        if (xe && ys) atomicAdd(gp + p.Wt + 1, g * bl.se);
                        gix += g * ((tf[ch].ne - tf[ch].nw) * (1.0f - ty) + (tf[ch].se - tf[ch].sw) * ty);
                        giy += g * ((tf[ch].sw - tf[ch].nw) * (1.0f - tx) + (tf[ch].se - tf[ch].ne) * tx);
                    }
                }
                const float g_gx = bl.in_x ? gix * ((float)p.Wt * 0.5f) : 0.0f;
                const float g_gy = bl.in_y ? giy * ((float)p.Ht * 0.5f) : 0.0f;
                const float g_u = 2.0f * g_gx, g_v = -2.0f * g_gy;

                // d coef / d normal -> unit face normal (features are the same normal on 3 corners)
                const float* l = c.lights;
                const float nx = nrm[0], ny = nrm[1], nz = nrm[2];
                const float dcx = l[1] * SH_C1 + l[4] * SH_C2 * ny + l[7] * SH_C4 * nz + l[8] * SH_C5 * 2.0f * nx;
                const float dcy = l[3] * SH_C1 + l[4] * SH_C2 * nx + l[5] * SH_C2 * nz - l[8] * SH_C5 * 2.0f * ny;
                const float dcz = l[2] * SH_C1 + l[5] * SH_C2 * ny + l[6] * SH_C3 * 2.0f * nz + l[7] * SH_C4 * nx;
                float* g = gacc + (size_t)best_f * 9;
                const float gn_scale = g_coef * tm;    // sum_i w_i * g_n
                if (gn_scale != 0.0f) {
                    atomicAdd(g + 6, gn_scale * dcx);
                    atomicAdd(g + 7, gn_scale * dcy);
                    atomicAdd(g + 8, gn_scale * dcz);
                }
                // hard rasteriser backward (DIBR_SPEC A.3) for the u,v channels
                if (g_u != 0.0f || g_v != 0.0f) {
                    const float k1 = bar.k1, k2 = bar.k2, k3 = bar.k3;
                    const float m = bar.m, pp = bar.p, n = bar.n, q = bar.q, s = bar.s, t = bar.t;
                    // numerators of dw1/d(.) and dw2/d(.) (common 1/k3^2 applied in dldI)
                    const float dw1dm = SUB(MUL(0.0f, k3), MUL(q, k1)),   dw1dn = SUB(MUL(-t, k3), MUL(-pp, k1));
                    const float dw1dp = SUB(MUL(0.0f, k3), MUL(-n, k1)),  dw1dq = SUB(MUL(s, k3), MUL(m, k1));
                    const float dw1ds = SUB(MUL(q, k3), MUL(0.0f, k1)),   dw1dt = SUB(MUL(-n, k3), MUL(0.0f, k1));
                    const float dw2dm = SUB(MUL(t, k3), MUL(q, k2)),      dw2dn = SUB(MUL(0.0f, k3), MUL(-pp, k2));
                    const float dw2dp = SUB(MUL(-s, k3), MUL(-n, k2)),    dw2dq = SUB(MUL(0.0f, k3), MUL(m, k2));
                    const float dw2ds = SUB(MUL(-pp, k3), MUL(0.0f, k2)), dw2dt = SUB(MUL(m, k3), MUL(0.0f, k2));
                    const float dw1dax = -ADD(ADD(dw1dm, dw1dn), dw1ds), dw1day = -ADD(ADD(dw1dp, dw1dq), dw1dt);
                    const float dw2dax = -ADD(ADD(dw2dm, dw2dn), dw2ds), dw2day = -ADD(ADD(dw2dp, dw2dq), dw2dt);
                    const float den = ADD(MUL(k3, k3), p.eps);
                    float gv[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
                    #pragma unroll
                    for (int d = 0; d < 2; ++d) {
                        const float gd = d == 0 ? g_u : g_v;
                        const float c0 = uv[d], c1 = uv[2 + d], c2 = uv[4 + d];
                        const float e1 = SUB(c1, c0), e2 = SUB(c2, c0);
                        const float dldI = DIV(MUL(p.multiplier, gd), den);
                        gv[0] += MUL(dldI, ADD(MUL(e1, dw1dax), MUL(e2, dw2dax)));
                        gv[1] += MUL(dldI, ADD(MUL(e1, dw1day), MUL(e2, dw2day)));
                        gv[2] += MUL(dldI, ADD(MUL(e1, dw1dm), MUL(e2, dw2dm)));
                        gv[3] += MUL(dldI, ADD(MUL(e1, dw1dp), MUL(e2, dw2dp)));
                        gv[4] += MUL(dldI, ADD(MUL(e1, dw1dn), MUL(e2, dw2dn)));
                        gv[5] += MUL(dldI, ADD(MUL(e1, dw1dq), MUL(e2, dw2dq)));
                    }
                    #pragma unroll
                    for (int i = 0; i < 6; ++i) atomicAdd(g + i, gv[i]);
                }
            }
        }
    }

    // ---- per-CTA partials: contour sum + 9 light gradients (summed deterministically later)
    float* pb = p.part_bwd + ((size_t)b * p.nparts + blockIdx.x) * 12;
    {
        const float s = block_sum(acc_contour, c.red);
        if (threadIdx.x == 0) pb[0] = s;
    }
    #pragma unroll
    for (int i = 0; i < 9; ++i) {
        const float s = block_sum(acc_l[i], c.red);
        if (threadIdx.x == 0) pb[1 + i] = s;
    }
}

}  // namespace

size_t mm_raster_smem_bytes(const mm_ctx* c) { return raster_smem(c->nwords, c->knum); }

cudaError_t mm_raster_configure(const mm_ctx* c) {
    const int bytes = (int)raster_smem(c->nwords, c->knum);
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_raster_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
    if ((e = cudaFuncSetAttribute(k_raster_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
    if ((e = cudaFuncSetAttribute(k_raster_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
    return cudaSuccess;
}

void mm_launch_raster_fwd(const mm_ctx* c, const mm_raster_params& p, bool with_loss, cudaStream_t s)
{
    const dim3 grid(c->nparts, p.B);
    const size_t smem = raster_smem(c->nwords, c->knum);
    if (with_loss) k_raster_fwd<true><<<grid, MM_THREADS, smem, s>>>(p);
    else           k_raster_fwd<false><<<grid, MM_THREADS, smem, s>>>(p);
}

void mm_launch_raster_bwd(const mm_ctx* c, const mm_raster_params& p, cudaStream_t s)
{
    const dim3 grid(c->nparts, p.B);
    k_raster_bwd<<<grid, MM_THREADS, raster_smem(c->nwords, c->knum), s>>>(p);
}
