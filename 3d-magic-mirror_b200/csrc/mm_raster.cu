// mm_raster.cu -- per-pixel stage: DIB-R hard visibility + soft silhouette + UV
// texture sampling + SH lighting + composite (+ loss partial sums), forward and
// backward.  Replaces kaolin dibr_rasterization / texture_mapping /
// spherical_harmonic_lighting and the ~40 elementwise torch kernels of
// networks.py:297-317, and their autograd.
//
// Work decomposition (B200: 148 SMs, 227 KB smem/SM):
//   grid = (bands, B).  One CTA owns a band of `st_rows` sub-tile rows of one image
//   (a sub-tile = 8x4 pixels = one warp).  The CTA
//     1. stages the image's whole face-record block (F*48 B, contiguous) into shared
//        memory with ONE TMA bulk copy (cp.async.bulk + mbarrier) -- no per-thread
//        global loads of geometry afterwards;
//     2. bins faces into per-sub-tile BITMASKS (bit f of sub-tile s set iff face f's
//        enlarged bbox can touch s): a bitmask keeps faces in index order for free,
//        which DIB-R's "first knum faces in index order" truncation needs, and needs
//        no compaction/scan;
//     3. each warp walks its sub-tiles; all 32 lanes visit the same face at the same
//        time (record reads are shared-memory broadcasts) and test their own pixel.
//   Backward re-derives the same per-pixel state from `face_idx` (saved) and the
//   same bitmasks instead of storing Kaolin's knum-deep side buffers
//   (B*H*W*30*(4+8+1) B = 307 MB at B=48,128^2).
#include "mm_device.cuh"

namespace {

struct SmemPlan {
    // byte offsets into dynamic shared memory
    size_t bar, rec, maskS, maskH, summ, lights, red, total;
};

__host__ __device__ inline SmemPlan smem_plan(int F, int nst, int nwords, int nsum, bool rec_in_smem) {
    SmemPlan s;
    size_t off = 0;
    s.bar = off;    off += 16;
    s.rec = off;    off += rec_in_smem ? (size_t)F * MM_REC_FLOATS * 4 : 0;
    s.maskS = off;  off += (size_t)nst * nwords * 4;
    s.maskH = off;  off += (size_t)nst * nwords * 4;
    s.summ = off;   off += (size_t)nst * nsum * 4;
    s.lights = off; off += 16 * 4;
    s.red = off;    off += 16 * 4;
    s.total = off;
    return s;
}

struct TileCtx {
    const float* rec;        // face records of this image (shared or global)
    uint32_t* maskS;         // [nst][nwords] all faces, enlarged bbox
    uint32_t* maskH;         // [nst][nwords] front faces, tight bbox
    uint32_t* summ;          // [nst][nsum]   non-zero words of maskS
    float* lights;           // 9
    float* red;              // MM_WARPS
    int nst, nsum, band_y0;
};

// Common prologue: stage records, bin faces.  Ends with a __syncthreads().
template <bool REC_SMEM>
__device__ __forceinline__ void tile_prologue(const mm_raster_params& p, int b, int band, unsigned char* smem, TileCtx& tc)
{
    const int nst = p.nstx * p.st_rows;
    const int nsum = (p.nwords + 31) >> 5;
    const SmemPlan sp = smem_plan(p.F, nst, p.nwords, nsum, REC_SMEM);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + sp.bar);
    float* srec = reinterpret_cast<float*>(smem + sp.rec);
    tc.maskS = reinterpret_cast<uint32_t*>(smem + sp.maskS);
    tc.maskH = reinterpret_cast<uint32_t*>(smem + sp.maskH);
    tc.summ = reinterpret_cast<uint32_t*>(smem + sp.summ);
    tc.lights = reinterpret_cast<float*>(smem + sp.lights);
    tc.red = reinterpret_cast<float*>(smem + sp.red);
    tc.nst = nst; tc.nsum = nsum;
    tc.band_y0 = band * p.st_rows * MM_ST_H;
    const float* grec = p.frec + (size_t)b * p.F * MM_REC_FLOATS;
    tc.rec = REC_SMEM ? srec : grec;

    if (REC_SMEM) {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            // one bulk copy per <= 64 KB chunk (all complete on the same mbarrier phase)
            const uint32_t total = (uint32_t)p.F * MM_REC_FLOATS * 4;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(total) : "memory");
            for (uint32_t o = 0; o < total; o += 32768u) {
                const uint32_t n = (total - o) < 32768u ? (total - o) : 32768u;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(reinterpret_cast<unsigned char*>(srec) + o)),
                               "l"(reinterpret_cast<const unsigned char*>(grec) + o), "r"(n), "r"(smem_u32(bar)) : "memory");
            }
        }
    }
    // zero the bitmasks while the copy is in flight
    const int nmask = nst * p.nwords;
    for (int i = threadIdx.x; i < nmask; i += MM_THREADS) { tc.maskS[i] = 0u; tc.maskH[i] = 0u; }
    for (int i = threadIdx.x; i < nst * nsum; i += MM_THREADS) tc.summ[i] = 0u;
    if (threadIdx.x < 9) tc.lights[threadIdx.x] = p.lights[b * 9 + threadIdx.x];
    __syncthreads();
    if (REC_SMEM) mbar_wait(bar, 0);

    // ---- binning: conservative pixel ranges of the enlarged / tight bbox -> sub-tile bits
    const float inv_sx = 1.0f / p.sx, inv_sy = 1.0f / p.sy;
    const int band_rows = p.st_rows * MM_ST_H;
    for (int f = threadIdx.x; f < p.F; f += MM_THREADS) {
        const FaceRec r = load_rec(tc.rec, f);
        const float xmin = fminf(fminf(r.ax, r.bx), r.cx), xmax = fmaxf(fmaxf(r.ax, r.bx), r.cx);
        const float ymin = fminf(fminf(r.ay, r.by), r.cy), ymax = fmaxf(fmaxf(r.ay, r.by), r.cy);
        // enlarged by blen and by half a pixel of slack (exact tests are redone per pixel)
        const float xl = xmin - p.blen, xh = xmax + p.blen, yl = ymin - p.blen, yh = ymax + p.blen;
        float fx_lo = (xl * inv_sx + (float)(p.W - 1)) * 0.5f;
        float fx_hi = (xh * inv_sx + (float)(p.W - 1)) * 0.5f;
        float fy_lo = ((float)(p.H - 1) - yh * inv_sy) * 0.5f;
        float fy_hi = ((float)(p.H - 1) - yl * inv_sy) * 0.5f;
        fx_lo = fminf(fmaxf(fx_lo, -4.0f), 1.0e6f); fx_hi = fminf(fmaxf(fx_hi, -4.0f), 1.0e6f);
        fy_lo = fminf(fmaxf(fy_lo, -4.0f), 1.0e6f); fy_hi = fminf(fmaxf(fy_hi, -4.0f), 1.0e6f);
        int ix0 = (int)floorf(fx_lo), ix1 = (int)ceilf(fx_hi);
        int iy0 = (int)floorf(fy_lo), iy1 = (int)ceilf(fy_hi);
        ix0 = max(ix0, 0); ix1 = min(ix1, p.W - 1);
        iy0 = max(iy0 - tc.band_y0, 0); iy1 = min(iy1 - tc.band_y0, band_rows - 1);
        if (ix0 > ix1 || iy0 > iy1) continue;
        const bool front = r.nz >= 0.0f;
        // tight bbox range (for the hard pass): shrink by blen in pixel units, conservatively
        const float bpx = p.blen * inv_sx * 0.5f, bpy = p.blen * inv_sy * 0.5f;
        const int hx0 = max((int)floorf(fx_lo + bpx), 0), hx1 = min((int)ceilf(fx_hi - bpx), p.W - 1);
        const int hy0 = max((int)floorf(fy_lo + bpy) - tc.band_y0, 0), hy1 = min((int)ceilf(fy_hi - bpy) - tc.band_y0, band_rows - 1);
        const uint32_t bit = 1u << (f & 31);
        const int wd = f >> 5;
        for (int sy = iy0 >> 2; sy <= (iy1 >> 2); ++sy) {
            for (int sxi = ix0 >> 3; sxi <= (ix1 >> 3); ++sxi) {
                const int st = sy * p.nstx + sxi;
                atomicOr(&tc.maskS[st * p.nwords + wd], bit);
                atomicOr(&tc.summ[st * nsum + (wd >> 5)], 1u << (wd & 31));
                if (front && sxi >= (hx0 >> 3) && sxi <= (hx1 >> 3) && sy >= (hy0 >> 2) && sy <= (hy1 >> 2))
                    atomicOr(&tc.maskH[st * p.nwords + wd], bit);
            }
        }
    }
    __syncthreads();
}

// Hard pass over one sub-tile (DIBR_SPEC A.2). Every lane walks the same faces.
__device__ __forceinline__ void hard_pass(const mm_raster_params& p, const TileCtx& tc, int st, float x0, float y0,
                                          int& best_f, float& bw0, float& bw1, float& bw2)
{
    float best_z = -INFINITY;
    best_f = -1; bw0 = bw1 = bw2 = 0.0f;
    for (int sw = 0; sw < tc.nsum; ++sw) {
        uint32_t smk = tc.summ[st * tc.nsum + sw];
        while (smk) {
            const int wd = (sw << 5) + __ffs(smk) - 1;
            smk &= smk - 1;
            uint32_t m = tc.maskH[st * p.nwords + wd];
            while (m) {
                const int f = (wd << 5) + __ffs(m) - 1;
                m &= m - 1;
                const FaceRec r = load_rec(tc.rec, f);
                float w0, w1, w2, zz;
                if (hard_test(r, x0, y0, p.eps, w0, w1, w2, zz)) {
                    if (!(zz <= best_z)) { best_z = zz; best_f = f; bw0 = w0; bw1 = w1; bw2 = w2; }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- forward
template <bool REC_SMEM, bool WITH_LOSS>
__global__ void __launch_bounds__(MM_THREADS)
k_raster_fwd(const mm_raster_params p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int b = blockIdx.y, band = blockIdx.x;
    TileCtx tc;
    tile_prologue<REC_SMEM>(p, b, band, smem, tc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lx = lane & 7, ly = lane >> 3;
    const size_t HW = (size_t)p.H * p.W;
    float acc_l1 = 0.0f, acc_n = 0.0f, acc_d = 0.0f;

    for (int st = warp; st < tc.nst; st += MM_WARPS) {
        const int sty = st / p.nstx, stx = st - sty * p.nstx;
        const int ix = stx * MM_ST_W + lx, iy = tc.band_y0 + sty * MM_ST_H + ly;
        const bool active = (ix < p.W) && (iy < p.H);
        const float x0 = pix_x(ix, p.W, p.sx), y0 = pix_y(iy, p.H, p.sy);

        int best_f; float w0, w1, w2;
        hard_pass(p, tc, st, x0, y0, best_f, w0, w1, w2);

        // ---- soft silhouette (DIBR_SPEC A.4) for uncovered pixels
        float soft = 1.0f;
        const bool need_soft = active && (best_f < 0);
        if (__any_sync(0xffffffffu, need_soft)) {
            float allprob = 1.0f;
            int kid = 0;
            for (int sw = 0; sw < tc.nsum; ++sw) {
                uint32_t smk = tc.summ[st * tc.nsum + sw];
                while (smk) {
                    const int wd = (sw << 5) + __ffs(smk) - 1;
                    smk &= smk - 1;
                    uint32_t m = tc.maskS[st * p.nwords + wd];
                    while (m) {
                        const int f = (wd << 5) + __ffs(m) - 1;
                        m &= m - 1;
                        const FaceRec r = load_rec(tc.rec, f);
                        if (need_soft && kid < p.knum && soft_bbox_test(r, x0, y0, p.blen)) {
                            int type;
                            const float d2 = soft_d2(r, x0, y0, p.multiplier, type);
                            const float prob = soft_prob(d2, p.sigmainv, p.multiplier);
                            allprob = allprob * (1.0f - prob);
                            ++kid;
                        }
                    }
                }
            }
            if (need_soft) soft = 1.0f - allprob;
        }

        if (!active) continue;
        const size_t pix = (size_t)iy * p.W + ix;

        // ---- shading (networks.py:303-314)
        float tm = 0.0f, nrm[3] = {0.0f, 0.0f, 0.0f}, tcol[3] = {0.0f, 0.0f, 0.0f};
        if (best_f >= 0) {
            const float* uvp = p.face_uvs + best_f * 6;
            // interpolation in the rasteriser's operation order (w0*c0 + w1*c1) + w2*c2, uncontracted
            const float u = interp3(w0, w1, w2, __ldg(uvp + 0), __ldg(uvp + 2), __ldg(uvp + 4));
            const float v = interp3(w0, w1, w2, __ldg(uvp + 1), __ldg(uvp + 3), __ldg(uvp + 5));
            const FaceRec r = load_rec(tc.rec, best_f);
            tm = ADD(ADD(w0, w1), w2);
            nrm[0] = interp3(w0, w1, w2, r.nx, r.nx, r.nx);
            nrm[1] = interp3(w0, w1, w2, r.ny, r.ny, r.ny);
            nrm[2] = interp3(w0, w1, w2, r.nz, r.nz, r.nz);
            Bilin bl;
            bilin_setup(u, v, p.Ht, p.Wt, bl);
            const float* tb = p.tex + (size_t)b * 3 * p.Ht * p.Wt;
            #pragma unroll
            for (int c = 0; c < 3; ++c) {
                const TexFetch t = tex_fetch(tb + (size_t)c * p.Ht * p.Wt, bl, p.Ht, p.Wt);
                tcol[c] = t.nw * bl.nw + t.ne * bl.ne + t.sw * bl.sw + t.se * bl.se;
            }
        }
        float bnd[9];
        sh_bands(nrm[0], nrm[1], nrm[2], bnd);
        const float coef = sh_coef(bnd, tc.lights);
        float img[3];
        #pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v;
            if (p.no_mask) {
                const float bgc = __ldg(p.bg + ((size_t)b * 3 + c) * HW + pix);
                v = (tcol[c] * tm + bgc * (1.0f - tm)) * coef;
            } else {
                v = tcol[c] * tm * coef + (1.0f - tm);
            }
            img[c] = clamp01(v);
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
            for (int c = 0; c < 3; ++c) acc_l1 += fabsf(l1_term(img[c], __ldg(g + c * HW), gm));
            const float mul = soft * gm;
            acc_n += mul;
            acc_d += (soft + gm) - mul;
        }
    }
    if (WITH_LOSS) {
        const float s0 = block_sum(acc_l1, tc.red);
        const float s1 = block_sum(acc_n, tc.red);
        const float s2 = block_sum(acc_d, tc.red);
        if (threadIdx.x == 0) {
            float* pf = p.part_fwd + ((size_t)b * p.nbands + band) * 4;
            pf[0] = s0; pf[1] = s1; pf[2] = s2; pf[3] = 0.0f;
        }
    }
}

// ---------------------------------------------------------------------------------------------- backward
// d(loss)/d(silhouette pixel) of the soft-IoU + contour terms (DIBR_SPEC A.7)
__device__ __forceinline__ float contour_c(float m, float mref) { return fabsf(m - mref); }

template <bool REC_SMEM>
__global__ void __launch_bounds__(MM_THREADS)
k_raster_bwd(const mm_raster_params p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int b = blockIdx.y, band = blockIdx.x;
    TileCtx tc;
    tile_prologue<REC_SMEM>(p, b, band, smem, tc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lx = lane & 7, ly = lane >> 3;
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

    // loss-gradient constants
    float k_img = 0.0f, k_iou = 0.0f, k_cont = 0.0f, Nb = 0.0f, De = 1.0f;
    if (p.analytic_loss) {
        k_img = p.loss_scale * p.image_weight / ((float)p.B * 3.0f * (float)HW);
        k_iou = p.loss_scale / (float)p.B;
        k_cont = p.loss_scale * p.contour / ((float)p.B * (float)HW);
        float Db = 0.0f;
        for (int k = 0; k < p.nbands; ++k) {      // fixed order: every CTA of the image gets the same sums
            Nb += p.part_fwd_in[((size_t)b * p.nbands + k) * 4 + 1];
            Db += p.part_fwd_in[((size_t)b * p.nbands + k) * 4 + 2];
        }
        De = Db + 1e-10f;
    }
    const float* rg = p.rgba + (size_t)b * 4 * HW;         // forward output (silhouette re-read)
    const float* gtb = p.gt ? p.gt + (size_t)b * 4 * HW : nullptr;
    const float* gup = p.g_rgba ? p.g_rgba + (size_t)b * 4 * HW : nullptr;
    float* gacc = p.gfacc + (size_t)b * p.F * 9;
    float* gtex = p.g_tex + (size_t)b * 3 * p.Ht * p.Wt;

    for (int st = warp; st < tc.nst; st += MM_WARPS) {
        const int sty = st / p.nstx, stx = st - sty * p.nstx;
        const int ix = stx * MM_ST_W + lx, iy = tc.band_y0 + sty * MM_ST_H + ly;
        const bool active = (ix < W) && (iy < H);
        const float x0 = pix_x(ix, W, p.sx), y0 = pix_y(iy, H, p.sy);
        const size_t pix = active ? (size_t)iy * W + ix : 0;

        int best_f = active ? p.face_idx_ws[(size_t)b * HW + pix] : -2;   // -2: inactive lane
        // ---- upstream gradient of the 4 output channels
        float g_img[3] = {0.0f, 0.0f, 0.0f}, g_soft = 0.0f;
        float soft = 0.0f;
        if (active) {
            soft = rg[3 * HW + pix];
            if (gup) { g_img[0] = gup[pix]; g_img[1] = gup[HW + pix]; g_img[2] = gup[2 * HW + pix]; g_soft = gup[3 * HW + pix]; }
            if (p.analytic_loss) {
                const float gm = __ldg(gtb + 3 * HW + pix);
                #pragma unroll
                for (int c = 0; c < 3; ++c)
                    g_img[c] += k_img * sgnf(l1_term(rg[c * HW + pix], __ldg(gtb + c * HW + pix), gm)) * gm;
                // soft IoU: -(1/B) * (gm*De - Nb*(1-gm)) / De^2
                g_soft += -k_iou * (gm * De - Nb * (1.0f - gm)) / (De * De);
                if (p.contour > 0.0f) {
                    const int ry = refrow[iy], rx = refcol[ix];
                    const size_t rp = (size_t)ry * W + rx;
                    const float mref = rg[3 * HW + rp], gref = __ldg(gtb + 3 * HW + rp);
                    const float cp = contour_c(soft, mref), cg = contour_c(gm, gref);
                    const float dlt = cp - cg;
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
        const bool need_soft = active && (best_f == -1) && (g_soft != 0.0f);
        if (__any_sync(0xffffffffu, need_soft)) {
            int kid = 0;
            const float one_m_all = 1.0f - soft;
            for (int sw = 0; sw < tc.nsum; ++sw) {
                uint32_t smk = tc.summ[st * tc.nsum + sw];
                while (smk) {
                    const int wd = (sw << 5) + __ffs(smk) - 1;
                    smk &= smk - 1;
                    uint32_t m = tc.maskS[st * p.nwords + wd];
                    while (m) {
                        const int f = (wd << 5) + __ffs(m) - 1;
                        m &= m - 1;
                        const FaceRec r = load_rec(tc.rec, f);
                        if (need_soft && kid < p.knum && soft_bbox_test(r, x0, y0, p.blen)) {
                            ++kid;
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
                }
            }
        }

        if (!active) continue;

        // ---- shading backward
        float tm = 0.0f, nrm[3] = {0.0f, 0.0f, 0.0f}, tcol[3] = {0.0f, 0.0f, 0.0f};
        FaceRec r;
        Bary bar;
        Bilin bl;
        TexFetch tf[3];
        float uv[6];
        if (best_f >= 0) {
            r = load_rec(tc.rec, best_f);
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
            for (int c = 0; c < 3; ++c) {
                tf[c] = tex_fetch(tb + (size_t)c * p.Ht * p.Wt, bl, p.Ht, p.Wt);
                tcol[c] = tf[c].nw * bl.nw + tf[c].ne * bl.ne + tf[c].sw * bl.sw + tf[c].se * bl.se;
            }
        }
        float bnd[9];
        sh_bands(nrm[0], nrm[1], nrm[2], bnd);
        const float coef = sh_coef(bnd, tc.lights);
        float g_coef = 0.0f, g_tcol[3];
        #pragma unroll
        for (int c = 0; c < 3; ++c) {
            float pre, bgc = 0.0f;
            if (p.no_mask) {
                bgc = __ldg(p.bg + ((size_t)b * 3 + c) * HW + pix);
                pre = (tcol[c] * tm + bgc * (1.0f - tm)) * coef;
            } else {
                pre = tcol[c] * tm * coef + (1.0f - tm);
            }
            const float g = (pre >= 0.0f && pre <= 1.0f) ? g_img[c] : 0.0f;     // torch.clamp backward
            g_tcol[c] = g * tm * coef;
            if (p.no_mask) {
                g_coef += g * (tcol[c] * tm + bgc * (1.0f - tm));
                if (p.g_bg) p.g_bg[((size_t)b * 3 + c) * HW + pix] = g * (1.0f - tm) * coef;
            } else {
                g_coef += g * (tcol[c] * tm);
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
            for (int c = 0; c < 3; ++c) {
                const float g = g_tcol[c];
                if (g != 0.0f) {
                    float* gp = gtex + ((size_t)c * p.Ht + bl.iy) * p.Wt + bl.ix;
                    atomicAdd(gp, g * bl.nw);
                    if (xe) atomicAdd(gp + 1, g * bl.ne);
                    if (ys) atomicAdd(gp + p.Wt, g * bl.sw);
                    if (xe && ys) atomicAdd(gp + p.Wt + 1, g * bl.se);
                    gix += g * ((tf[c].ne - tf[c].nw) * (1.0f - ty) + (tf[c].se - tf[c].sw) * ty);
                    giy += g * ((tf[c].sw - tf[c].nw) * (1.0f - tx) + (tf[c].se - tf[c].ne) * tx);
                }
            }
            const float g_gx = bl.in_x ? gix * ((float)p.Wt * 0.5f) : 0.0f;
            const float g_gy = bl.in_y ? giy * ((float)p.Ht * 0.5f) : 0.0f;
            const float g_u = 2.0f * g_gx, g_v = -2.0f * g_gy;

            // d coef / d normal -> unit face normal (features are the same normal on 3 corners)
            const float* l = tc.lights;
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

    // ---- per-CTA partials: contour sum + 9 light gradients (summed deterministically later)
    float* pb = p.part_bwd + ((size_t)b * p.nbands + band) * 12;
    {
        const float s = block_sum(acc_contour, tc.red);
        if (threadIdx.x == 0) pb[0] = s;
    }
    #pragma unroll
    for (int i = 0; i < 9; ++i) {
        const float s = block_sum(acc_l[i], tc.red);
        if (threadIdx.x == 0) pb[1 + i] = s;
    }
}

}  // namespace

static size_t raster_smem_bytes(const mm_ctx* c) {
    const int nst = c->nstx * c->st_rows;
    const int nsum = (c->nwords + 31) >> 5;
    return smem_plan(c->F, nst, c->nwords, nsum, c->rec_in_smem != 0).total;
}

size_t mm_raster_smem_bytes(const mm_ctx* c) { return raster_smem_bytes(c); }

cudaError_t mm_raster_configure(const mm_ctx* c) {
    const int bytes = (int)raster_smem_bytes(c);
    cudaError_t e;
    if (c->rec_in_smem) {
        if ((e = cudaFuncSetAttribute(k_raster_fwd<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
        if ((e = cudaFuncSetAttribute(k_raster_fwd<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
        if ((e = cudaFuncSetAttribute(k_raster_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
    } else {
        if ((e = cudaFuncSetAttribute(k_raster_fwd<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
        if ((e = cudaFuncSetAttribute(k_raster_fwd<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
        if ((e = cudaFuncSetAttribute(k_raster_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
    }
    return cudaSuccess;
}

void mm_launch_raster_fwd(const mm_ctx* c, const mm_raster_params& p, bool with_loss, cudaStream_t s)
{
    const dim3 grid(c->nbands, p.B);
    const size_t smem = raster_smem_bytes(c);
    if (c->rec_in_smem) {
        if (with_loss) k_raster_fwd<true, true><<<grid, MM_THREADS, smem, s>>>(p);
        else           k_raster_fwd<true, false><<<grid, MM_THREADS, smem, s>>>(p);
    } else {
        if (with_loss) k_raster_fwd<false, true><<<grid, MM_THREADS, smem, s>>>(p);
        else           k_raster_fwd<false, false><<<grid, MM_THREADS, smem, s>>>(p);
    }
}

void mm_launch_raster_bwd(const mm_ctx* c, const mm_raster_params& p, cudaStream_t s)
{
    const dim3 grid(c->nbands, p.B);
    const size_t smem = raster_smem_bytes(c);
    if (c->rec_in_smem) k_raster_bwd<true><<<grid, MM_THREADS, smem, s>>>(p);
    else                k_raster_bwd<false><<<grid, MM_THREADS, smem, s>>>(p);
}
