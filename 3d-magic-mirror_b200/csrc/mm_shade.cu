// mm_shade.cu -- the STREAMING half of the per-pixel stage: shading (UV texture sampling, SH lighting, composite,
// clamp), the recon_data loss terms, and their backward.  Replaces kaolin texture_mapping /
// spherical_harmonic_lighting, the ~40 elementwise torch kernels of networks.py:303-317 and :364-390, and their
// autograd.  Visibility and the soft silhouette come from the geometry kernels (mm_raster.cu) as two 64-bit words
// per pixel: the atomicMax-resolved (depth, face) key and the fixed-point log-product accumulator.
//
// One thread per pixel, one warp per 8x4-pixel sub-tile (a warp's accesses are four 32-byte row segments, i.e. whole
// sectors), ONE warp per CTA: covered tiles (texture gathers, gradient atomics) take ~10x longer than background
// tiles, and in a multi-warp CTA the background warps would park at a barrier waiting for them (ncu: 47 % of the stall
// samples of the 8-warp version).  Per-image sums (L1, IoU, contour, light gradients) leave each warp as fixed-point
// 64-bit integer atomics: order-independent, hence deterministic, and barrier-free.
#include "mm_device.cuh"

namespace {

#define FULL 0xffffffffu

__device__ __forceinline__ float contour_c(float m, float mref) { return fabsf(m - mref); }

__device__ __forceinline__ float warp_sum(float v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------- forward
template <bool WITH_LOSS>
__global__ void __launch_bounds__(32)
k_shade_fwd(const mm_raster_params p)
{
    mm_pdl_prologue();
    __shared__ float s_lights[16];
    const int b = blockIdx.y, lane = threadIdx.x;
    if (lane < 9) s_lights[lane] = p.lights[b * 9 + lane];
    __syncwarp();
    const size_t HW = (size_t)p.H * p.W;
    const float* rec = p.frec + (size_t)b * p.F * MM_REC_FLOATS;
    float acc_l1 = 0.0f, acc_n = 0.0f, acc_d = 0.0f;
    const int st = blockIdx.x;
    {
        const int sty = st / p.nstx, stx = st - sty * p.nstx;
        const int ix = stx * MM_ST_W + (lane & 7), iy = sty * MM_ST_H + (lane >> 3);
        const bool active = (ix < p.W) && (iy < p.H);
        int best_f = -1;
        float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f, soft = 0.0f;
        if (active) {
            const size_t pix0 = (size_t)iy * p.W + ix;
            best_f = key_face(p.zbuf[(size_t)b * HW + pix0]);
            soft = (best_f >= 0) ? 1.0f : lacc_soft(p.lacc[(size_t)b * HW + pix0]);
            if (best_f >= 0) {                                       // the winner's weights, same instruction sequence as the hard pass
                const FaceRec r = load_rec(rec, best_f);
                Bary bb;
                bary_eval(r, pix_x(ix, p.W, p.sx), pix_y(iy, p.H, p.sy), p.eps, bb);
                w0 = bb.w0; w1 = bb.w1; w2 = bb.w2;
            }
        }
        if (active) {
            const size_t pix = (size_t)iy * p.W + ix;
            // ---- shading (networks.py:303-314)
            float tm = 0.0f, nrm[3] = {0.0f, 0.0f, 0.0f}, tcol[3] = {0.0f, 0.0f, 0.0f};
            if (best_f >= 0) {
                const float* uvp = p.face_uvs + best_f * 6;
                // interpolation in the rasteriser's operation order (w0*c0 + w1*c1) + w2*c2, uncontracted
                const float u = interp3(w0, w1, w2, __ldg(uvp + 0), __ldg(uvp + 2), __ldg(uvp + 4));
                const float v = interp3(w0, w1, w2, __ldg(uvp + 1), __ldg(uvp + 3), __ldg(uvp + 5));
                const FaceRec r = load_rec(rec, best_f);
                tm = ADD(ADD(w0, w1), w2);
                nrm[0] = interp3(w0, w1, w2, r.nx, r.nx, r.nx);
                nrm[1] = interp3(w0, w1, w2, r.ny, r.ny, r.ny);
                nrm[2] = interp3(w0, w1, w2, r.nz, r.nz, r.nz);
                Bilin bl;
                bilin_setup(u, v, p.Ht, p.Wt, bl);
                const float* tb = p.tex + (size_t)b * 3 * p.Htp * p.Wt;
                #pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const TexFetch t = tex_fetch(tb + (size_t)ch * p.Htp * p.Wt, bl, p.Ht, p.Wt, p.Htp);
                    tcol[ch] = tex_blend(t, bl);
                }
            }
            float bnd[9];
            sh_bands(nrm[0], nrm[1], nrm[2], bnd);
            const float coef = sh_coef(bnd, s_lights);
            float img[3];
            #pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float bgc = p.no_mask ? __ldg(p.bg + ((size_t)b * 3 + ch) * HW + pix) : 0.0f;
                img[ch] = clamp01(composite_pre(p.no_mask, tcol[ch], tm, bgc, coef));
            }
            float* out = p.rgba + (size_t)b * 4 * HW + pix;
            out[0] = img[0]; out[HW] = img[1]; out[2 * HW] = img[2]; out[3 * HW] = soft;
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
    if (WITH_LOSS) {                     // per-image sums: fixed-point integer atomics (deterministic, no barrier)
        const float s0 = warp_sum(acc_l1), s1 = warp_sum(acc_n), s2 = warp_sum(acc_d);
        if (lane == 0) {
            if (s0 != 0.0f) fx_add(p.img_fwd + b * 4 + 0, s0, MM_FX_LOSS);
            if (s1 != 0.0f) fx_add(p.img_fwd + b * 4 + 1, s1, MM_FX_LOSS);
            if (s2 != 0.0f) fx_add(p.img_fwd + b * 4 + 2, s2, MM_FX_LOSS);
        }
    }
}

// ---------------------------------------------------------------------------------------------- backward
__global__ void __launch_bounds__(32, 28)
k_shade_bwd(const mm_raster_params p)
{
    mm_pdl_prologue();
    __shared__ float s_lights[16];
    const int b = blockIdx.y, lane = threadIdx.x;
    if (lane < 9) s_lights[lane] = p.lights[b * 9 + lane];
    __syncwarp();
    const size_t HW = (size_t)p.H * p.W;
    const int H = p.H, W = p.W;
    const int32_t* refrow = p.tab;
    const int32_t* rowlo = p.tab + H;
    const int32_t* rowhi = p.tab + 2 * H;
    const int32_t* refcol = p.tab + 3 * H;
    const int32_t* collo = p.tab + 3 * H + W;
    const int32_t* colhi = p.tab + 3 * H + 2 * W;
    const float* rec = p.frec + (size_t)b * p.F * MM_REC_FLOATS;
    float acc_contour = 0.0f, acc_gc = 0.0f;
    float acc_l[9];
    #pragma unroll
    for (int i = 0; i < 9; ++i) acc_l[i] = 0.0f;
    bool tile_covered = false;
    // loss-gradient constants; the per-image IoU sums were reduced by the forward kernel (fixed order)
    float k_img = 0.0f, k_iou = 0.0f, k_cont = 0.0f, Nb = 0.0f, De = 1.0f;
    if (p.analytic_loss) {
        k_img = p.loss_scale * p.image_weight / ((float)p.B * 3.0f * (float)HW);
        k_iou = p.loss_scale / (float)p.B;
        k_cont = p.loss_scale * p.contour / ((float)p.B * (float)HW);
        Nb = fx_get(p.img_fwd + b * 4 + 1, MM_FX_LOSS);
        De = fx_get(p.img_fwd + b * 4 + 2, MM_FX_LOSS) + 1e-10f;
    }
    const float* rg = p.rgba + (size_t)b * 4 * HW;         // forward output
    const float* gtb = p.gt ? p.gt + (size_t)b * 4 * HW : nullptr;
    const float* gup = p.g_rgba ? p.g_rgba + (size_t)b * 4 * HW : nullptr;
    float* gacc = p.gfacc + (size_t)b * p.F * MM_GF;
    float* gtex = p.g_tex + (size_t)b * 3 * p.Htp * p.Wt;
    const int st = blockIdx.x;
    {
        const int sty = st / p.nstx, stx = st - sty * p.nstx;
        const int ix = stx * MM_ST_W + (lane & 7), iy = sty * MM_ST_H + (lane >> 3);
        const bool active = (ix < W) && (iy < H);
        const float x0 = pix_x(ix, W, p.sx), y0 = pix_y(iy, H, p.sy);
        const size_t pix = active ? (size_t)iy * W + ix : 0;
        const int best_f = active ? key_face(p.zbuf[(size_t)b * HW + pix]) : -1;
        const bool any_covered = __any_sync(FULL, best_f >= 0);
        tile_covered = any_covered;
        // ---- upstream gradient of the 4 output channels
        float g_img[3] = {0.0f, 0.0f, 0.0f}, g_soft = 0.0f;
        float soft = 0.0f, gm_lane = 0.0f;
        const bool fast4 = ((H & 3) == 0) && ((W & 3) == 0);
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
                if (p.contour > 0.0f && !fast4) {
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
                gm_lane = gm;
            }
        }
        // contour term, fast path: H and W are multiples of 4, so the 8x4 sub-tile holds two complete 4x4 contour
        // blocks (lanes with lx < 4 / lx >= 4) whose reference pixels are lanes 0 and 4: everything is exchanged
        // with shuffles instead of 2 + 32 dependent global loads per reference pixel.
        if (p.analytic_loss && p.contour > 0.0f && fast4) {
            const int ref_lane = lane & 4;
            const float mref = __shfl_sync(FULL, soft, ref_lane), gref = __shfl_sync(FULL, gm_lane, ref_lane);
            const float dlt = active ? contour_c(soft, mref) - contour_c(gm_lane, gref) : 0.0f;
            acc_contour += dlt * dlt;
            const float own = 2.0f * dlt * sgnf(soft - mref);
            float t = -own;                                     // what this pixel contributes to its reference pixel
            t += __shfl_xor_sync(FULL, t, 1); t += __shfl_xor_sync(FULL, t, 2);
            t += __shfl_xor_sync(FULL, t, 8); t += __shfl_xor_sync(FULL, t, 16);
            if (active) g_soft += k_cont * (own + ((lane == ref_lane) ? t : 0.0f));
        }

        // hand d(loss)/d(silhouette) to the geometry backward (only uncovered pixels that saw a candidate can use it)
        if (active && best_f < 0 && soft > 0.0f) p.gsoft[(size_t)b * HW + pix] = g_soft;

        if (active) {
            // ---- shading backward
            float tm = 0.0f, nrm[3] = {0.0f, 0.0f, 0.0f}, tcol[3] = {0.0f, 0.0f, 0.0f};
            FaceRec r;
            Bary bar;
            Bilin bl;
            TexFetch tf[3];
            float uv[6];
            if (best_f >= 0) {
                r = load_rec(rec, best_f);
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
                const float* tb = p.tex + (size_t)b * 3 * p.Htp * p.Wt;
                #pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    tf[ch] = tex_fetch(tb + (size_t)ch * p.Htp * p.Wt, bl, p.Ht, p.Wt, p.Htp);
                    tcol[ch] = tex_blend(tf[ch], bl);
                }
            }
            float bnd[9];
            sh_bands(nrm[0], nrm[1], nrm[2], bnd);
            const float coef = sh_coef(bnd, s_lights);
            float g_coef = 0.0f, g_tcol[3];
            #pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float bgc = p.no_mask ? __ldg(p.bg + ((size_t)b * 3 + ch) * HW + pix) : 0.0f;
                const float pre = composite_pre(p.no_mask, tcol[ch], tm, bgc, coef);
                const float g = (pre >= 0.0f && pre <= 1.0f) ? g_img[ch] : 0.0f;     // torch.clamp backward
                g_tcol[ch] = g * tm * coef;
                if (p.no_mask) {
                    g_coef += g * (tcol[ch] * tm + bgc * (1.0f - tm));
                    if (p.g_bg) p.g_bg[((size_t)b * 3 + ch) * HW + pix] = g * (1.0f - tm) * coef;
                } else {
                    g_coef += g * (tcol[ch] * tm);
                }
            }
            if (any_covered) {
                #pragma unroll
                for (int i = 0; i < 9; ++i) acc_l[i] += g_coef * bnd[i];
            } else {
                acc_gc += g_coef;                  // no normal anywhere in the tile: bands are the constants (C0,0,..,-C3B,0,0)
            }

            if (best_f >= 0) {
                // texture gradient + d/d(u,v)
                float gix = 0.0f, giy = 0.0f;
                const bool xe = (bl.ix + 1) < p.Wt, ys = (bl.iy + 1) < p.Ht;
                const int tr0 = tex_row(bl.iy, p.Ht, p.Htp), tr1 = tex_row(ys ? bl.iy + 1 : bl.iy, p.Ht, p.Htp);
                const float tx = bl.x - (float)bl.ix, ty = bl.y - (float)bl.iy;
                #pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const float g = g_tcol[ch];
                    if (g != 0.0f) {
                        float* gp = gtex + ((size_t)ch * p.Htp + tr0) * p.Wt + bl.ix;
                        float* gq = gtex + ((size_t)ch * p.Htp + tr1) * p.Wt + bl.ix;
                        atomicAdd(gp, g * bl.nw);
                        if (xe) atomicAdd(gp + 1, g * bl.ne);
                        if (ys) atomicAdd(gq, g * bl.sw);
                        if (xe && ys) atomicAdd(gq + 1, g * bl.se);
                        gix += g * ((tf[ch].ne - tf[ch].nw) * (1.0f - ty) + (tf[ch].se - tf[ch].sw) * ty);
                        giy += g * ((tf[ch].sw - tf[ch].nw) * (1.0f - tx) + (tf[ch].se - tf[ch].ne) * tx);
                    }
                }
                const float g_gx = bl.in_x ? gix * ((float)p.Wt * 0.5f) : 0.0f;
                const float g_gy = bl.in_y ? giy * ((float)p.Ht * 0.5f) : 0.0f;
                const float g_u = 2.0f * g_gx, g_v = -2.0f * g_gy;

                // d coef / d normal -> unit face normal (features are the same normal on 3 corners)
                const float* l = s_lights;
                const float nx = nrm[0], ny = nrm[1], nz = nrm[2];
                const float dcx = l[1] * SH_C1 + l[4] * SH_C2 * ny + l[7] * SH_C4 * nz + l[8] * SH_C5 * 2.0f * nx;
                const float dcy = l[3] * SH_C1 + l[4] * SH_C2 * nx + l[5] * SH_C2 * nz - l[8] * SH_C5 * 2.0f * ny;
                const float dcz = l[2] * SH_C1 + l[5] * SH_C2 * ny + l[6] * SH_C3 * 2.0f * nz + l[7] * SH_C4 * nx;
                float* g = gacc + (size_t)best_f * MM_GF;
                const float gn_scale = g_coef * tm;    // sum_i w_i * g_n
                if (gn_scale != 0.0f) {
                    red_add_v4(g + 8, gn_scale * dcx, gn_scale * dcy, gn_scale * dcz, 0.0f);
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
                    red_add_corners(g, gv);
                }
            }
        }
    }
    // ---- per-image sums (contour, 9 light gradients): fixed-point integer atomics (deterministic, no barrier)
    const float sc = warp_sum(acc_contour);
    if (lane == 0 && sc != 0.0f) fx_add(p.img_bwd + b * 12, sc, MM_FX_LOSS);
    if (tile_covered) {
        #pragma unroll
        for (int i = 0; i < 9; ++i) {
            const float si = warp_sum(acc_l[i]);
            if (lane == 0 && si != 0.0f) fx_add(p.img_bwd + b * 12 + 1 + i, si, MM_FX_GRAD);
        }
    } else {                              // background tile: one reduction instead of nine
        const float sg = warp_sum(acc_gc);
        if (lane == 0 && sg != 0.0f) {
            fx_add(p.img_bwd + b * 12 + 1 + 0, sg * SH_C0, MM_FX_GRAD);
            fx_add(p.img_bwd + b * 12 + 1 + 6, sg * (-SH_C3B), MM_FX_GRAD);
        }
    }
}

}  // namespace

void mm_launch_shade_fwd(const mm_ctx* c, const mm_raster_params& p, bool with_loss, cudaStream_t s)
{
    const dim3 grid(c->nst, p.B);
    mm_launch(with_loss ? k_shade_fwd<true> : k_shade_fwd<false>, grid, dim3(32), 0, s, g_mm_pdl != 0, p);
}

void mm_launch_shade_bwd(const mm_ctx* c, const mm_raster_params& p, cudaStream_t s)
{
    const dim3 grid(c->nst, p.B);
    mm_launch(k_shade_bwd, grid, dim3(32), 0, s, false, p);      // first kernel after the memsets of the backward
}
