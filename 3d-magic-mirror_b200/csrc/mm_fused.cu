// mm_fused.cu -- the per-pixel SHADING stage of every entry point: UV texture sampling, SH lighting, composite, clamp
// (kaolin texture_mapping / spherical_harmonic_lighting + networks.py:303-317), the recon_data loss terms
// (networks.py:364-390) and their backward.  ONE kernel body, three modes:
//
//   SHADE_FUSED  mm_render_compare_fwd_bwd: shading forward, the L1 / IoU / contour partial sums AND the complete RGB-side
//                backward in one pass.  Of the whole backward only d(loss)/d(silhouette) depends on per-image sums (the IoU
//                ratio): every RGB-side gradient (texture, background, lights, face normals, hard-rasteriser uv term) is
//                a function of the pixel alone, so it is formed while the pixel is being shaded.
//   SHADE_FWD    mm_render_forward: the forward alone (+ the optional imnormal / face_idx outputs).
//   SHADE_BWD    mm_render_backward: the backward alone, re-deriving the forward per pixel.  The upstream gradient is any
//                mix of a materialised g_rgba and the ANALYTIC recon_data gradient (lazy fusion: recon_data's backward hands
//                over (gt, weights, a device scalar) instead of a (B,4,H,W) tensor; the IoU sums are in the workspace).
// Sharing the body makes the images of the three entry points bit-identical by construction.  The forward modes also re-do the
// truncated pixels of the soft pass (no launch of their own), and every mode takes its strips in the order of the shading
// schedule the soft pass wrote (longest first): see k_shade.
//   k_gsoft      H or W not a multiple of 4 only: d(loss)/d(silhouette) in its own pass (contour term through index tables).
// The geometry backward (mm_raster.cu) consumes `gsoft`.
#include "mm_device.cuh"
#include "mm_soft_fwd.cuh"
#pragma nv_diag_suppress 128    // ("loop is not reachable": the forward-only instantiation leaves the covered-pixel loop early)

namespace {

#ifndef FULL
#define FULL 0xffffffffu
#endif
#define FT_W MM_SH_TW     // tile of one warp: 16 x 8 pixels, lane = (ly = lane >> 2, lx = lane & 3), 4 pixels per lane
#define FT_H MM_SH_TH

__device__ __forceinline__ float warp_sum(float v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// 4 consecutive floats of a plane row: one 16-byte access when the row start is 16-byte aligned (VEC), else guarded scalars
template <bool VEC>
__device__ __forceinline__ void load4(const float* __restrict__ p, int n, float (&v)[4]) {
    if (VEC) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        #pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (j < n) ? __ldg(p + j) : 0.0f;
    }
}
template <bool VEC>
__device__ __forceinline__ void store4(float* __restrict__ p, int n, const float (&v)[4]) {
    if (VEC) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        #pragma unroll
        for (int j = 0; j < 4; ++j) if (j < n) p[j] = v[j];
    }
}

// a truncated pixel whose exact accumulator has not been stored yet (see k_shade): spin on the word, served by L2
__device__ __noinline__ unsigned long long lacc_wait_exact(const unsigned long long* a) {
    unsigned long long v;
    unsigned spins = 0;                               // (seconds of waiting = a broken dispatch order: trap, do not hang)
    for (;;) {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(a) : "memory");
        if (lacc_count(v) == (int)MM_LACC_OVF) return v;
        __nanosleep(200);
        if (++spins > (1u << 23)) __trap();
    }
}

// ---------------------------------------------------------------------------------------------- fused shading
// A CTA of FUSED_WARPS warps owns a strip of FUSED_WARPS 16x8-pixel tiles.
//   pass 1 (streaming): every thread handles 4 consecutive pixels as float4 accesses issued up front, shades them AS
//           BACKGROUND (texmask = 0, normal = 0), accumulates the loss sums and stores the outputs.  Covered pixels are
//           appended to a shared-memory list.
//   pass 2 (dense): the CTA's covered pixels, ONE PER LANE whatever their position: shading forward + the whole RGB-side
//           backward; the pixel's outputs are overwritten with scalar stores.
// The first version ran the covered path inside the 4-pixel loop of pass 1: 77 % lane utilisation, ~150 registers (8 resident
// warps per SM) and ~840 dependent warp-instructions per pixel quad at 2 warps per scheduler -- 46 us, latency-bound
// (profiles/r1_notes.md).  Splitting the passes lets pass 2 run compacted and at a smaller register footprint.
#define FUSED_WARPS MM_SH_WARPS
#ifndef FUSED_MINB
#define FUSED_MINB 5         // 96 registers: 5 CTAs per SM.  Round 1 measured register caps as slower; with the lights and the
#endif                       // schedule in shared memory the only spills at 96 are the five loss accumulators (outside the loops),
                             // and the fifth CTA per SM is worth 0.6 us at cfg-2, 8 us at cfg-5 (profiles/r2_notes.md section 10)
#define FUSED_THREADS (32 * FUSED_WARPS)
#define FUSED_TILE_PX (FT_W * FT_H)

struct ShadeSmem {
    uint32_t list[FUSED_WARPS * FUSED_TILE_PX];      // face << 10 | warp << 7 | lane << 2 | j   (F <= 65535, <= 8 warps)
};

enum { SHADE_FUSED = 0, SHADE_FWD = 1, SHADE_BWD = 2 };

template <bool VEC, bool HAS_GUP, int MODE>
// (the caller has loaded the image's lights into s_lights, zeroed s_count and passed a block barrier)
__device__ __forceinline__ void shade_role(const mm_raster_params& p, ShadeSmem& sm, const float* s_lights, int& s_count, const int bx, const int b)
{
    constexpr bool OUT = MODE != SHADE_BWD;        // stores the image (rgba [+ imnormal, face_idx])
    constexpr bool SUMS = MODE == SHADE_FUSED;     // accumulates the loss sums (L1, IoU, contour)
    constexpr bool GRAD = MODE != SHADE_FWD;       // forms the RGB-side backward and d(loss)/d(silhouette)
    const bool analytic = GRAD && p.analytic_loss; // the recon_data gradient is formed in-kernel from gt
    uint32_t* s_list = sm.list;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int H = p.H, W = p.W;
    const size_t HW = (size_t)H * W;
    const int ntx = (W + FT_W - 1) / FT_W;
    const float lscale = analytic ? eff_loss_scale(p) : 0.0f;
    const float k_img = lscale * p.image_weight / ((float)p.B * 3.0f * (float)HW);
    float coef_bg;                                                          // sh_coef of a zero normal, same op sequence
    { float bnd0[9]; sh_bands(0.0f, 0.0f, 0.0f, bnd0); coef_bg = sh_coef(bnd0, s_lights); }

    float acc_l1 = 0.0f, acc_n = 0.0f, acc_d = 0.0f, acc_gc = 0.0f, acc_c = 0.0f;
    {
        const int tile = bx * FUSED_WARPS + warp;
        const int ty = tile / ntx, tx = tile - ty * ntx;
        const int iy = ty * FT_H + (lane >> 2), ix0 = tx * FT_W + (lane & 3) * 4;
        const int n = (iy < H) ? min(4, W - ix0) : 0;             // valid pixels of this lane (<= 0: none)
        const bool active = n > 0;
        const size_t pix0 = active ? (size_t)iy * W + ix0 : 0;
        int face[4] = {-1, -1, -1, -1};
        float soft[4] = {0.0f, 0.0f, 0.0f, 0.0f}, gmv[4] = {0.0f, 0.0f, 0.0f, 0.0f}, gs[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (active) {
            // ---- every streamed input of the 4 pixels, issued up front (nothing that might wait, and no call, in between: the
            // exact-value wait of a truncated pixel below is an opaque call the compiler will not move loads across)
            const unsigned long long* zb = p.zbuf + (size_t)b * HW + pix0;
            const unsigned long long* la = p.lacc + (size_t)b * HW + pix0;
            unsigned long long z[4], l[4] = {0ull, 0ull, 0ull, 0ull};
            if (VEC) {
                const ulonglong2 z0 = *reinterpret_cast<const ulonglong2*>(zb), z1 = *reinterpret_cast<const ulonglong2*>(zb + 2);
                z[0] = z0.x; z[1] = z0.y; z[2] = z1.x; z[3] = z1.y;
                const ulonglong2 l0 = *reinterpret_cast<const ulonglong2*>(la), l1 = *reinterpret_cast<const ulonglong2*>(la + 2);
                l[0] = l0.x; l[1] = l0.y; l[2] = l1.x; l[3] = l1.y;
            } else {
                #pragma unroll
                for (int j = 0; j < 4; ++j) { z[j] = (j < n) ? zb[j] : 0ull; l[j] = (j < n) ? la[j] : 0ull; }
            }
            float bgv[3][4], gtv[4][4], gup[3][4];
            #pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                if (analytic || SUMS) load4<VEC>(p.gt + ((size_t)b * 4 + ch) * HW + pix0, n, gtv[ch]);
                else { gtv[ch][0] = gtv[ch][1] = gtv[ch][2] = gtv[ch][3] = 0.0f; }
            }
            #pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                if (p.no_mask) load4<VEC>(p.bg + ((size_t)b * 3 + ch) * HW + pix0, n, bgv[ch]);
                else { bgv[ch][0] = bgv[ch][1] = bgv[ch][2] = bgv[ch][3] = 0.0f; }
                if (HAS_GUP) load4<VEC>(p.g_rgba + ((size_t)b * 4 + ch) * HW + pix0, n, gup[ch]);
                else { gup[ch][0] = gup[ch][1] = gup[ch][2] = gup[ch][3] = 0.0f; }
            }
            if (HAS_GUP) load4<VEC>(p.g_rgba + ((size_t)b * 4 + 3) * HW + pix0, n, gs);      // upstream d/d(silhouette)
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                face[j] = (j < n) ? key_face(z[j]) : -1;
                if (MODE != SHADE_BWD && j < n && face[j] < 0 && lacc_count(l[j]) > p.knum && lacc_count(l[j]) != (int)MM_LACC_OVF)
                    l[j] = lacc_wait_exact(la + j);                  // truncated pixel, exact value on its way
                soft[j] = (face[j] >= 0) ? 1.0f : lacc_soft(l[j]);
            }
            #pragma unroll
            for (int j = 0; j < 4; ++j) gmv[j] = gtv[3][j];
            // channel by channel, each plane stored as soon as it is complete (keeps the live register set small: the kernel's
            // occupancy is register-bound); covered pixels are overwritten by pass 2
            float* out = p.rgba + (size_t)b * 4 * HW + pix0;
            float g_coef[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            #pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float img[4], gbg[4];
                #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float gm = gtv[3][j];
                    // the covered-path expression with tcol = tm = 0 (every mode runs the same operation sequence)
                    const float pre = composite_pre(p.no_mask, 0.0f, 0.0f, bgv[ch][j], coef_bg);
                    const float v = clamp01(pre);
                    img[j] = v;
                    const float lt = l1_term(v, gtv[ch][j], gm);
                    if (SUMS && face[j] < 0) acc_l1 += fabsf(lt);
                    float g = gup[ch][j] + k_img * sgnf(lt) * gm;
                    g = (pre >= 0.0f && pre <= 1.0f) ? g : 0.0f;
                    gbg[j] = g * coef_bg;
                    if (p.no_mask) g_coef[j] += g * bgv[ch][j];
                }
                if (OUT) store4<VEC>(out + ch * HW, n, img);
                if (GRAD && p.g_bg) store4<VEC>(p.g_bg + ((size_t)b * 3 + ch) * HW + pix0, n, gbg);
            }
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float gm = gtv[3][j];
                if (GRAD && face[j] < 0 && j < n) acc_gc += g_coef[j];
                if (SUMS && j < n) {                           // IoU partial sums (kaolin mask_iou)
                    const float mul = soft[j] * gm;
                    acc_n += mul;
                    acc_d += (soft[j] + gm) - mul;
                }
            }
            if (OUT) store4<VEC>(out + 3 * HW, n, soft);
            if (MODE == SHADE_FWD) {
                if (p.face_idx_out) {
                    int32_t* fo = p.face_idx_out + (size_t)b * HW + pix0;
                    if (VEC) *reinterpret_cast<int4*>(fo) = make_int4(face[0], face[1], face[2], face[3]);
                    else { for (int j = 0; j < 4; ++j) if (j < n) fo[j] = face[j]; }
                }
                if (p.imnormal) {                              // background: zero normal; covered pixels overwritten by pass 2
                    float* no = p.imnormal + ((size_t)b * HW + pix0) * 3;
                    if (VEC) {
                        const float4 z4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        float4* n4 = reinterpret_cast<float4*>(no);
                        n4[0] = z4; n4[1] = z4; n4[2] = z4;
                    } else { for (int j = 0; j < 3 * n; ++j) no[j] = 0.0f; }
                }
            }
        }
        // ---- d(loss)/d(silhouette) handed to the geometry backward: upstream + (analytic loss) the contour term (DIBR_SPEC A.7).
        // H and W are multiples of 4 whenever the contour term is formed here (gsoft_iou_pending), so the 16x8 tile holds whole
        // 4x4 contour blocks: a block = lanes differing in lane bits 2,3 (4 rows) with the same lx; its reference pixel is pixel 0
        // of the lane with (ly & 3) == 0.  The IoU term is added by the consumers (gsoft_at).  Without an analytic loss the
        // upstream gradient alone is stored, whatever the image size.
        if (GRAD && (p.gsoft_iou_pending || !analytic)) {
            if (analytic && p.contour > 0.0f) {
                const float k_cont = lscale * p.contour / ((float)p.B * (float)HW);
                const int ref_lane = lane & ~12;
                const float mref = __shfl_sync(FULL, soft[0], ref_lane), gref = __shfl_sync(FULL, gmv[0], ref_lane);
                float tsum = 0.0f;
                #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float dlt = fabsf(soft[j] - mref) - fabsf(gmv[j] - gref);
                    acc_c += dlt * dlt;                       // lanes outside the image hold zeros
                    const float own = 2.0f * dlt * sgnf(soft[j] - mref);
                    gs[j] += k_cont * own;
                    tsum -= own;                              // what this pixel contributes to its reference pixel
                }
                tsum += __shfl_xor_sync(FULL, tsum, 4);
                tsum += __shfl_xor_sync(FULL, tsum, 8);
                if ((lane & 12) == 0) gs[0] += k_cont * tsum;
            }
            if (active) store4<VEC>(p.gsoft + (size_t)b * HW + pix0, n, gs);
        }
        // ---- covered pixels -> shared list (warp-aggregated append, 4 ballots)
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool cov = face[j] >= 0;
            const uint32_t m = __ballot_sync(FULL, cov);
            if (m) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&s_count, __popc(m));
                base = __shfl_sync(FULL, base, 0);
                if (cov) s_list[base + __popc(m & ((1u << lane) - 1u))] =
                    ((uint32_t)face[j] << 10) | ((uint32_t)warp << 7) | ((uint32_t)lane << 2) | (uint32_t)j;
            }
        }
    }
    __syncthreads();

    // ---- pass 2: the CTA's covered pixels, one per lane
    const int count = s_count;
    float acc_l[9];
    #pragma unroll
    for (int i = 0; i < 9; ++i) acc_l[i] = 0.0f;
    const float* rec = p.frec + (size_t)b * p.F * MM_REC_FLOATS;
    const float* tb = p.tex + (size_t)b * 3 * p.Htp * p.Wt;
    float* gacc = GRAD ? p.gfacc + (size_t)b * p.F * MM_GF : nullptr;
    float* gtex = GRAD ? p.g_tex + (size_t)b * 3 * p.Htp * p.Wt : nullptr;
    // (a texel-prefetch sub-pass ahead of this loop was measured: 43.7 -> 46.3 us, the kernel is not bound by that miss)
    #pragma unroll 1
    for (int i = threadIdx.x; i < count; i += FUSED_THREADS) {
        const uint32_t e = s_list[i];
        const int f = (int)(e >> 10);
        const int tile = bx * FUSED_WARPS + (int)((e >> 7) & 7u);
        const int el = (int)((e >> 2) & 31u);
        const int ty = tile / ntx, tx = tile - ty * ntx;
        const int iy = ty * FT_H + (el >> 2), ix = tx * FT_W + (el & 3) * 4 + (int)(e & 3u);
        const size_t pix = (size_t)iy * W + ix;
        float gm = 0.0f, bgj[3], gtj[3], guj[3];
        if (analytic || SUMS) gm = __ldg(p.gt + ((size_t)b * 4 + 3) * HW + pix);
        #pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            gtj[ch] = (analytic || SUMS) ? __ldg(p.gt + ((size_t)b * 4 + ch) * HW + pix) : 0.0f;
            bgj[ch] = p.no_mask ? __ldg(p.bg + ((size_t)b * 3 + ch) * HW + pix) : 0.0f;
            guj[ch] = HAS_GUP ? __ldg(p.g_rgba + ((size_t)b * 4 + ch) * HW + pix) : 0.0f;
        }
        const FaceRec r = load_rec(rec, f);
        Bary bar;
        bary_eval(r, pix_x(ix, W, p.sx), pix_y(iy, H, p.sy), p.eps, bar);
        float u, v;
        {
            float uvf[6];
            const float* uvp = p.face_uvs + f * 6;
            #pragma unroll
            for (int k = 0; k < 6; ++k) uvf[k] = __ldg(uvp + k);
            u = interp3(bar.w0, bar.w1, bar.w2, uvf[0], uvf[2], uvf[4]);
            v = interp3(bar.w0, bar.w1, bar.w2, uvf[1], uvf[3], uvf[5]);
        }
        const float tm = ADD(ADD(bar.w0, bar.w1), bar.w2);
        const float nx = interp3(bar.w0, bar.w1, bar.w2, r.nx, r.nx, r.nx);
        const float ny = interp3(bar.w0, bar.w1, bar.w2, r.ny, r.ny, r.ny);
        const float nz = interp3(bar.w0, bar.w1, bar.w2, r.nz, r.nz, r.nz);
        Bilin bl;
        bilin_setup(u, v, p.Ht, p.Wt, bl);
        TexFetch tf[3];
        float tcol[3];
        #pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            tf[ch] = tex_fetch(tb + (size_t)ch * p.Htp * p.Wt, bl, p.Ht, p.Wt, p.Htp);
            tcol[ch] = tex_blend(tf[ch], bl);
        }
        float bnd[9];
        sh_bands(nx, ny, nz, bnd);
        const float coef = sh_coef(bnd, s_lights);
        float g_coef = 0.0f, g_tcol[3];
        float* out = p.rgba + (size_t)b * 4 * HW + pix;
        #pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float pre = composite_pre(p.no_mask, tcol[ch], tm, bgj[ch], coef);
            const float val = clamp01(pre);
            if (OUT) out[ch * HW] = val;
            if (GRAD) {
                const float lt = l1_term(val, gtj[ch], gm);
                if (SUMS) acc_l1 += fabsf(lt);
                float g = guj[ch] + k_img * sgnf(lt) * gm;
                g = (pre >= 0.0f && pre <= 1.0f) ? g : 0.0f;                     // torch.clamp backward
                g_tcol[ch] = g * tm * coef;
                if (p.g_bg) p.g_bg[((size_t)b * 3 + ch) * HW + pix] = g * (1.0f - tm) * coef;
                g_coef += p.no_mask ? g * (tcol[ch] * tm + bgj[ch] * (1.0f - tm)) : g * (tcol[ch] * tm);
            }
        }
        if (MODE == SHADE_FWD && p.imnormal) {
            float* no = p.imnormal + ((size_t)b * HW + pix) * 3;
            no[0] = nx; no[1] = ny; no[2] = nz;
        }
        if (!GRAD) continue;
        #pragma unroll
        for (int k = 0; k < 9; ++k) acc_l[k] += g_coef * bnd[k];

        // texture gradient + d/d(u,v)
        float gix = 0.0f, giy = 0.0f;
        const bool xe = (bl.ix + 1) < p.Wt, ys = (bl.iy + 1) < p.Ht;
        // (even texel column, even row length, 8-byte aligned planes: half of the covered pixels; 12 -> 6 reductions for them)
        const bool pair = xe && p.gtex_pair && (bl.ix & 1) == 0;
        const int tr0 = tex_row(bl.iy, p.Ht, p.Htp), tr1 = tex_row(ys ? bl.iy + 1 : bl.iy, p.Ht, p.Htp);
        const float txf = bl.x - (float)bl.ix, tyf = bl.y - (float)bl.iy;
        #pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float g = g_tcol[ch];
            if (g != 0.0f) {
                float* gp = gtex + ((size_t)ch * p.Htp + tr0) * p.Wt + bl.ix;
                float* gq = gtex + ((size_t)ch * p.Htp + tr1) * p.Wt + bl.ix;
                if (pair) {                          // the (ix, ix + 1) texels as ONE 8-byte vector reduction
                    red_add_v2(gp, g * bl.nw, g * bl.ne);
                    if (ys) red_add_v2(gq, g * bl.sw, g * bl.se);
                } else {
                    atomicAdd(gp, g * bl.nw);
                    if (xe) atomicAdd(gp + 1, g * bl.ne);
                    if (ys) atomicAdd(gq, g * bl.sw);
                    if (xe && ys) atomicAdd(gq + 1, g * bl.se);
                }
                gix += g * ((tf[ch].ne - tf[ch].nw) * (1.0f - tyf) + (tf[ch].se - tf[ch].sw) * tyf);
                giy += g * ((tf[ch].sw - tf[ch].nw) * (1.0f - txf) + (tf[ch].se - tf[ch].ne) * txf);
            }
        }
        const float g_gx = bl.in_x ? gix * ((float)p.Wt * 0.5f) : 0.0f;
        const float g_gy = bl.in_y ? giy * ((float)p.Ht * 0.5f) : 0.0f;
        const float g_u = 2.0f * g_gx, g_v = -2.0f * g_gy;

        // d coef / d normal -> unit face normal (the three corners carry the same normal)
        const float* l = s_lights;
        const float dcx = l[1] * SH_C1 + l[4] * SH_C2 * ny + l[7] * SH_C4 * nz + l[8] * SH_C5 * 2.0f * nx;
        const float dcy = l[3] * SH_C1 + l[4] * SH_C2 * nx + l[5] * SH_C2 * nz - l[8] * SH_C5 * 2.0f * ny;
        const float dcz = l[2] * SH_C1 + l[5] * SH_C2 * ny + l[6] * SH_C3 * 2.0f * nz + l[7] * SH_C4 * nx;
        float* ga = gacc + (size_t)f * MM_GF;
        const float gn_scale = g_coef * tm;
        if (gn_scale != 0.0f) red_add_v4(ga + 8, gn_scale * dcx, gn_scale * dcy, gn_scale * dcz, 0.0f);
        // hard rasteriser backward (DIBR_SPEC A.3) for the u,v channels.  Gradients carry a tolerance (unlike the visibility
        // decisions), so this block is written for instruction count, not for the reference's rounding sequence: the two
        // channels are folded into A1 = sum_d dLdI_d (c1_d - c0_d), A2 = sum_d dLdI_d (c2_d - c0_d) first, and the
        // structurally-zero partials are dropped.
        if (g_u != 0.0f || g_v != 0.0f) {
            // the record, the barycentric set-up and the face's uvs are RE-derived here (L1 hits + ~15 flops) instead of being
            // kept alive across the texture section: ~20 registers less at the kernel's widest point
            const FaceRec r2 = load_rec(rec, f);
            const float m = r2.bx - r2.ax, pp = r2.by - r2.ay, nn = r2.cx - r2.ax, q = r2.cy - r2.ay;
            const float sb = pix_x(ix, W, p.sx) - r2.ax, t = pix_y(iy, H, p.sy) - r2.ay;
            const float k1 = sb * q - nn * t, k2 = m * t - sb * pp, k3 = m * q - nn * pp;
            float uv[6];
            #pragma unroll
            for (int k = 0; k < 6; ++k) uv[k] = __ldg(p.face_uvs + f * 6 + k);
            const float rden = __fdividef(p.multiplier, k3 * k3 + p.eps);
            const float A1 = rden * (g_u * (uv[2] - uv[0]) + g_v * (uv[3] - uv[1]));
            const float A2 = rden * (g_u * (uv[4] - uv[0]) + g_v * (uv[5] - uv[1]));
            // numerators of dw1/d(.) and dw2/d(.):  w1 = k1/k3, w2 = k2/k3
            const float d1m = -q * k1,          d2m = t * k3 - q * k2;
            const float d1n = pp * k1 - t * k3, d2n = pp * k2;
            const float d1p = nn * k1,          d2p = nn * k2 - sb * k3;
            const float d1q = sb * k3 - m * k1, d2q = -m * k2;
            const float d1s = q * k3,           d2s = -pp * k3;
            const float d1t = -nn * k3,         d2t = m * k3;
            const float gbx = A1 * d1m + A2 * d2m, gby = A1 * d1p + A2 * d2p;      // d/d(bx), d/d(by)
            const float gcx = A1 * d1n + A2 * d2n, gcy = A1 * d1q + A2 * d2q;      // d/d(cx), d/d(cy)
            const float gsx = A1 * d1s + A2 * d2s, gsy = A1 * d1t + A2 * d2t;      // d/d(s), d/d(t)
            red_add_v4(ga, -(gbx + gcx + gsx), -(gby + gcy + gsy), gbx, gby);
            red_add_v2(ga + 4, gcx, gcy);
        }
    }
    if (!GRAD && !SUMS) return;

    // ---- per-image sums: warp shuffle -> fixed-point integer atomics (order-independent, hence deterministic)
    if (SUMS) {
        const float s0 = warp_sum(acc_l1), s1 = warp_sum(acc_n), s2 = warp_sum(acc_d);
        if (p.gsoft_iou_pending && p.contour > 0.0f) {
            const float s3 = warp_sum(acc_c);
            if (lane == 0 && s3 != 0.0f) fx_add(p.img_bwd + b * 12, s3, MM_FX_LOSS);
        }
        if (lane == 0) {
            if (s0 != 0.0f) fx_add(p.img_fwd + b * 4 + 0, s0, MM_FX_LOSS);
            if (s1 != 0.0f) fx_add(p.img_fwd + b * 4 + 1, s1, MM_FX_LOSS);
            if (s2 != 0.0f) fx_add(p.img_fwd + b * 4 + 2, s2, MM_FX_LOSS);
        }
    }
    if (GRAD) {
        acc_l[0] += acc_gc * SH_C0;                  // background pixels: bands = (C0, 0, .., -C3B, 0, 0)
        acc_l[6] += acc_gc * (-SH_C3B);
        #pragma unroll
        for (int i = 0; i < 9; ++i) {
            if (count > 0 || i == 0 || i == 6) {                // background strip: only bands 0 and 6 are non-zero
                const float si = warp_sum(acc_l[i]);
                if (lane == 0 && si != 0.0f) fx_add(p.img_bwd + b * 12 + 1 + i, si, MM_FX_GRAD);
            }
        }
    }
}

// grid (B * p.nstrips): one CTA per strip, taken in SCHEDULE order: the soft pass classed every strip by the rounds its dense
// pass will need (shade_sched_produce, mm_soft_fwd.cuh); CTA i -- CTAs are dispatched in index order -- takes the i-th strip of
// the longest-class-first order, so the few-microsecond background strips fill the end of the kernel instead of a 20 us strip
// of a near-camera image starting in the last wave.
// Forward modes (SHADE_FUSED / SHADE_FWD) also own the TRUNCATED pixels of the soft pass (more than knum candidates: DIB-R keeps
// the first knum in face order, DIBR_SPEC A.4).  If there are any (ovf_count[0], final when this kernel starts; fetched together
// with the schedule counters), the first p.novf CTAs of the grid first re-do their share of them exactly, one CTA per pixel
// (soft_ovf_role), which replaces the pixel's accumulator by the exact word (count field MM_LACC_OVF).  Nobody waits at CTA
// level: a lane of pass 1 that meets an accumulator with count > knum that is not exact yet spins on THAT word
// (lacc_wait_exact).  The CTA it waits for is one of the grid's first p.novf, which are resident together, so the wait ends.
// This used to be a kernel of its own between the soft pass and the shading: 5 us of launch + drain latency for ~40 pixels per
// step at cfg-2.
struct ShadeOvfSmem { uint32_t mask[OVF_MAX_WORDS]; int kept[MM_MAX_KNUM]; };

template <bool VEC, bool HAS_GUP, int MODE>
__global__ void __launch_bounds__(FUSED_THREADS, FUSED_MINB)
k_shade(const mm_raster_params p)
{
    static_assert(FUSED_THREADS == OVF_THREADS, "the overflow role runs with the shading kernel's CTA shape");
    mm_pdl_prologue((p.pdl_late & 8) != 0);
    __shared__ union { ShadeSmem sm; ShadeOvfSmem ov; } u;
    __shared__ float s_lights[16];
    __shared__ uint32_t s_n[8];                      // strips per class (0..4), [5] = truncated pixels
    __shared__ int s_sid;
    __shared__ int s_count;                          // covered pixels of the CTA's strip (shade_role)
    const int wid = blockIdx.x * FUSED_WARPS + (threadIdx.x >> 5);
    (void)wid;
    MM_PROF_MARK(p.prof, 3, wid, 0);
    if (threadIdx.x < 5) s_n[threadIdx.x] = p.sched_n[threadIdx.x];
    if (threadIdx.x == 5) s_n[5] = (MODE != SHADE_BWD) ? p.ovf_count[0] : 0u;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    if (MODE != SHADE_BWD && s_n[5] > blockIdx.x && (int)blockIdx.x < p.novf)
        soft_ovf_role<false>(p, u.ov.mask, u.ov.kept, s_n[5], blockIdx.x, p.novf);     // (every pixel ends with a block barrier)
    {                                                // the blockIdx.x-th strip, longest class first: its id and its image's lights
        uint32_t r = blockIdx.x;
        int k = 4;
        while (k > 0 && r >= s_n[k]) { r -= s_n[k]; --k; }
        const uint32_t* ent = p.sched_list + ((size_t)k * gridDim.x + r) * MM_SCHED_WORDS;
        if (threadIdx.x == 0) s_sid = (int)min(ent[0], gridDim.x - 1u);      // (clamped: a backward on a workspace without a forward reads garbage, not out of bounds)
        else if (threadIdx.x < 10) s_lights[threadIdx.x - 1] = __uint_as_float(ent[threadIdx.x]);
    }
    __syncthreads();
    const int b = s_sid / p.nstrips, bx = s_sid - b * p.nstrips;
    MM_PROF_MARK(p.prof, 3, wid, 1);
    shade_role<VEC, HAS_GUP, MODE>(p, u.sm, s_lights, s_count, bx, b);
    MM_PROF_MARK(p.prof, 3, wid, 2);
}

// ---------------------------------------------------------------------------------------------- d(loss)/d(silhouette)
// H or W not a multiple of 4: the contour term's nearest-down / nearest-up reference of a pixel (DIBR_SPEC A.7) is not
// tile-local; this pass forms d(loss)/d(silhouette) (upstream + IoU + contour) per pixel through the ctx's index tables and,
// in the fused step, the contour loss sum.  (Multiples of 4 never come here: the shading kernel does it on the fly.)
__global__ void __launch_bounds__(128)
k_gsoft(const mm_raster_params p)
{
    mm_pdl_prologue();
    const int b = blockIdx.y;
    const int H = p.H, W = p.W;
    const size_t HW = (size_t)H * W;
    // the silhouette, from the workspace (the image tensor itself is not an input of the backward)
    const unsigned long long* zb = p.zbuf + (size_t)b * HW;
    const unsigned long long* la = p.lacc + (size_t)b * HW;
    auto alpha_at = [&](size_t i) { return zb[i] ? 1.0f : lacc_soft(la[i]); };
    const float* gmask = p.gt + (size_t)b * 4 * HW + 3 * HW;
    const float* gup = p.g_rgba ? p.g_rgba + (size_t)b * 4 * HW + 3 * HW : nullptr;
    float* gs = p.gsoft + (size_t)b * HW;
    __shared__ float s_k[4];               // per-image constants (fixed-point -> float conversions are ~100 instructions: once per CTA)
    if (threadIdx.x == 0) {
        const float Nb0 = fx_get(p.img_fwd + b * 4 + 1, MM_FX_LOSS);
        const float De0 = fx_get(p.img_fwd + b * 4 + 2, MM_FX_LOSS) + 1e-10f;
        s_k[0] = Nb0; s_k[1] = De0; s_k[2] = 1.0f / (De0 * De0); s_k[3] = eff_loss_scale(p);
    }
    __syncthreads();
    const float Nb = s_k[0], De = s_k[1], inv_de2 = s_k[2], lscale = s_k[3];
    const float k_iou = lscale / (float)p.B;
    const float k_cont = lscale * p.contour / ((float)p.B * (float)HW);
    float acc_c = 0.0f;
    const int32_t* refrow = p.tab;
    const int32_t* rowlo = p.tab + H;
    const int32_t* rowhi = p.tab + 2 * H;
    const int32_t* refcol = p.tab + 3 * H;
    const int32_t* collo = p.tab + 3 * H + W;
    const int32_t* colhi = p.tab + 3 * H + 2 * W;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < H * W) {
        const float gm = gmask[i], m = alpha_at(i);
        float g = (gup ? gup[i] : 0.0f) - k_iou * (gm * De - Nb * (1.0f - gm)) * inv_de2;
        if (p.contour > 0.0f) {
            const int iy = i / W, ix = i - iy * W;
            const size_t rp = (size_t)refrow[iy] * W + refcol[ix];
            const float mref = alpha_at(rp), gref = gmask[rp];
            const float dlt = fabsf(m - mref) - fabsf(gm - gref);
            acc_c += dlt * dlt;
            float gc = 2.0f * dlt * sgnf(m - mref);
            for (int yy = rowlo[iy]; yy < rowhi[iy]; ++yy)
                for (int xx = collo[ix]; xx < colhi[ix]; ++xx) {
                    const size_t q = (size_t)yy * W + xx;
                    const float mq = alpha_at(q), gq = gmask[q];
                    const float dq = fabsf(mq - m) - fabsf(gq - gm);
                    gc -= 2.0f * dq * sgnf(mq - m);
                }
            g += k_cont * gc;
        }
        gs[i] = g;
    }
    if (p.contour > 0.0f && p.analytic_loss == 1) {           // fused step only (== 1): the contour LOSS sum is formed here too
        const float sc = warp_sum(acc_c);
        if ((threadIdx.x & 31) == 0 && sc != 0.0f) fx_add(p.img_bwd + b * 12, sc, MM_FX_LOSS);
    }
}

}  // namespace

// VEC (16-byte accesses) needs rows that start 16-byte aligned: W a multiple of 4 AND every plane pointer 16-byte aligned
// (a torch view with an odd storage offset, or any other ABI client, may hand over less: scalar template then)
static bool aligned16(const void* q) { return (((uintptr_t)q) & 15) == 0; }

cudaError_t mm_launch_shade(const mm_ctx* c, const mm_raster_params& p, int mode, cudaStream_t s)
{
    const dim3 grid(p.nstrips * p.B);
    const bool vec = (p.W & 3) == 0 && aligned16(p.zbuf) && aligned16(p.lacc) && aligned16(p.gt) && aligned16(p.bg) &&
                     aligned16(p.g_rgba) && aligned16(p.rgba) && aligned16(p.g_bg) && aligned16(p.gsoft) &&
                     aligned16(p.imnormal) && aligned16(p.face_idx_out);
    const bool gup = p.g_rgba != nullptr;
    void (*k)(mm_raster_params) = nullptr;
    if (mode == SHADE_FUSED)
        k = vec ? (gup ? k_shade<true, true, SHADE_FUSED> : k_shade<true, false, SHADE_FUSED>)
                : (gup ? k_shade<false, true, SHADE_FUSED> : k_shade<false, false, SHADE_FUSED>);
    else if (mode == SHADE_FWD)
        k = vec ? k_shade<true, false, SHADE_FWD> : k_shade<false, false, SHADE_FWD>;
    else
        k = vec ? (gup ? k_shade<true, true, SHADE_BWD> : k_shade<true, false, SHADE_BWD>)
                : (gup ? k_shade<false, true, SHADE_BWD> : k_shade<false, false, SHADE_BWD>);
    return mm_launch(k, grid, dim3(FUSED_THREADS), 0, s, c->pdl != 0, p);
}

cudaError_t mm_launch_gsoft(const mm_ctx* c, const mm_raster_params& p, cudaStream_t s)
{
    return mm_launch(k_gsoft, dim3((p.H * p.W + 127) / 128, p.B), dim3(128), 0, s, c->pdl != 0, p);
}
