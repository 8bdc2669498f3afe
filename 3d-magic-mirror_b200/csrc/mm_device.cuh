// mm_device.cuh -- per-pixel device functions shared by the raster forward and
// backward kernels.  Geometry decisions (bbox tests, barycentric weights, depth,
// soft-silhouette distances) are written with explicit round-to-nearest intrinsics
// in the operation order of docs/DIBR_SPEC.md so that nvcc can neither contract
// them into FMAs nor reorder them: `face_idx` has to be bit-identical to the
// oracle's on identical face records, and the DIB-R distance formulas cancel badly
// enough in fp32 (C = x2*y1 - x1*y2 on 1000-scaled coordinates) that a different
// rounding sequence shows up at the 1e-3 level in the silhouette probability.
#pragma once
#include "mm_common.cuh"

#define MUL(a, b) __fmul_rn((a), (b))
#define ADD(a, b) __fadd_rn((a), (b))
#define SUB(a, b) __fsub_rn((a), (b))
#define DIV(a, b) __fdiv_rn((a), (b))

struct FaceRec {
    float ax, ay, bx, by, cx, cy, az, bz, cz, nx, ny, nz;
};

__device__ __forceinline__ FaceRec load_rec(const float* __restrict__ rec_base, int f) {
    const float4* p = reinterpret_cast<const float4*>(rec_base) + f * 3;
    const float4 r0 = p[0], r1 = p[1], r2 = p[2];
    FaceRec r;
    r.ax = r0.x; r.ay = r0.y; r.bx = r0.z; r.by = r0.w;
    r.cx = r1.x; r.cy = r1.y; r.az = r1.z; r.bz = r1.w;
    r.cz = r2.x; r.nx = r2.y; r.ny = r2.z; r.nz = r2.w;
    return r;
}

// DIBR_SPEC A.1: pixel centre in multiplier-scaled NDC
__device__ __forceinline__ float pix_x(int ix, int W, float sx) { return MUL(sx, (float)(2 * ix + 1 - W)); }
__device__ __forceinline__ float pix_y(int iy, int H, float sy) { return MUL(sy, (float)(H - 2 * iy - 1)); }

struct Bary {
    float k1, k2, k3, w0, w1, w2;
    float m, p, n, q, s, t;
};

__device__ __forceinline__ void bary_eval(const FaceRec& r, float x0, float y0, float eps, Bary& b) {
    b.m = SUB(r.bx, r.ax); b.p = SUB(r.by, r.ay);
    b.n = SUB(r.cx, r.ax); b.q = SUB(r.cy, r.ay);
    b.s = SUB(x0, r.ax);   b.t = SUB(y0, r.ay);
    b.k1 = SUB(MUL(b.s, b.q), MUL(b.n, b.t));
    b.k2 = SUB(MUL(b.m, b.t), MUL(b.s, b.p));
    b.k3 = SUB(MUL(b.m, b.q), MUL(b.n, b.p));
    const float den = ADD(b.k3, eps);
    b.w1 = DIV(b.k1, den);
    b.w2 = DIV(b.k2, den);
    b.w0 = SUB(SUB(1.0f, b.w1), b.w2);
}

// bary_eval + the reference's inside test (all three weights >= 0), with the two IEEE divisions skipped when the SIGN
// of a quotient already decides "outside": for finite k and den of opposite sign, k/den < 0 unless it underflows to -0
// (which the reference's `w < 0` test would NOT reject) -- the magnitude guards exclude that, so the decision is exact.
__device__ __forceinline__ bool bary_eval_inside(const FaceRec& r, float x0, float y0, float eps, Bary& b) {
    b.m = SUB(r.bx, r.ax); b.p = SUB(r.by, r.ay);
    b.n = SUB(r.cx, r.ax); b.q = SUB(r.cy, r.ay);
    b.s = SUB(x0, r.ax);   b.t = SUB(y0, r.ay);
    b.k1 = SUB(MUL(b.s, b.q), MUL(b.n, b.t));
    b.k2 = SUB(MUL(b.m, b.t), MUL(b.s, b.p));
    b.k3 = SUB(MUL(b.m, b.q), MUL(b.n, b.p));
    const float den = ADD(b.k3, eps);
    const float aden = fabsf(den);
    if (aden > 1e-18f && aden < 1e18f) {                 // |k| > 1e-18 and |den| < 1e18  =>  |k/den| > 1e-36 > FLT_MIN
        const float sg = den > 0.0f ? 1.0f : -1.0f;
        if (b.k1 * sg < -1e-18f || b.k2 * sg < -1e-18f) return false;
    }
    b.w1 = DIV(b.k1, den);
    b.w2 = DIV(b.k2, den);
    b.w0 = SUB(SUB(1.0f, b.w1), b.w2);
    return !(b.w0 < 0.0f || b.w1 < 0.0f || b.w2 < 0.0f);
}

// feature interpolation exactly as the rasteriser writes it: (w0*c0 + w1*c1) + w2*c2, no FMA contraction
__device__ __forceinline__ float interp3(float w0, float w1, float w2, float c0, float c1, float c2) {
    return ADD(ADD(MUL(w0, c0), MUL(w1, c1)), MUL(w2, c2));
}

// ------------------------------------------------------------------ exact pixel rectangles of a face's bbox
struct PixRange { int ix0, ix1, iy0, iy1; };

// Conservative pixel-index range, clipped to the image, of the scaled-NDC box [xl,xh) x [yl,yh)
// (the exact half-open tests are redone per pixel).  Returns false if empty.
template <class P>
__device__ __forceinline__ bool pix_range(const P& p, float xl, float xh, float yl, float yh, PixRange& r)
{
    const float inv_sx = 1.0f / p.sx, inv_sy = 1.0f / p.sy;
    float fx_lo = (xl * inv_sx + (float)(p.W - 1)) * 0.5f;
    float fx_hi = (xh * inv_sx + (float)(p.W - 1)) * 0.5f;
    float fy_lo = ((float)(p.H - 1) - yh * inv_sy) * 0.5f;
    float fy_hi = ((float)(p.H - 1) - yl * inv_sy) * 0.5f;
    // NaN/Inf coordinates (vertex on the camera plane) stay conservative: treat as "everywhere"
    if (!(fx_lo == fx_lo) || !(fx_hi == fx_hi)) { fx_lo = -4.0f; fx_hi = 1.0e6f; }
    if (!(fy_lo == fy_lo) || !(fy_hi == fy_hi)) { fy_lo = -4.0f; fy_hi = 1.0e6f; }
    fx_lo = fminf(fmaxf(fx_lo, -4.0f), 1.0e6f); fx_hi = fminf(fmaxf(fx_hi, -4.0f), 1.0e6f);
    fy_lo = fminf(fmaxf(fy_lo, -4.0f), 1.0e6f); fy_hi = fminf(fmaxf(fy_hi, -4.0f), 1.0e6f);
    r.ix0 = max((int)floorf(fx_lo), 0);
    r.ix1 = min((int)ceilf(fx_hi), p.W - 1);
    r.iy0 = max((int)floorf(fy_lo), 0);
    r.iy1 = min((int)ceilf(fy_hi), p.H - 1);
    return r.ix0 <= r.ix1 && r.iy0 <= r.iy1;
}

// EXACT pixel rectangle of a face's bbox (tight, or enlarged by boxlen: DIBR_SPEC A.4) under the reference's half-open
// fp32 test  xmin <= px < xmax, ymin <= py < ymax: conservative float->int estimate, then <= 2 correction steps per side
// with the very comparison the reference uses (pixel centres are monotone in the index, so the exact set is a rectangle).
template <class P>
__device__ __forceinline__ void exact_rect(const P& p, const FaceRec& r, bool enlarged,
                                           int& ix0, int& ix1, int& iy0, int& iy1)
{
    float xmin = fminf(fminf(r.ax, r.bx), r.cx), xmax = fmaxf(fmaxf(r.ax, r.bx), r.cx);
    float ymin = fminf(fminf(r.ay, r.by), r.cy), ymax = fmaxf(fmaxf(r.ay, r.by), r.cy);
    if (enlarged) { xmin = SUB(xmin, p.blen); xmax = ADD(xmax, p.blen); ymin = SUB(ymin, p.blen); ymax = ADD(ymax, p.blen); }
    PixRange pr;
    ix0 = 0; ix1 = -1; iy0 = 0; iy1 = -1;
    if (pix_range(p, xmin, xmax, ymin, ymax, pr)) {
        ix0 = pr.ix0; ix1 = pr.ix1; iy0 = pr.iy0; iy1 = pr.iy1;
        while (ix0 <= ix1 && pix_x(ix0, p.W, p.sx) < xmin) ++ix0;
        while (ix1 >= ix0 && pix_x(ix1, p.W, p.sx) >= xmax) --ix1;
        while (iy0 <= iy1 && pix_y(iy0, p.H, p.sy) >= ymax) ++iy0;      // y decreases with the row index
        while (iy1 >= iy0 && pix_y(iy1, p.H, p.sy) < ymin) --iy1;
    }
}

// The two exact rectangles of a face -- tight (hard pass; EMPTY for a back face, DIBR_SPEC A.2: the hard pass sees front faces
// only) and enlarged by boxlen (soft pass) -- are computed ONCE, by the vertex stage (one thread per face, where the record is in
// registers anyway), and travel as four 16-bit pairs: x = tight ix0 | ix1 << 16, y = tight iy0 | iy1 << 16, z / w = enlarged.
// The hard and the soft pass used to redo this per warp with 8 of 32 lanes active: ~2.3 M of their 8.3 M / 9.7 M
// warp-instructions each at cfg-2, in kernels that are instruction-issue bound.  An empty rectangle is stored as (1, 0).
__device__ __forceinline__ uint32_t rect_pack(int lo, int hi, bool empty) {
    return empty ? (1u | (0u << 16)) : ((uint32_t)lo | ((uint32_t)hi << 16));
}
template <class P>
__device__ __forceinline__ uint4 face_rects(const P& p, const FaceRec& r) {
    int ix0, ix1, iy0, iy1;
    uint4 o;
    ix0 = 0; ix1 = -1; iy0 = 0; iy1 = -1;
    if (r.nz >= 0.0f) exact_rect(p, r, false, ix0, ix1, iy0, iy1);
    bool empty = ix1 < ix0 || iy1 < iy0;
    o.x = rect_pack(ix0, ix1, empty); o.y = rect_pack(iy0, iy1, empty);
    exact_rect(p, r, true, ix0, ix1, iy0, iy1);
    empty = ix1 < ix0 || iy1 < iy0;
    o.z = rect_pack(ix0, ix1, empty); o.w = rect_pack(iy0, iy1, empty);
    return o;
}
__device__ __forceinline__ void rect_unpack(uint32_t xw, uint32_t yw, int& ix0, int& ix1, int& iy0, int& iy1) {
    ix0 = (int)(xw & 0xffffu); ix1 = (int)(xw >> 16); iy0 = (int)(yw & 0xffffu); iy1 = (int)(yw >> 16);
}

// ---- soft-silhouette distance (DIBR_SPEC A.4): squared distance to each edge (perpendicular if the foot lies on the segment,
// else "far") and to each vertex, minimum + its type 0..5.  The cancellation-prone quantities (A, B, C, up, down) keep the
// reference's exact operation order; what differs from the oracle's literal statement is only HOW the well-conditioned tail
// is evaluated:
//   * the "foot of the perpendicular lies on the segment" test is done on the un-divided foot (X - x1*dn)(X - x2*dn) +
//     (Y - y1*dn)(Y - y2*dn) > 0  (== direct * dn^2): two IEEE divisions less per edge; the decision can only differ
//     where direct ~ 0, i.e. where the perpendicular and the vertex distance coincide (the min is continuous there);
//   * up^2 / dn uses the 2-ulp fast division, exp the hardware ex2 path.
// Measured against the exact oracle: |soft - oracle| <= 1e-6 (tests assert 2e-6).
__device__ __forceinline__ float edge_d2_fast(float x1, float y1, float x2, float y2, float x0, float y0, float far) {
    const float A = SUB(y2, y1), B = SUB(x1, x2), C = SUB(MUL(x2, y1), MUL(x1, y2));
    const float up = ADD(ADD(MUL(A, x0), MUL(B, y0)), C);
    const float dn = ADD(ADD(MUL(A, A), MUL(B, B)), 1e-10f);
    const float X = SUB(SUB(MUL(MUL(B, B), x0), MUL(MUL(A, B), y0)), MUL(A, C));
    const float Y = SUB(SUB(MUL(MUL(A, A), y0), MUL(MUL(A, B), x0)), MUL(B, C));
    const float direct = (X - x1 * dn) * (X - x2 * dn) + (Y - y1 * dn) * (Y - y2 * dn);
    return direct > 0.0f ? far : __fdividef(up * up, dn);
}

__device__ __forceinline__ float soft_d2_fast(const FaceRec& r, float x0, float y0, float mult, int& type) {
    const float far = 4.0f * mult * mult;
    float d = edge_d2_fast(r.ax, r.ay, r.bx, r.by, x0, y0, far);
    type = 0;
    float v = edge_d2_fast(r.bx, r.by, r.cx, r.cy, x0, y0, far);
    if (d > v) { d = v; type = 1; }
    v = edge_d2_fast(r.cx, r.cy, r.ax, r.ay, x0, y0, far);
    if (d > v) { d = v; type = 2; }
    v = ADD(MUL(SUB(x0, r.ax), SUB(x0, r.ax)), MUL(SUB(y0, r.ay), SUB(y0, r.ay)));
    if (d > v) { d = v; type = 3; }
    v = ADD(MUL(SUB(x0, r.bx), SUB(x0, r.bx)), MUL(SUB(y0, r.by), SUB(y0, r.by)));
    if (d > v) { d = v; type = 4; }
    v = ADD(MUL(SUB(x0, r.cx), SUB(x0, r.cx)), MUL(SUB(y0, r.cy), SUB(y0, r.cy)));
    if (d > v) { d = v; type = 5; }
    return d;
}

// p = exp(-sigmainv * d2 / mult^2); kz = sigmainv / mult / mult (host-computed)
__device__ __forceinline__ float soft_prob_fast(float d2, float kz) { return __expf(-d2 * kz); }

// ------------------------------------------------------------------ shading
// kaolin texture_mapping == grid_sample(bilinear, align_corners=False, padding_mode='border') with v flipped
struct Bilin {
    int ix, iy;            // north-west texel
    float nw, ne, sw, se;  // weights
    float x, y;            // clipped, unnormalised coordinates
    bool in_x, in_y;       // clip gradient masks
};

__device__ __forceinline__ void bilin_setup(float u, float v, int Ht, int Wt, Bilin& s) {
    // explicit roundings: the fused and the unfused kernels must produce bit-identical images, so nothing here is left
    // to the compiler's FMA-contraction choices
    const float gx = SUB(MUL(u, 2.0f), 1.0f);
    const float gy = -SUB(MUL(v, 2.0f), 1.0f);
    float x = MUL(SUB(MUL(ADD(gx, 1.0f), (float)Wt), 1.0f), 0.5f);
    float y = MUL(SUB(MUL(ADD(gy, 1.0f), (float)Ht), 1.0f), 0.5f);
    s.in_x = (x > 0.0f) && (x < (float)(Wt - 1));
    s.in_y = (y > 0.0f) && (y < (float)(Ht - 1));
    x = fminf((float)(Wt - 1), fmaxf(x, 0.0f));
    y = fminf((float)(Ht - 1), fmaxf(y, 0.0f));
    const float fx = floorf(x), fy = floorf(y);
    s.ix = (int)fx; s.iy = (int)fy;
    s.x = x; s.y = y;
    const float tx = SUB(x, fx), ty = SUB(y, fy);
    s.nw = MUL(SUB(1.0f, tx), SUB(1.0f, ty));
    s.ne = MUL(tx, SUB(1.0f, ty));
    s.sw = MUL(SUB(1.0f, tx), ty);
    s.se = MUL(tx, ty);
}

struct TexFetch {
    float nw, ne, sw, se;
};

// the reference's TextureEncoder emits cat([t, t.flip(2)], dim=2) (model_res.py:609-610).  With a mirrored
// texture the caller hands over t alone ([B,3,Htp,Wt], Htp = Ht/2) and logical atlas row r lives in physical row
// r < Htp ? r : Ht-1-r; same texels, same weights, half the texture bytes.  Htp == Ht for a plain atlas (identity map).
__device__ __forceinline__ int tex_row(int r, int Ht, int Htp) { return r < Htp ? r : Ht - 1 - r; }

__device__ __forceinline__ TexFetch tex_fetch(const float* __restrict__ plane, const Bilin& s, int Ht, int Wt, int Htp) {
    TexFetch t;
    const bool xe = (s.ix + 1) < Wt, ys = (s.iy + 1) < Ht;
    const float* p0 = plane + (size_t)tex_row(s.iy, Ht, Htp) * Wt + s.ix;
    const float* p1 = plane + (size_t)tex_row(ys ? s.iy + 1 : s.iy, Ht, Htp) * Wt + s.ix;
    t.nw = __ldg(p0);
    t.ne = xe ? __ldg(p0 + 1) : 0.0f;
    t.sw = ys ? __ldg(p1) : 0.0f;
    t.se = (xe && ys) ? __ldg(p1 + 1) : 0.0f;
    return t;
}

// bilinear blend of the 4 fetched texels, fixed FMA chain
__device__ __forceinline__ float tex_blend(const TexFetch& t, const Bilin& b) {
    return __fmaf_rn(t.se, b.se, __fmaf_rn(t.sw, b.sw, __fmaf_rn(t.ne, b.ne, MUL(t.nw, b.nw))));
}

// networks.py:307-313 composite before the clamp, fixed operation order
__device__ __forceinline__ float composite_pre(bool no_mask, float tcol, float tm, float bgc, float coef) {
    return no_mask ? MUL(__fmaf_rn(tcol, tm, MUL(bgc, SUB(1.0f, tm))), coef)
                   : __fmaf_rn(MUL(tcol, tm), coef, SUB(1.0f, tm));
}

// kaolin spherical_harmonic_lighting (9 bands)
#define SH_C0 0.28209479177f
#define SH_C1 0.4886025119f
#define SH_C2 1.09254843059f
#define SH_C3 0.94617469575f
#define SH_C3B 0.31539156525f
#define SH_C4 0.77254840404f
#define SH_C5 0.38627420202f

__device__ __forceinline__ void sh_bands(float x, float y, float z, float* bnd) {
    bnd[0] = SH_C0;
    bnd[1] = MUL(SH_C1, x);
    bnd[2] = MUL(SH_C1, z);
    bnd[3] = MUL(SH_C1, y);
    bnd[4] = MUL(SH_C2, MUL(x, y));
    bnd[5] = MUL(SH_C2, MUL(y, z));
    bnd[6] = __fmaf_rn(SH_C3, MUL(z, z), -SH_C3B);
    bnd[7] = MUL(SH_C4, MUL(x, z));
    bnd[8] = MUL(SH_C5, __fmaf_rn(x, x, -MUL(y, y)));
}

__device__ __forceinline__ float sh_coef(const float* bnd, const float* l) {
    float c = 0.0f;
    #pragma unroll
    for (int i = 0; i < 9; ++i) c = __fmaf_rn(bnd[i], l[i], c);
    return c;
}

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

// ------------------------------------------------------------------ loss (DIBR_SPEC A.7; networks.py:364-390)
__device__ __forceinline__ float l1_term(float pred, float gt, float gm) {
    const float one_m = 1.0f - gm;
    const float a = pred * gm + one_m;
    const float b = gt * gm + one_m;
    return a - b;          // caller takes |.| / sign(.)
}

__device__ __forceinline__ float sgnf(float v) { return (v > 0.0f) ? 1.0f : ((v < 0.0f) ? -1.0f : 0.0f); }

// block-wide sum for MM_THREADS threads; `red` holds MM_WARPS floats
__device__ __forceinline__ float block_sum(float v, float* red) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.0f;
    #pragma unroll
    for (int i = 0; i < MM_WARPS; ++i) r += red[i];
    return r;
}

// ------------------------------------------------------------------ order-independent (deterministic) per-image sums
// Loss terms and light gradients are accumulated per image with 64-bit INTEGER atomics on fixed-point values: integer
// addition is associative, so the sums do not depend on the order in which warps arrive (float atomics would make the
// loss differ from run to run in the last bits) and no block barrier / last-CTA pass is needed.
// Scales: 2^40 for sums of non-negative loss terms (range +-8.3e6, resolution 9e-13), 2^44 for gradient sums
// (range +-5.2e5, resolution 6e-14).
#define MM_FX_LOSS 1099511627776.0
#define MM_FX_GRAD 17592186044416.0
__device__ __forceinline__ void fx_add(long long* acc, float v, double scale) {
    atomicAdd(reinterpret_cast<unsigned long long*>(acc), (unsigned long long)__double2ll_rn((double)v * scale));
}
__device__ __forceinline__ float fx_get(const long long* acc, double scale) { return (float)((double)(*acc) / scale); }

// d(loss)/d(silhouette) of pixel `pix` of image b as the geometry backward consumes it.  In the fused step the shading kernel
// stores the tile-local part (upstream gradient + contour term); the IoU term -(1/B) (gm De - Nb (1 - gm)) / De^2 needs the
// complete per-image sums and is added here, by the consumer (DIBR_SPEC A.7).
// loss_scale of the analytic loss gradient: the host-side factor times, when given, a DEVICE scalar (the upstream gradient
// of the loss that autograd hands to recon_data's backward -- lazy fusion, no host round trip)
__device__ __forceinline__ float eff_loss_scale(const mm_raster_params& p) {
    return p.loss_scale_dev ? p.loss_scale * __ldg(p.loss_scale_dev) : p.loss_scale;
}

__device__ __forceinline__ float gsoft_at(const mm_raster_params& p, int b, size_t pix) {
    const size_t HW = (size_t)p.H * p.W;
    float g = p.gsoft[(size_t)b * HW + pix];
    if (p.gsoft_iou_pending) {
        const float gm = __ldg(p.gt + ((size_t)b * 4 + 3) * HW + pix);
        const float Nb = fx_get(p.img_fwd + b * 4 + 1, MM_FX_LOSS);
        const float De = fx_get(p.img_fwd + b * 4 + 2, MM_FX_LOSS) + 1e-10f;
        g += -(eff_loss_scale(p) / (float)p.B) * (gm * De - Nb * (1.0f - gm)) / (De * De);
    }
    return g;
}

// ------------------------------------------------------------------ per-pixel soft-silhouette accumulator
// One 64-bit word per pixel, updated with ONE integer atomicAdd per (pixel, face) candidate:
//   bits 63..16  sum of log(1 - p_k) in fixed point (scale 2^32, two's complement; |sum| < 2^15)
//   bits 15..0   number of candidates seen (0xFFFF = "truncated at knum: exact ordered value stored by the overflow pass")
// Integer addition is associative, so the silhouette does not depend on the order in which face threads arrive
// (a float atomicAdd would make the image differ from run to run in the last bit).
#define MM_LACC_SCALE 4294967296.0
#define MM_LACC_OVF   0xFFFFull
__device__ __forceinline__ unsigned long long lacc_term(float log1m_p) {
    const long long fx = __double2ll_rn((double)fmaxf(log1m_p, -80.0f) * MM_LACC_SCALE);
    return ((unsigned long long)fx << 16) + 1ull;
}
__device__ __forceinline__ int lacc_count(unsigned long long a) { return (int)(a & 0xFFFFull); }
__device__ __forceinline__ float lacc_logsum(unsigned long long a) { return (float)((double)((long long)a >> 16) / MM_LACC_SCALE); }
__device__ __forceinline__ unsigned long long lacc_exact(float log_allprob) {
    const long long fx = __double2ll_rn((double)fmaxf(log_allprob, -2400.0f) * MM_LACC_SCALE);
    return ((unsigned long long)fx << 16) | MM_LACC_OVF;
}
// silhouette value of an uncovered pixel: 1 - prod(1 - p_k) = -expm1(sum log(1 - p_k))
__device__ __forceinline__ float lacc_soft(unsigned long long a) { return (a == 0ull) ? 0.0f : -expm1f(lacc_logsum(a)); }

// visibility buffer: (order-preserving depth << 32) | ~face ; atomicMax == "largest z, then smallest face index",
// i.e. the reference's "strictly greater z replaces, first face wins ties" scan (DIBR_SPEC A.2).  0 = uncovered.
__device__ __forceinline__ unsigned long long depth_key(float z, int f) {
    const uint32_t b = __float_as_uint(z);
    const uint32_t zo = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)zo << 32) | (unsigned long long)(0xffffffffu - (uint32_t)f);
}
__device__ __forceinline__ int key_face(unsigned long long key) { return key ? (int)(0xffffffffu - (uint32_t)(key & 0xffffffffull)) : -1; }

// ------------------------------------------------------------------ vector reductions into the per-face accumulators
// sm_90+: red.global.add.v4.f32 / .v2.f32 (SASS REDG.E.ADD.F32x4 / x2) -- one L2 reduction for a 16 / 8-byte aligned group
// instead of one per float.
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
// d/d(the six image-plane corner coordinates) of face record `g` (gfacc + face * MM_GF)
__device__ __forceinline__ void red_add_corners(float* g, const float (&v)[6]) {
    red_add_v4(g, v[0], v[1], v[2], v[3]);
    red_add_v2(g + 4, v[4], v[5]);
}
