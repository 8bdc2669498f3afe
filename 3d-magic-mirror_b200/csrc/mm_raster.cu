// mm_raster.cu -- the GEOMETRY half of the per-pixel stage: DIB-R hard visibility and soft silhouette, forward and
// backward.  Replaces kaolin dibr_rasterization's four CUDA kernels (call site networks.py:297-299).  Shading and
// the loss live in mm_shade.cu.
//
// Kaolin (and the first three designs tried here, see profiles/r1_notes.md and docs/history/) GATHER: every pixel walks
// a face list.  The template meshes have 1280 faces of ~3x3 pixels each, so a pixel-side walk is almost all
// bookkeeping: ncu showed ~4500 warp-instructions per 8x4-pixel tile for ~100 useful (pixel, face) pairs.  This file
// SCATTERS instead: the unit of work is a face.
//
//   k_hard        4 lanes per (image, face): rasterise the face's tight bbox (rows interleaved over the 4 lanes), exact
//                 DIB-R inside test + depth, resolve visibility with ONE 64-bit atomicMax per covered pixel on a packed
//                 (order-preserving depth << 32 | ~face) key.  max == "largest z, then smallest face index" == the
//                 reference's ordered scan with its strictly-greater test, so `face_idx` is bit-exact and independent
//                 of thread order.  No binning, no face lists, no shared memory.
//   k_soft_fwd    4 lanes per (image, face), all faces: walk the bbox enlarged by `boxlen`; for every UNCOVERED pixel
//                 inside (exact half-open test) evaluate the DIB-R distance / probability once and fold
//                 log(1 - p) and a candidate count into the pixel's 64-bit accumulator with ONE integer atomicAdd
//                 (fixed point => order independent => deterministic).  A pixel whose count reaches knum + 1 is
//                 appended to the overflow list.
//   k_soft_ovf    rare path (far cameras): DIB-R keeps only the FIRST knum candidates in face-index order.  One warp per
//                 overflowed pixel replays the reference's ordered scan over all faces (32 faces per step, ballot keeps
//                 the order) and stores the exact truncated product; the backward variant scatters the gradients of
//                 exactly those knum candidates.
//   k_soft_bwd    4 lanes per (image, face): same walk; the face's 6 corner gradients are accumulated in REGISTERS over
//                 all its pixels, combined over the 4 lanes with shuffles, and leave as <= 6 atomics per face.
//   Backward re-derives every probability from the face records instead of storing Kaolin's knum-deep side buffers
//   (B*H*W*30*(4+8+1) B = 307 MB at B=48,128^2).
#include "mm_device.cuh"

namespace {

#define FULL 0xffffffffu
#define LANES_PER_FACE 4

struct PixRange { int ix0, ix1, iy0, iy1; };

// Conservative pixel-index range, clipped to the image, of the scaled-NDC box [xl,xh) x [yl,yh)
// (the exact half-open tests are redone per pixel).  Returns false if empty.
__device__ __forceinline__ bool pix_range(const mm_raster_params& p, float xl, float xh, float yl, float yh, PixRange& r)
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

// ---------------------------------------------------------------------------------------------- hard pass
__global__ void __launch_bounds__(256)
k_hard(const mm_raster_params p)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int q = gid & (LANES_PER_FACE - 1);
    const int fid = gid / LANES_PER_FACE;
    if (fid >= p.B * p.F) return;
    const int b = fid / p.F, f = fid - b * p.F;
    const FaceRec r = load_rec(p.frec + (size_t)b * p.F * MM_REC_FLOATS, f);
    if (!(r.nz >= 0.0f)) return;                                   // DIBR_SPEC A.2: front faces only
    const float xmin = fminf(fminf(r.ax, r.bx), r.cx), xmax = fmaxf(fmaxf(r.ax, r.bx), r.cx);
    const float ymin = fminf(fminf(r.ay, r.by), r.cy), ymax = fmaxf(fmaxf(r.ay, r.by), r.cy);
    PixRange pr;
    if (!pix_range(p, xmin, xmax, ymin, ymax, pr)) return;
    unsigned long long* zb = p.zbuf + (size_t)b * p.H * p.W;
    for (int iy = pr.iy0 + q; iy <= pr.iy1; iy += LANES_PER_FACE) {
        const float py = pix_y(iy, p.H, p.sy);
        if (py < ymin || py >= ymax) continue;
        for (int ix = pr.ix0; ix <= pr.ix1; ++ix) {
            float w0, w1, w2, zz;
            if (hard_test(r, pix_x(ix, p.W, p.sx), py, p.eps, w0, w1, w2, zz))
                atomicMax(zb + (size_t)iy * p.W + ix, depth_key(zz, f));
        }
    }
}

// ---------------------------------------------------------------------------------------------- soft pass, forward
__global__ void __launch_bounds__(256)
k_soft_fwd(const mm_raster_params p)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int q = gid & (LANES_PER_FACE - 1);
    const int fid = gid / LANES_PER_FACE;
    if (fid >= p.B * p.F) return;
    const int b = fid / p.F, f = fid - b * p.F;
    const FaceRec r = load_rec(p.frec + (size_t)b * p.F * MM_REC_FLOATS, f);
    const float xmin = SUB(fminf(fminf(r.ax, r.bx), r.cx), p.blen), xmax = ADD(fmaxf(fmaxf(r.ax, r.bx), r.cx), p.blen);
    const float ymin = SUB(fminf(fminf(r.ay, r.by), r.cy), p.blen), ymax = ADD(fmaxf(fmaxf(r.ay, r.by), r.cy), p.blen);
    PixRange pr;
    if (!pix_range(p, xmin, xmax, ymin, ymax, pr)) return;
    const size_t HW = (size_t)p.H * p.W;
    const unsigned long long* zb = p.zbuf + (size_t)b * HW;
    unsigned long long* la = p.lacc + (size_t)b * HW;
    const float kz = p.sigmainv / p.multiplier / p.multiplier;
    for (int iy = pr.iy0 + q; iy <= pr.iy1; iy += LANES_PER_FACE) {
        const float py = pix_y(iy, p.H, p.sy);
        if (py < ymin || py >= ymax) continue;
        for (int ix = pr.ix0; ix <= pr.ix1; ++ix) {
            const float px = pix_x(ix, p.W, p.sx);
            if (px < xmin || px >= xmax) continue;
            const size_t pix = (size_t)iy * p.W + ix;
            if (zb[pix] != 0ull) continue;                          // covered pixels get soft = 1, no candidates
            int type;
            const float d2 = soft_d2_fast(r, px, py, p.multiplier, type);
            const float prob = soft_prob_fast(d2, kz);
            const unsigned long long old = atomicAdd(la + pix, lacc_term(log1pf(-prob)));
            if (lacc_count(old) == p.knum) {                        // this is candidate knum+1: the pixel needs the ordered pass
                const uint32_t slot = atomicAdd(p.ovf_count, 1u);
                p.ovf_list[slot] = (uint32_t)((size_t)b * HW + pix);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- soft pass, backward
// one (pixel, face) candidate's contribution to the face's 6 corner gradients (DIBR_SPEC A.5, fast tail)
__device__ __forceinline__ void soft_pair_grad(const mm_raster_params& p, const FaceRec& r, float px, float py, float kz,
                                               float inv_mult, float g_soft, float one_m_all, float (&ga)[6])
{
    int type;
    const float d2s = soft_d2_fast(r, px, py, p.multiplier, type);
    const float prob = soft_prob_fast(d2s, kz);
    // dLdz = -sigmainv * dLdp * (1-allprob) / (1-prob+1e-6) * prob
    const float dLdz = __fdividef(-p.sigmainv * g_soft * one_m_all, (1.0f - prob) + 1e-6f) * prob * inv_mult;
    float v[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (type >= 3) {
        const int i = type - 3;
        const float x1 = (i == 0) ? r.ax : ((i == 1) ? r.bx : r.cx);
        const float y1 = (i == 0) ? r.ay : ((i == 1) ? r.by : r.cy);
        const float gx = dLdz * 2.0f * (x1 - px), gy = dLdz * 2.0f * (y1 - py);
        #pragma unroll
        for (int k = 0; k < 3; ++k) { v[2 * k] = (i == k) ? gx : 0.0f; v[2 * k + 1] = (i == k) ? gy : 0.0f; }
    } else {
        const int i = type, j = (type == 2) ? 0 : type + 1;
        const float x1 = (i == 0) ? r.ax : ((i == 1) ? r.bx : r.cx);
        const float y1 = (i == 0) ? r.ay : ((i == 1) ? r.by : r.cy);
        const float x2 = (j == 0) ? r.ax : ((j == 1) ? r.bx : r.cx);
        const float y2 = (j == 0) ? r.ay : ((j == 1) ? r.by : r.cy);
        const float A = SUB(y2, y1), Bc = SUB(x1, x2), C = SUB(MUL(x2, y1), MUL(x1, y2));
        const float up = ADD(ADD(MUL(A, px), MUL(Bc, py)), C);
        const float rdn = __fdividef(1.0f, ADD(ADD(MUL(A, A), MUL(Bc, Bc)), 1e-10f));
        const float d2 = up * up * rdn;
        const float dzdA = 2.0f * (px * up - d2 * A) * rdn;
        const float dzdB = 2.0f * (py * up - d2 * Bc) * rdn;
        const float dzdC = 2.0f * up * rdn;
        const float g1x = dLdz * (dzdB - y2 * dzdC), g1y = dLdz * (x2 * dzdC - dzdA);
        const float g2x = dLdz * (y1 * dzdC - dzdB), g2y = dLdz * (dzdA - x1 * dzdC);
        #pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[2 * k] = (i == k) ? g1x : ((j == k) ? g2x : 0.0f);
            v[2 * k + 1] = (i == k) ? g1y : ((j == k) ? g2y : 0.0f);
        }
    }
    #pragma unroll
    for (int k = 0; k < 6; ++k) ga[k] += v[k];
}

__global__ void __launch_bounds__(256)
k_soft_bwd(const mm_raster_params p)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int q = gid & (LANES_PER_FACE - 1);
    const int fid = gid / LANES_PER_FACE;
    const bool live = fid < p.B * p.F;
    const int b = live ? fid / p.F : 0, f = live ? fid - b * p.F : 0;
    float ga[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (live) {
        const FaceRec r = load_rec(p.frec + (size_t)b * p.F * MM_REC_FLOATS, f);
        const float xmin = SUB(fminf(fminf(r.ax, r.bx), r.cx), p.blen), xmax = ADD(fmaxf(fmaxf(r.ax, r.bx), r.cx), p.blen);
        const float ymin = SUB(fminf(fminf(r.ay, r.by), r.cy), p.blen), ymax = ADD(fmaxf(fmaxf(r.ay, r.by), r.cy), p.blen);
        PixRange pr;
        if (pix_range(p, xmin, xmax, ymin, ymax, pr)) {
            const size_t HW = (size_t)p.H * p.W;
            const unsigned long long* zb = p.zbuf + (size_t)b * HW;
            const unsigned long long* la = p.lacc + (size_t)b * HW;
            const float* gs = p.gsoft + (size_t)b * HW;
            const float* alpha = p.rgba + (size_t)b * 4 * HW + 3 * HW;
            const float kz = p.sigmainv / p.multiplier / p.multiplier;
            const float inv_mult = 1.0f / p.multiplier;
            for (int iy = pr.iy0 + q; iy <= pr.iy1; iy += LANES_PER_FACE) {
                const float py = pix_y(iy, p.H, p.sy);
                if (py < ymin || py >= ymax) continue;
                for (int ix = pr.ix0; ix <= pr.ix1; ++ix) {
                    const float px = pix_x(ix, p.W, p.sx);
                    if (px < xmin || px >= xmax) continue;
                    const size_t pix = (size_t)iy * p.W + ix;
                    if (zb[pix] != 0ull) continue;
                    if (lacc_count(la[pix]) == (int)MM_LACC_OVF) continue;       // truncated pixel: ordered pass owns it
                    const float g = gs[pix];
                    const float soft = alpha[pix];
                    if (g == 0.0f || !(soft > 0.0f)) continue;
                    soft_pair_grad(p, r, px, py, kz, inv_mult, g, 1.0f - soft, ga);
                }
            }
        }
    }
    // combine the 4 lanes of the face, one set of atomics per face
    #pragma unroll
    for (int k = 0; k < 6; ++k) {
        ga[k] += __shfl_xor_sync(FULL, ga[k], 1);
        ga[k] += __shfl_xor_sync(FULL, ga[k], 2);
    }
    if (live && q == 0) {
        float* g = p.gfacc + ((size_t)b * p.F + f) * 9;
        #pragma unroll
        for (int k = 0; k < 6; ++k) if (ga[k] != 0.0f) atomicAdd(g + k, ga[k]);
    }
}

// ---------------------------------------------------------------------------------------------- overflow (ordered) pass
// One warp per pixel that saw more than knum candidates: replay DIBR_SPEC A.4 literally -- faces in index order, first
// knum whose enlarged bbox holds the pixel.  32 faces per step; the ballot keeps the order.
template <bool BWD>
__global__ void __launch_bounds__(256)
k_soft_ovf(const mm_raster_params p)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n = *p.ovf_count;
    const size_t HW = (size_t)p.H * p.W;
    const float kz = p.sigmainv / p.multiplier / p.multiplier;
    const float inv_mult = 1.0f / p.multiplier;
    for (uint32_t e = warp; e < n; e += nwarps) {
        const uint32_t pg = p.ovf_list[e];
        const int b = (int)(pg / HW);
        const int pix = (int)(pg - (size_t)b * HW);
        const int iy = pix / p.W, ix = pix - iy * p.W;
        const float px = pix_x(ix, p.W, p.sx), py = pix_y(iy, p.H, p.sy);
        const float* rec = p.frec + (size_t)b * p.F * MM_REC_FLOATS;
        float g = 0.0f, one_m_all = 0.0f;
        if (BWD) {
            g = p.gsoft[pg];
            const float soft = p.rgba[(size_t)b * 4 * HW + 3 * HW + pix];
            one_m_all = 1.0f - soft;
            if (g == 0.0f || !(soft > 0.0f)) continue;          // warp-uniform
        }
        int kid = 0;
        float allprob = 1.0f;
        for (int f0 = 0; f0 < p.F && kid < p.knum; f0 += 32) {
            const int f = f0 + lane;
            FaceRec r;
            bool hit = false;
            if (f < p.F) { r = load_rec(rec, f); hit = soft_bbox_test(r, px, py, p.blen); }
            uint32_t m = __ballot_sync(FULL, hit);
            const int room = p.knum - kid;
            if (__popc(m) > room) {                              // keep the `room` lowest set bits
                uint32_t k2 = 0u, h = m;
                #pragma unroll 1
                for (int a = 0; a < room; ++a) { const uint32_t low = h & (0u - h); k2 |= low; h ^= low; }
                m = k2;
            }
            const bool mine = (m >> lane) & 1u;
            if (BWD) {
                if (mine) {
                    float ga[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
                    soft_pair_grad(p, r, px, py, kz, inv_mult, g, one_m_all, ga);
                    float* gf = p.gfacc + ((size_t)b * p.F + f) * 9;
                    #pragma unroll
                    for (int k = 0; k < 6; ++k) if (ga[k] != 0.0f) atomicAdd(gf + k, ga[k]);
                }
            } else {
                float prob = 0.0f;
                if (mine) { int type; prob = soft_prob_fast(soft_d2_fast(r, px, py, p.multiplier, type), kz); }
                uint32_t mm = m;
                #pragma unroll 1
                while (mm) {                                     // the reference's ordered product
                    const int j = __ffs(mm) - 1;
                    mm &= mm - 1;
                    allprob = allprob * (1.0f - __shfl_sync(FULL, prob, j));
                }
            }
            kid += __popc(m);
        }
        if (!BWD && lane == 0) p.lacc[pg] = lacc_exact(allprob > 0.0f ? logf(allprob) : -2400.0f);
    }
}

}  // namespace

void mm_launch_geom_fwd(const mm_ctx* c, const mm_raster_params& p, cudaStream_t s)
{
    const int threads = p.B * c->F * LANES_PER_FACE;
    const int grid = (threads + 255) / 256;
    k_hard<<<grid, 256, 0, s>>>(p);
    k_soft_fwd<<<grid, 256, 0, s>>>(p);
    k_soft_ovf<false><<<c->num_sms * 2, 256, 0, s>>>(p);
}

void mm_launch_geom_bwd(const mm_ctx* c, const mm_raster_params& p, cudaStream_t s)
{
    const int threads = p.B * c->F * LANES_PER_FACE;
    const int grid = (threads + 255) / 256;
    k_soft_bwd<<<grid, 256, 0, s>>>(p);
    k_soft_ovf<true><<<c->num_sms * 2, 256, 0, s>>>(p);
}

size_t mm_raster_smem_bytes(const mm_ctx* c) { (void)c; return 0; }
cudaError_t mm_raster_configure(const mm_ctx* c) { (void)c; return cudaSuccess; }
