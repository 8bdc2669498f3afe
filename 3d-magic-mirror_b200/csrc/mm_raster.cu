// mm_raster.cu -- the GEOMETRY half of the per-pixel stage: DIB-R hard visibility and soft silhouette, forward and
// backward.  Replaces kaolin dibr_rasterization's four CUDA kernels (call site networks.py:297-299).  Shading and
// the loss live in mm_shade.cu.
//
// Kaolin (and the first three designs tried here, see profiles/r1_notes.md and docs/history/) GATHER: every pixel walks
// a face list.  The template meshes have 1280 faces of ~3x3 pixels each, so a pixel-side walk is almost all
// bookkeeping: ncu showed ~4500 warp-instructions per 8x4-pixel tile for ~100 useful (pixel, face) pairs.  This file
// SCATTERS instead: the unit of work is a face.
//
//   k_scatter_hard  a CTA owns 64 faces (8 warps x 8, from all image groups of the batch); the (face, pixel) pairs of their EXACT
//                 tight-bbox rectangles (computed once by the vertex stage: frect) are numbered and dealt to the 256 threads; a
//                 pair runs the exact DIB-R inside test + depth and resolves visibility with ONE 64-bit atomicMax on a packed
//                 (order-preserving depth << 32 | ~face) key.  max == "largest z, then smallest face index" == the reference's
//                 ordered scan with its strictly-greater test, so `face_idx` is bit-exact and independent of thread order.  No
//                 binning, no face lists.
//   k_soft_fwd     (mm_soft_fwd.cuh) all faces: walk the bbox enlarged by `boxlen` against the coverage bitmap; for every
//                 UNCOVERED pixel inside (exact half-open test) evaluate the DIB-R distance / probability once and fold
//                 log(1 - p) and a candidate count into the pixel's 64-bit accumulator with ONE integer atomicAdd (fixed
//                 point => order independent => deterministic); the pairs go to a list the backward replays.  A pixel whose
//                 count reaches knum + 1 is appended to the overflow list.
//   (truncated pixels, far cameras: DIB-R keeps only the FIRST knum candidates in face-index order.  soft_ovf_role, in
//                 mm_soft_fwd.cuh, replays the reference's ordered scan over all faces, one CTA per pixel, and stores the exact
//                 truncated product; the forward runs it inside the shading kernel, there is no launch of its own.)
//   k_soft_bwd     ONE launch: CTAs >= novf replay the pair list (one pair per lane, vector REDs into the per-face
//                 accumulators; if the list overflowed its buffer: the bbox walk again, scatter_warp_bwd), the rest redo
//                 the truncated pixels and scatter the gradients of exactly those knum candidates.
//   Backward re-derives every probability from the face records instead of storing Kaolin's knum-deep side buffers
//   (B*H*W*30*(4+8+1) B = 307 MB at B=48,128^2).
#include "mm_device.cuh"
#include "mm_soft_fwd.cuh"

namespace {

// ---------------------------------------------------------------------------------------------- the pair engine
// FPW = 8 faces per warp; every face comes with its EXACT pixel rectangle (frect: the reference's half-open fp32 bbox test
// turned into index ranges by the vertex stage), so the (face, pixel) pairs of a group of faces can be NUMBERED and dealt to
// the lanes round-robin: balanced whatever the face sizes, and no per-pixel bbox test is left.
//   hard pass (k_scatter_hard): pairs of a CTA's 64 faces over its 256 threads; every pair runs the inside test + depth
//              (2 IEEE divisions) and, if inside, one atomicMax + one atomicOr.
//   soft backward FALLBACK (scatter_warp_bwd; only if the forward's pair list overflowed its buffer): pairs of a warp's 8
//              faces over its 32 lanes; the loop only filters (uncovered and not truncated: two 8-byte loads),
//              ballot-compacts the qualifying pairs into a small shared-memory queue, and whenever 32 are waiting the
//              whole warp evaluates them, one pair per lane.
#define FPW 8

struct WarpQ {
    uint32_t q[64];          // pending pairs: slot << 24 | iy << 12 | ix
    float rec[12][FPW];      // the warp's 8 face records
    int img[FPW];            // image of each slot
    int face[FPW];           // face index of each slot
    int ix0[FPW], iy0[FPW], w[FPW];   // exact pixel rectangle of each slot's bbox (origin, width)
    int pre[FPW];            // first pair number of each slot
    float rw[FPW];           // 1 / width
    float facc[6][FPW];      // backward: corner-gradient accumulators per slot
};

__device__ __forceinline__ FaceRec slot_rec(const WarpQ& wq, int slot) {
    FaceRec r;
    r.ax = wq.rec[0][slot]; r.ay = wq.rec[1][slot]; r.bx = wq.rec[2][slot]; r.by = wq.rec[3][slot];
    r.cx = wq.rec[4][slot]; r.cy = wq.rec[5][slot]; r.az = wq.rec[6][slot]; r.bz = wq.rec[7][slot];
    r.cz = wq.rec[8][slot]; r.nx = r.ny = r.nz = 0.0f;
    return r;
}

// one queued pair of the soft backward's fallback walk, evaluated by one lane (no warp collectives in here: the tail of the
// queue runs divergent)
__device__ __forceinline__ void eval_pair_bwd(const mm_raster_params& p, WarpQ& wq, uint32_t e, float kz, float inv_mult)
{
    const int slot = (int)(e >> 24), iy = (int)((e >> 12) & 0xfffu), ix = (int)(e & 0xfffu);
    const int b = wq.img[slot];
    const size_t HW = (size_t)p.H * p.W;
    const size_t pix = (size_t)iy * p.W + ix;
    const float g = gsoft_at(p, b, pix);
    const float soft = lacc_soft(p.lacc[(size_t)b * HW + pix]);      // (candidates are uncovered pixels)
    if (g == 0.0f || !(soft > 0.0f)) return;
    float ga[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    soft_pair_grad(p, slot_rec(wq, slot), pix_x(ix, p.W, p.sx), pix_y(iy, p.H, p.sy), kz, inv_mult, g, 1.0f - soft, ga);
    #pragma unroll
    for (int k = 0; k < 6; ++k) if (ga[k] != 0.0f) atomicAdd(&wq.facc[k][slot], ga[k]);
}

__device__ __forceinline__ void scatter_warp_bwd(const mm_raster_params& p, WarpQ& wq, const int gwarp, const int nwarps)
{
    const int lane = threadIdx.x & 31;
    const size_t HW = (size_t)p.H * p.W;
    const float kz = p.sigmainv / p.multiplier / p.multiplier;
    const float inv_mult = 1.0f / p.multiplier;
    // ---- set-up: lanes 0..7 each own one face of the warp: record -> smem, enlarged rectangle
    int npx = 0;
    if (lane < FPW) {
        const int slot = lane;
        const int fid = slot * nwarps + gwarp;       // (faces dealt with a stride of the warp count: 8 different images)
        int ix0 = 0, ix1 = -1, iy0 = 0, iy1 = -1;
        if (fid < p.B * p.F) {
            const int b = fid / p.F, f = fid - b * p.F;
            const FaceRec r = load_rec(p.frec + (size_t)b * p.F * MM_REC_FLOATS, f);
            wq.rec[0][slot] = r.ax; wq.rec[1][slot] = r.ay; wq.rec[2][slot] = r.bx; wq.rec[3][slot] = r.by;
            wq.rec[4][slot] = r.cx; wq.rec[5][slot] = r.cy; wq.rec[6][slot] = r.az; wq.rec[7][slot] = r.bz;
            wq.rec[8][slot] = r.cz; wq.img[slot] = b; wq.face[slot] = f;
            const uint4 rc = p.frect[fid];
            rect_unpack(rc.z, rc.w, ix0, ix1, iy0, iy1);
        } else { wq.img[slot] = 0; wq.face[slot] = 0; }
        const int w = ix1 - ix0 + 1, h = iy1 - iy0 + 1;
        npx = (w > 0 && h > 0) ? w * h : 0;
        wq.ix0[slot] = ix0; wq.iy0[slot] = iy0; wq.w[slot] = w > 0 ? w : 1;
        wq.rw[slot] = __frcp_rn((float)(w > 0 ? w : 1));
        #pragma unroll
        for (int k = 0; k < 6; ++k) wq.facc[k][slot] = 0.0f;
    }
    // exclusive prefix of the 8 pixel counts (every lane keeps all of them: the slot search below is 8 compares)
    int pre[FPW + 1];
    pre[0] = 0;
    #pragma unroll
    for (int sl = 0; sl < FPW; ++sl) pre[sl + 1] = pre[sl] + __shfl_sync(FULL, npx, sl);
    const int total = pre[FPW];
    if (lane < FPW) {
        int mine = 0;
        #pragma unroll
        for (int sl = 1; sl < FPW; ++sl) mine = (lane == sl) ? pre[sl] : mine;
        wq.pre[lane] = mine;
    }
    __syncwarp();

    // ---- the warp's (face, pixel) pairs, dealt to the 32 lanes round-robin
    int qn = 0;
    const uint32_t lt = (1u << lane) - 1u;
    #pragma unroll 1
    for (int k0 = 0; k0 < total; k0 += 32) {
        const int k = k0 + lane;
        bool cand = false;
        uint32_t entry = 0u;
        if (k < total) {
            int slot = 0;
            #pragma unroll
            for (int sl = 1; sl < FPW; ++sl) slot += (k >= pre[sl]) ? 1 : 0;
            const int local = k - wq.pre[slot];
            const int w = wq.w[slot];
            int dy = (int)(((float)local + 0.5f) * wq.rw[slot]);
            int dx = local - dy * w;
            if (dx < 0) { --dy; dx += w; } else if (dx >= w) { ++dy; dx -= w; }
            const int ix = wq.ix0[slot] + dx, iy = wq.iy0[slot] + dy;
            entry = ((uint32_t)slot << 24) | ((uint32_t)iy << 12) | (uint32_t)ix;
            const size_t pg = (size_t)wq.img[slot] * HW + (size_t)iy * p.W + ix;
            cand = (p.zbuf[pg] == 0ull) && lacc_count(p.lacc[pg]) != (int)MM_LACC_OVF;    // uncovered, not truncated
        }
        const uint32_t m = __ballot_sync(FULL, cand);
        if (cand) wq.q[qn + __popc(m & lt)] = entry;
        qn += __popc(m);
        __syncwarp();
        if (qn >= 32) {
            const uint32_t e = wq.q[lane];
            const uint32_t carry = wq.q[32 + lane];
            __syncwarp();
            eval_pair_bwd(p, wq, e, kz, inv_mult);
            qn -= 32;
            if (lane < qn) wq.q[lane] = carry;
            __syncwarp();
        }
    }
    if (lane < qn) eval_pair_bwd(p, wq, wq.q[lane], kz, inv_mult);
    __syncwarp();
    for (int idx = lane; idx < 6 * FPW; idx += 32) {
        const int kk = idx / FPW, sl = idx - kk * FPW;
        const int fg = sl * nwarps + gwarp;
        if (fg < p.B * p.F) {
            const float v = wq.facc[kk][sl];
            if (v != 0.0f) atomicAdd(p.gfacc + (size_t)fg * MM_GF + kk, v);
        }
    }
}

// hard pass kernel.  Set-up per warp (lanes 0..7: record -> smem, the face's exact tight rectangle); the (face, pixel)
// pairs of the CTA's 64 faces are numbered TOGETHER and dealt to its 256 threads: a warp's 8 faces hold between 24 and 655
// pairs at cfg-2 (median 91) and a warp's run time follows its pair count (per-warp timeline, profiles/r2_notes.md: correlation
// 0.67, slowest warp 10 us in a kernel whose median warp takes 4.6 us) -- and the kernel, one wave, lasts as long as its
// slowest warp.  Over 64 faces the counts even out.  The slot search is a 6-step binary search over the shared prefix.
#define HARD_WARPS 8
__global__ void __launch_bounds__(32 * HARD_WARPS)
k_scatter_hard(const mm_raster_params p)
{
    mm_pdl_prologue((p.pdl_late & 1) != 0);
    __shared__ WarpQ s_wq[HARD_WARPS];
    __shared__ int s_pre[HARD_WARPS * FPW + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // warp w of CTA c is "warp" w * gridDim.x + c of the batch-wide dealing: the CTA's 8 warps sit in 8 different image groups,
    // so its 64 faces come from (up to) 64 different images and every CTA sees the same mix of near and far cameras (with 8
    // CONSECUTIVE warps a CTA held 8 neighbouring faces of the same 8 images: 795 pairs at the median, 3143 at the maximum)
    const int nwarps = (p.B * p.F + FPW - 1) / FPW;
    const int gwarp = (warp * (int)gridDim.x + (int)blockIdx.x < nwarps) ? warp * (int)gridDim.x + (int)blockIdx.x : nwarps;
    MM_PROF_MARK(p.prof, 1, gwarp, 0);
    // ---- set-up: lanes 0..7 of every warp own one face each (dealt with a stride of the warp count: 8 different images)
    if (lane < FPW) {
        WarpQ& wq = s_wq[warp];
        const int slot = lane;
        const int fid = slot * nwarps + gwarp;
        int ix0 = 0, ix1 = -1, iy0 = 0, iy1 = -1;
        wq.img[slot] = 0; wq.face[slot] = 0;
        if (gwarp < nwarps && fid < p.B * p.F) {
            const int b = fid / p.F, f = fid - b * p.F;
            const FaceRec r = load_rec(p.frec + (size_t)b * p.F * MM_REC_FLOATS, f);
            wq.rec[0][slot] = r.ax; wq.rec[1][slot] = r.ay; wq.rec[2][slot] = r.bx; wq.rec[3][slot] = r.by;
            wq.rec[4][slot] = r.cx; wq.rec[5][slot] = r.cy; wq.rec[6][slot] = r.az; wq.rec[7][slot] = r.bz;
            wq.rec[8][slot] = r.cz; wq.img[slot] = b; wq.face[slot] = f;
            const uint4 rc = p.frect[fid];                 // exact tight rectangle (empty for a back face), from the vertex stage
            rect_unpack(rc.x, rc.y, ix0, ix1, iy0, iy1);
        }
        const int w = ix1 - ix0 + 1, h = iy1 - iy0 + 1;
        wq.ix0[slot] = ix0; wq.iy0[slot] = iy0; wq.w[slot] = w > 0 ? w : 1;
        wq.rw[slot] = __frcp_rn((float)(w > 0 ? w : 1));
        s_pre[1 + warp * FPW + slot] = (w > 0 && h > 0) ? w * h : 0;
    }
    __syncthreads();
    // ---- inclusive prefix of the 64 pair counts (warp 0, two values per lane)
    if (warp == 0) {
        int v0 = s_pre[1 + lane], v1 = s_pre[33 + lane];
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t0 = __shfl_up_sync(FULL, v0, o), t1 = __shfl_up_sync(FULL, v1, o);
            if (lane >= o) { v0 += t0; v1 += t1; }
        }
        const int half = __shfl_sync(FULL, v0, 31);
        s_pre[1 + lane] = v0; s_pre[33 + lane] = half + v1;
        if (lane == 0) s_pre[0] = 0;
    }
    __syncthreads();
    const int total = s_pre[HARD_WARPS * FPW];
#ifdef MM_PROF
    if (p.prof && lane == 0 && gwarp < 16384) p.prof[((size_t)1 * 16384 + gwarp) * 4 + 3] = (unsigned long long)total;
#endif
    // ---- the CTA's pairs, dealt to its 256 threads round-robin
    #pragma unroll 1
    for (int k = threadIdx.x; k < total; k += 32 * HARD_WARPS) {
        int g = 0;
        #pragma unroll
        for (int step = HARD_WARPS * FPW / 2; step > 0; step >>= 1) g += (s_pre[g + step] <= k) ? step : 0;
        WarpQ& wq = s_wq[g >> 3];
        const int slot = g & 7;
        const int local = k - s_pre[g];
        const int w = wq.w[slot];
        int dy = (int)(((float)local + 0.5f) * wq.rw[slot]);
        int dx = local - dy * w;
        if (dx < 0) { --dy; dx += w; } else if (dx >= w) { ++dy; dx -= w; }
        const int ix = wq.ix0[slot] + dx, iy = wq.iy0[slot] + dy;
        // exact DIB-R inside test + depth, one atomicMax on the packed (depth, ~face) key, one bit of the coverage bitmap
        Bary bb;
        const FaceRec fr = slot_rec(wq, slot);
        if (!bary_eval_inside(fr, pix_x(ix, p.W, p.sx), pix_y(iy, p.H, p.sy), p.eps, bb)) continue;
        const float zz = ADD(ADD(MUL(bb.w0, fr.az), MUL(bb.w1, fr.bz)), MUL(bb.w2, fr.cz));
        const int b = wq.img[slot];
        atomicMax(p.zbuf + ((size_t)b * p.H + iy) * p.W + ix, depth_key(zz, wq.face[slot]));
        atomicOr(p.cov + ((size_t)b * p.H + iy) * p.covw + (ix >> 5), 1u << (ix & 31));
    }
    // The hard pass is issue-bound and leaves the memory system idle: it clears the step's texture-gradient buffer on the side --
    // at the END of the CTA: nothing waits for these stores, and in front of the set-up they held every CTA's first loads back
    // (ncu source page: 15 % of the kernel's stall samples on this loop when it came first).
    if (p.nclr) {
        const size_t nthreads = (size_t)gridDim.x * blockDim.x;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.nclr; i += nthreads) p.clr[i] = z;
    }
    MM_PROF_MARK(p.prof, 1, gwarp, 2);
}

// ---------------------------------------------------------------------------------------------- soft pass forward
// (engine in mm_soft_fwd.cuh, shared with the fused step's merged soft + shading kernel)
#ifndef SF_MINB
#define SF_MINB 10
#endif
__global__ void __launch_bounds__(32 * SF_WARPS, SF_MINB)
k_soft_fwd(const mm_raster_params p)
{
    mm_pdl_prologue((p.pdl_late & 2) != 0);
    __shared__ SoftQ s_wq[SF_WARPS];
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    MM_PROF_MARK(p.prof, 2, gwarp, 0);
    soft_fwd_role(p, s_wq[threadIdx.x >> 5], gwarp);
    MM_PROF_MARK(p.prof, 2, gwarp, 2);
}

// ---------------------------------------------------------------------------------------------- soft pass backward, list-driven
// Replays the forward's dense candidate list: one pair per lane, no bbox walk, no filtering.
#define SB_THREADS 128
__device__ __forceinline__ void soft_bwd_list_role(const mm_raster_params& p, WarpQ* s_wq, const int vblock, const int nvblocks)
{
    const uint32_t n = p.ovf_count[1];
    if (n > p.plist_cap) {                  // the forward's pair list overflowed its buffer (never at the template meshes'
        const int nwarps = (p.B * p.F + FPW - 1) / FPW;           // shapes): redo the bbox walk with the filtering pair engine
        const int wstride = (nvblocks * SB_THREADS) >> 5;
        for (int gw = (vblock * SB_THREADS + threadIdx.x) >> 5; gw < nwarps; gw += wstride) {
            scatter_warp_bwd(p, s_wq[threadIdx.x >> 5], gw, nwarps);
            __syncwarp();
        }
        return;
    }
    const int lane = threadIdx.x & 31;
    const size_t HW = (size_t)p.H * p.W;
    const float kz = p.sigmainv / p.multiplier / p.multiplier;
    const float inv_mult = 1.0f / p.multiplier;
    const uint32_t stride = (uint32_t)nvblocks * SB_THREADS;
    for (uint32_t i0 = ((uint32_t)vblock * SB_THREADS + threadIdx.x) & ~31u; i0 < n; i0 += stride) {
        const uint32_t i = i0 + lane;
        float ga[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        uint32_t fg = 0u;
        bool live = false;
        if (i < n) {
            const unsigned long long e = p.plist[i];
            fg = (uint32_t)(e >> 32);
            const int iy = (int)((e >> 12) & 0xfffu), ix = (int)(e & 0xfffu);
            const int b = (int)(fg / (uint32_t)p.F);
            const size_t pg = (size_t)b * HW + (size_t)iy * p.W + ix;
            // the face record is fetched TOGETHER with the pixel's words (its address depends on the list entry alone): one
            // round trip less in a kernel that is a chain of them
            const float4* q4 = reinterpret_cast<const float4*>(p.frec) + (size_t)fg * 3;
            const float4 c0 = __ldg(q4), c1 = __ldg(q4 + 1);
            // the silhouette value comes from the workspace's accumulator (bit-identical to the image's alpha plane, which the
            // backward therefore does not need: the caller may have edited the image in place)
            const unsigned long long acc = p.lacc[pg];
            const float g = gsoft_at(p, b, (size_t)iy * p.W + ix);
            const float soft = lacc_soft(acc);
            if (g != 0.0f && soft > 0.0f && lacc_count(acc) != (int)MM_LACC_OVF) {
                FaceRec r;
                r.ax = c0.x; r.ay = c0.y; r.bx = c0.z; r.by = c0.w; r.cx = c1.x; r.cy = c1.y;
                r.az = r.bz = r.cz = r.nx = r.ny = r.nz = 0.0f;
                soft_pair_grad(p, r, pix_x(ix, p.W, p.sx), pix_y(iy, p.H, p.sy), kz, inv_mult, g, 1.0f - soft, ga);
                live = true;
            }
        }
        // the six corner gradients of the pair's face leave as one 16-byte + one 8-byte vector RED (a warp-wide RED is ONE
        // instruction: pre-combining lanes that share a face with match_any + shuffles costs more than it saves)
        if (live) red_add_corners(p.gfacc + (size_t)fg * MM_GF, ga);
    }
}

// backward: ONE launch for the two independent halves of the soft-silhouette backward -- the CTAs from `novf` on replay the
// pair list, the rest redo the truncated pixels (both only add into the per-face accumulators)
__global__ void __launch_bounds__(SB_THREADS)
k_soft_bwd(const mm_raster_params p, const int novf)
{
    mm_pdl_prologue((p.pdl_late & 16) != 0);
    __shared__ WarpQ s_wq[SB_THREADS / 32];
    __shared__ uint32_t s_mask[OVF_MAX_WORDS];
    __shared__ int s_kept[MM_MAX_KNUM];
    MM_PROF_MARK(p.prof, 4, blockIdx.x * (SB_THREADS / 32) + (threadIdx.x >> 5), 0);
    // The truncated pixels are dealt to ALL CTAs of the grid (entry e to CTA e mod grid): each costs a CTA ~3 us of pure latency
    // (~8 us with 5120 faces).  The first `novf` CTAs have nothing else to do and come FIRST in the grid, so the usual handful
    // of entries starts at once (as the kernel's last CTAs they were its tail); the list CTAs behind them take their share
    // after their part of the pair list, which only matters when there are thousands (far cameras, sphere2: 16 k).
    if ((int)blockIdx.x >= novf) soft_bwd_list_role(p, s_wq, blockIdx.x - novf, gridDim.x - novf);
    const uint32_t ntrunc = p.ovf_count[0];
    if (ntrunc > blockIdx.x) soft_ovf_role<true>(p, s_mask, s_kept, ntrunc, blockIdx.x, gridDim.x);
    MM_PROF_MARK(p.prof, 4, blockIdx.x * (SB_THREADS / 32) + (threadIdx.x >> 5), 2);
}

}  // namespace

cudaError_t mm_launch_geom_fwd(const mm_ctx* c, const mm_raster_params& p, cudaStream_t s)
{
    const bool pdl = c->pdl != 0;
    const int warps = (p.B * c->F + FPW - 1) / FPW;
    cudaError_t e = mm_launch(k_scatter_hard, dim3((warps + HARD_WARPS - 1) / HARD_WARPS), dim3(32 * HARD_WARPS), 0, s, pdl, p);
    if (e != cudaSuccess) return e;
    const int nw = (p.B * c->F + SF_FPW - 1) / SF_FPW;
    return mm_launch(k_soft_fwd, dim3((nw + SF_WARPS - 1) / SF_WARPS), dim3(32 * SF_WARPS), 0, s, pdl, p);
}

cudaError_t mm_launch_geom_bwd(const mm_ctx* c, const mm_raster_params& p, cudaStream_t s)
{
    static_assert(SB_THREADS == OVF_THREADS, "the merged backward kernel runs both roles with one CTA shape");
    // (swept in round 2: 16 + 8 CTAs per SM 0.0930 ms, 8 + 8 0.0930, 8 + 2 0.0923, 16 + 2 0.0927, 4 + 2 0.0939)
    const int nlist = c->num_sms * 8, novf = c->num_sms * 2;
    return mm_launch(k_soft_bwd, dim3(nlist + novf), dim3(SB_THREADS), 0, s, c->pdl != 0, p, novf);
}

