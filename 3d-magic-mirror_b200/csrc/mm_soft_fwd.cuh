// mm_soft_fwd.cuh -- the soft silhouette (DIBR_SPEC A.4) as device functions shared by mm_raster.cu and mm_fused.cu:
//   soft_fwd_role        the forward walk of k_soft_fwd (candidate search on the coverage bitmap, evaluation, pair list)
//   shade_sched_produce  the shading schedule k_soft_fwd writes on the side (strips classed by their covered pixels)
//   soft_pair_grad       one candidate's gradient (k_soft_bwd)
//   soft_ovf_role        the exact ordered re-scan of a truncated pixel: forward inside the shading kernel, backward in k_soft_bwd
#pragma once
#include "mm_device.cuh"

namespace {

#ifndef FULL
#define FULL 0xffffffffu
#endif

// ---------------------------------------------------------------------------------------------- soft pass forward
// The candidates of the soft pass are the UNCOVERED pixels inside a face's enlarged bbox -- for most faces (the interior of
// the object) there are none.  The pair engine above found that out with one 8-byte zbuf load + ~50 bookkeeping
// instructions PER BBOX PIXEL (ncu r1c: 65 % of the kernel's 18.8 M warp-instructions).  Here every lane owns one face and
// walks the rows of its rectangle against the coverage bitmap the hard pass left behind: one 32-bit word per (row, 32-pixel
// column block) tells which of its pixels are uncovered.  The set bits become queue entries (warp scan + per-lane bit
// loop); whenever 32 are waiting the whole warp evaluates them, one candidate per lane, exactly as before.
#define SF_QCAP (1024 + 64)
#define SF_DCAP 192          // evaluated candidates staged per warp before they go to the global pair list

struct SoftQ {
    uint32_t q[SF_QCAP];     // pending candidates: slot << 24 | iy << 12 | ix
    uint32_t done[SF_DCAP];  // evaluated candidates (same encoding) waiting for their slots in the pair list
    float rec[6][8];         // the warp's 8 faces: image-plane corners
    int img[8];
    int face[8];
};

// One candidate: distance, probability, ONE integer atomicAdd into the pixel's accumulator.  Returns the accumulator's previous
// word; the caller looks at it one round LATER (soft_fwd_check), so that the atomic's round trip overlaps the next round's
// arithmetic instead of ending every round (the slowest warps of the kernel are the ones with five and more rounds).
__device__ __forceinline__ unsigned long long soft_fwd_eval(const mm_raster_params& p, const SoftQ& wq, uint32_t e, float kz)
{
    const int slot = (int)(e >> 24), iy = (int)((e >> 12) & 0xfffu), ix = (int)(e & 0xfffu);
    const size_t pg = (size_t)wq.img[slot] * p.H * p.W + (size_t)iy * p.W + ix;
    FaceRec r;
    r.ax = wq.rec[0][slot]; r.ay = wq.rec[1][slot]; r.bx = wq.rec[2][slot]; r.by = wq.rec[3][slot];
    r.cx = wq.rec[4][slot]; r.cy = wq.rec[5][slot];
    r.az = r.bz = r.cz = r.nx = r.ny = r.nz = 0.0f;
    int type;
    const float d2 = soft_d2_fast(r, pix_x(ix, p.W, p.sx), pix_y(iy, p.H, p.sy), p.multiplier, type);
    const float prob = soft_prob_fast(d2, kz);
    return atomicAdd(p.lacc + pg, lacc_term(log1pf(-prob)));
}
__device__ __forceinline__ void soft_fwd_check(const mm_raster_params& p, const SoftQ& wq, unsigned long long old, uint32_t e)
{
    if (lacc_count(old) == p.knum) {                            // candidate knum+1: the pixel needs the ordered pass
        const int slot = (int)(e >> 24), iy = (int)((e >> 12) & 0xfffu), ix = (int)(e & 0xfffu);
        const uint32_t s2 = atomicAdd(p.ovf_count, 1u);
        p.ovf_list[s2] = (uint32_t)((size_t)wq.img[slot] * p.H * p.W + (size_t)iy * p.W + ix);
    }
}

// the pair-list entry of candidate e of this warp
__device__ __forceinline__ void soft_fwd_store_pair(const mm_raster_params& p, const SoftQ& wq, uint32_t e, uint32_t at)
{
    if (at < p.plist_cap) {
        const int slot = (int)(e >> 24);
        const unsigned long long fg = (unsigned long long)((size_t)wq.img[slot] * p.F + wq.face[slot]);
        p.plist[at] = (fg << 32) | (unsigned long long)(e & 0xffffffu);
    }
}
// the staged candidates go to the global pair list: ONE atomicAdd for up to SF_DCAP of them (it was one per round of 32: a
// second round trip at the end of every round); all lanes call it
__device__ __forceinline__ void soft_fwd_flush(const mm_raster_params& p, const SoftQ& wq, int dn, int lane)
{
    if (dn == 0) return;
    uint32_t base = 0u;
    if (lane == 0) base = atomicAdd(p.ovf_count + 1, (uint32_t)dn);
    base = __shfl_sync(FULL, base, 0);
    for (int i = lane; i < dn; i += 32) soft_fwd_store_pair(p, wq, wq.done[i], base + (uint32_t)i);
}

// ---------------------------------------------------------------------------------------------- shading schedule
// The shading kernel's CTAs cost between ~3 us (a strip of background) and ~20 us (a strip full of covered pixels: four rounds
// of its dense pass), and with ~2.6 CTAs per resident slot the grid-order dispatch left a 15 us drain behind the last wave
// (per-warp timeline, profiles/r2_notes.md; list-scheduling the measured run times longest-first: 33.5 -> 26.9 us).  So the
// soft pass, which runs on the final coverage bitmap anyway, classes every strip by the rounds it will need -- one warp per
// strip: 32 lanes = 4 tiles x 8 rows, one 16-bit slice of a bitmap word each -- and the shading CTAs take the strips longest
// class first.
__device__ __forceinline__ void shade_sched_produce(const mm_raster_params& p, const int widx, const int nwarps, const int lane)
{
    static_assert(MM_SH_WARPS * MM_SH_TH == 32 && MM_SH_TW == 16, "one lane per (tile, row), one 16-bit slice of a bitmap word each");
    const int ntx = (p.W + MM_SH_TW - 1) / MM_SH_TW, ntiles = ntx * ((p.H + MM_SH_TH - 1) / MM_SH_TH);
    const int nsid = p.B * p.nstrips;
    for (int sid = widx; sid < nsid; sid += nwarps) {
        const int b = sid / p.nstrips, s = sid - b * p.nstrips;
        const int tile = s * MM_SH_WARPS + (lane >> 3);
        int cnt = 0;
        if (tile < ntiles) {
            const int ty = tile / ntx, tx = tile - ty * ntx;
            const int iy = ty * MM_SH_TH + (lane & 7), x0 = tx * MM_SH_TW;
            if (iy < p.H) cnt = __popc((p.cov[((size_t)b * p.H + iy) * p.covw + (x0 >> 5)] >> (x0 & 31)) & 0xffffu);
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
        // entry: strip id + the image's lights (lanes 1..9), so that a shading CTA gets both with one round trip
        const float lt = (lane >= 1 && lane <= 9) ? p.lights[b * 9 + lane - 1] : 0.0f;
        const int k = min(4, (cnt + 32 * MM_SH_WARPS - 1) / (32 * MM_SH_WARPS));
        uint32_t pos = 0u;
        if (lane == 0) pos = atomicAdd(p.sched_n + k, 1u);
        pos = __shfl_sync(FULL, pos, 0);
        if (lane < 10) p.sched_list[((size_t)k * nsid + pos) * MM_SCHED_WORDS + lane] = lane == 0 ? (uint32_t)sid : __float_as_uint(lt);
    }
}

#define SF_WARPS 4
#define SF_FPW 8             // faces per warp: 4 lanes share a face and take its (row, column-block) segments round-robin, so the
                             // dependent chain of bitmap loads per warp is 4x shorter and there are 4x more warps to overlap it
__device__ __forceinline__ void soft_fwd_role(const mm_raster_params& p, SoftQ& wq, const int gwarp)
{
    const int lane = threadIdx.x & 31;
    const int nwarps = (p.B * p.F + SF_FPW - 1) / SF_FPW;
    if (gwarp >= nwarps) return;
    shade_sched_produce(p, gwarp, nwarps, lane);
    const float kz = p.sigmainv / p.multiplier / p.multiplier;

    // ---- set-up: lane = (slot, sub); faces dealt with a stride of the warp count (a warp mixes 8 images: balanced)
    const int slot = lane >> 2, sub = lane & 3;
    const int fid = slot * nwarps + gwarp;
    int ix0 = 0, ix1 = -1, iy0 = 0, iy1 = -1, b = 0;
    if (fid < p.B * p.F) {
        b = fid / p.F;
        const int f = fid - b * p.F;
        const float4* q4 = reinterpret_cast<const float4*>(p.frec) + (size_t)fid * 3;
        const float4 c0 = __ldg(q4), c1 = __ldg(q4 + 1);
        FaceRec r;
        r.ax = c0.x; r.ay = c0.y; r.bx = c0.z; r.by = c0.w; r.cx = c1.x; r.cy = c1.y;
        r.az = r.bz = r.cz = r.nx = r.ny = r.nz = 0.0f;
        if (sub == 0) {
            wq.rec[0][slot] = r.ax; wq.rec[1][slot] = r.ay; wq.rec[2][slot] = r.bx; wq.rec[3][slot] = r.by;
            wq.rec[4][slot] = r.cx; wq.rec[5][slot] = r.cy; wq.img[slot] = b; wq.face[slot] = f;
        }
        const uint4 rc = p.frect[fid];                        // exact enlarged rectangle, from the vertex stage
        rect_unpack(rc.z, rc.w, ix0, ix1, iy0, iy1);
    }
    __syncwarp();
    // ---- (row, 32-column block) segments of the face's rectangle, row-major; this lane takes segments sub, sub+4, ..
    const int wd0 = ix0 >> 5;
    const int nwd = (ix1 >= ix0 && iy1 >= iy0) ? (ix1 >> 5) - wd0 + 1 : 0;
    const int nseg_face = nwd * (iy1 - iy0 + 1);
    const int nseg = (nseg_face - sub + 3) >> 2;
    const uint32_t* covb = p.cov + (size_t)b * p.H * p.covw;
    int maxseg = nseg;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxseg = max(maxseg, __shfl_xor_sync(FULL, maxseg, o));
    int qn = 0, dn = 0;
#ifdef MM_PROF
    int ncand = 0;
#endif
    unsigned long long pend_old = 0ull;      // the previous round's accumulator word and candidate of this lane
    uint32_t pend_e = 0u;
    bool pend = false;
    // running (row, word) of this lane's segment `it * 4 + sub`, advanced by 4 segments per iteration without a division
    int row = iy0, wd = wd0 + sub;
    if (nwd > 0) while (wd >= wd0 + nwd) { wd -= nwd; ++row; }
    #pragma unroll 1
    for (int it = 0; it < maxseg; ++it) {
        uint32_t bits = 0u;
        if (it < nseg) {
            const int lo = max(ix0 - (wd << 5), 0), hi = min(ix1 - (wd << 5), 31);         // column range inside this word
            const uint32_t colmask = (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo);
            bits = ~__ldg(covb + (size_t)row * p.covw + wd) & colmask;                       // uncovered pixels of the segment
        }
        if (__any_sync(FULL, bits != 0u)) {
            // exclusive scan of the per-lane candidate counts -> queue offsets
            const int c = __popc(bits);
            int incl = c;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
            int pos = qn + incl - c;
            const uint32_t hdr = ((uint32_t)slot << 24) | ((uint32_t)row << 12) | (uint32_t)(wd << 5);
            while (bits) {
                const int j = __ffs(bits) - 1;
                bits &= bits - 1u;
                wq.q[pos++] = hdr + (uint32_t)j;
            }
            qn += __shfl_sync(FULL, incl, 31);
#ifdef MM_PROF
            ncand += __shfl_sync(FULL, incl, 31);
#endif
            __syncwarp();
            while (qn >= 32) {                                   // evaluate from the top of the queue: no shifting
                qn -= 32;
                const uint32_t e = wq.q[qn + lane];
                const unsigned long long old = soft_fwd_eval(p, wq, e, kz);
                if (pend) soft_fwd_check(p, wq, pend_old, pend_e);
                pend_old = old; pend_e = e; pend = true;
                if (dn + 32 > SF_DCAP) { __syncwarp(); soft_fwd_flush(p, wq, dn, lane); dn = 0; __syncwarp(); }
                wq.done[dn + lane] = e;
                dn += 32;
            }
            __syncwarp();
        }
        if (it < nseg) {
            if (nwd == 1) row += 4;
            else { wd += 4; while (wd >= wd0 + nwd) { wd -= nwd; ++row; } }
        }
    }
    // The end of the warp is a chain of round trips; they are made to overlap: the slots of ALL staged candidates (including the
    // tail about to be evaluated) are requested from the global pair list FIRST, the tail is evaluated under that round trip,
    // then the pairs are stored, and the last looks at the accumulators' previous words come behind everything.
    const int ntot = dn + qn;
    uint32_t base = 0u;
    if (lane == 0 && ntot > 0) base = atomicAdd(p.ovf_count + 1, (uint32_t)ntot);
    unsigned long long last_old = 0ull;
    uint32_t last_e = 0u;
    const bool last = lane < qn;
    if (last) {
        last_e = wq.q[lane];
        last_old = soft_fwd_eval(p, wq, last_e, kz);
    }
    if (ntot > 0) {
        base = __shfl_sync(FULL, base, 0);
        for (int i = lane; i < dn; i += 32) soft_fwd_store_pair(p, wq, wq.done[i], base + (uint32_t)i);
        if (last) soft_fwd_store_pair(p, wq, last_e, base + (uint32_t)(dn + lane));
    }
    if (pend) soft_fwd_check(p, wq, pend_old, pend_e);
    if (last) soft_fwd_check(p, wq, last_old, last_e);
#ifdef MM_PROF
    if (p.prof && lane == 0 && gwarp < 16384) p.prof[((size_t)2 * 16384 + gwarp) * 4 + 3] = ((unsigned long long)maxseg << 32) | (unsigned)ncand;
#endif
}

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



// ---------------------------------------------------------------------------------------------- overflow (ordered) pass
// DIB-R keeps only the FIRST knum candidates in face-index order (DIBR_SPEC A.4).  Pixels that saw more are re-done
// here literally.  One CTA per overflowed pixel:
//   phase 1  all F enlarged-bbox tests (each thread owns F/128 faces, OVF_UNROLL loads in flight; the test is read off the exact
//            rectangles of frect: 8 bytes and four integer compares per face); the per-warp ballots land in shared memory as
//            hit words in face order.
//   phase 2  warp 0 keeps the first knum set bits (running count over the words) and writes the kept faces, in order, to a
//            shared list; then -- still warp 0, no further block barrier -- lane k evaluates candidate k, and the ordered
//            product (forward) is folded with shuffles exactly in the reference's order; backward: lane k scatters the
//            gradient of candidate k.
// A handful of pixels per step take this path (far cameras): pure latency, ~2 round trips per pixel.  Forward: run by the
// shading kernel's CTAs (mm_fused.cu); backward: by the tail CTAs of k_soft_bwd.
#define OVF_THREADS 128
#define OVF_MAX_WORDS 2048          // F <= 65535
#define OVF_UNROLL 8

// (this CTA takes entries first, first + stride, .. of the `count` listed pixels)
template <bool BWD>
__device__ __forceinline__ void soft_ovf_role(const mm_raster_params& p, uint32_t* s_mask, int* s_kept, const uint32_t count,
                                              const int first, const int stride)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t HW = (size_t)p.H * p.W;
    const uint32_t n = count;
    const float kz = p.sigmainv / p.multiplier / p.multiplier;
    const float inv_mult = 1.0f / p.multiplier;
    const int nw = (p.F + 31) >> 5;
    const int niter = (p.F + OVF_THREADS - 1) / OVF_THREADS;              // faces per thread
    for (uint32_t e = (uint32_t)first; e < n; e += (uint32_t)stride) {
        const size_t pg = p.ovf_list[e];
        const int b = (int)(pg / HW);
        const int pix = (int)(pg - (size_t)b * HW);
        const int iy = pix / p.W, ix = pix - iy * p.W;
        const float px = pix_x(ix, p.W, p.sx), py = pix_y(iy, p.H, p.sy);
        const float4* rec4 = reinterpret_cast<const float4*>(p.frec + (size_t)b * p.F * MM_REC_FLOATS);
        float g = 0.0f, one_m_all = 0.0f;
        if (BWD) {
            g = gsoft_at(p, b, (size_t)pix);
            const float soft = lacc_soft(p.lacc[pg]);           // (the exact word stored by the forward)
            one_m_all = 1.0f - soft;
            if (g == 0.0f || !(soft > 0.0f)) continue;          // block-uniform
        }
        // ---- phase 1: hit words in face order (word = f >> 5).  "The pixel is inside the face's enlarged bbox" is read off the
        // EXACT enlarged rectangle the vertex stage left in frect -- by construction the same decision as the reference's
        // half-open fp32 test on the record, for 8 bytes and four integer compares per face instead of 32 bytes and the
        // min / max / add chain; OVF_UNROLL independent loads in flight
        const uint4* rects = p.frect + (size_t)b * p.F;
        for (int j0 = 0; j0 < niter; j0 += OVF_UNROLL) {
            uint2 rc[OVF_UNROLL];
            #pragma unroll
            for (int u = 0; u < OVF_UNROLL; ++u) {
                const int f = (j0 + u) * OVF_THREADS + threadIdx.x;
                rc[u] = (j0 + u < niter && f < p.F) ? __ldg(reinterpret_cast<const uint2*>(rects + f) + 1) : make_uint2(1u, 1u);   // (.z, .w; (1, 0) = empty)
            }
            #pragma unroll
            for (int u = 0; u < OVF_UNROLL; ++u) {
                const int j = j0 + u;
                const bool hit = ix >= (int)(rc[u].x & 0xffffu) && ix <= (int)(rc[u].x >> 16) &&
                                 iy >= (int)(rc[u].y & 0xffffu) && iy <= (int)(rc[u].y >> 16);
                const uint32_t m = __ballot_sync(FULL, hit);
                const int word = j * (OVF_THREADS / 32) + warp;
                if (lane == 0 && j < niter && word < nw) s_mask[word] = m;
            }
        }
        __syncthreads();
        // ---- phase 2 (warp 0): first knum set bits over all words -> ordered list of kept faces
        if (warp == 0) {
            int seen = 0;
            for (int w0 = 0; w0 < nw && seen < p.knum; w0 += 32) {
                const int wd = w0 + lane;
                uint32_t m = (wd < nw) ? s_mask[wd] : 0u;
                const int c = __popc(m);
                int incl = c;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
                int pos = seen + incl - c;                       // candidates before this word
                while (m && pos < p.knum) {                       // this word's bits, in face order
                    const int bit = __ffs(m) - 1;
                    m &= m - 1u;
                    s_kept[pos++] = (wd << 5) + bit;
                }
                seen += __shfl_sync(FULL, incl, 31);
            }
            const int nk = min(seen, p.knum);
            __syncwarp();
            // ---- the kept candidates, one per lane (knum <= 64: two rounds at most)
            float allprob = 1.0f;
            for (int k0 = 0; k0 < nk; k0 += 32) {
                const int k = k0 + lane;
                const bool mine = k < nk;
                const int f = mine ? s_kept[k] : 0;
                FaceRec r;
                {
                    const float4 c0 = __ldg(rec4 + (size_t)f * 3), c1 = __ldg(rec4 + (size_t)f * 3 + 1);
                    r.ax = c0.x; r.ay = c0.y; r.bx = c0.z; r.by = c0.w; r.cx = c1.x; r.cy = c1.y;
                    r.az = r.bz = r.cz = r.nx = r.ny = r.nz = 0.0f;
                }
                if (BWD) {
                    if (mine) {
                        float ga[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
                        soft_pair_grad(p, r, px, py, kz, inv_mult, g, one_m_all, ga);
                        red_add_corners(p.gfacc + ((size_t)b * p.F + f) * MM_GF, ga);
                    }
                } else {
                    float prob = 0.0f;
                    if (mine) { int type; prob = soft_prob_fast(soft_d2_fast(r, px, py, p.multiplier, type), kz); }
                    const int cnt = min(32, nk - k0);
                    #pragma unroll 1
                    for (int q = 0; q < cnt; ++q)                // the reference's ordered product
                        allprob = allprob * (1.0f - __shfl_sync(FULL, prob, q));
                }
            }
            if (!BWD && lane == 0) p.lacc[pg] = lacc_exact(allprob > 0.0f ? logf(allprob) : -2400.0f);
        }
        __syncthreads();
    }
}

}  // namespace
