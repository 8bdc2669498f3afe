// mm_raster.cu -- the GEOMETRY half of the per-pixel stage: DIB-R hard visibility and soft silhouette, forward and
// backward.  Replaces kaolin dibr_rasterization's four CUDA kernels (call site networks.py:297-299).  Shading and
// the loss live in mm_shade.cu; this file only visits sub-tiles whose face list is non-empty.
//
// Work decomposition (B200: 148 SMs):
//   grid = (G, B) PERSISTENT single-warp CTAs, G*B ~ 24 warps per SM.  The G warps of image b pull 8x4-pixel
//   sub-tiles (lane = pixel) from the image's compacted list of non-empty sub-tiles through an atomic ticket.
//   Silhouette tiles cost ~100x a tile with one face; in a multi-warp CTA finished warps park at the final barrier
//   (ncu: 66 % of stall samples), so the unit of scheduling is one warp and the queue balances tiles individually.
//     1. TMA prefetch: while a tile is processed, the next tile's two bitmask rows ("tile face lists", written by the
//        vertex stage) are already in flight -- cp.async.bulk into the other half of a double buffer, completion on
//        an mbarrier.
//     2. Pair engine (see pair_batch): faces -> hit words -> ranked dense pair list -> per-pair arithmetic ->
//        per-pixel ordered fold.  Bit-exact DIB-R decisions (mm_device.cuh).
//   Backward re-derives the per-pixel candidate lists from the same masks instead of storing Kaolin's knum-deep
//   side buffers (B*H*W*30*(4+8+1) B = 307 MB at B=48,128^2).
#include "mm_device.cuh"

namespace {

#define FULL 0xffffffffu

// per-warp scratch in shared memory (one sub-tile = one warp)
struct WarpScratch {
    uint32_t fq[64];               // face queue: set bits of the mask row compacted into dense batches of 32
    uint32_t fid[32];              // face ids of the current batch (slot j = lane j's face)
    uint32_t hit[32];              // per pixel: which slots (faces) of the current batch hit it
    uint32_t cnt[32];              // per pixel: candidates seen so far (DIB-R's knum cap)
    float gs[32];                  // bwd: upstream gradient of the silhouette per pixel
    float oma[32];                 // bwd: 1 - soft per pixel
    float rec[9][32];              // the batch's face records (ax ay bx by cx cy az bz cz), column j = slot j
    float facc[6][32];             // bwd: per-face corner-gradient accumulators of the batch
    uint32_t pr[1024];             // pair list: (slot << 5 | pixel), overwritten in place by the pair's result
    long long dbg[4];              // debug cycle counters (mm_debug_set_profile_buffer)
};

struct CtaCtx {
    const uint32_t* mS;     // current sub-tile's S mask row (shared memory, TMA-staged)
    const uint32_t* mH;     // current sub-tile's H mask row
    WarpScratch* ws;        // this warp's scratch
    const float* rec;       // face records of the current tile's image (global, read through L1)
    int b, st, stx, sty, ix, iy;   // current tile: image, sub-tile, pixel of this lane
    bool active;
    // persistent-warp state
    uint64_t* bar;          // two mbarriers (double-buffered mask rows)
    uint32_t* buf;          // [2][2*nwords]
    uint32_t* ticket;       // global ticket counter of this pass
    int total;              // length of the batch-wide list of non-empty sub-tiles
    int first;              // 1 until the warp has taken its static first tile
    int nb, nxt, k;
    uint32_t phase0, phase1;
};

// dynamic smem: | 2 mbarriers 16 B | 2 x (S row + H row) | WarpScratch |
__host__ __device__ inline size_t raster_smem(int nwords, int knum) {
    (void)knum;
    return 16 + 4 * (size_t)nwords * 4 + sizeof(WarpScratch);
}

// Next entry of the batch-wide work list: the first tile of every warp is static (its own index), the rest come from
// one atomic ticket, so images with few non-empty tiles (far camera) cost nothing and warps never idle while any
// image still has work.  Returns st = -1 when the list is exhausted.
__device__ __forceinline__ void fetch_tile(const mm_raster_params& p, CtaCtx& c, int lane, int& ob, int& ost) {
    int t = 0;
    if (c.first) { t = (int)blockIdx.x; c.first = 0; }
    else {
        if (lane == 0) t = (int)atomicAdd(c.ticket, 1u) + (int)gridDim.x;
        t = __shfl_sync(FULL, t, 0);
    }
    ob = 0; ost = -1;
    if (t >= c.total) return;
    uint32_t e = 0u;
    if (lane == 0) e = p.glist[t];
    e = __shfl_sync(FULL, e, 0);
    ob = (int)(e >> 16); ost = (int)(e & 0xffffu);
}

__device__ __forceinline__ void issue_masks(const mm_raster_params& p, CtaCtx& c, int b, int st, int k, int lane) {
    if (lane == 0) {
        const uint32_t bytes = (uint32_t)p.nwords * 4;
        const size_t off = ((size_t)b * p.nst + st) * p.nwords;
        uint64_t* bar = c.bar + k;
        uint32_t* dst = c.buf + (size_t)k * 2 * p.nwords;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(2 * bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(p.maskS + off), "r"(bytes), "r"(smem_u32(bar)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst + p.nwords)), "l"(p.maskH + off), "r"(bytes), "r"(smem_u32(bar)) : "memory");
    }
}

__device__ __forceinline__ void warp_init(const mm_raster_params& p, unsigned char* smem, CtaCtx& c, int which /* 0 fwd, 1 bwd */)
{
    const int lane = threadIdx.x & 31;
    c.bar = reinterpret_cast<uint64_t*>(smem);
    c.buf = reinterpret_cast<uint32_t*>(smem + 16);
    c.ws = reinterpret_cast<WarpScratch*>(c.buf + 4 * p.nwords);
    c.ticket = p.gctr + 1 + which;
    c.total = (int)p.gctr[0];
    c.first = 1;
    if (lane == 0) { mbar_init(c.bar, 1); mbar_init(c.bar + 1, 1); }
    __syncwarp();
    c.k = 0; c.phase0 = 0u; c.phase1 = 0u;
    fetch_tile(p, c, lane, c.b, c.st);
    if (c.st >= 0) issue_masks(p, c, c.b, c.st, 0, lane);
}

// Called at the top of every loop iteration: prefetch the next tile's masks, wait for the current ones.
__device__ __forceinline__ void tile_begin(const mm_raster_params& p, CtaCtx& c, int lane)
{
    fetch_tile(p, c, lane, c.nb, c.nxt);
    if (c.nxt >= 0) issue_masks(p, c, c.nb, c.nxt, c.k ^ 1, lane);
    if (c.k == 0) { mbar_wait(c.bar, c.phase0); c.phase0 ^= 1u; }
    else          { mbar_wait(c.bar + 1, c.phase1); c.phase1 ^= 1u; }
    c.mS = c.buf + (size_t)c.k * 2 * p.nwords;
    c.mH = c.mS + p.nwords;
    c.rec = p.frec + (size_t)c.b * p.F * MM_REC_FLOATS;
    c.sty = c.st / p.nstx; c.stx = c.st - c.sty * p.nstx;
    c.ix = c.stx * MM_ST_W + (lane & 7);
    c.iy = c.sty * MM_ST_H + (lane >> 3);
    c.active = (c.ix < p.W) && (c.iy < p.H);
}

__device__ __forceinline__ void tile_end(CtaCtx& c)
{
    __syncwarp();
    c.b = c.nb;
    c.st = c.nxt;
    c.k ^= 1;
}

// Compacts the set bits of one sub-tile mask row (face-index order) into dense batches of 32 faces and calls
// fn(f) once per batch with ALL lanes converged: lane j gets the j-th face of the batch, or -1.
template <typename Fn>
__device__ __forceinline__ void for_each_batch(const uint32_t* row, int nwords, int lane, uint32_t* fq, Fn fn)
{
    int qn = 0;
    const uint32_t lt = (1u << lane) - 1u;
    #pragma unroll 1
    for (int wd0 = 0; wd0 < nwords; wd0 += 32) {
        const uint32_t w = (wd0 + lane < nwords) ? row[wd0 + lane] : 0u;
        uint32_t nz = __ballot_sync(FULL, w != 0u);
        #pragma unroll 1
        while (nz) {
            const int src = __ffs(nz) - 1;
            nz &= nz - 1;
            const uint32_t m = __shfl_sync(FULL, w, src);
            if ((m >> lane) & 1u) fq[qn + __popc(m & lt)] = (uint32_t)(((wd0 + src) << 5) + lane);
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32) {
                const int f = (int)fq[lane];
                const uint32_t carry = fq[32 + lane];
                __syncwarp();
                fn(f);
                qn -= 32;
                if (lane < qn) fq[lane] = carry;
                __syncwarp();
            }
        }
    }
    if (qn > 0) fn(lane < qn ? (int)fq[lane] : -1);
}

__device__ __forceinline__ bool mask_empty(const uint32_t* row, int nwords, int lane)
{
    uint32_t any = 0u;
    #pragma unroll 1
    for (int wd0 = 0; wd0 < nwords; wd0 += 32) any |= (wd0 + lane < nwords) ? row[wd0 + lane] : 0u;
    return __ballot_sync(FULL, any != 0u) == 0u;
}

// Conservative pixel-index range, clipped to the sub-tile, of the scaled-NDC box [xl,xh) x [yl,yh)
// (the exact half-open tests are redone per pixel).  Returns false if empty.
struct PixRange { int ix0, ix1, iy0, iy1; };
__device__ __forceinline__ bool pix_range(const mm_raster_params& p, const CtaCtx& c, float xl, float xh, float yl, float yh,
                                          PixRange& r)
{
    const float inv_sx = 1.0f / p.sx, inv_sy = 1.0f / p.sy;
    float fx_lo = (xl * inv_sx + (float)(p.W - 1)) * 0.5f;
    float fx_hi = (xh * inv_sx + (float)(p.W - 1)) * 0.5f;
    float fy_lo = ((float)(p.H - 1) - yh * inv_sy) * 0.5f;
    float fy_hi = ((float)(p.H - 1) - yl * inv_sy) * 0.5f;
    if (!(fx_lo == fx_lo) || !(fx_hi == fx_hi)) { fx_lo = -4.0f; fx_hi = 1.0e6f; }
    if (!(fy_lo == fy_lo) || !(fy_hi == fy_hi)) { fy_lo = -4.0f; fy_hi = 1.0e6f; }
    fx_lo = fminf(fmaxf(fx_lo, -4.0f), 1.0e6f); fx_hi = fminf(fmaxf(fx_hi, -4.0f), 1.0e6f);
    fy_lo = fminf(fmaxf(fy_lo, -4.0f), 1.0e6f); fy_hi = fminf(fmaxf(fy_hi, -4.0f), 1.0e6f);
    const int bx = c.stx * MM_ST_W, by = c.sty * MM_ST_H;
    r.ix0 = max((int)floorf(fx_lo), bx);
    r.ix1 = min(min((int)ceilf(fx_hi), bx + MM_ST_W - 1), p.W - 1);
    r.iy0 = max((int)floorf(fy_lo), by);
    r.iy1 = min(min((int)ceilf(fy_hi), by + MM_ST_H - 1), p.H - 1);
    return r.ix0 <= r.ix1 && r.iy0 <= r.iy1;
}

// ------------------------------------------------------------------------------------------------------------
// The pair engine: one batch of <= 32 faces against the 32 pixels of the sub-tile, in four warp-synchronous phases.
//   ph1 (lanes = faces)  mark(f): the lane parks its face record in ws->rec and ORs bit `lane` into ws->hit[pixel]
//                        for every pixel of the sub-tile inside the face's (tight or enlarged) bbox -- exact
//                        half-open fp32 test, DIBR_SPEC A.2 / A.4.
//   ph2 (lanes = pixels) batches arrive in face-index order and slots inside a batch are in face-index order, so
//                        cnt[pixel] + rank-in-hit-word is the pair's position in the reference's ordered scan; pairs
//                        beyond `cap` are dropped (DIB-R's order-dependent knum truncation).  Accepted pairs go to a
//                        dense list; each pixel's pairs are CONTIGUOUS and in face order.
//   ph3 (lanes = pairs)  eval(slot, pixel) -> float: the expensive arithmetic (barycentrics + 2 IEEE divisions, or
//                        distance + exp) runs once per accepted pair with all 32 lanes busy, whatever the shape of
//                        the face/pixel incidence.
//   ph4 (lanes = pixels) scan(slot, value): every pixel folds ITS pairs in face order -- a strictly-greater depth
//                        test or a running product -- i.e. exactly the reference's sequential loop over faces,
//                        without atomics, so `face_idx` ties and the silhouette product keep the reference's order.
template <typename MarkFn, typename EvalFn, typename ScanFn>
__device__ __forceinline__ int pair_batch(const mm_raster_params& p, WarpScratch* ws, int lane, int f, int cap,
                                          MarkFn mark, EvalFn eval, ScanFn scan)
{
    ws->fid[lane] = (uint32_t)f;
    mark(f);
    __syncwarp();
    uint32_t keep = ws->hit[lane];
    const int base = (int)ws->cnt[lane];
    const int nh = __popc(keep);
    ws->cnt[lane] = (uint32_t)(base + nh);
    ws->hit[lane] = 0u;
    int allowed = cap - base;
    allowed = allowed < 0 ? 0 : allowed;
    if (nh > allowed) {                          // keep only the `allowed` lowest set bits (rare: > knum candidates)
        uint32_t k2 = 0u, h = keep;
        #pragma unroll 1
        for (int a = 0; a < allowed; ++a) { const uint32_t low = h & (0u - h); k2 |= low; h ^= low; }
        keep = k2;
    }
    const int nk = __popc(keep);
    int incl = nk;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
    const int total = __shfl_sync(FULL, incl, 31);
    const int pos0 = incl - nk;
    {
        int pos = pos0;
        uint32_t k2 = keep;
        #pragma unroll 1
        while (k2) { const int j = __ffs(k2) - 1; k2 &= k2 - 1; ws->pr[pos++] = (uint32_t)((j << 5) | lane); }
    }
    __syncwarp();
    #pragma unroll 1
    for (int i = lane; i < total; i += 32) {
        const uint32_t e = ws->pr[i];
        ws->pr[i] = __float_as_uint(eval((int)(e >> 5), (int)(e & 31u)));
    }
    __syncwarp();
    {
        int pos = pos0;
        uint32_t k2 = keep;
        #pragma unroll 1
        while (k2) { const int j = __ffs(k2) - 1; k2 &= k2 - 1; scan(j, __uint_as_float(ws->pr[pos++])); }
    }
    __syncwarp();
    (void)p;
    return total;
}

// ph1 helper: park the record, then mark the pixels of the sub-tile whose centre lies in [xmin,xmax) x [ymin,ymax)
__device__ __forceinline__ void mark_box(const mm_raster_params& p, const CtaCtx& c, int lane, const FaceRec& r,
                                         float xmin, float xmax, float ymin, float ymax, uint32_t need)
{
    WarpScratch* ws = c.ws;
    ws->rec[0][lane] = r.ax; ws->rec[1][lane] = r.ay; ws->rec[2][lane] = r.bx; ws->rec[3][lane] = r.by;
    ws->rec[4][lane] = r.cx; ws->rec[5][lane] = r.cy; ws->rec[6][lane] = r.az; ws->rec[7][lane] = r.bz; ws->rec[8][lane] = r.cz;
    PixRange pr;
    if (!pix_range(p, c, xmin, xmax, ymin, ymax, pr)) return;
    const int bx = c.stx * MM_ST_W, by = c.sty * MM_ST_H;
    #pragma unroll 1
    for (int iy = pr.iy0; iy <= pr.iy1; ++iy) {
        const float py = pix_y(iy, p.H, p.sy);
        if (py < ymin || py >= ymax) continue;
        #pragma unroll 1
        for (int ix = pr.ix0; ix <= pr.ix1; ++ix) {
            const int pl = (iy - by) * MM_ST_W + (ix - bx);
            const float px = pix_x(ix, p.W, p.sx);
            if (((need >> pl) & 1u) && !(px < xmin || px >= xmax)) atomicOr(&ws->hit[pl], 1u << lane);
        }
    }
}

__device__ __forceinline__ FaceRec slot_rec(const WarpScratch* ws, int j) {
    FaceRec r;
    r.ax = ws->rec[0][j]; r.ay = ws->rec[1][j]; r.bx = ws->rec[2][j]; r.by = ws->rec[3][j];
    r.cx = ws->rec[4][j]; r.cy = ws->rec[5][j]; r.az = ws->rec[6][j]; r.bz = ws->rec[7][j]; r.cz = ws->rec[8][j];
    r.nx = r.ny = r.nz = 0.0f;
    return r;
}

// Hard pass (DIBR_SPEC A.2): front faces, tight bbox, barycentric inside test, strictly-greater depth, first face wins
// ties.  The winner's weights are recomputed at the end by the pixel's own lane (same instruction sequence).
__device__ __forceinline__ void hard_pass(const mm_raster_params& p, const CtaCtx& c, int lane, float x0, float y0,
                                          int& best_f, float& bw0, float& bw1, float& bw2)
{
    WarpScratch* ws = c.ws;
    ws->hit[lane] = 0u; ws->cnt[lane] = 0u;
    __syncwarp();
    const int bx = c.stx * MM_ST_W, by = c.sty * MM_ST_H;
    float best_z = -INFINITY;
    int bf = -1;
    for_each_batch(c.mH, p.nwords, lane, ws->fq, [&](int f) {
        pair_batch(p, ws, lane, f, 0x7fffffff,
            [&](int ff) {
                if (ff < 0) return;
                const FaceRec r = load_rec(c.rec, ff);
                if (!(r.nz >= 0.0f)) return;
                mark_box(p, c, lane, r, fminf(fminf(r.ax, r.bx), r.cx), fmaxf(fmaxf(r.ax, r.bx), r.cx),
                         fminf(fminf(r.ay, r.by), r.cy), fmaxf(fmaxf(r.ay, r.by), r.cy), FULL);
            },
            [&](int j, int pl) -> float {
                const FaceRec r = slot_rec(ws, j);
                Bary b;
                bary_eval(r, pix_x(bx + (pl & 7), p.W, p.sx), pix_y(by + (pl >> 3), p.H, p.sy), p.eps, b);
                if (b.w0 < 0.0f || b.w1 < 0.0f || b.w2 < 0.0f) return -INFINITY;
                return ADD(ADD(MUL(b.w0, r.az), MUL(b.w1, r.bz)), MUL(b.w2, r.cz));
            },
            [&](int j, float zz) {
                if (!(zz <= best_z)) { best_z = zz; bf = (int)ws->fid[j]; }
            });
    });
    best_f = bf; bw0 = bw1 = bw2 = 0.0f;
    if (bf >= 0) {
        const FaceRec r = load_rec(c.rec, bf);
        Bary b;
        bary_eval(r, x0, y0, p.eps, b);
        bw0 = b.w0; bw1 = b.w1; bw2 = b.w2;
    }
}

// Soft pass skeleton shared by forward and backward (DIBR_SPEC A.4/A.5): all faces (no back-face test), bbox
// enlarged by blen, first knum candidates per pixel in face order.  `need` = pixels that take part.
template <typename EvalFn, typename ScanFn, typename EndFn>
__device__ __forceinline__ void soft_pass(const mm_raster_params& p, const CtaCtx& c, int lane, uint32_t need,
                                          EvalFn eval, ScanFn scan, EndFn end)
{
    WarpScratch* ws = c.ws;
    ws->hit[lane] = 0u; ws->cnt[lane] = 0u;
    __syncwarp();
    for_each_batch(c.mS, p.nwords, lane, ws->fq, [&](int f) {
        const long long t0 = p.prof ? clock64() : 0;
        const int total = pair_batch(p, ws, lane, f, p.knum,
            [&](int ff) {
                if (ff < 0) return;
                const FaceRec r = load_rec(c.rec, ff);
                mark_box(p, c, lane, r, SUB(fminf(fminf(r.ax, r.bx), r.cx), p.blen), ADD(fmaxf(fmaxf(r.ax, r.bx), r.cx), p.blen),
                         SUB(fminf(fminf(r.ay, r.by), r.cy), p.blen), ADD(fmaxf(fmaxf(r.ay, r.by), r.cy), p.blen), need);
            },
            eval, scan);
        if (p.prof && lane == 0) { ws->dbg[1] += clock64() - t0; ws->dbg[3] += total; }
        end(f);
        __syncwarp();
    });
}

// ---------------------------------------------------------------------------------------------- forward
// Writes, for every pixel of every non-empty sub-tile: face_idx (workspace) and the soft silhouette (alpha plane of rgba).
__global__ void __launch_bounds__(MM_RTHREADS, MM_RMINB)
k_geom_fwd(const mm_raster_params p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    CtaCtx c;
    const int lane = threadIdx.x & 31;
    warp_init(p, smem, c, 0);
    const size_t HW = (size_t)p.H * p.W;
    const float kz = p.sigmainv / p.multiplier / p.multiplier;

    while (c.st >= 0) {
        const long long t_start = p.prof ? clock64() : 0;
        tile_begin(p, c, lane);
        const int b = c.b;
        const float x0 = pix_x(c.ix, p.W, p.sx), y0 = pix_y(c.iy, p.H, p.sy);
        int best_f = -1;
        float w0, w1, w2, soft = 0.0f;
        if (p.prof && lane == 0) { c.ws->dbg[0] = 0; c.ws->dbg[1] = 0; c.ws->dbg[2] = 0; c.ws->dbg[3] = 0; }
        const long long th = p.prof ? clock64() : 0;
        hard_pass(p, c, lane, x0, y0, best_f, w0, w1, w2);
        if (p.prof && lane == 0) c.ws->dbg[0] = clock64() - th;
        const uint32_t need = __ballot_sync(FULL, c.active && (best_f < 0));
        if (need) {
            const int bx = c.stx * MM_ST_W, by = c.sty * MM_ST_H;
            float allprob = 1.0f;
            soft_pass(p, c, lane, need,
                      [&](int j, int pl) -> float {
                          int type;
                          const FaceRec r = slot_rec(c.ws, j);
                          const float d2 = soft_d2_fast(r, pix_x(bx + (pl & 7), p.W, p.sx), pix_y(by + (pl >> 3), p.H, p.sy),
                                                        p.multiplier, type);
                          return soft_prob_fast(d2, kz);
                      },
                      [&](int, float prob) { allprob = allprob * (1.0f - prob); },      // the reference's ordered product
                      [&](int) {});
            soft = 1.0f - allprob;
        }
        if (best_f >= 0) soft = 1.0f;
        if (c.active) {
            const size_t pix = (size_t)c.iy * p.W + c.ix;
            p.face_idx_ws[(size_t)b * HW + pix] = best_f;
            p.rgba[(size_t)b * 4 * HW + 3 * HW + pix] = soft;
        }
        if (p.prof && lane == 0) {
            long long* pr = p.prof + ((size_t)b * p.nst + c.st) * 8;
            int ns = 0, nh = 0;
            #pragma unroll 1
            for (int i = 0; i < p.nwords; ++i) { ns += __popc(c.mS[i]); nh += __popc(c.mH[i]); }
            pr[0] = clock64() - t_start; pr[2] = ns; pr[3] = nh;
            pr[4] = c.ws->dbg[0]; pr[5] = c.ws->dbg[1]; pr[6] = 0; pr[7] = c.ws->dbg[3];
        }
        tile_end(c);
    }
}

// ---------------------------------------------------------------------------------------------- backward
// Soft-silhouette backward (DIBR_SPEC A.5): consumes d(loss)/d(silhouette) per pixel (`gsoft`, written by the shading
// backward) for the uncovered pixels of non-empty sub-tiles and scatters into the per-face accumulators.
__global__ void __launch_bounds__(MM_RTHREADS, MM_RMINB)
k_geom_bwd(const mm_raster_params p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    CtaCtx c;
    const int lane = threadIdx.x & 31;
    warp_init(p, smem, c, 1);
    const size_t HW = (size_t)p.H * p.W;
    const int H = p.H, W = p.W;

    while (c.st >= 0) {
        const long long t_start = p.prof ? clock64() : 0;
        tile_begin(p, c, lane);
        const int b = c.b;
        float* gacc = p.gfacc + (size_t)b * p.F * 9;
        const bool active = c.active;
        const size_t pix = active ? (size_t)c.iy * W + c.ix : 0;
        const int best_f = active ? p.face_idx_ws[(size_t)b * HW + pix] : -2;      // -2: inactive lane
        float soft = 0.0f, g_soft = 0.0f;
        if (active && best_f == -1) {
            soft = p.rgba[(size_t)b * 4 * HW + 3 * HW + pix];
            g_soft = p.gsoft[(size_t)b * HW + pix];
        }
        const uint32_t need = __ballot_sync(FULL, active && (best_f == -1) && (g_soft != 0.0f) && (soft > 0.0f));
        if (need) {
            c.ws->gs[lane] = g_soft;
            c.ws->oma[lane] = 1.0f - soft;
            #pragma unroll
            for (int k = 0; k < 6; ++k) c.ws->facc[k][lane] = 0.0f;
            const float kz = p.sigmainv / p.multiplier / p.multiplier;
            const float inv_mult = 1.0f / p.multiplier;
            const int bx = c.stx * MM_ST_W, by = c.sty * MM_ST_H;
            soft_pass(p, c, lane, need,
                      [&](int j, int pl) -> float {
                          int type;
                          const FaceRec r = slot_rec(c.ws, j);
                          const float px = pix_x(bx + (pl & 7), W, p.sx), py = pix_y(by + (pl >> 3), H, p.sy);
                          const float d2s = soft_d2_fast(r, px, py, p.multiplier, type);
                          const float prob = soft_prob_fast(d2s, kz);
                          // dLdz = -sigmainv * dLdp * (1-allprob) / (1-prob+1e-6) * prob   (DIBR_SPEC A.5)
                          const float dLdz = __fdividef(-p.sigmainv * c.ws->gs[pl] * c.ws->oma[pl], (1.0f - prob) + 1e-6f) * prob * inv_mult;
                          if (type >= 3) {
                              const int i = type - 3;
                              const float x1 = (i == 0) ? r.ax : ((i == 1) ? r.bx : r.cx);
                              const float y1 = (i == 0) ? r.ay : ((i == 1) ? r.by : r.cy);
                              atomicAdd(&c.ws->facc[2 * i][j], dLdz * 2.0f * (x1 - px));
                              atomicAdd(&c.ws->facc[2 * i + 1][j], dLdz * 2.0f * (y1 - py));
                          } else {
                              const int i = type, i2 = (type == 2) ? 0 : type + 1;
                              const float x1 = (i == 0) ? r.ax : ((i == 1) ? r.bx : r.cx);
                              const float y1 = (i == 0) ? r.ay : ((i == 1) ? r.by : r.cy);
                              const float x2 = (i2 == 0) ? r.ax : ((i2 == 1) ? r.bx : r.cx);
                              const float y2 = (i2 == 0) ? r.ay : ((i2 == 1) ? r.by : r.cy);
                              const float A = SUB(y2, y1), Bc = SUB(x1, x2), C = SUB(MUL(x2, y1), MUL(x1, y2));
                              const float up = ADD(ADD(MUL(A, px), MUL(Bc, py)), C);
                              const float rdn = __fdividef(1.0f, ADD(ADD(MUL(A, A), MUL(Bc, Bc)), 1e-10f));
                              const float d2 = up * up * rdn;
                              const float dzdA = 2.0f * (px * up - d2 * A) * rdn;
                              const float dzdB = 2.0f * (py * up - d2 * Bc) * rdn;
                              const float dzdC = 2.0f * up * rdn;
                              atomicAdd(&c.ws->facc[2 * i][j], dLdz * (dzdB - y2 * dzdC));
                              atomicAdd(&c.ws->facc[2 * i + 1][j], dLdz * (x2 * dzdC - dzdA));
                              atomicAdd(&c.ws->facc[2 * i2][j], dLdz * (y1 * dzdC - dzdB));
                              atomicAdd(&c.ws->facc[2 * i2 + 1][j], dLdz * (dzdA - x1 * dzdC));
                          }
                          return 0.0f;
                      },
                      [&](int, float) {},
                      [&](int f) {
                          if (f >= 0) {
                              float* g = gacc + (size_t)f * 9;
                              #pragma unroll
                              for (int k = 0; k < 6; ++k) {
                                  const float v = c.ws->facc[k][lane];
                                  if (v != 0.0f) { atomicAdd(g + k, v); c.ws->facc[k][lane] = 0.0f; }
                              }
                          }
                      });
        }

        if (p.prof && lane == 0) p.prof[((size_t)b * p.nst + c.st) * 8 + 1] = clock64() - t_start;
        tile_end(c);
    }
    // leave the workspace reusable for another backward on the same forward: the last warp to finish resets the ticket
    if (lane == 0) {
        __threadfence();
        const uint32_t done = atomicAdd(p.gctr + 3, 1u);
        if (done == gridDim.x - 1) { p.gctr[2] = 0u; p.gctr[3] = 0u; }
    }
}

}  // namespace

size_t mm_raster_smem_bytes(const mm_ctx* c) { return raster_smem(c->nwords, c->knum); }

cudaError_t mm_raster_configure(const mm_ctx* c) {
    const int bytes = (int)raster_smem(c->nwords, c->knum);
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_geom_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
    if ((e = cudaFuncSetAttribute(k_geom_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
    return cudaSuccess;
}

void mm_launch_geom_fwd(const mm_ctx* c, const mm_raster_params& p, cudaStream_t s)
{
    const dim3 grid(mm_raster_parts(c, p.B));
    k_geom_fwd<<<grid, MM_RTHREADS, raster_smem(c->nwords, c->knum), s>>>(p);
}

void mm_launch_geom_bwd(const mm_ctx* c, const mm_raster_params& p, cudaStream_t s)
{
    const dim3 grid(mm_raster_parts(c, p.B));
    k_geom_bwd<<<grid, MM_RTHREADS, raster_smem(c->nwords, c->knum), s>>>(p);
}
