// mm_raster.cu -- per-pixel stage: DIB-R hard visibility + soft silhouette + UV
// texture sampling + SH lighting + composite (+ loss partial sums), forward and
// backward.  Replaces kaolin dibr_rasterization / texture_mapping /
// spherical_harmonic_lighting and the ~40 elementwise torch kernels of
// networks.py:297-317, and their autograd.
//
// Work decomposition (B200: 148 SMs):
//   grid = (nparts, B); one CTA = MM_RWARPS warps (default 1), one warp per sub-tile (8x4 pixels, one lane per
//   pixel).  Single-warp CTAs let the hardware scheduler balance sub-tiles individually: silhouette tiles cost
//   100x an empty tile, and in a multi-warp CTA the finished warps would park at the final barrier.
//     1. The CTA stages the bitmask rows ("tile face lists") of its sub-tiles -- one
//        contiguous block per mask -- into shared memory with TMA bulk copies
//        (cp.async.bulk + mbarrier).  The masks were produced by the vertex stage.
//     2. Hard pass, FACE-parallel: the set bits of the H mask are compacted into dense
//        batches of 32 faces; each lane loads ONE face record and rasterises it over the
//        few pixels of the sub-tile its bbox touches, resolving visibility with a packed
//        (depth, ~face) atomicMax per pixel in shared memory.  Work is proportional to
//        sum |bbox ∩ sub-tile| instead of faces x 32 pixels, there is no serial
//        load->test chain across faces, and far-camera images (all 1280 faces inside a
//        handful of sub-tiles) stop being a critical path.  Bit-exact DIB-R arithmetic
//        (see mm_device.cuh): max z / smallest index == the reference's ordered scan.
//     3. Soft pass, only if some pixel is uncovered, also face-parallel over the S mask:
//        lanes mark per-pixel hit words (exact half-open enlarged-bbox test); the rank of
//        a (pixel, face) pair among the pixel's candidates (batch order == lane order ==
//        face-index order) enforces DIB-R's "first knum faces" cap; accepted pairs
//        evaluate the distance/exp code once each.
//   Backward re-derives the same per-pixel state from `face_idx` (saved) and the same
//   masks instead of storing Kaolin's knum-deep side buffers
//   (B*H*W*30*(4+8+1) B = 307 MB at B=48,128^2).
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
    float* lights;          // 9
    float* red;             // unused with single-warp CTAs
    const float* rec;       // face records of this image (global, read through L1)
    int st, stx, sty, ix, iy;
    bool st_valid, active;
    // persistent-warp state
    uint64_t* bar;          // two mbarriers (double-buffered mask rows)
    uint32_t* buf;          // [2][2*nwords]
    uint32_t* queue;        // this image's sub-tile work counter
    int nxt, k;
    uint32_t phase0, phase1;
};

// dynamic smem: | 2 mbarriers 16 B | 2 x (S row + H row) | WarpScratch |
__host__ __device__ inline size_t raster_smem(int nwords, int knum) {
    (void)knum;
    return 16 + 4 * (size_t)nwords * 4 + sizeof(WarpScratch);
}

// The raster kernels run PERSISTENT single-warp CTAs: grid = (G, B), the G warps of image b pull 8x4-pixel sub-tiles
// from a per-image atomic work counter.  Silhouette tiles cost ~100x an empty tile, so static tile->warp assignment
// would leave most warps idle; the counter balances them, and the per-tile fixed cost (barrier set-up, light
// vector, loss partials, image-level reduction ticket) is paid once per warp instead of once per tile.
// While a tile is being processed the NEXT tile's two bitmask rows ("tile face lists") are already in flight:
// a TMA bulk copy (cp.async.bulk) into the other half of a double buffer, completion tracked by an mbarrier.
__device__ __forceinline__ int fetch_tile(CtaCtx& c, int lane) {
    int t = 0;
    if (lane == 0) t = (int)atomicAdd(c.queue, 1u);
    return __shfl_sync(FULL, t, 0);
}

__device__ __forceinline__ void issue_masks(const mm_raster_params& p, CtaCtx& c, int st, int k, int lane) {
    if (lane == 0) {
        const uint32_t bytes = (uint32_t)p.nwords * 4;
        const size_t off = ((size_t)blockIdx.y * p.nst + st) * p.nwords;
        uint64_t* bar = c.bar + k;
        uint32_t* dst = c.buf + (size_t)k * 2 * p.nwords;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(2 * bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(p.maskS + off), "r"(bytes), "r"(smem_u32(bar)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst + p.nwords)), "l"(p.maskH + off), "r"(bytes), "r"(smem_u32(bar)) : "memory");
    }
}

__device__ __forceinline__ void warp_init(const mm_raster_params& p, unsigned char* smem, float* s_lights, float* s_red,
                                          CtaCtx& c, int which /* 0 fwd, 1 bwd */)
{
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    c.bar = reinterpret_cast<uint64_t*>(smem);
    c.buf = reinterpret_cast<uint32_t*>(smem + 16);
    c.ws = reinterpret_cast<WarpScratch*>(c.buf + 4 * p.nwords);
    c.lights = s_lights; c.red = s_red;
    c.rec = p.frec + (size_t)b * p.F * MM_REC_FLOATS;
    c.queue = p.tickets + b * 4 + which;
    if (lane == 0) { mbar_init(c.bar, 1); mbar_init(c.bar + 1, 1); }
    if (lane < 9) s_lights[lane] = p.lights[b * 9 + lane];
    __syncwarp();
    c.k = 0; c.phase0 = 0u; c.phase1 = 0u;
    c.st = fetch_tile(c, lane);
    if (c.st < p.nst) issue_masks(p, c, c.st, 0, lane);
}

// Called at the top of every loop iteration: prefetch the next tile's masks, wait for the current ones.
__device__ __forceinline__ void tile_begin(const mm_raster_params& p, CtaCtx& c, int lane)
{
    c.nxt = fetch_tile(c, lane);
    if (c.nxt < p.nst) issue_masks(p, c, c.nxt, c.k ^ 1, lane);
    if (c.k == 0) { mbar_wait(c.bar, c.phase0); c.phase0 ^= 1u; }
    else          { mbar_wait(c.bar + 1, c.phase1); c.phase1 ^= 1u; }
    c.mS = c.buf + (size_t)c.k * 2 * p.nwords;
    c.mH = c.mS + p.nwords;
    c.st_valid = true;
    c.sty = c.st / p.nstx; c.stx = c.st - c.sty * p.nstx;
    c.ix = c.stx * MM_ST_W + (lane & 7);
    c.iy = c.sty * MM_ST_H + (lane >> 3);
    c.active = (c.ix < p.W) && (c.iy < p.H);
}

__device__ __forceinline__ void tile_end(CtaCtx& c)
{
    __syncwarp();
    c.st = c.nxt;
    c.k ^= 1;
}

// sum over the raster CTA (MM_RWARPS warps); a single-warp CTA needs no barrier at all
__device__ __forceinline__ float rblock_sum(float v, float* red) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    if (MM_RWARPS == 1) return v;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.0f;
    #pragma unroll
    for (int i = 0; i < MM_RWARPS; ++i) r += red[i];
    return r;
}

// Per-image reduction without a second kernel and without float atomics: every CTA publishes its partial
// sums, takes a ticket, and the LAST CTA of the image sums all partials in a fixed order (deterministic).
// Returns true in the CTA that did the reduction.  The ticket resets itself so the workspace can be reused.
template <int NV>
__device__ __forceinline__ bool image_reduce_last(const float (&v)[NV], float* part /* [nparts][STRIDE] of this image */,
                                                  int stride, int nparts, uint32_t* ticket, float* out, int lane)
{
    __shared__ uint32_t s_ticket;
    if (threadIdx.x == 0) {
        #pragma unroll
        for (int i = 0; i < NV; ++i) part[(size_t)blockIdx.x * stride + i] = v[i];
        __threadfence();
        s_ticket = atomicAdd(ticket, 1u);
    }
    __syncthreads();
    if (s_ticket != (uint32_t)(nparts - 1)) return false;
    __threadfence();
    if (threadIdx.x < 32) {
        float acc[NV];
        #pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = 0.0f;
        #pragma unroll 1
        for (int k = lane; k < nparts; k += 32) {
            #pragma unroll
            for (int i = 0; i < NV; ++i) acc[i] += __ldcg(part + (size_t)k * stride + i);
        }
        #pragma unroll
        for (int i = 0; i < NV; ++i) {
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(FULL, acc[i], o);
        }
        if (lane == 0) {
            #pragma unroll
            for (int i = 0; i < NV; ++i) out[i] = acc[i];
            *ticket = 0u;            // arrival counter
            *(ticket - 2) = 0u;      // this pass's work queue (tickets[b] = {fwd queue, bwd queue, fwd arrivals, bwd arrivals})
        }
    }
    return true;
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
template <bool WITH_LOSS>
__global__ void __launch_bounds__(MM_RTHREADS, MM_RMINB)
k_raster_fwd(const mm_raster_params p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ float s_lights[16];
    __shared__ float s_red[MM_RWARPS];
    CtaCtx c;
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    warp_init(p, smem, s_lights, s_red, c, 0);
    const size_t HW = (size_t)p.H * p.W;
    float acc_l1 = 0.0f, acc_n = 0.0f, acc_d = 0.0f;

    while (c.st < p.nst) {
        const long long t_start = p.prof ? clock64() : 0;
        tile_begin(p, c, lane);
        const float x0 = pix_x(c.ix, p.W, p.sx), y0 = pix_y(c.iy, p.H, p.sy);
        int best_f = -1;
        float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f, soft = 0.0f;
        if (p.prof && lane == 0) { c.ws->dbg[0] = 0; c.ws->dbg[1] = 0; c.ws->dbg[2] = 0; c.ws->dbg[3] = 0; }
        if (!mask_empty(c.mS, p.nwords, lane)) {
            const long long th = p.prof ? clock64() : 0;
            hard_pass(p, c, lane, x0, y0, best_f, w0, w1, w2);
            if (p.prof && lane == 0) c.ws->dbg[0] = clock64() - th;
            const uint32_t need = __ballot_sync(FULL, c.active && (best_f < 0));
            if (need) {
                const float kz = p.sigmainv / p.multiplier / p.multiplier;
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
    if (WITH_LOSS) {
        const float v[4] = {rblock_sum(acc_l1, c.red), rblock_sum(acc_n, c.red), rblock_sum(acc_d, c.red), 0.0f};
        image_reduce_last<4>(v, p.part_fwd + (size_t)b * gridDim.x * 4, 4, gridDim.x, p.tickets + b * 4 + 2, p.img_fwd + b * 4, lane);
    }
}

// ---------------------------------------------------------------------------------------------- backward
__device__ __forceinline__ float contour_c(float m, float mref) { return fabsf(m - mref); }

__global__ void __launch_bounds__(MM_RTHREADS, 16)
k_raster_bwd(const mm_raster_params p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ float s_lights[16];
    __shared__ float s_red[MM_RWARPS];
    CtaCtx c;
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    warp_init(p, smem, s_lights, s_red, c, 1);
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

    // loss-gradient constants; the per-image IoU sums were reduced by the forward kernel (fixed order)
    float k_img = 0.0f, k_iou = 0.0f, k_cont = 0.0f, Nb = 0.0f, De = 1.0f;
    if (p.analytic_loss) {
        k_img = p.loss_scale * p.image_weight / ((float)p.B * 3.0f * (float)HW);
        k_iou = p.loss_scale / (float)p.B;
        k_cont = p.loss_scale * p.contour / ((float)p.B * (float)HW);
        Nb = p.img_fwd[b * 4 + 1];
        De = p.img_fwd[b * 4 + 2] + 1e-10f;
    }
    const float* rg = p.rgba + (size_t)b * 4 * HW;         // forward output (silhouette re-read)
    const float* gtb = p.gt ? p.gt + (size_t)b * 4 * HW : nullptr;
    const float* gup = p.g_rgba ? p.g_rgba + (size_t)b * 4 * HW : nullptr;
    float* gacc = p.gfacc + (size_t)b * p.F * 9;
    float* gtex = p.g_tex + (size_t)b * 3 * p.Ht * p.Wt;

    while (c.st < p.nst) {
        const long long t_start = p.prof ? clock64() : 0;
        tile_begin(p, c, lane);
        const int ix = c.ix, iy = c.iy;
        const bool active = c.active;
        const float x0 = pix_x(ix, W, p.sx), y0 = pix_y(iy, H, p.sy);
        const size_t pix = active ? (size_t)iy * W + ix : 0;

        const int best_f = active ? p.face_idx_ws[(size_t)b * HW + pix] : -2;   // -2: inactive lane
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

        // ---- soft silhouette backward (DIBR_SPEC A.5): uncovered pixels only; face-parallel, so each lane owns ONE
        // face per batch and accumulates that face's 6 corner gradients in registers -> 6 atomics per (face, sub-tile)
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
        if (p.prof && lane == 0) p.prof[((size_t)b * p.nst + c.st) * 8 + 1] = clock64() - t_start;
        tile_end(c);
    }

    // ---- per-CTA partials: contour sum + 9 light gradients, reduced per image by the last CTA (fixed order)
    float v[10];
    v[0] = rblock_sum(acc_contour, c.red);
    #pragma unroll
    for (int i = 0; i < 9; ++i) v[1 + i] = rblock_sum(acc_l[i], c.red);
    image_reduce_last<10>(v, p.part_bwd + (size_t)b * gridDim.x * 12, 12, gridDim.x, p.tickets + b * 4 + 3, p.img_bwd + b * 12, lane);
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
    const dim3 grid(mm_raster_parts(c, p.B), p.B);
    const size_t smem = raster_smem(c->nwords, c->knum);
    if (with_loss) k_raster_fwd<true><<<grid, MM_RTHREADS, smem, s>>>(p);
    else           k_raster_fwd<false><<<grid, MM_RTHREADS, smem, s>>>(p);
}

void mm_launch_raster_bwd(const mm_ctx* c, const mm_raster_params& p, cudaStream_t s)
{
    const dim3 grid(mm_raster_parts(c, p.B), p.B);
    k_raster_bwd<<<grid, MM_RTHREADS, raster_smem(c->nwords, c->knum), s>>>(p);
}
