// mm_band.cu -- the WHOLE forward geometry of an image band in ONE kernel, in shared memory.
//
// Replaces, for image sizes whose band fits in shared memory, the four dependent launches k_vertex_fwd -> k_scatter_hard ->
// k_soft_fwd -> k_soft_ovf_fwd (56 of the step's 121 us at B=48, 128^2; ncu r1m): those kernels are latency chains on
// global-memory atomics and bitmap loads, each followed by a launch boundary and a single-wave tail.  Here a CTA OWNS a band
// of image rows of one image:
//
//   * rows are dealt to the NB bands of an image in interleaved groups of 4 (group g -> band g % NB), so an object in the
//     middle of the frame loads all bands alike; a band's R local rows live in shared memory as the 64-bit visibility buffer
//     and the 64-bit soft-silhouette accumulator of mm_raster.cu -- same keys, same fixed-point terms, hence the SAME
//     bit-exact face_idx and order-independent silhouette -- but the atomics are shared-memory atomics of ONE CTA: the owner
//     computes, nothing is exchanged between CTAs (64-bit atomicMax on distributed shared memory does not work on this
//     part: tools/probes/dsmem_atomics.cu);
//   * every CTA transforms the image's V vertices itself (642 x ~40 flops: cheaper than a launch boundary), finds the faces
//     whose enlarged bbox touches its rows (one conservative test per face) and rasterises only those;
//   * hard pass: the (face, row) segments of 32 faces at a time are numbered and dealt to the 32 lanes; a lane walks its
//     segment in x with the row-invariant part of the reference's barycentric expressions hoisted (t, n*t, m*t, k3, k3+eps:
//     the very same rounded products, so the inside decision and the depth are bit-identical to evaluating
//     bary_eval_inside per pixel) -- ~12 instructions per outside pixel instead of ~70 for the pair decode + full test;
//   * soft pass: (face, row, 32-column word) segments against the band's coverage bitmap (built from the visibility buffer
//     after one CTA barrier), candidates compacted through a small queue and evaluated one per lane, as in mm_soft_fwd.cuh;
//   * pixels with more than knum candidates are re-done in face order by the whole CTA (DIB-R keeps the FIRST knum);
//   * the band's rows of zbuf / lacc are written out once, coalesced; the candidate list and the truncated-pixel list the
//     backward replays go to global memory as before.
// Face records, face normals, the zeroed backward accumulators and (fused step) the cleared texture gradient are produced on
// the side, each band taking its share.
#include "mm_device.cuh"
#include "mm_soft_fwd.cuh"
#include "mm_camera.cuh"

#include <mutex>

namespace {

#define BD_THREADS 256
#define BD_WARPS (BD_THREADS / 32)
#define BD_QCAP 64

struct BandParams {
    mm_raster_params p;
    const int32_t* faces;
    const float* vertices; const float* azim; const float* elev; const float* dist; const float* bias;
    float proj_x, proj_y;
    float* frec_out; float* vimg; float* face_normals; float* gfacc_zero;
    int nb_shift;            // log2(bands per image)
    int R;                   // local rows per band (multiple of 4, <= 252)
    long long* prof;         // MM_BAND_PROF=1: [CTA][8] clock64 stamps at the phase boundaries (diagnostics), else NULL
};

struct WarpStage {
    float rec[9][32];        // ax ay bx by cx cy (scaled image plane) az bz cz
    int ix0[32], ix1[32], l0[32], face[32], nwd[32];
    int pre[33];             // exclusive prefix of the 32 faces' segment counts (+ total)
    uint32_t q[BD_QCAP];     // pending soft candidates: owner lane << 20 | local row << 12 | ix
};

__device__ __forceinline__ int band_row_y(int l, int shift, int band) { return ((((l >> 2) << shift) + band) << 2) + (l & 3); }
// first local row of `band` whose image row is >= iy
__device__ __forceinline__ int band_l_lo(int iy, int shift, int band) {
    const int mask = (1 << shift) - 1, g = iy >> 2, r = g & mask;
    if (r == band) return ((g >> shift) << 2) + (iy & 3);
    return ((g + ((band - r) & mask)) >> shift) << 2;
}
// last local row of `band` whose image row is <= iy (-1: none)
__device__ __forceinline__ int band_l_hi(int iy, int shift, int band) {
    const int mask = (1 << shift) - 1, g = iy >> 2, r = g & mask;
    if (r == band) return ((g >> shift) << 2) + (iy & 3);
    const int g2 = g - ((r - band) & mask);
    return g2 < 0 ? -1 : ((g2 >> shift) << 2) + 3;
}

// owner of segment k: the largest j with pre[j] <= k (faces without segments have pre[j] == pre[j+1] and are skipped)
__device__ __forceinline__ int seg_owner(const int* pre, int k) {
    int j = 0;
    #pragma unroll
    for (int st = 16; st > 0; st >>= 1) if (pre[j + st] <= k) j += st;
    return j;
}

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int& total) {
    int incl = v;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
    total = __shfl_sync(FULL, incl, 31);
    return incl - v;
}

__global__ void __launch_bounds__(BD_THREADS, 3)
k_raster_band(const BandParams q)
{
    mm_pdl_prologue();
    const mm_raster_params& p = q.p;
    extern __shared__ __align__(16) unsigned char smraw[];
    const int W = p.W, H = p.H, V = p.V, F = p.F, R = q.R, covw = p.covw, shift = q.nb_shift;
    const int band = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npix = R * W, nw = (F + 31) >> 5;
    const size_t HW = (size_t)H * W;
    unsigned long long* s_z = reinterpret_cast<unsigned long long*>(smraw);
    unsigned long long* s_l = s_z + npix;
    float* s_vc = reinterpret_cast<float*>(s_l + npix);          // 3V camera space
    float* s_vi = s_vc + 3 * V;                                  // 2V image plane, already x multiplier
    uint32_t* s_cov = reinterpret_cast<uint32_t*>(s_vi + 2 * V); // R * covw
    uint32_t* s_ovf = s_cov + R * covw;                          // R * covw: pixels that saw more than knum candidates
    uint32_t* s_mask = s_ovf + R * covw;                         // nw
    int* s_kept = reinterpret_cast<int*>(s_mask + nw);           // MM_MAX_KNUM
    float* sT = reinterpret_cast<float*>(s_kept + MM_MAX_KNUM);  // 12 (+4)
    int* s_cnt = reinterpret_cast<int*>(sT + 16);                // [0] relevant faces
    WarpStage* s_ws = reinterpret_cast<WarpStage*>(s_cnt + 4);
    uint16_t* s_list = reinterpret_cast<uint16_t*>(s_ws + BD_WARPS);   // F
    const float kz = p.sigmainv / p.multiplier / p.multiplier;
    const uint32_t lt = (1u << lane) - 1u;
    long long* prof = q.prof ? q.prof + ((size_t)b * gridDim.x + band) * 8 : nullptr;
#define BD_STAMP(i) do { if (prof && tid == 0) prof[i] = clock64(); } while (0)
    BD_STAMP(0);

    // ---------------------------------------------------------------- P0: clears, camera, vertex transform
    for (int i = tid; i < npix; i += BD_THREADS) { s_z[i] = 0ull; s_l[i] = 0ull; }
    for (int i = tid; i < R * covw; i += BD_THREADS) { s_cov[i] = 0u; s_ovf[i] = 0u; }
    if (tid == 0) {
        Cam c;
        camera_setup(q.azim[b], q.elev[b], q.dist[b], q.bias[b * 2], q.bias[b * 2 + 1], c);
        for (int i = 0; i < 12; ++i) sT[i] = c.T[i];
        s_cnt[0] = 0;
    }
    if (band == 0 && tid < 16) {                       // the per-image fixed-point accumulators start at zero
        if (tid < 4) p.img_fwd[b * 4 + tid] = 0;
        if (tid < 12) p.img_bwd[b * 12 + tid] = 0;
    }
    if (p.nclr) {                                      // fused step: the texture-gradient output, cleared on the side
        const size_t nthreads = (size_t)gridDim.x * gridDim.y * BD_THREADS;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (size_t i = ((size_t)b * gridDim.x + band) * BD_THREADS + tid; i < p.nclr; i += nthreads) p.clr[i] = z;
    }
    __syncthreads();
    {
        const float* vb = q.vertices + (size_t)b * V * 3;
        for (int v = tid; v < V; v += BD_THREADS) {
            float cx, cy, cz, xi, yi;
            project_vertex(sT, q.proj_x, q.proj_y, vb[v * 3], vb[v * 3 + 1], vb[v * 3 + 2], cx, cy, cz, xi, yi);
            s_vc[v * 3] = cx; s_vc[v * 3 + 1] = cy; s_vc[v * 3 + 2] = cz;
            s_vi[v * 2] = __fmul_rn(xi, p.multiplier); s_vi[v * 2 + 1] = __fmul_rn(yi, p.multiplier);
            if (band == 0 && q.vimg) { q.vimg[((size_t)b * V + v) * 2] = xi; q.vimg[((size_t)b * V + v) * 2 + 1] = yi; }
        }
    }
    __syncthreads();
    BD_STAMP(1);

    // ---------------------------------------------------------------- PA: faces relevant to this band; this band's share of the records
    for (int f0 = 0; f0 < F; f0 += BD_THREADS) {
        const int f = f0 + tid;
        bool rel = false;
        if (f < F) {
            const int i0 = __ldg(q.faces + f * 3), i1 = __ldg(q.faces + f * 3 + 1), i2 = __ldg(q.faces + f * 3 + 2);
            const float ax = s_vi[i0 * 2], ay = s_vi[i0 * 2 + 1], bx = s_vi[i1 * 2], by = s_vi[i1 * 2 + 1];
            const float cx = s_vi[i2 * 2], cy = s_vi[i2 * 2 + 1];
            if ((f & ((1 << shift) - 1)) == band) {
                float nx, ny, nz;
                face_normal(s_vc[i0 * 3], s_vc[i0 * 3 + 1], s_vc[i0 * 3 + 2], s_vc[i1 * 3], s_vc[i1 * 3 + 1], s_vc[i1 * 3 + 2],
                            s_vc[i2 * 3], s_vc[i2 * 3 + 1], s_vc[i2 * 3 + 2], nx, ny, nz);
                float4* rec = reinterpret_cast<float4*>(q.frec_out + ((size_t)b * F + f) * MM_REC_FLOATS);
                rec[0] = make_float4(ax, ay, bx, by);
                rec[1] = make_float4(cx, cy, s_vc[i0 * 3 + 2], s_vc[i1 * 3 + 2]);
                rec[2] = make_float4(s_vc[i2 * 3 + 2], nx, ny, nz);
                if (q.face_normals) {
                    float* fn = q.face_normals + ((size_t)b * F + f) * 3;
                    fn[0] = nx; fn[1] = ny; fn[2] = nz;
                }
                float4* g = reinterpret_cast<float4*>(q.gfacc_zero + ((size_t)b * F + f) * MM_GF);
                g[0] = g[1] = g[2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
            // conservative pixel range of the ENLARGED bbox (a superset of what either pass will touch)
            const float xmin = fminf(fminf(ax, bx), cx) - p.blen, xmax = fmaxf(fmaxf(ax, bx), cx) + p.blen;
            const float ymin = fminf(fminf(ay, by), cy) - p.blen, ymax = fmaxf(fmaxf(ay, by), cy) + p.blen;
            PixRange pr;
            if (pix_range(p, xmin, xmax, ymin, ymax, pr))
                rel = band_l_lo(pr.iy0, shift, band) <= min(band_l_hi(pr.iy1, shift, band), R - 1);
        }
        const uint32_t m = __ballot_sync(FULL, rel);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_cnt[0], __popc(m));
            base = __shfl_sync(FULL, base, 0);
            if (rel) s_list[base + __popc(m & lt)] = (uint16_t)f;
        }
    }
    __syncthreads();
    const int nL = s_cnt[0];
    WarpStage& ws = s_ws[warp];
    BD_STAMP(2);

    // ---------------------------------------------------------------- P1: hard visibility (front faces), row segments
    for (int base = warp * 32; base < nL; base += BD_THREADS) {
        const int idx = base + lane;
        int nseg = 0;
        if (idx < nL) {
            const int f = s_list[idx];
            const int i0 = __ldg(q.faces + f * 3), i1 = __ldg(q.faces + f * 3 + 1), i2 = __ldg(q.faces + f * 3 + 2);
            FaceRec r;
            r.ax = s_vi[i0 * 2]; r.ay = s_vi[i0 * 2 + 1]; r.bx = s_vi[i1 * 2]; r.by = s_vi[i1 * 2 + 1];
            r.cx = s_vi[i2 * 2]; r.cy = s_vi[i2 * 2 + 1];
            r.az = s_vc[i0 * 3 + 2]; r.bz = s_vc[i1 * 3 + 2]; r.cz = s_vc[i2 * 3 + 2];
            face_normal(s_vc[i0 * 3], s_vc[i0 * 3 + 1], r.az, s_vc[i1 * 3], s_vc[i1 * 3 + 1], r.bz,
                        s_vc[i2 * 3], s_vc[i2 * 3 + 1], r.cz, r.nx, r.ny, r.nz);
            int ix0 = 0, ix1 = -1, iy0 = 0, iy1 = -1, l0 = 0;
            if (r.nz >= 0.0f) {                                          // DIBR_SPEC A.2: the hard pass sees front faces only
                exact_rect(p, r, false, ix0, ix1, iy0, iy1);
                if (ix0 <= ix1 && iy0 <= iy1) {
                    l0 = band_l_lo(iy0, shift, band);
                    nseg = max(0, min(band_l_hi(iy1, shift, band), R - 1) - l0 + 1);
                }
            }
            ws.rec[0][lane] = r.ax; ws.rec[1][lane] = r.ay; ws.rec[2][lane] = r.bx; ws.rec[3][lane] = r.by;
            ws.rec[4][lane] = r.cx; ws.rec[5][lane] = r.cy; ws.rec[6][lane] = r.az; ws.rec[7][lane] = r.bz;
            ws.rec[8][lane] = r.cz;
            ws.ix0[lane] = ix0; ws.ix1[lane] = ix1; ws.l0[lane] = l0; ws.face[lane] = f;
        }
        int total;
        const int excl = warp_excl_scan(nseg, lane, total);
        ws.pre[lane] = excl;
        if (lane == 31) ws.pre[32] = total;
        __syncwarp();
        #pragma unroll 1
        for (int k0 = 0; k0 < total; k0 += 32) {
            const int k = k0 + lane;
            if (k >= total) continue;
            const int j = seg_owner(ws.pre, k);
            const int l = ws.l0[j] + (k - ws.pre[j]);
            const int y = band_row_y(l, shift, band);
            const float ax = ws.rec[0][j], ay = ws.rec[1][j], az = ws.rec[6][j], bz = ws.rec[7][j], cz = ws.rec[8][j];
            // the row-invariant part of bary_eval_inside, same operations in the same order (bit-identical decisions)
            const float m = SUB(ws.rec[2][j], ax), pp = SUB(ws.rec[3][j], ay);
            const float n = SUB(ws.rec[4][j], ax), qq = SUB(ws.rec[5][j], ay);
            const float t = SUB(pix_y(y, H, p.sy), ay);
            const float nt = MUL(n, t), mt = MUL(m, t);
            const float den = ADD(SUB(MUL(m, qq), MUL(n, pp)), p.eps);
            const float aden = fabsf(den);
            const bool guard = aden > 1e-18f && aden < 1e18f;
            const float sg = den > 0.0f ? 1.0f : -1.0f;
            const int face = ws.face[j], ix1 = ws.ix1[j];
            unsigned long long* zrow = s_z + l * W;
            #pragma unroll 1
            for (int ix = ws.ix0[j]; ix <= ix1; ++ix) {
                const float s = SUB(pix_x(ix, W, p.sx), ax);
                const float k1 = SUB(MUL(s, qq), nt), k2 = SUB(mt, MUL(s, pp));
                if (guard && (k1 * sg < -1e-18f || k2 * sg < -1e-18f)) continue;       // sign decides: see bary_eval_inside
                const float w1 = DIV(k1, den), w2 = DIV(k2, den);
                const float w0 = SUB(SUB(1.0f, w1), w2);
                if (w0 < 0.0f || w1 < 0.0f || w2 < 0.0f) continue;
                const float zz = ADD(ADD(MUL(w0, az), MUL(w1, bz)), MUL(w2, cz));
                atomicMax(zrow + ix, depth_key(zz, face));
            }
        }
        __syncwarp();
    }
    __syncthreads();
    BD_STAMP(3);

    // ---------------------------------------------------------------- coverage bitmap of the band
    for (int wi = warp; wi < R * covw; wi += BD_WARPS) {
        const int l = wi / covw, ix = (wi - l * covw) * 32 + lane;
        const uint32_t m = __ballot_sync(FULL, ix < W && s_z[l * W + ix] != 0ull);
        if (lane == 0) s_cov[wi] = m;
    }
    __syncthreads();

    // ---------------------------------------------------------------- P2: soft silhouette (all faces), (row, word) segments
    for (int base = warp * 32; base < nL; base += BD_THREADS) {
        const int idx = base + lane;
        int nseg = 0;
        if (idx < nL) {
            const int f = s_list[idx];
            const int i0 = __ldg(q.faces + f * 3), i1 = __ldg(q.faces + f * 3 + 1), i2 = __ldg(q.faces + f * 3 + 2);
            FaceRec r;
            r.ax = s_vi[i0 * 2]; r.ay = s_vi[i0 * 2 + 1]; r.bx = s_vi[i1 * 2]; r.by = s_vi[i1 * 2 + 1];
            r.cx = s_vi[i2 * 2]; r.cy = s_vi[i2 * 2 + 1];
            r.az = r.bz = r.cz = r.nx = r.ny = r.nz = 0.0f;
            int ix0, ix1, iy0, iy1, l0 = 0, nwd = 1;
            exact_rect(p, r, true, ix0, ix1, iy0, iy1);
            if (ix0 <= ix1 && iy0 <= iy1) {
                l0 = band_l_lo(iy0, shift, band);
                nwd = (ix1 >> 5) - (ix0 >> 5) + 1;
                nseg = max(0, min(band_l_hi(iy1, shift, band), R - 1) - l0 + 1) * nwd;
            }
            ws.rec[0][lane] = r.ax; ws.rec[1][lane] = r.ay; ws.rec[2][lane] = r.bx; ws.rec[3][lane] = r.by;
            ws.rec[4][lane] = r.cx; ws.rec[5][lane] = r.cy;
            ws.ix0[lane] = ix0; ws.ix1[lane] = ix1; ws.l0[lane] = l0; ws.face[lane] = f; ws.nwd[lane] = nwd;
        }
        int total;
        const int excl = warp_excl_scan(nseg, lane, total);
        ws.pre[lane] = excl;
        if (lane == 31) ws.pre[32] = total;
        __syncwarp();
        int qn = 0;
        // one queued candidate, evaluated by one lane (no warp collectives inside: the tail of the queue runs divergent)
        auto eval = [&](uint32_t e) {
            const int j = (int)(e >> 20), l = (int)((e >> 12) & 0xffu), ix = (int)(e & 0xfffu);
            FaceRec r;
            r.ax = ws.rec[0][j]; r.ay = ws.rec[1][j]; r.bx = ws.rec[2][j]; r.by = ws.rec[3][j];
            r.cx = ws.rec[4][j]; r.cy = ws.rec[5][j];
            r.az = r.bz = r.cz = r.nx = r.ny = r.nz = 0.0f;
            int type;
            const float d2 = soft_d2_fast(r, pix_x(ix, W, p.sx), pix_y(band_row_y(l, shift, band), H, p.sy), p.multiplier, type);
            const float prob = soft_prob_fast(d2, kz);
            const unsigned long long old = atomicAdd(s_l + l * W + ix, lacc_term(log1pf(-prob)));
            if (lacc_count(old) == p.knum) atomicOr(&s_ovf[l * covw + (ix >> 5)], 1u << (ix & 31));   // candidate knum+1
        };
        // append the n candidates just evaluated to the global pair list the backward replays (all lanes call it)
        auto record = [&](uint32_t e, int n) {
            uint32_t gb = 0u;
            if (lane == 0) gb = atomicAdd(p.ovf_count + 1, (uint32_t)n);
            gb = __shfl_sync(FULL, gb, 0);
            if (lane < n && gb + (uint32_t)lane < p.plist_cap) {
                const int j = (int)(e >> 20), l = (int)((e >> 12) & 0xffu);
                const unsigned long long fg = (unsigned long long)((size_t)b * F + ws.face[j]);
                p.plist[gb + lane] = (fg << 32) | ((unsigned long long)band_row_y(l, shift, band) << 12) | (unsigned long long)(e & 0xfffu);
            }
        };
        #pragma unroll 1
        for (int k0 = 0; k0 < total; k0 += 32) {
            const int k = k0 + lane;
            uint32_t bits = 0u, hdr = 0u;
            if (k < total) {
                const int j = seg_owner(ws.pre, k);
                const int loc = k - ws.pre[j], nwd = ws.nwd[j];
                const int li = (nwd == 1) ? loc : loc / nwd;
                const int l = ws.l0[j] + li, wd = (ws.ix0[j] >> 5) + (loc - li * nwd);
                const int lo = max(ws.ix0[j] - (wd << 5), 0), hi = min(ws.ix1[j] - (wd << 5), 31);
                bits = ~s_cov[l * covw + wd] & (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo);      // uncovered pixels of the segment
                hdr = ((uint32_t)j << 20) | ((uint32_t)l << 12) | (uint32_t)(wd << 5);
            }
            // rounds: every lane contributes its lowest pending bit, so at most 32 entries join the queue per round
            while (__any_sync(FULL, bits != 0u)) {
                const bool has = bits != 0u;
                const uint32_t m = __ballot_sync(FULL, has);
                if (has) {
                    const int jb = __ffs(bits) - 1;
                    bits &= bits - 1u;
                    ws.q[qn + __popc(m & lt)] = hdr + (uint32_t)jb;
                }
                qn += __popc(m);
                __syncwarp();
                if (qn >= 32) {
                    qn -= 32;
                    const uint32_t e = ws.q[qn + lane];
                    eval(e);
                    record(e, 32);
                }
                __syncwarp();
            }
        }
        if (qn > 0) {                                    // the stage is re-used by the next 32 faces: drain the queue
            const uint32_t e = lane < qn ? ws.q[lane] : 0u;
            if (lane < qn) eval(e);
            record(e, qn);
        }
        __syncwarp();
    }
    __syncthreads();
    BD_STAMP(4);

    // ---------------------------------------------------------------- P3: truncated pixels, exact ordered re-scan (rare)
    for (int wi = 0; wi < R * covw; ++wi) {
        uint32_t om = s_ovf[wi];                          // uniform over the CTA
        while (om) {
            const int bit = __ffs(om) - 1;
            om &= om - 1u;
            const int l = wi / covw, ix = (wi - l * covw) * 32 + bit;
            const int y = band_row_y(l, shift, band);
            const float px = pix_x(ix, W, p.sx), py = pix_y(y, H, p.sy);
            for (int f0 = 0; f0 < F; f0 += BD_THREADS) {      // enlarged-bbox hit words in face order
                const int f = f0 + tid;
                bool hit = false;
                if (f < F) {
                    const int i0 = __ldg(q.faces + f * 3), i1 = __ldg(q.faces + f * 3 + 1), i2 = __ldg(q.faces + f * 3 + 2);
                    FaceRec r;
                    r.ax = s_vi[i0 * 2]; r.ay = s_vi[i0 * 2 + 1]; r.bx = s_vi[i1 * 2]; r.by = s_vi[i1 * 2 + 1];
                    r.cx = s_vi[i2 * 2]; r.cy = s_vi[i2 * 2 + 1];
                    hit = soft_bbox_test(r, px, py, p.blen);
                }
                const uint32_t m = __ballot_sync(FULL, hit);
                const int word = (f0 >> 5) + warp;
                if (lane == 0 && word < nw) s_mask[word] = m;
            }
            __syncthreads();
            if (warp == 0) {
                int seen = 0;
                for (int w0 = 0; w0 < nw && seen < p.knum; w0 += 32) {       // first knum set bits -> ordered list of kept faces
                    const int wd = w0 + lane;
                    uint32_t m = (wd < nw) ? s_mask[wd] : 0u;
                    const int c = __popc(m);
                    int tot;
                    int pos = seen + warp_excl_scan(c, lane, tot);
                    while (m && pos < p.knum) {
                        const int bb = __ffs(m) - 1;
                        m &= m - 1u;
                        s_kept[pos++] = (wd << 5) + bb;
                    }
                    seen += tot;
                }
                const int nk = min(seen, p.knum);
                __syncwarp();
                float allprob = 1.0f;
                for (int k0 = 0; k0 < nk; k0 += 32) {                        // the kept candidates, one per lane
                    const int k = k0 + lane;
                    float prob = 0.0f;
                    if (k < nk) {
                        const int f = s_kept[k];
                        const int i0 = __ldg(q.faces + f * 3), i1 = __ldg(q.faces + f * 3 + 1), i2 = __ldg(q.faces + f * 3 + 2);
                        FaceRec r;
                        r.ax = s_vi[i0 * 2]; r.ay = s_vi[i0 * 2 + 1]; r.bx = s_vi[i1 * 2]; r.by = s_vi[i1 * 2 + 1];
                        r.cx = s_vi[i2 * 2]; r.cy = s_vi[i2 * 2 + 1];
                        r.az = r.bz = r.cz = r.nx = r.ny = r.nz = 0.0f;
                        int type;
                        prob = soft_prob_fast(soft_d2_fast(r, px, py, p.multiplier, type), kz);
                    }
                    const int cnt = min(32, nk - k0);
                    #pragma unroll 1
                    for (int qq = 0; qq < cnt; ++qq)                         // the reference's ordered product
                        allprob = allprob * (1.0f - __shfl_sync(FULL, prob, qq));
                }
                if (lane == 0) {
                    s_l[l * W + ix] = lacc_exact(allprob > 0.0f ? logf(allprob) : -2400.0f);
                    const uint32_t e = atomicAdd(p.ovf_count, 1u);           // the backward redoes these pixels the same way
                    p.ovf_list[e] = (uint32_t)((size_t)b * HW + (size_t)y * W + ix);
                }
            }
            __syncthreads();
        }
    }

    BD_STAMP(5);
    // ---------------------------------------------------------------- P4: the band's rows of zbuf / lacc -> global
    {
        unsigned long long* zb = p.zbuf + (size_t)b * HW;
        unsigned long long* la = p.lacc + (size_t)b * HW;
        if ((W & 1) == 0) {
            const int W2 = W >> 1;
            for (int i = tid; i < R * W2; i += BD_THREADS) {
                const int l = i / W2, x2 = i - l * W2;
                const int y = band_row_y(l, shift, band);
                if (y < H) {
                    const size_t g = (size_t)y * W + 2 * x2;
                    *reinterpret_cast<ulonglong2*>(zb + g) = *reinterpret_cast<const ulonglong2*>(s_z + l * W + 2 * x2);
                    *reinterpret_cast<ulonglong2*>(la + g) = *reinterpret_cast<const ulonglong2*>(s_l + l * W + 2 * x2);
                }
            }
        } else {
            for (int i = tid; i < npix; i += BD_THREADS) {
                const int l = i / W, ix = i - l * W;
                const int y = band_row_y(l, shift, band);
                if (y < H) { zb[(size_t)y * W + ix] = s_z[i]; la[(size_t)y * W + ix] = s_l[i]; }
            }
        }
    }
    if (prof && tid == 0) { prof[6] = clock64(); prof[7] = nL; }
}

size_t band_smem(const mm_ctx* c, int R) {
    const size_t covw = (size_t)(c->W + 31) / 32, nw = (size_t)(c->F + 31) / 32;
    return (size_t)R * c->W * 16 + (size_t)c->V * 5 * 4 + 2 * (size_t)R * covw * 4 + nw * 4 + MM_MAX_KNUM * 4 + 16 * 4 + 4 * 4 +
           sizeof(WarpStage) * BD_WARPS + (((size_t)c->F * 2 + 15) & ~(size_t)15);
}

}  // namespace

// Band geometry of a ctx: bands per image (power of two) and local rows per band such that a band's two 64-bit planes take
// ~32 KB; returns false when the configuration does not fit (the four-kernel path is used instead).
bool mm_band_config(const mm_ctx* c, size_t smem_optin, int* nb_shift, int* R, size_t* smem)
{
    const int ngroups = (c->H + 3) / 4;
    int gmax = (2048 / (c->W > 0 ? c->W : 1)) / 4;             // groups of 4 rows per band
    if (gmax < 1) gmax = 1;
    int shift = 0;
    while ((ngroups + (1 << shift) - 1) / (1 << shift) > gmax && shift < 10) ++shift;
    const int G = (ngroups + (1 << shift) - 1) / (1 << shift);
    const int rows = 4 * G;
    if (rows > 252 || c->W > 4095 || c->F > 65535) return false;
    const size_t need = band_smem(c, rows);
    if (need > smem_optin || need > 110 * 1024) return false;      // keep at least two CTAs per SM
    *nb_shift = shift; *R = rows; *smem = need;
    return true;
}

cudaError_t mm_band_set_smem(int device, size_t bytes)
{
    static std::mutex mu;
    static size_t cur[64] = {0};
    std::lock_guard<std::mutex> lock(mu);
    const int d = (device >= 0 && device < 64) ? device : 0;
    if (bytes > cur[d]) {
        cudaError_t e = cudaFuncSetAttribute(k_raster_band, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        cur[d] = bytes;
    }
    return cudaSuccess;
}

cudaError_t mm_launch_band_fwd(const mm_ctx* c, const mm_raster_params& p, const float* vertices, const float* azim,
                               const float* elev, const float* dist, const float* bias, float* frec, float* vimg,
                               float* face_normals, float* gfacc_zero, cudaStream_t s)
{
    BandParams q;
    q.p = p;
    q.faces = c->d_faces;
    q.vertices = vertices; q.azim = azim; q.elev = elev; q.dist = dist; q.bias = bias;
    q.proj_x = c->proj_x; q.proj_y = c->proj_y;
    q.frec_out = frec; q.vimg = vimg; q.face_normals = face_normals; q.gfacc_zero = gfacc_zero;
    q.nb_shift = c->band_shift; q.R = c->band_rows;
    q.prof = (c->band_prof && (size_t)p.B * (1u << c->band_shift) <= 8192) ? c->band_prof : nullptr;
    return mm_launch(k_raster_band, dim3(1 << c->band_shift, p.B), dim3(BD_THREADS), c->band_smem, s, false, q);
}
