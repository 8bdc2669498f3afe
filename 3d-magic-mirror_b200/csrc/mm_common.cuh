// mm_common.cuh -- shared definitions for libmagicmirror.so (sm_100a only).
//
// Data layout in HBM (all fp32 unless noted), B = batch, V vertices, F faces:
//   ctx (library-owned, static per mesh):
//     faces      [F,3]   int32        (DiffRender.faces,    networks.py:196,253)
//     face_uvs   [F,3,2]              (DiffRender.face_uvs, networks.py:201,254)
//     refrow/rowlo/rowhi [H], refcol/collo/colhi [W] int32: nearest-down/nearest-up
//                index tables of the contour term (networks.py:381-382)
//   workspace (caller-owned, per call; mm_ws_make):
//     frect      [B,F] uint4  the face's exact pixel rectangles, tight and enlarged (mm_device.cuh: face_rects), 16-bit packed
//     frec       [B,F,12]  face records: (ax,ay,bx,by | cx,cy,az,bz | cz,nx,ny,nz), xy already multiplied by `multiplier`,
//                          z camera-space, n = unit face normal in camera space.  48 B = 3 x float4 per face.
//     zbuf       [B,H,W] u64  visibility buffer: (order-preserving depth << 32 | ~face), atomicMax-resolved; 0 = uncovered
//     lacc       [B,H,W] u64  soft-silhouette accumulator of uncovered pixels: fixed-point sum log(1-p) << 16 | count
//     cov        [B,H,ceil(W/32)] u32 coverage bitmap written by the hard pass (atomicOr): the soft pass finds the UNCOVERED
//                             pixels of a face's enlarged bbox with one word load per row instead of one zbuf load per pixel
//     ovf_count  [4] u32   {truncated pixels (more than knum soft candidates), candidate pairs recorded, -, -}
//     sched_n    [8] u32   shading schedule: number of 4-tile strips whose covered pixels need k = 0..4 rounds of the dense pass
//                          (zbuf, lacc, cov, ovf_count and sched_n are contiguous: k_vertex_fwd clears them in one range)
//     ovf_list   [B*H*W] u32  the truncated pixels, re-done exactly, in face order, by the shading kernel's overflow role
//     sched_list [5,B*nstrips,16] u32  the strips of each class, written by the soft pass from the coverage bitmap: strip id
//                             (image * nstrips + strip) and the image's 9 lights (what a shading CTA needs before it can
//                             start: one round trip for both); the shading CTAs take them longest class first (mm_fused.cu)
//     plist      [2*B*H*W] u64 the (face, pixel) candidate pairs the forward soft pass evaluated, (image*F+face) << 32 |
//                             iy << 12 | ix: the backward soft pass replays this dense list
//     gsoft      [B,H,W]      d(loss)/d(silhouette) per pixel, handed from the shading stage to the geometry backward
//     vimg       [B,V,2]   unscaled image-plane xy (debug export / parity tests)
//     gfacc      [B,F,12]  backward accumulators: d/d(fvi) (6, unscaled) | pad (2) | d/d(unit normal) (3) | pad (1)
//     img_fwd    [B,4]     per-image sums (L1, N, D, contour) as fixed-point 64-bit integers (order-independent atomics)
//     ticket     [4] u32   last-CTA ticket of the stand-alone recon_data forward (cleared together with img_fwd)
//     img_bwd    [B,12]    per-image sums (contour, 9 light gradients, -, -), same scheme
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#define MM_THREADS      256          // threads per CTA of the loss / vertex kernels
#define MM_WARPS        (MM_THREADS / 32)
#ifndef MM_VTHREADS
#define MM_VTHREADS     512          // threads per CTA of the vertex-stage kernels
#endif
#define MM_REC_FLOATS   12
#define MM_GF           12           // floats per face of the backward accumulator `gfacc`: d/d corners (6), pad (2), d/d unit
                                     // normal (3), pad (1) -- 48 B, so the three groups are 16/8/16-byte aligned for vector REDs
#define MM_MAX_KNUM     64
// shading kernel geometry: a warp owns a 16 x 8-pixel tile, a CTA a strip of 4 consecutive tiles (row-major tile numbering)
#define MM_SH_TW        16
#define MM_SH_TH        8
#define MM_SH_WARPS     4
#define MM_SCHED_WORDS  16           // words per entry of the shading schedule: strip id, the image's 9 lights, padding (64 B)
static inline int mm_shade_strips(int H, int W) {
    return (((W + MM_SH_TW - 1) / MM_SH_TW) * ((H + MM_SH_TH - 1) / MM_SH_TH) + MM_SH_WARPS - 1) / MM_SH_WARPS;
}

struct mm_ctx {
    int device;
    int V, F, H, W;
    float proj_x, proj_y;
    float sigmainv, boxlen, multiplier, eps;
    int knum;
    // derived
    float sx, sy;            // multiplier / W, multiplier / H  (fp32 division, DIBR_SPEC A.1)
    float blen;              // boxlen * multiplier
    int nparts_recon;        // CTAs per image of the stand-alone recon_data kernels
    int nchunks;             // vertex forward: CTAs per image (each emits 1/nchunks of the face records)
    size_t smem_vertex_fwd;  // dynamic smem bytes of the vertex forward kernel
    int num_sms;
    int pdl;                 // 1 = dependent kernels are launched with programmatic stream serialization (default; MM_PDL=0 disables)
    int pdl_late;            // see mm_raster_params.pdl_late (MM_PDL_LATE)
    unsigned long long* prof;     // MM_PROF builds: device buffer of per-warp time stamps (mm_debug_profile), else NULL
    unsigned plist_cap_max;  // test hook (MM_PLIST_CAP): caps the forward's pair list so that the backward's fallback path runs
    // device arrays
    int32_t* d_faces;        // [F,3]
    float*   d_face_uvs;     // [F,6]
    int32_t* d_tab;          // [3*H + 3*W] contour tables: refrow,rowlo,rowhi,refcol,collo,colhi
    // mesh-regulariser topology (mm_ctx_set_regularizer_topology; NULL until set)
    int reg_E; float reg_ratio;
    int32_t *d_edges, *d_edge2faces, *d_flip, *d_lap_off, *d_lap_col;
    float *d_sign_init, *d_lap_val;
    int32_t *d_lapT_off, *d_lapT_row;   // the same Laplacian transposed (column j -> rows i), for x @ lpl (mm_template.cu)
    float* d_lapT_val;
    // measurement hook (mm_ctx_set_timing)
    int timing;
    cudaEvent_t ev[8];
};

struct mm_ws_layout {
    size_t frec, frect, zbuf, lacc, cov, ovf_count, sched_n, ovf_list, sched_list, plist, gsoft, vimg, gfacc, img_fwd, ticket, img_bwd, total;
};

static inline size_t mm_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static inline mm_ws_layout mm_ws_make(const mm_ctx* c, int B) {
    mm_ws_layout L;
    size_t off = 0;
    L.frec = off;     off = mm_align_up(off + (size_t)B * c->F * MM_REC_FLOATS * 4, 256);
    L.frect = off;    off = mm_align_up(off + (size_t)B * c->F * 16, 256);
    // zbuf .. sched_n are contiguous: one range, cleared at the start of every forward
    L.zbuf = off;     off = mm_align_up(off + (size_t)B * c->H * c->W * 8, 256);
    L.lacc = off;     off = mm_align_up(off + (size_t)B * c->H * c->W * 8, 256);
    L.cov = off;      off = mm_align_up(off + (size_t)B * c->H * ((c->W + 31) / 32) * 4, 256);
    L.ovf_count = off; off = off + 16;
    L.sched_n = off;  off = mm_align_up(off + 32, 256);
    L.ovf_list = off; off = mm_align_up(off + (size_t)B * c->H * c->W * 4, 256);
    L.sched_list = off; off = mm_align_up(off + (size_t)5 * B * mm_shade_strips(c->H, c->W) * MM_SCHED_WORDS * 4, 256);
    L.plist = off;    off = mm_align_up(off + (size_t)2 * B * c->H * c->W * 8, 256);
    L.gsoft = off;    off = mm_align_up(off + (size_t)B * c->H * c->W * 4, 256);
    L.vimg = off;     off = mm_align_up(off + (size_t)B * c->V * 2 * 4, 256);
    L.gfacc = off;    off = mm_align_up(off + (size_t)B * c->F * MM_GF * 4, 256);
    // img_fwd and ticket are contiguous: the stand-alone recon_data forward clears them with one memset
    L.img_fwd = off;  off = off + (size_t)B * 4 * 8;
    L.ticket = off;   off = mm_align_up(off + 16, 256);
    L.img_bwd = off;  off = mm_align_up(off + (size_t)B * 12 * 8, 256);
    L.total = off;
    return L;
}

// parameters shared by the raster kernels (passed by value)
struct mm_raster_params {
    int B, V, F, H, W, Ht, Wt;
    int Htp;                 // physical texture rows: Ht, or Ht/2 for a mirrored texture (mm_ctx_set_texture_mirror)
    int knum;
    float sx, sy, blen, multiplier, eps, sigmainv;
    int no_mask;
    int pdl_late;            // bit k: kernel k releases its dependents at CTA exit (0 hard, 1 soft_fwd, 3 shade, 4 soft_bwd)
    const float* frec;       // [B,F,12]
    const uint4* frect;      // [B,F] exact pixel rectangles (tight | enlarged)
    unsigned long long* zbuf;     // [B,H,W]
    unsigned long long* lacc;     // [B,H,W]
    uint32_t* cov;           // [B,H,ceil(W/32)] coverage bitmap: bit set = some front face covers the pixel (hard pass, atomicOr)
    int covw;                // words per bitmap row
    uint32_t* ovf_list;      // [B*H*W] truncated pixels (global pixel index)
    uint32_t* ovf_count;     // [4]: {truncated pixels, candidate pairs recorded, -, -}
    uint32_t* sched_n;       // [8] strips per class of the shading schedule
    uint32_t* sched_list;    // [5, B*nstrips, MM_SCHED_WORDS]
    int nstrips, novf;       // shading: strips (= CTAs) per image; how many CTAs (the first of the grid) double as the overflow role
    unsigned long long* prof;     // MM_PROF builds: per-warp time stamps (tools/probes/timeline.py), else NULL
    unsigned long long* plist;    // [plist_cap]
    uint32_t plist_cap;

    float* gsoft;            // [B,H,W]
    int gsoft_iou_pending;   // 1: `gsoft` holds upstream + contour terms only; consumers add the IoU term (per-image sums) on the fly
    const float* face_uvs;   // [F,6]
    const float* tex;        // [B,3,Htp,Wt]
    const float* lights;     // [B,9]
    const float* bg;         // [B,3,H,W] or NULL
    const float* gt;         // [B,4,H,W] or NULL
    const int32_t* tab;      // contour tables
    float* rgba;             // [B,4,H,W]
    float* imnormal;         // [B,H,W,3] or NULL
    int32_t* face_idx_out;   // [B,H,W] or NULL
    long long* img_fwd;      // [B,4]  fixed-point (mm_device.cuh)
    long long* img_bwd;      // [B,12] fixed-point
    // backward
    const float* g_rgba;     // [B,4,H,W] or NULL
    float image_weight, contour, loss_scale;
    const float* loss_scale_dev;  // optional DEVICE scalar multiplied into loss_scale (the upstream gradient of the loss, lazy fusion)
    int analytic_loss;
    float* gfacc;            // [B,F,MM_GF]
    float* g_tex;            // [B,3,Htp,Wt]
    int gtex_pair;           // 1: Wt even and g_tex 8-byte aligned (texel pairs may leave as red.v2)
    float* g_bg;             // [B,3,H,W] or NULL
    uint4* clr; size_t nclr;  // buffer the hard pass clears on the side (fused step: the texture-gradient output), 16-byte units
};

// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// A step is ~11 small dependent kernels; back to back in a stream each dependency costs ~2-3 us of launch latency + drain.
// With programmatic stream serialization the next kernel's CTAs are scheduled while the previous kernel is still running and
// park at `griddepcontrol.wait`, which returns once the previous grid has completed and its memory is visible.  EVERY kernel of
// the library executes mm_pdl_prologue() first, in every thread, before any early return: that keeps the dependency transitive
// (kernel N+1 waits for N, and N cannot finish before N-1 has) and makes the attribute safe on any launch.
#ifdef __CUDACC__
__device__ __forceinline__ void mm_pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
// The same for a kernel whose grid is SEVERAL waves deep: the dependents are released when a CTA exits (a CTA that never
// executes launch_dependents counts as having done so at exit) instead of at its first instruction, so that the next kernel's
// CTAs -- which can only park at their own wait -- do not take the SM slots this kernel's later waves need.  `late` is a launch
// parameter (experiment switch MM_PDL_LATE, bit per kernel).
__device__ __forceinline__ void mm_pdl_prologue(bool late) {
    if (!late) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
// ---- per-warp time stamps (MM_PROF builds only; tools/probes/timeline.py): prof[((kernel * 16384 + warp) * 4) + slot]
#ifdef MM_PROF
__device__ __forceinline__ void mm_prof_mark(unsigned long long* prof, int kid, int wid, int slot) {
    if (prof && (threadIdx.x & 31) == 0 && wid < 16384) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        prof[((size_t)kid * 16384 + wid) * 4 + slot] = t;
    }
}
#define MM_PROF_MARK(prof, kid, wid, slot) mm_prof_mark(prof, kid, wid, slot)
#else
#define MM_PROF_MARK(prof, kid, wid, slot) do { } while (0)
#endif
template <typename... KArgs, typename... Args>
static inline cudaError_t mm_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl,
                                    Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif
// launchers (defined in the .cu files); every one returns the launch's cudaError_t
cudaError_t mm_launch_vertex_fwd(const mm_ctx* c, int B, const float* vertices, const float* azim, const float* elev,
                          const float* dist, const float* bias, float* frec,
                          float* vimg, float* face_normals, float* gfacc_zero, long long* img_fwd, long long* img_bwd,
                          void* clr0, size_t bytes0, void* clr1, size_t bytes1, uint4* frect, cudaStream_t s);
cudaError_t mm_launch_vertex_bwd(const mm_ctx* c, int B, const float* vertices, const float* azim, const float* elev,
                          const float* dist, const float* bias, float* gfacc, const float* g_face_normals,
                          long long* img_bwd, float* g_vertices, float* g_azim, float* g_elev, float* g_dist,
                          float* g_bias, float* g_lights, float* loss, const long long* img_fwd, float image_weight,
                          float contour, cudaStream_t s);
cudaError_t mm_launch_template_fwd(const mm_ctx* c, int N, int h, int w, const float* x, const float* tmpl, float* local,
                                   float* ndiff, cudaStream_t s);
cudaError_t mm_launch_template_bwd(const mm_ctx* c, int N, int h, int w, const float* tmpl, const float* g_local,
                                   const float* g_ndiff, float* g_x, cudaStream_t s);
int mm_template_max_plane(void);
cudaError_t mm_launch_geom_fwd(const mm_ctx* c, const mm_raster_params& p, cudaStream_t s);
cudaError_t mm_launch_geom_bwd(const mm_ctx* c, const mm_raster_params& p, cudaStream_t s);
// shading: mode 0 = fused (forward + loss sums + RGB-side backward), 1 = forward only, 2 = backward only
cudaError_t mm_launch_shade(const mm_ctx* c, const mm_raster_params& p, int mode, cudaStream_t s);
cudaError_t mm_launch_gsoft(const mm_ctx* c, const mm_raster_params& p, cudaStream_t s);
cudaError_t mm_launch_recon_fwd(const mm_ctx* c, int B, const float* pred, const float* gt, float image_weight, float contour,
                         long long* img_fwd, unsigned* ticket, float* loss, float* iou_out, cudaStream_t s);
cudaError_t mm_launch_recon_bwd(const mm_ctx* c, int B, const float* pred, const float* gt, const long long* img_fwd,
                         float image_weight, float contour, float loss_scale, const float* loss_scale_dev, float* g_pred,
                         cudaStream_t s);
size_t mm_vertex_smem_fwd(const mm_ctx* c);
size_t mm_vertex_smem_bwd(int V);
cudaError_t mm_vertex_set_smem(int device, size_t fwd, size_t bwd);
cudaError_t mm_launch_meshreg_fwd(const mm_ctx* c, int B, const float* delta, const float* vertices, const float* fn, float temp,
                           float eps, int flip_l1, unsigned mask, float* partials, unsigned* ticket, float* terms, cudaStream_t s);
cudaError_t mm_launch_meshreg_bwd(const mm_ctx* c, int B, const float* delta, const float* vertices, const float* fn, float temp,
                           float eps, int flip_l1, unsigned mask, const float* g_terms, float* g_delta, float* g_vertices,
                           float* g_fn, cudaStream_t s);
cudaError_t mm_launch_export_faces(const mm_ctx* c, int B, const float* frec, const float* vimg, float* fvi, float* fvz,
                            float* fnz, cudaStream_t s);
cudaError_t mm_launch_texflow_fwd(const mm_ctx* c, int B, int C, int Hi, int Wi, int Ho, int Wo, int concat, const float* img,
                                  const float* flow, float* out, cudaStream_t s);
cudaError_t mm_launch_texflow_bwd(const mm_ctx* c, int B, int C, int Hi, int Wi, int Ho, int Wo, int concat, const float* img,
                                  const float* flow, const float* g_out, float* g_img, float* g_flow, cudaStream_t s);
