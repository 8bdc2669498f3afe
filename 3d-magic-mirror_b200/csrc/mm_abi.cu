// mm_abi.cu -- the extern "C" surface declared in include/magicmirror.h.
// Host-side orchestration only: argument validation, workspace carving, launches.
#include "../../include/magicmirror.h"
#include "mm_common.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define MM_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return fail(MM_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                               \
    } while (0)

#define MM_REQUIRE(cond, msg) \
    do { if (!(cond)) return fail(MM_E_INVALID, "invalid argument: %s", msg); } while (0)

// torch F.interpolate(mode='nearest') source index: min(floor(dst * scale), in - 1), scale = in/out in fp32
int nearest_src(int dst, int in_size, int out_size) {
    const float scale = (float)in_size / (float)out_size;
    const int s = (int)floorf((float)dst * scale);
    return s < in_size - 1 ? s : in_size - 1;
}

// refidx[y] = down(up(y)); lo/hi[r] = contiguous range of y with refidx[y] == r (empty unless r is a reference)
void contour_tables(int n, int32_t* ref, int32_t* lo, int32_t* hi) {
    const int n4 = n / 4;
    for (int y = 0; y < n; ++y) { lo[y] = 0; hi[y] = 0; }
    for (int y = 0; y < n; ++y) {
        int r = y;
        if (n4 > 0) r = nearest_src(nearest_src(y, n4, n), n, n4);
        ref[y] = r;
    }
    for (int y = 0; y < n; ++y) {
        const int r = ref[y];
        if (hi[r] == 0) { lo[r] = y; hi[r] = y + 1; }
        else            { hi[r] = y + 1; }
    }
}

#define MM_LAUNCH(call, what)                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            (void)cudaGetLastError();        /* clear the (non-sticky) launch error */          \
            return fail(MM_E_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e_));    \
        }                                                                                       \
    } while (0)

// the ctx's device must be the calling thread's current device (a kernel launched on another device's stream fails late and
// obscurely); cudaGetDevice is a thread-local read
int check_device(const mm_ctx* c) {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != c->device)
        return fail(MM_E_INVALID, "ctx is bound to device %d but the current device is %d (cudaSetDevice / torch.cuda.device first)",
                    c->device, cur);
    return MM_OK;
}

int check_ws(const mm_ctx* c, int B, const void* ws, size_t bytes) {
    if (!ws) return fail(MM_E_INVALID, "invalid argument: workspace is NULL");
    if (((uintptr_t)ws & 255) != 0) return fail(MM_E_INVALID, "invalid argument: workspace must be 256-byte aligned");
    const size_t need = mm_ws_make(c, B).total;
    if (bytes < need)
        return fail(MM_E_INVALID, "invalid argument: workspace holds %zu bytes, batch %d needs %zu (mm_workspace_bytes)", bytes, B, need);
    return MM_OK;
}

#define MM_COMMON_CHECKS(c, B, ws, bytes)                                   \
    do {                                                                    \
        MM_REQUIRE((c) && (B) > 0 && (B) <= 65535, "ctx / B (1..65535)");   \
        if (int r_ = check_device(c)) return r_;                            \
        if (int r_ = check_ws((c), (B), (ws), (bytes))) return r_;          \
    } while (0)

void fill_params(const mm_ctx* c, int B, int Ht, int Wt, int tex_mirror, int no_mask, mm_raster_params& p) {
    memset(&p, 0, sizeof(p));
    p.B = B; p.V = c->V; p.F = c->F; p.H = c->H; p.W = c->W; p.Ht = Ht; p.Wt = Wt;
    p.Htp = tex_mirror ? Ht / 2 : Ht;
    p.knum = c->knum;
    p.sx = c->sx; p.sy = c->sy; p.blen = c->blen; p.multiplier = c->multiplier; p.eps = c->eps; p.sigmainv = c->sigmainv;
    p.no_mask = no_mask;
    p.pdl_late = c->pdl_late;
    p.covw = (c->W + 31) / 32;
    p.nstrips = mm_shade_strips(c->H, c->W);
    // the first CTAs of the shading grid double as the overflow role when there are truncated pixels (far cameras); never
    // more than are resident together (lanes of later CTAs wait for their results)
    // (3 per SM: the shading kernels keep 4-5 CTAs per SM resident, the margin is for whatever else shares the GPU)
    p.novf = c->num_sms * 3;
    if (p.novf > p.nstrips * B) p.novf = p.nstrips * B;
    p.prof = c->prof;
    p.face_uvs = c->d_face_uvs;
    p.tab = c->d_tab;
}

void set_ws(const mm_ctx* c, const mm_ws_layout& L, char* ws, mm_raster_params& p) {
    p.frec = (const float*)(ws + L.frec); p.frect = (const uint4*)(ws + L.frect);
    p.zbuf = (unsigned long long*)(ws + L.zbuf); p.lacc = (unsigned long long*)(ws + L.lacc);
    p.cov = (uint32_t*)(ws + L.cov);
    p.ovf_list = (uint32_t*)(ws + L.ovf_list); p.ovf_count = (uint32_t*)(ws + L.ovf_count);
    p.sched_n = (uint32_t*)(ws + L.sched_n); p.sched_list = (uint32_t*)(ws + L.sched_list);
    p.gsoft = (float*)(ws + L.gsoft);
    p.plist = (unsigned long long*)(ws + L.plist); p.plist_cap = (uint32_t)((L.gsoft - L.plist) / 8);
    if (c->plist_cap_max && p.plist_cap > c->plist_cap_max) p.plist_cap = c->plist_cap_max;
    p.img_fwd = (long long*)(ws + L.img_fwd); p.img_bwd = (long long*)(ws + L.img_bwd);
    p.gfacc = (float*)(ws + L.gfacc);
}

// vertex stage; the kernel also clears the visibility buffer, the silhouette accumulators, the coverage bitmap, the overflow
// counters (one contiguous range) and the per-face backward accumulators for the rest of the step
cudaError_t launch_vertex_fwd(const mm_ctx* c, int B, const mm_ws_layout& L, char* ws, const float* vertices, const float* azim,
                              const float* elev, const float* dist, const float* bias, float* face_normals, cudaStream_t s) {
    // zbuf .. sched_n are contiguous and 256-byte aligned: one clear range (16-byte units)
    const size_t bytes0 = mm_align_up((L.sched_n + 32) - L.zbuf, 16);
    return mm_launch_vertex_fwd(c, B, vertices, azim, elev, dist, bias, (float*)(ws + L.frec), (float*)(ws + L.vimg), face_normals,
                                (float*)(ws + L.gfacc), (long long*)(ws + L.img_fwd), (long long*)(ws + L.img_bwd),
                                ws + L.zbuf, bytes0, nullptr, 0, (uint4*)(ws + L.frect), s);
}

// forward geometry of a batch: face records + visibility buffer + soft-silhouette accumulators + candidate lists (vertex stage ->
// hard pass -> soft pass; truncated pixels are re-done inside the shading kernel).  p.clr / p.nclr (fused step) name a buffer
// that is cleared on the side.
int launch_geometry_forward(const mm_ctx* c, int B, const mm_ws_layout& L, char* ws, const mm_raster_params& p,
                            const float* vertices, const float* azim, const float* elev, const float* dist, const float* bias,
                            float* face_normals, cudaStream_t s) {
    MM_LAUNCH(launch_vertex_fwd(c, B, L, ws, vertices, azim, elev, dist, bias, face_normals, s), "vertex_fwd");
    if (c->timing) cudaEventRecord(c->ev[1], s);
    MM_LAUNCH(mm_launch_geom_fwd(c, p, s), "geom_fwd");
    return MM_OK;
}

}  // namespace

extern "C" {

int mm_abi_version(void) { return MM_ABI_VERSION; }
const char* mm_last_error(void) { return g_err; }

int mm_ctx_create(mm_ctx** out, int device, int V, int F, const int32_t* faces_host, const float* face_uvs_host,
                  int H, int W, float proj_x, float proj_y, float sigmainv, float boxlen, int knum,
                  float multiplier, float eps)
{
    MM_REQUIRE(out != nullptr, "out");
    *out = nullptr;
    MM_REQUIRE(V > 0 && F > 0 && H > 0 && W > 0, "V, F, H, W must be positive");
    MM_REQUIRE(H <= 4095 && W <= 4095, "H, W must be <= 4095 (12-bit pixel coordinates in the scatter queue)");
    MM_REQUIRE(faces_host && face_uvs_host, "faces_host / face_uvs_host");
    MM_REQUIRE(knum > 0 && knum <= MM_MAX_KNUM, "knum out of range");
    MM_REQUIRE(multiplier > 0.0f, "multiplier");
    for (int i = 0; i < F * 3; ++i)
        if (faces_host[i] < 0 || faces_host[i] >= V) return fail(MM_E_INVALID, "faces[%d] = %d out of [0,%d)", i, faces_host[i], V);
    int ndev = 0;
    MM_CUDA(cudaGetDeviceCount(&ndev));
    MM_REQUIRE(device >= 0 && device < ndev, "device index");
    int prev = -1;
    MM_CUDA(cudaGetDevice(&prev));
    MM_CUDA(cudaSetDevice(device));
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev};
    cudaDeviceProp prop;
    MM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(MM_E_UNSUPPORTED, "libmagicmirror is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);

    mm_ctx* c = new mm_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device; c->V = V; c->F = F; c->H = H; c->W = W;
    c->proj_x = proj_x; c->proj_y = proj_y;
    c->sigmainv = sigmainv; c->boxlen = boxlen; c->multiplier = multiplier; c->eps = eps; c->knum = knum;
    c->sx = multiplier / (float)W;
    c->sy = multiplier / (float)H;
    c->blen = boxlen * multiplier;
    c->nparts_recon = (H * W + 2047) / 2048 < 1 ? 1 : (H * W + 2047) / 2048;     // ~2048 pixels per recon CTA
    c->num_sms = prop.multiProcessorCount;
    if (F > 65535) { delete c; return fail(MM_E_UNSUPPORTED, "F=%d exceeds the 16-bit face ids of the soft-pass lists", F); }
    const size_t smem_max = prop.sharedMemPerBlockOptin;
    // vertex stage: CTAs per image (each recomputes the vertex transform and emits 1/nchunks of the face records)
    c->nchunks = 8;
    c->pdl = 1;
    if (const char* e = getenv("MM_PDL")) c->pdl = atoi(e) != 0;
    // measured (profiles/r2_notes.md section 4): releasing the dependents at CTA exit instead of at the first instruction in the
    // four forward raster kernels (the parked CTAs of the next kernel no longer take slots), but EARLY in k_soft_bwd so that
    // k_vertex_bwd's input-only prologue (cold loads, camera chain, vertex transform) runs under it: 0.1093 -> 0.1045 ms
    c->pdl_late = 15;
    if (const char* e = getenv("MM_PDL_LATE")) c->pdl_late = atoi(e);
    if (const char* e = getenv("MM_PLIST_CAP")) c->plist_cap_max = (unsigned)atoi(e);
    if (const char* e = getenv("MM_VCHUNKS")) { const int v = atoi(e); if (v > 0 && v <= 32) c->nchunks = v; }
    c->smem_vertex_fwd = mm_vertex_smem_fwd(c);
    const size_t vs_f = c->smem_vertex_fwd, vs_b = mm_vertex_smem_bwd(V);
    if (vs_f > smem_max || vs_b > smem_max) {
        const size_t need = vs_f > vs_b ? vs_f : vs_b;
        delete c;
        return fail(MM_E_UNSUPPORTED, "V=%d F=%d W=%d needs %zu B of shared memory per CTA (> %zu)", V, F, W, need, smem_max);
    }
    if (cudaError_t e = mm_vertex_set_smem(device, vs_f, vs_b)) {
        delete c;
        return fail(MM_E_CUDA, "cudaFuncSetAttribute(vertex kernels) failed: %s", cudaGetErrorString(e));
    }
    std::vector<int32_t> tab(3 * (size_t)H + 3 * (size_t)W);
    contour_tables(H, tab.data(), tab.data() + H, tab.data() + 2 * H);
    contour_tables(W, tab.data() + 3 * H, tab.data() + 3 * H + W, tab.data() + 3 * H + 2 * W);
    if (cudaMalloc(&c->d_faces, (size_t)F * 3 * 4) != cudaSuccess ||
        cudaMalloc(&c->d_face_uvs, (size_t)F * 6 * 4) != cudaSuccess ||
        cudaMalloc(&c->d_tab, tab.size() * 4) != cudaSuccess) {
        const cudaError_t e = cudaGetLastError();
        mm_ctx_destroy(c);
        return fail(MM_E_CUDA, "cudaMalloc of ctx tables failed: %s", cudaGetErrorString(e));
    }
    if (cudaMemcpy(c->d_faces, faces_host, (size_t)F * 3 * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(c->d_face_uvs, face_uvs_host, (size_t)F * 6 * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(c->d_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        const cudaError_t e = cudaGetLastError();
        mm_ctx_destroy(c);
        return fail(MM_E_CUDA, "upload of ctx tables failed: %s", cudaGetErrorString(e));
    }
    *out = c;
    return MM_OK;
}

int mm_ctx_destroy(mm_ctx* c) {
    if (!c) return MM_OK;
    cudaFree(c->d_edges); cudaFree(c->d_edge2faces); cudaFree(c->d_flip); cudaFree(c->d_sign_init);
    cudaFree(c->d_lap_off); cudaFree(c->d_lap_col); cudaFree(c->d_lap_val);
    cudaFree(c->d_lapT_off); cudaFree(c->d_lapT_row); cudaFree(c->d_lapT_val);
    cudaFree(c->d_faces);
    cudaFree(c->d_face_uvs);
    cudaFree(c->d_tab);
    for (int i = 0; i < 8; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    delete c;
    return MM_OK;
}

int mm_ctx_get_int(const mm_ctx* c, const char* key) {
    if (!c || !key) return -1;
    const int geom = 3;                                     // vertex + hard + soft
    if (!strcmp(key, "fused_kernels")) return geom + 3;     // + shading, soft backward, vertex backward
    if (!strcmp(key, "api_kernels")) return geom + 1 + 1 + 3;   // + shade fwd | recon | shade bwd, soft bwd, vertex bwd
    return -1;
}

size_t mm_workspace_bytes(const mm_ctx* c, int B) {
    if (!c || B <= 0) return 0;
    return mm_ws_make(c, B).total;
}

int mm_render_forward(mm_ctx* c, int B, const float* vertices, const float* azim, const float* elev, const float* dist,
                      const float* bias, const float* tex, int Ht, int Wt, int tex_mirror, const float* lights, const float* bg,
                      int no_mask, float* rgba, float* face_normals, float* imnormal, int32_t* face_idx,
                      void* workspace, size_t workspace_bytes, void* stream)
{
    MM_COMMON_CHECKS(c, B, workspace, workspace_bytes);
    MM_REQUIRE(vertices && azim && elev && dist && bias && tex && lights, "NULL input");
    MM_REQUIRE(Ht > 0 && Wt > 0, "texture size");
    MM_REQUIRE(!tex_mirror || (Ht & 1) == 0, "a mirrored texture needs an even logical height Ht");
    MM_REQUIRE(!no_mask || bg, "no_mask=1 requires bg");
    MM_REQUIRE(rgba, "rgba");
    cudaStream_t s = (cudaStream_t)stream;
    const mm_ws_layout L = mm_ws_make(c, B);
    char* ws = (char*)workspace;
    mm_raster_params p;
    fill_params(c, B, Ht, Wt, tex_mirror, no_mask, p);
    set_ws(c, L, ws, p);
    p.tex = tex; p.lights = lights; p.bg = bg;
    p.rgba = rgba; p.imnormal = imnormal; p.face_idx_out = face_idx;
    if (int r = launch_geometry_forward(c, B, L, ws, p, vertices, azim, elev, dist, bias, face_normals, s)) return r;
    MM_LAUNCH(mm_launch_shade(c, p, 1, s), "shade_fwd");
    return MM_OK;
}

int mm_render_backward(mm_ctx* c, int B, const float* vertices, const float* azim, const float* elev, const float* dist,
                       const float* bias, const float* tex, int Ht, int Wt, int tex_mirror, const float* lights, const float* bg,
                       int no_mask, const float* rgba, const float* g_rgba, const float* g_face_normals,
                       const float* recon_gt, float image_weight, float contour, float loss_scale, const float* loss_scale_dev,
                       float* g_vertices, float* g_azim, float* g_elev, float* g_dist, float* g_bias, float* g_tex,
                       float* g_lights, float* g_bg, void* workspace, size_t workspace_bytes, void* stream)
{
    MM_COMMON_CHECKS(c, B, workspace, workspace_bytes);
    MM_REQUIRE(vertices && azim && elev && dist && bias && tex && lights, "NULL input");
    MM_REQUIRE(!no_mask || bg, "no_mask=1 requires bg");
    (void)rgba;                  // (kept in the signature; the backward reads the silhouette from the workspace since ABI v3)
    MM_REQUIRE(Ht > 0 && Wt > 0 && (!tex_mirror || (Ht & 1) == 0), "texture size (even Ht for a mirrored texture)");
    MM_REQUIRE(g_vertices && g_azim && g_elev && g_dist && g_bias && g_tex && g_lights, "NULL gradient output");
    cudaStream_t s = (cudaStream_t)stream;
    const mm_ws_layout L = mm_ws_make(c, B);
    char* ws = (char*)workspace;
    const size_t HW = (size_t)c->H * c->W;
    MM_CUDA(cudaMemsetAsync(g_tex, 0, (size_t)B * 3 * (tex_mirror ? Ht / 2 : Ht) * Wt * 4, s));
    if (g_bg && (!no_mask || !(g_rgba || recon_gt))) MM_CUDA(cudaMemsetAsync(g_bg, 0, (size_t)B * 3 * HW * 4, s));
    mm_raster_params p;
    fill_params(c, B, Ht, Wt, tex_mirror, no_mask, p);
    set_ws(c, L, ws, p);
    p.tex = tex; p.lights = lights; p.bg = bg;
    p.g_rgba = g_rgba;
    p.g_tex = g_tex; p.g_bg = no_mask ? g_bg : nullptr;
    p.gtex_pair = ((Wt & 1) == 0 && ((uintptr_t)g_tex & 7) == 0) ? 1 : 0;
    if (recon_gt) {
        p.gt = recon_gt; p.image_weight = image_weight; p.contour = contour; p.loss_scale = loss_scale;
        p.loss_scale_dev = loss_scale_dev;
        p.analytic_loss = 2;                                                       // gradient only: the loss itself is recon_data's
        p.gsoft_iou_pending = ((c->H & 3) == 0 && (c->W & 3) == 0) ? 1 : 0;
    }
    if (g_rgba || recon_gt) {
        // (no programmatic launch across the memset nodes: the first kernel starts after them)
        mm_ctx c0 = *c; c0.pdl = 0;
        MM_LAUNCH(mm_launch_shade(&c0, p, 2, s), "shade_bwd");
        if (recon_gt && !p.gsoft_iou_pending) MM_LAUNCH(mm_launch_gsoft(c, p, s), "gsoft");
        MM_LAUNCH(mm_launch_geom_bwd(c, p, s), "geom_bwd");
    }
    // (without any image gradient the per-face accumulators are still zero: only g_face_normals flows)
    MM_LAUNCH(mm_launch_vertex_bwd(c, B, vertices, azim, elev, dist, bias, p.gfacc, g_face_normals, p.img_bwd, g_vertices,
                                   g_azim, g_elev, g_dist, g_bias, g_lights, nullptr, nullptr, 0.0f, 0.0f, s), "vertex_bwd");
    return MM_OK;
}

int mm_face_normals_forward(mm_ctx* c, int B, const float* vertices, const float* azim, const float* elev, const float* dist,
                            const float* bias, float* face_normals, void* workspace, size_t workspace_bytes, void* stream)
{
    MM_COMMON_CHECKS(c, B, workspace, workspace_bytes);
    MM_REQUIRE(vertices && azim && elev && dist && bias && face_normals, "NULL argument");
    const mm_ws_layout L = mm_ws_make(c, B);
    char* ws = (char*)workspace;
    MM_LAUNCH(mm_launch_vertex_fwd(c, B, vertices, azim, elev, dist, bias, (float*)(ws + L.frec), (float*)(ws + L.vimg), face_normals,
                                   (float*)(ws + L.gfacc), (long long*)(ws + L.img_fwd), (long long*)(ws + L.img_bwd), nullptr, 0,
                                   nullptr, 0, nullptr, (cudaStream_t)stream), "vertex_fwd");
    return MM_OK;
}

int mm_face_normals_backward(mm_ctx* c, int B, const float* vertices, const float* azim, const float* elev, const float* dist,
                             const float* bias, const float* g_face_normals, float* g_vertices, float* g_azim, float* g_elev,
                             float* g_dist, float* g_bias, void* workspace, size_t workspace_bytes, void* stream)
{
    MM_COMMON_CHECKS(c, B, workspace, workspace_bytes);
    MM_REQUIRE(vertices && azim && elev && dist && bias && g_face_normals, "NULL argument");
    MM_REQUIRE(g_vertices && g_azim && g_elev && g_dist && g_bias, "NULL gradient output");
    const mm_ws_layout L = mm_ws_make(c, B);
    char* ws = (char*)workspace;
    // the per-face accumulators were zeroed by the forward on this workspace and are left zeroed by every backward
    MM_LAUNCH(mm_launch_vertex_bwd(c, B, vertices, azim, elev, dist, bias, (float*)(ws + L.gfacc), g_face_normals,
                                   (long long*)(ws + L.img_bwd), g_vertices, g_azim, g_elev, g_dist, g_bias, nullptr, nullptr,
                                   nullptr, 0.0f, 0.0f, (cudaStream_t)stream), "vertex_bwd");
    return MM_OK;
}

int mm_recon_data_forward(mm_ctx* c, int B, const float* pred, const float* gt, float image_weight, float contour,
                          float* loss, float* iou_sums, void* workspace, size_t workspace_bytes, void* stream)
{
    MM_COMMON_CHECKS(c, B, workspace, workspace_bytes);
    MM_REQUIRE(pred && gt && loss, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    const mm_ws_layout L = mm_ws_make(c, B);
    char* ws = (char*)workspace;
    MM_CUDA(cudaMemsetAsync(ws + L.img_fwd, 0, (L.ticket + 16) - L.img_fwd, s));      // per-image sums + the ticket
    MM_LAUNCH(mm_launch_recon_fwd(c, B, pred, gt, image_weight, contour, (long long*)(ws + L.img_fwd), (unsigned*)(ws + L.ticket),
                                  loss, iou_sums, s), "recon_fwd");
    return MM_OK;
}

int mm_recon_data_backward(mm_ctx* c, int B, const float* pred, const float* gt, float image_weight, float contour,
                           float loss_scale, const float* loss_scale_dev, float* g_pred, void* workspace, size_t workspace_bytes,
                           void* stream)
{
    MM_COMMON_CHECKS(c, B, workspace, workspace_bytes);
    MM_REQUIRE(pred && gt && g_pred, "NULL argument");
    const mm_ws_layout L = mm_ws_make(c, B);
    char* ws = (char*)workspace;
    MM_LAUNCH(mm_launch_recon_bwd(c, B, pred, gt, (const long long*)(ws + L.img_fwd), image_weight, contour, loss_scale,
                                  loss_scale_dev, g_pred, (cudaStream_t)stream), "recon_bwd");
    return MM_OK;
}

int mm_render_compare_fwd_bwd(mm_ctx* c, int B, const float* vertices, const float* azim, const float* elev,
                              const float* dist, const float* bias, const float* tex, int Ht, int Wt, int tex_mirror,
                              const float* lights, const float* bg, int no_mask, const float* gt, float image_weight,
                              float contour, float loss_scale, const float* g_rgba_extra, const float* g_face_normals,
                              float* rgba,
                              float* face_normals, float* loss, float* g_vertices, float* g_azim, float* g_elev,
                              float* g_dist, float* g_bias, float* g_tex, float* g_lights, float* g_bg,
                              void* workspace, size_t workspace_bytes, void* stream)
{
    MM_COMMON_CHECKS(c, B, workspace, workspace_bytes);
    MM_REQUIRE(vertices && azim && elev && dist && bias && tex && lights && gt, "NULL input");
    MM_REQUIRE(Ht > 0 && Wt > 0, "texture size");
    MM_REQUIRE(!tex_mirror || (Ht & 1) == 0, "a mirrored texture needs an even logical height Ht");
    MM_REQUIRE(!no_mask || bg, "no_mask=1 requires bg");
    MM_REQUIRE(rgba && loss, "rgba / loss");
    MM_REQUIRE(g_vertices && g_azim && g_elev && g_dist && g_bias && g_tex && g_lights, "NULL gradient output");
    cudaStream_t s = (cudaStream_t)stream;
    const mm_ws_layout L = mm_ws_make(c, B);
    char* ws = (char*)workspace;
    const size_t HW = (size_t)c->H * c->W;
    if (g_bg && !no_mask) MM_CUDA(cudaMemsetAsync(g_bg, 0, (size_t)B * 3 * HW * 4, s));
    if (c->timing) cudaEventRecord(c->ev[0], s);
    const size_t gtex_bytes = (size_t)B * 3 * (tex_mirror ? Ht / 2 : Ht) * Wt * 4;
    const bool gtex_side = (((uintptr_t)g_tex & 15) == 0) && ((gtex_bytes & 15) == 0);     // cleared by the hard pass on the side
    if (!gtex_side) MM_CUDA(cudaMemsetAsync(g_tex, 0, gtex_bytes, s));
    mm_raster_params p;
    fill_params(c, B, Ht, Wt, tex_mirror, no_mask, p);
    set_ws(c, L, ws, p);
    p.tex = tex; p.lights = lights; p.bg = bg; p.gt = gt;
    p.rgba = rgba;
    p.g_rgba = g_rgba_extra;
    p.image_weight = image_weight; p.contour = contour; p.loss_scale = loss_scale;
    p.analytic_loss = 1;
    p.g_tex = g_tex; p.g_bg = no_mask ? g_bg : nullptr;
    p.gtex_pair = ((Wt & 1) == 0 && ((uintptr_t)g_tex & 7) == 0) ? 1 : 0;
    // H, W multiples of 4: the contour term is tile-local, so d(loss)/d(silhouette) minus its IoU term is emitted by the
    // shading kernel and the geometry backward adds the IoU term from the per-image sums on the fly
    p.gsoft_iou_pending = ((c->H & 3) == 0 && (c->W & 3) == 0) ? 1 : 0;
    if (gtex_side) { p.clr = (uint4*)g_tex; p.nclr = gtex_bytes / 16; }
    if (int r = launch_geometry_forward(c, B, L, ws, p, vertices, azim, elev, dist, bias, face_normals, s)) return r;
    p.clr = nullptr; p.nclr = 0;
    if (c->timing) cudaEventRecord(c->ev[2], s);
    MM_LAUNCH(mm_launch_shade(c, p, 0, s), "shade_fused");       // shading forward + loss sums + the whole RGB-side backward
    if (c->timing) cudaEventRecord(c->ev[3], s);
    if (!p.gsoft_iou_pending) MM_LAUNCH(mm_launch_gsoft(c, p, s), "gsoft");    // general sizes: own pass (index tables)
    if (c->timing) cudaEventRecord(c->ev[4], s);
    MM_LAUNCH(mm_launch_geom_bwd(c, p, s), "geom_bwd");
    if (c->timing) cudaEventRecord(c->ev[5], s);
    // vertex backward; its last CTA also finalises the loss scalars (img_fwd / img_bwd are re-zeroed by the next vertex_fwd)
    MM_LAUNCH(mm_launch_vertex_bwd(c, B, vertices, azim, elev, dist, bias, p.gfacc, g_face_normals, p.img_bwd, g_vertices, g_azim,
                                   g_elev, g_dist, g_bias, g_lights, loss, p.img_fwd, image_weight, contour, s), "vertex_bwd");
    if (c->timing) { cudaEventRecord(c->ev[6], s); cudaEventRecord(c->ev[7], s); }
    return MM_OK;
}

int mm_ctx_set_regularizer_topology(mm_ctx* c, int E, const int32_t* edges_host, const int32_t* edge2faces_host,
                                    const int32_t* flip_index_host, const float* sign_init_host, int nnz,
                                    const int32_t* lap_row_off_host, const int32_t* lap_col_host, const float* lap_val_host,
                                    float ratio)
{
    MM_REQUIRE(c, "ctx");
    MM_REQUIRE(E > 0 && nnz > 0 && ratio > 0.0f, "E, nnz, ratio must be positive");
    MM_REQUIRE(edges_host && edge2faces_host && flip_index_host && sign_init_host && lap_row_off_host && lap_col_host && lap_val_host,
               "NULL topology array");
    for (int i = 0; i < E * 2; ++i) {
        if (edges_host[i] < 0 || edges_host[i] >= c->V) return fail(MM_E_INVALID, "edges[%d] = %d out of [0,%d)", i, edges_host[i], c->V);
        if (edge2faces_host[i] < 0 || edge2faces_host[i] >= c->F) return fail(MM_E_INVALID, "edge2faces[%d] = %d out of [0,%d)", i, edge2faces_host[i], c->F);
    }
    for (int i = 0; i < c->V; ++i)
        if (flip_index_host[i] < 0 || flip_index_host[i] >= c->V) return fail(MM_E_INVALID, "flip_index[%d] out of range", i);
    MM_REQUIRE(lap_row_off_host[0] == 0 && lap_row_off_host[c->V] == nnz, "laplacian row offsets");
    for (int i = 0; i < c->V; ++i) MM_REQUIRE(lap_row_off_host[i] <= lap_row_off_host[i + 1], "laplacian row offsets must be non-decreasing");
    for (int i = 0; i < nnz; ++i)
        if (lap_col_host[i] < 0 || lap_col_host[i] >= c->V) return fail(MM_E_INVALID, "laplacian column %d out of range", i);
    if (int r = check_device(c)) return r;
    cudaFree(c->d_edges); cudaFree(c->d_edge2faces); cudaFree(c->d_flip); cudaFree(c->d_sign_init);
    cudaFree(c->d_lap_off); cudaFree(c->d_lap_col); cudaFree(c->d_lap_val);
    cudaFree(c->d_lapT_off); cudaFree(c->d_lapT_row); cudaFree(c->d_lapT_val);
    c->d_lapT_off = c->d_lapT_row = nullptr; c->d_lapT_val = nullptr;
    c->d_edges = c->d_edge2faces = c->d_flip = c->d_lap_off = c->d_lap_col = nullptr;
    c->d_sign_init = c->d_lap_val = nullptr;
    MM_CUDA(cudaMalloc(&c->d_edges, (size_t)E * 2 * 4));
    MM_CUDA(cudaMalloc(&c->d_edge2faces, (size_t)E * 2 * 4));
    MM_CUDA(cudaMalloc(&c->d_flip, (size_t)c->V * 4));
    MM_CUDA(cudaMalloc(&c->d_sign_init, (size_t)c->V * 4));
    MM_CUDA(cudaMalloc(&c->d_lap_off, (size_t)(c->V + 1) * 4));
    MM_CUDA(cudaMalloc(&c->d_lap_col, (size_t)nnz * 4));
    MM_CUDA(cudaMalloc(&c->d_lap_val, (size_t)nnz * 4));
    MM_CUDA(cudaMemcpy(c->d_edges, edges_host, (size_t)E * 2 * 4, cudaMemcpyHostToDevice));
    MM_CUDA(cudaMemcpy(c->d_edge2faces, edge2faces_host, (size_t)E * 2 * 4, cudaMemcpyHostToDevice));
    MM_CUDA(cudaMemcpy(c->d_flip, flip_index_host, (size_t)c->V * 4, cudaMemcpyHostToDevice));
    MM_CUDA(cudaMemcpy(c->d_sign_init, sign_init_host, (size_t)c->V * 4, cudaMemcpyHostToDevice));
    MM_CUDA(cudaMemcpy(c->d_lap_off, lap_row_off_host, (size_t)(c->V + 1) * 4, cudaMemcpyHostToDevice));
    MM_CUDA(cudaMemcpy(c->d_lap_col, lap_col_host, (size_t)nnz * 4, cudaMemcpyHostToDevice));
    MM_CUDA(cudaMemcpy(c->d_lap_val, lap_val_host, (size_t)nnz * 4, cudaMemcpyHostToDevice));
    {   // transpose (counting sort by column; rows stay ascending inside a column): x @ lpl walks columns (mm_template.cu)
        std::vector<int32_t> toff(c->V + 1, 0), trow(nnz);
        std::vector<float> tval(nnz);
        for (int k = 0; k < nnz; ++k) ++toff[lap_col_host[k] + 1];
        for (int j = 0; j < c->V; ++j) toff[j + 1] += toff[j];
        std::vector<int32_t> fill(toff.begin(), toff.end() - 1);
        for (int i = 0; i < c->V; ++i)
            for (int k = lap_row_off_host[i]; k < lap_row_off_host[i + 1]; ++k) {
                const int d = fill[lap_col_host[k]]++;
                trow[d] = i; tval[d] = lap_val_host[k];
            }
        MM_CUDA(cudaMalloc(&c->d_lapT_off, (size_t)(c->V + 1) * 4));
        MM_CUDA(cudaMalloc(&c->d_lapT_row, (size_t)nnz * 4));
        MM_CUDA(cudaMalloc(&c->d_lapT_val, (size_t)nnz * 4));
        MM_CUDA(cudaMemcpy(c->d_lapT_off, toff.data(), (size_t)(c->V + 1) * 4, cudaMemcpyHostToDevice));
        MM_CUDA(cudaMemcpy(c->d_lapT_row, trow.data(), (size_t)nnz * 4, cudaMemcpyHostToDevice));
        MM_CUDA(cudaMemcpy(c->d_lapT_val, tval.data(), (size_t)nnz * 4, cudaMemcpyHostToDevice));
    }
    c->reg_E = E; c->reg_ratio = ratio;
    return MM_OK;
}

int mm_mesh_reg_forward(mm_ctx* c, int B, const float* delta_vertices, const float* vertices, const float* face_normals,
                        float temp, float eps, int flip_l1, unsigned term_mask, float* terms, float* scratch, void* stream)
{
    MM_REQUIRE(c && B > 0, "ctx / B");
    if (int r = check_device(c)) return r;
    MM_REQUIRE(c->d_edges, "mm_ctx_set_regularizer_topology has not been called");
    MM_REQUIRE(terms && scratch, "terms / scratch");
    MM_REQUIRE(!(term_mask & (1u | 64u | 128u)) || delta_vertices, "laplacian / deform / flip terms need delta_vertices");
    MM_REQUIRE(!(term_mask & (4u | 8u | 16u | 32u)) || vertices, "edge / depth terms need vertices");
    MM_REQUIRE(!(term_mask & 2u) || face_normals, "flat term needs face_normals");
    // the last-CTA ticket lives in the caller's per-call scratch (behind the B*8 partial sums), not in the ctx
    unsigned* ticket = (unsigned*)(scratch + (size_t)B * 8);
    MM_CUDA(cudaMemsetAsync(ticket, 0, 4, (cudaStream_t)stream));
    MM_LAUNCH(mm_launch_meshreg_fwd(c, B, delta_vertices, vertices, face_normals, temp, eps, flip_l1, term_mask, scratch, ticket,
                                    terms, (cudaStream_t)stream), "meshreg_fwd");
    return MM_OK;
}

int mm_mesh_reg_backward(mm_ctx* c, int B, const float* delta_vertices, const float* vertices, const float* face_normals,
                         float temp, float eps, int flip_l1, unsigned term_mask, const float* g_terms, float* g_delta,
                         float* g_vertices, float* g_face_normals, void* stream)
{
    MM_REQUIRE(c && B > 0, "ctx / B");
    if (int r = check_device(c)) return r;
    MM_REQUIRE(c->d_edges, "mm_ctx_set_regularizer_topology has not been called");
    MM_REQUIRE(g_terms, "g_terms");
    MM_REQUIRE(!(term_mask & (1u | 64u | 128u)) || (delta_vertices && g_delta), "laplacian / deform / flip terms need delta_vertices and g_delta");
    MM_REQUIRE(!(term_mask & (4u | 8u | 16u | 32u)) || (vertices && g_vertices), "edge / depth terms need vertices and g_vertices");
    MM_REQUIRE(!(term_mask & 2u) || (face_normals && g_face_normals), "flat term needs face_normals and g_face_normals");
    MM_LAUNCH(mm_launch_meshreg_bwd(c, B, delta_vertices, vertices, face_normals, temp, eps, flip_l1, term_mask, g_terms, g_delta,
                                    g_vertices, g_face_normals, (cudaStream_t)stream), "meshreg_bwd");
    return MM_OK;
}

int mm_template_features_forward(mm_ctx* c, int N, int h, int w, const float* x, const float* template_xyz, float* local,
                                 float* neighbor_diff, void* stream)
{
    MM_REQUIRE(c && N > 0 && h > 0 && w > 0, "ctx / N / h / w");
    if (int r = check_device(c)) return r;
    MM_REQUIRE(x && template_xyz && local, "NULL argument");
    MM_REQUIRE(!neighbor_diff || c->d_lapT_off, "neighbor_diff needs mm_ctx_set_regularizer_topology (the Laplacian)");
    MM_LAUNCH(mm_launch_template_fwd(c, N, h, w, x, template_xyz, local, neighbor_diff, (cudaStream_t)stream), "template_features_fwd");
    return MM_OK;
}

int mm_template_features_backward(mm_ctx* c, int N, int h, int w, const float* template_xyz, const float* g_local,
                                  const float* g_neighbor_diff, float* g_x, void* stream)
{
    MM_REQUIRE(c && N > 0 && h > 0 && w > 0, "ctx / N / h / w");
    if (int r = check_device(c)) return r;
    MM_REQUIRE(template_xyz && g_x && (g_local || g_neighbor_diff), "NULL argument");
    MM_REQUIRE(!g_neighbor_diff || c->d_lap_off, "g_neighbor_diff needs mm_ctx_set_regularizer_topology (the Laplacian)");
    if ((long long)h * w > mm_template_max_plane())
        return fail(MM_E_UNSUPPORTED, "feature plane %dx%d exceeds %d elements (shared-memory accumulator of the backward)", h, w,
                    mm_template_max_plane());
    MM_LAUNCH(mm_launch_template_bwd(c, N, h, w, template_xyz, g_local, g_neighbor_diff, g_x, (cudaStream_t)stream), "template_features_bwd");
    return MM_OK;
}

int mm_texture_flow_forward(mm_ctx* c, int B, int C, int Hi, int Wi, int Ho, int Wo, int concat, const float* img,
                            const float* flow, float* out, void* stream)
{
    MM_REQUIRE(c && B > 0 && C > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "ctx / sizes");
    if (int r = check_device(c)) return r;
    MM_REQUIRE(img && flow && out, "NULL argument");
    MM_LAUNCH(mm_launch_texflow_fwd(c, B, C, Hi, Wi, Ho, Wo, concat, img, flow, out, (cudaStream_t)stream), "texture_flow_fwd");
    return MM_OK;
}

int mm_texture_flow_backward(mm_ctx* c, int B, int C, int Hi, int Wi, int Ho, int Wo, int concat, const float* img,
                             const float* flow, const float* g_out, float* g_img, float* g_flow, void* stream)
{
    MM_REQUIRE(c && B > 0 && C > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "ctx / sizes");
    if (int r = check_device(c)) return r;
    MM_REQUIRE(img && flow && g_out && g_img && g_flow, "NULL argument");
    MM_CUDA(cudaMemsetAsync(g_img, 0, (size_t)B * C * Hi * Wi * 4, (cudaStream_t)stream));
    MM_LAUNCH(mm_launch_texflow_bwd(c, B, C, Hi, Wi, Ho, Wo, concat, img, flow, g_out, g_img, g_flow, (cudaStream_t)stream),
              "texture_flow_bwd");
    return MM_OK;
}

int mm_ctx_set_timing(mm_ctx* c, int enable) {
    MM_REQUIRE(c, "ctx");
    if (int r = check_device(c)) return r;
    if (enable && !c->ev[0]) {
        for (int i = 0; i < 8; ++i) MM_CUDA(cudaEventCreate(&c->ev[i]));
    }
    c->timing = enable ? 1 : 0;
    return MM_OK;
}

int mm_ctx_get_timing(mm_ctx* c, float* ms_host, int capacity) {
    MM_REQUIRE(c && ms_host && capacity >= 7, "ctx / ms_host / capacity >= 7");
    MM_REQUIRE(c->timing && c->ev[0], "timing not enabled");
    MM_CUDA(cudaEventSynchronize(c->ev[7]));
    for (int i = 0; i < 7; ++i) MM_CUDA(cudaEventElapsedTime(&ms_host[i], c->ev[i], c->ev[i + 1]));
    return 7;
}

size_t mm_debug_workspace_offset(const mm_ctx* c, int B, const char* block)
{
    if (!c || B <= 0 || !block) return (size_t)-1;
    const mm_ws_layout L = mm_ws_make(c, B);
    const struct { const char* name; size_t off; } tab[] = {
        {"frec", L.frec}, {"frect", L.frect}, {"zbuf", L.zbuf}, {"lacc", L.lacc}, {"cov", L.cov}, {"ovf_count", L.ovf_count},
        {"sched_n", L.sched_n}, {"ovf_list", L.ovf_list}, {"sched_list", L.sched_list}, {"plist", L.plist}, {"gsoft", L.gsoft},
        {"gfacc", L.gfacc}, {"img_fwd", L.img_fwd}, {"img_bwd", L.img_bwd}};
    for (const auto& t : tab) if (!strcmp(block, t.name)) return t.off;
    return (size_t)-1;
}

#ifdef MM_PROF
// MM_PROF builds only (not part of the ABI): device buffer [6][16384][4] u64 of per-warp time stamps, or NULL
int mm_debug_profile(mm_ctx* c, unsigned long long* buf) { if (!c) return MM_E_INVALID; c->prof = buf; return MM_OK; }
#endif

int mm_debug_export_faces(mm_ctx* c, int B, const void* workspace, size_t workspace_bytes, float* fvi, float* fvz, float* fnz,
                          void* stream)
{
    MM_COMMON_CHECKS(c, B, workspace, workspace_bytes);
    const mm_ws_layout L = mm_ws_make(c, B);
    const char* ws = (const char*)workspace;
    MM_LAUNCH(mm_launch_export_faces(c, B, (const float*)(ws + L.frec), (const float*)(ws + L.vimg), fvi, fvz, fnz,
                                     (cudaStream_t)stream), "export_faces");
    return MM_OK;
}

}  // extern "C"
