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

int check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return fail(MM_E_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return MM_OK;
}

void fill_params(const mm_ctx* c, int B, int Ht, int Wt, int no_mask, mm_raster_params& p) {
    memset(&p, 0, sizeof(p));
    p.B = B; p.V = c->V; p.F = c->F; p.H = c->H; p.W = c->W; p.Ht = Ht; p.Wt = Wt;
    p.Htp = c->tex_mirror ? Ht / 2 : Ht;
    p.nstx = c->nstx; p.nsty = c->nsty; p.nst = c->nst; p.knum = c->knum;
    p.sx = c->sx; p.sy = c->sy; p.blen = c->blen; p.multiplier = c->multiplier; p.eps = c->eps; p.sigmainv = c->sigmainv;
    p.no_mask = no_mask;
    p.covw = (c->W + 31) / 32;
    p.face_uvs = c->d_face_uvs;
    p.tab = c->d_tab;
}

void set_ws(const mm_ctx* c, const mm_ws_layout& L, char* ws, mm_raster_params& p) {
    p.frec = (const float*)(ws + L.frec);
    p.zbuf = (unsigned long long*)(ws + L.zbuf); p.lacc = (unsigned long long*)(ws + L.lacc);
    p.cov = (uint32_t*)(ws + L.cov);
    p.ovf_list = (uint32_t*)(ws + L.ovf_list); p.ovf_count = (uint32_t*)(ws + L.ovf_count);
    p.gsoft = (float*)(ws + L.gsoft);
    p.plist = (unsigned long long*)(ws + L.plist); p.plist_cap = (uint32_t)((L.gsoft - L.plist) / 8);
    if (c->plist_cap_max && p.plist_cap > c->plist_cap_max) p.plist_cap = c->plist_cap_max;
    p.img_fwd = (long long*)(ws + L.img_fwd); p.img_bwd = (long long*)(ws + L.img_bwd);
    p.gfacc = (float*)(ws + L.gfacc);
}

// vertex stage + the single memset that clears the visibility buffer, the silhouette accumulators and the overflow counter
int launch_vertex_fwd(const mm_ctx* c, int B, const mm_ws_layout& L, char* ws, const float* vertices, const float* azim,
                      const float* elev, const float* dist, const float* bias, float* face_normals, bool zero_gfacc,
                      void* clr1, size_t bytes1, cudaStream_t s) {
    // zbuf .. ovf_count are contiguous and 256-byte aligned: one clear range (16-byte units, tail padded inside the workspace)
    const size_t bytes0 = mm_align_up((L.ovf_count + 16) - L.zbuf, 16);
    if (clr1 && (((uintptr_t)clr1 & 15) != 0 || (bytes1 & 15) != 0)) {          // unaligned caller buffer: plain memset
        if (cudaMemsetAsync(clr1, 0, bytes1, s) != cudaSuccess) return 1;
        clr1 = nullptr; bytes1 = 0;
    }
    mm_launch_vertex_fwd(c, B, vertices, azim, elev, dist, bias, (float*)(ws + L.frec), (float*)(ws + L.vimg), face_normals,
                         zero_gfacc ? (float*)(ws + L.gfacc) : nullptr, (long long*)(ws + L.img_fwd), (long long*)(ws + L.img_bwd),
                         ws + L.zbuf, bytes0, clr1, bytes1, s);
    return 0;
}

}  // namespace

int g_mm_pdl = 1;

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
    MM_CUDA(cudaSetDevice(device));
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
    c->nstx = (W + MM_ST_W - 1) / MM_ST_W;
    c->nsty = (H + MM_ST_H - 1) / MM_ST_H;
    c->nst = c->nstx * c->nsty;
    c->nparts_recon = (H * W + 4095) / 4096 < 1 ? 1 : (H * W + 4095) / 4096;     // ~4096 pixels per recon CTA
    c->num_sms = prop.multiProcessorCount;
    if (c->nst > 65535) { delete c; return fail(MM_E_UNSUPPORTED, "image too large: %d sub-tiles exceed the 16-bit work-list ids", c->nst); }
    if (F > 65535) { delete c; return fail(MM_E_UNSUPPORTED, "F=%d exceeds the 16-bit face ids of the soft-pass lists", F); }
    const size_t smem_max = prop.sharedMemPerBlockOptin;
    // vertex stage: CTAs per image (each recomputes the vertex transform and emits 1/nchunks of the face records)
    c->nchunks = 8;
    if (const char* e = getenv("MM_PDL")) g_mm_pdl = atoi(e) != 0;
    if (const char* e = getenv("MM_PLIST_CAP")) c->plist_cap_max = (unsigned)atoi(e);
    // measured at cfg-2: 0.119 ms/step split vs 0.117 sequential -- both roles are latency-bound and share one register budget
    // (128/thread from the shading role), so side by side they only trade warps; kept selectable for when shading slims down
    c->split = 0;
    if (const char* e = getenv("MM_SPLIT")) c->split = atoi(e) != 0;
    c->parts = 1;
    if (const char* e = getenv("MM_PARTS")) { const int v = atoi(e); if (v >= 1 && v <= MM_MAX_PARTS) c->parts = v; }
    if (const char* e = getenv("MM_VCHUNKS")) { const int v = atoi(e); if (v > 0 && v <= 32) c->nchunks = v; }
    c->smem_vertex_fwd = mm_vertex_smem_fwd(c);
    const size_t vs_f = c->smem_vertex_fwd, vs_b = mm_vertex_smem_bwd(V);
    if (vs_f > smem_max || vs_b > smem_max) {
        const size_t need = vs_f > vs_b ? vs_f : vs_b;
        delete c;
        return fail(MM_E_UNSUPPORTED, "V=%d F=%d W=%d needs %zu B of shared memory per CTA (> %zu)", V, F, W, need, smem_max);
    }
    mm_vertex_set_smem(vs_f, vs_b);

    std::vector<int32_t> tab(3 * (size_t)H + 3 * (size_t)W);
    contour_tables(H, tab.data(), tab.data() + H, tab.data() + 2 * H);
    contour_tables(W, tab.data() + 3 * H, tab.data() + 3 * H + W, tab.data() + 3 * H + 2 * W);
    if (cudaMalloc(&c->d_faces, (size_t)F * 3 * 4) != cudaSuccess ||
        cudaMalloc(&c->d_face_uvs, (size_t)F * 6 * 4) != cudaSuccess ||
        cudaMalloc(&c->d_tab, tab.size() * 4) != cudaSuccess) {
        mm_ctx_destroy(c);
        return fail(MM_E_CUDA, "cudaMalloc of ctx tables failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    MM_CUDA(cudaMemcpy(c->d_faces, faces_host, (size_t)F * 3 * 4, cudaMemcpyHostToDevice));
    MM_CUDA(cudaMemcpy(c->d_face_uvs, face_uvs_host, (size_t)F * 6 * 4, cudaMemcpyHostToDevice));
    MM_CUDA(cudaMemcpy(c->d_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    *out = c;
    return MM_OK;
}

int mm_ctx_destroy(mm_ctx* c) {
    if (!c) return MM_OK;
    cudaFree(c->d_edges); cudaFree(c->d_edge2faces); cudaFree(c->d_flip); cudaFree(c->d_sign_init);
    cudaFree(c->d_lap_off); cudaFree(c->d_lap_col); cudaFree(c->d_lap_val); cudaFree(c->d_reg_ticket);
    cudaFree(c->d_lapT_off); cudaFree(c->d_lapT_row); cudaFree(c->d_lapT_val);
    cudaFree(c->d_faces);
    cudaFree(c->d_face_uvs);
    cudaFree(c->d_tab);
    for (int i = 0; i < 8; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->part_fork) cudaEventDestroy(c->part_fork);
    for (int i = 0; i < MM_MAX_PARTS - 1; ++i) {
        if (c->part_join[i]) cudaEventDestroy(c->part_join[i]);
        if (c->part_stream[i]) cudaStreamDestroy(c->part_stream[i]);
    }
    delete c;
    return MM_OK;
}

size_t mm_workspace_bytes(const mm_ctx* c, int B) {
    if (!c || B <= 0) return 0;
    // enough for the unsplit layout and for every split of the fused step into up to MM_MAX_PARTS sub-batches
    size_t need = mm_ws_make(c, B).total;
    for (int parts = 2; parts <= MM_MAX_PARTS && parts <= B; ++parts) {
        size_t sum = 0;
        for (int i = 0; i < parts; ++i) sum += mm_align_up(mm_ws_make(c, B / parts + (i < B % parts ? 1 : 0)).total, 256);
        sum += 256;                                        // the parts' loss scalars
        if (sum > need) need = sum;
    }
    return need;
}

int mm_render_forward(mm_ctx* c, int B, const float* vertices, const float* azim, const float* elev, const float* dist,
                      const float* bias, const float* tex, int Ht, int Wt, const float* lights, const float* bg,
                      int no_mask, float* rgba, float* face_normals, float* imnormal, int32_t* face_idx,
                      void* workspace, void* stream)
{
    MM_REQUIRE(c && B > 0 && B <= 65535, "ctx / B (1..65535)");
    MM_REQUIRE(vertices && azim && elev && dist && bias && tex && lights, "NULL input");
    MM_REQUIRE(Ht > 0 && Wt > 0, "texture size");
    MM_REQUIRE(!c->tex_mirror || (Ht & 1) == 0, "a mirrored texture needs an even logical height Ht");
    MM_REQUIRE(!no_mask || bg, "no_mask=1 requires bg");
    MM_REQUIRE(rgba && workspace, "rgba / workspace");
    cudaStream_t s = (cudaStream_t)stream;
    const mm_ws_layout L = mm_ws_make(c, B);
    char* ws = (char*)workspace;
    launch_vertex_fwd(c, B, L, ws, vertices, azim, elev, dist, bias, face_normals, false, nullptr, 0, s);
    if (int r = check_launch("vertex_fwd")) return r;
    mm_raster_params p;
    fill_params(c, B, Ht, Wt, no_mask, p);
    set_ws(c, L, ws, p);
    p.tex = tex; p.lights = lights; p.bg = bg;
    p.rgba = rgba; p.imnormal = imnormal; p.face_idx_out = face_idx;
    mm_launch_geom_fwd(c, p, s);
    if (int r = check_launch("geom_fwd")) return r;
    mm_launch_shade_fwd(c, p, false, s);
    return check_launch("shade_fwd");
}

int mm_render_backward(mm_ctx* c, int B, const float* vertices, const float* azim, const float* elev, const float* dist,
                       const float* bias, const float* tex, int Ht, int Wt, const float* lights, const float* bg,
                       int no_mask, const float* rgba, const float* g_rgba, const float* g_face_normals,
                       float* g_vertices, float* g_azim, float* g_elev, float* g_dist, float* g_bias, float* g_tex,
                       float* g_lights, float* g_bg, void* workspace, void* stream)
{
    MM_REQUIRE(c && B > 0, "ctx / B");
    MM_REQUIRE(vertices && azim && elev && dist && bias && tex && lights, "NULL input");
    MM_REQUIRE(!no_mask || bg, "no_mask=1 requires bg");
    MM_REQUIRE(rgba && g_rgba && workspace, "rgba / g_rgba / workspace");
    MM_REQUIRE(Ht > 0 && Wt > 0 && (!c->tex_mirror || (Ht & 1) == 0), "texture size (even Ht for a mirrored texture)");
    MM_REQUIRE(g_vertices && g_azim && g_elev && g_dist && g_bias && g_tex && g_lights, "NULL gradient output");
    cudaStream_t s = (cudaStream_t)stream;
    const mm_ws_layout L = mm_ws_make(c, B);
    char* ws = (char*)workspace;
    const size_t HW = (size_t)c->H * c->W;
    MM_CUDA(cudaMemsetAsync(ws + L.gfacc, 0, (size_t)B * c->F * MM_GF * 4, s));
    MM_CUDA(cudaMemsetAsync(g_tex, 0, (size_t)B * 3 * (c->tex_mirror ? Ht / 2 : Ht) * Wt * 4, s));
    if (g_bg && !no_mask) MM_CUDA(cudaMemsetAsync(g_bg, 0, (size_t)B * 3 * HW * 4, s));
    mm_raster_params p;
    fill_params(c, B, Ht, Wt, no_mask, p);
    set_ws(c, L, ws, p);
    p.tex = tex; p.lights = lights; p.bg = bg;
    p.rgba = const_cast<float*>(rgba);
    p.g_rgba = g_rgba;
    p.analytic_loss = 0;
    p.g_tex = g_tex; p.g_bg = no_mask ? g_bg : nullptr;
    mm_launch_shade_bwd(c, p, s);
    if (int r = check_launch("shade_bwd")) return r;
    mm_launch_geom_bwd(c, p, s);
    if (int r = check_launch("geom_bwd")) return r;
    mm_launch_vertex_bwd(c, B, vertices, azim, elev, dist, bias, p.gfacc, g_face_normals, p.img_bwd, 1, g_vertices,
                         g_azim, g_elev, g_dist, g_bias, g_lights, nullptr, nullptr, 0.0f, 0.0f, s);
    return check_launch("vertex_bwd");
}

int mm_face_normals_forward(mm_ctx* c, int B, const float* vertices, const float* azim, const float* elev, const float* dist,
                            const float* bias, float* face_normals, void* workspace, void* stream)
{
    MM_REQUIRE(c && B > 0 && B <= 65535, "ctx / B (1..65535)");
    MM_REQUIRE(vertices && azim && elev && dist && bias && face_normals && workspace, "NULL argument");
    const mm_ws_layout L = mm_ws_make(c, B);
    char* ws = (char*)workspace;
    mm_launch_vertex_fwd(c, B, vertices, azim, elev, dist, bias, (float*)(ws + L.frec), (float*)(ws + L.vimg), face_normals,
                         nullptr, (long long*)(ws + L.img_fwd), (long long*)(ws + L.img_bwd), nullptr, 0, nullptr, 0,
                         (cudaStream_t)stream);
    return check_launch("vertex_fwd");
}

int mm_face_normals_backward(mm_ctx* c, int B, const float* vertices, const float* azim, const float* elev, const float* dist,
                             const float* bias, const float* g_face_normals, float* g_vertices, float* g_azim, float* g_elev,
                             float* g_dist, float* g_bias, void* workspace, void* stream)
{
    MM_REQUIRE(c && B > 0, "ctx / B");
    MM_REQUIRE(vertices && azim && elev && dist && bias && g_face_normals && workspace, "NULL argument");
    MM_REQUIRE(g_vertices && g_azim && g_elev && g_dist && g_bias, "NULL gradient output");
    cudaStream_t s = (cudaStream_t)stream;
    const mm_ws_layout L = mm_ws_make(c, B);
    char* ws = (char*)workspace;
    MM_CUDA(cudaMemsetAsync(ws + L.gfacc, 0, (size_t)B * c->F * MM_GF * 4, s));
    mm_launch_vertex_bwd(c, B, vertices, azim, elev, dist, bias, (const float*)(ws + L.gfacc), g_face_normals,
                         (long long*)(ws + L.img_bwd), 0, g_vertices, g_azim, g_elev, g_dist, g_bias, nullptr, nullptr, nullptr,
                         0.0f, 0.0f, s);
    return check_launch("vertex_bwd");
}

int mm_recon_data_forward(mm_ctx* c, int B, const float* pred, const float* gt, float image_weight, float contour,
                          float* loss, float* iou_sums, void* workspace, void* stream)
{
    MM_REQUIRE(c && B > 0, "ctx / B");
    MM_REQUIRE(pred && gt && loss && workspace, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    const mm_ws_layout L = mm_ws_make(c, B);
    char* ws = (char*)workspace;
    mm_launch_recon_fwd(c, B, pred, gt, contour, (float*)(ws + L.part_fwd), s);
    if (int r = check_launch("recon_fwd")) return r;
    mm_launch_image_reduce(c, B, c->nparts_recon, (const float*)(ws + L.part_fwd), (long long*)(ws + L.img_fwd), s);
    mm_launch_loss_finalize(c, B, (const long long*)(ws + L.img_fwd), nullptr, image_weight, contour, loss, iou_sums, s);
    return check_launch("loss_finalize");
}

int mm_recon_data_backward(mm_ctx* c, int B, const float* pred, const float* gt, float image_weight, float contour,
                           float loss_scale, float* g_pred, void* workspace, void* stream)
{
    MM_REQUIRE(c && B > 0, "ctx / B");
    MM_REQUIRE(pred && gt && g_pred && workspace, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    const mm_ws_layout L = mm_ws_make(c, B);
    char* ws = (char*)workspace;
    // the IoU sums are re-derived so that the call does not depend on workspace state of an earlier forward
    mm_launch_recon_fwd(c, B, pred, gt, 0.0f, (float*)(ws + L.part_fwd), s);
    if (int r = check_launch("recon_fwd")) return r;
    mm_launch_image_reduce(c, B, c->nparts_recon, (const float*)(ws + L.part_fwd), (long long*)(ws + L.img_fwd), s);
    mm_launch_recon_bwd(c, B, pred, gt, (const long long*)(ws + L.img_fwd), image_weight, contour, loss_scale, g_pred, s);
    return check_launch("recon_bwd");
}

}  // extern "C"

// one fused step over B images on stream s (the whole of mm_render_compare_fwd_bwd when the batch is not split)
static int fused_step(mm_ctx* c, int B, const float* vertices, const float* azim, const float* elev,
                      const float* dist, const float* bias, const float* tex, int Ht, int Wt,
                      const float* lights, const float* bg, int no_mask, const float* gt, float image_weight,
                      float contour, float loss_scale, const float* g_rgba_extra, const float* g_face_normals,
                      float* rgba,
                      float* face_normals, float* loss, float* g_vertices, float* g_azim, float* g_elev,
                      float* g_dist, float* g_bias, float* g_tex, float* g_lights, float* g_bg,
                      void* workspace, cudaStream_t s)
{
    const mm_ws_layout L = mm_ws_make(c, B);
    char* ws = (char*)workspace;
    const size_t HW = (size_t)c->H * c->W;
    if (g_bg && !no_mask) MM_CUDA(cudaMemsetAsync(g_bg, 0, (size_t)B * 3 * HW * 4, s));
    if (c->timing) cudaEventRecord(c->ev[0], s);
    const size_t gtex_bytes = (size_t)B * 3 * (c->tex_mirror ? Ht / 2 : Ht) * Wt * 4;
    const bool gtex_side = (((uintptr_t)g_tex & 15) == 0) && ((gtex_bytes & 15) == 0);     // cleared by the hard pass on the side
    if (!gtex_side) MM_CUDA(cudaMemsetAsync(g_tex, 0, gtex_bytes, s));
    launch_vertex_fwd(c, B, L, ws, vertices, azim, elev, dist, bias, face_normals, true, nullptr, 0, s);
    if (int r = check_launch("vertex_fwd")) return r;
    if (c->timing) cudaEventRecord(c->ev[1], s);
    mm_raster_params p;
    fill_params(c, B, Ht, Wt, no_mask, p);
    set_ws(c, L, ws, p);
    p.tex = tex; p.lights = lights; p.bg = bg; p.gt = gt;
    p.rgba = rgba;
    p.g_rgba = g_rgba_extra;
    p.image_weight = image_weight; p.contour = contour; p.loss_scale = loss_scale;
    p.analytic_loss = 1;
    p.g_tex = g_tex; p.g_bg = no_mask ? g_bg : nullptr;
    // H, W multiples of 4: the contour term is tile-local, so d(loss)/d(silhouette) minus its IoU term is emitted together with
    // the alpha plane and the geometry backward adds the IoU term from the per-image sums on the fly
    p.gsoft_iou_pending = ((c->H & 3) == 0 && (c->W & 3) == 0) ? 1 : 0;
    if (gtex_side) { p.clr = (uint4*)g_tex; p.nclr = gtex_bytes / 16; }
    if (c->split) {
        // hard pass | soft forward || RGB shading fwd+bwd (one launch) | overflow pass | final silhouette (alpha, IoU sums, gsoft)
        mm_launch_hard(c, p, s);
        p.clr = nullptr; p.nclr = 0;
        if (int r = check_launch("hard")) return r;
        if (c->timing) cudaEventRecord(c->ev[2], s);
        mm_launch_soft_shade(c, p, s);
        if (int r = check_launch("soft_shade")) return r;
        if (c->timing) cudaEventRecord(c->ev[3], s);
        mm_launch_soft_ovf_fwd(c, p, s);
        mm_launch_alpha(c, p, s);
        if (int r = check_launch("alpha")) return r;
    } else {
        mm_launch_geom_fwd(c, p, s);
        p.clr = nullptr; p.nclr = 0;
        if (int r = check_launch("geom_fwd")) return r;
        if (c->timing) cudaEventRecord(c->ev[2], s);
        mm_launch_shade_fused(c, p, s);           // shading forward + loss sums + the whole RGB-side backward
        if (int r = check_launch("shade_fused")) return r;
        if (c->timing) cudaEventRecord(c->ev[3], s);
    }
    if (!p.gsoft_iou_pending) {
        mm_launch_gsoft(c, p, s);                 // general sizes: d(loss)/d(silhouette) in its own pass (index tables)
        if (int r = check_launch("gsoft")) return r;
    }
    if (c->timing) cudaEventRecord(c->ev[4], s);
    mm_launch_geom_bwd(c, p, s);
    if (int r = check_launch("geom_bwd")) return r;
    if (c->timing) cudaEventRecord(c->ev[5], s);
    // vertex backward; its last CTA also finalises the loss scalars (img_fwd / img_bwd are re-zeroed by the next vertex_fwd)
    mm_launch_vertex_bwd(c, B, vertices, azim, elev, dist, bias, p.gfacc, g_face_normals, p.img_bwd, 0, g_vertices, g_azim,
                         g_elev, g_dist, g_bias, g_lights, loss, p.img_fwd, image_weight, contour, s);
    if (c->timing) { cudaEventRecord(c->ev[6], s); cudaEventRecord(c->ev[7], s); }
    return check_launch("vertex_bwd");
}

// loss[k] = sum_i w[i] * part_loss[i][k]: every term of recon_data is a mean over the batch, so the whole-batch value is the
// image-count-weighted mean of the sub-batch values
__global__ void k_combine_loss(const float* __restrict__ part_loss, int nparts, float4 w, float* __restrict__ loss)
{
    mm_pdl_prologue();
    const int k = threadIdx.x;
    if (k >= 4) return;
    const float ww[4] = {w.x, w.y, w.z, w.w};
    float a = 0.0f;
    for (int i = 0; i < nparts; ++i) a += ww[i] * part_loss[i * 4 + k];
    loss[k] = a;
}

static inline int part_images(int B, int parts, int i) { return B / parts + (i < B % parts ? 1 : 0); }

// Sub-batches of the fused step run as independent kernel chains on the caller's stream + (parts-1) side streams.  Images are
// independent through render and loss, every kernel of the chain is latency- rather than throughput-bound at the
// benchmark's batch size, and the chain is strictly sequential: two half-size chains side by side fill the issue slots and
// the launch / tail gaps one chain leaves idle.  Results are those of the unsplit call (gradients: loss_scale * B_i / B per
// part; loss: weighted mean of the parts' scalars).
static int split_parts(const mm_ctx* c, int B) {
    if (c->parts <= 1 || c->timing) return 1;              // the per-kernel timing hook measures the unsplit chain
    int parts = c->parts;
    while (parts > 1 && B / parts < 8) --parts;            // below 8 images a sub-batch does not fill one wave of any kernel
    return parts;
}

extern "C" {

int mm_render_compare_fwd_bwd(mm_ctx* c, int B, const float* vertices, const float* azim, const float* elev,
                              const float* dist, const float* bias, const float* tex, int Ht, int Wt,
                              const float* lights, const float* bg, int no_mask, const float* gt, float image_weight,
                              float contour, float loss_scale, const float* g_rgba_extra, const float* g_face_normals,
                              float* rgba,
                              float* face_normals, float* loss, float* g_vertices, float* g_azim, float* g_elev,
                              float* g_dist, float* g_bias, float* g_tex, float* g_lights, float* g_bg,
                              void* workspace, void* stream)
{
    MM_REQUIRE(c && B > 0, "ctx / B");
    MM_REQUIRE(vertices && azim && elev && dist && bias && tex && lights && gt, "NULL input");
    MM_REQUIRE(Ht > 0 && Wt > 0, "texture size");
    MM_REQUIRE(!c->tex_mirror || (Ht & 1) == 0, "a mirrored texture needs an even logical height Ht");
    MM_REQUIRE(!no_mask || bg, "no_mask=1 requires bg");
    MM_REQUIRE(rgba && loss && workspace, "rgba / loss / workspace");
    MM_REQUIRE(g_vertices && g_azim && g_elev && g_dist && g_bias && g_tex && g_lights, "NULL gradient output");
    cudaStream_t s = (cudaStream_t)stream;
    const int parts = split_parts(c, B);
    if (parts == 1)
        return fused_step(c, B, vertices, azim, elev, dist, bias, tex, Ht, Wt, lights, bg, no_mask, gt, image_weight, contour,
                          loss_scale, g_rgba_extra, g_face_normals, rgba, face_normals, loss, g_vertices, g_azim, g_elev,
                          g_dist, g_bias, g_tex, g_lights, g_bg, workspace, s);
    if (!c->part_fork) {                                   // side streams and fork / join events, once per ctx
        MM_CUDA(cudaSetDevice(c->device));
        MM_CUDA(cudaEventCreateWithFlags(&c->part_fork, cudaEventDisableTiming));
        for (int i = 0; i < MM_MAX_PARTS - 1; ++i) {
            MM_CUDA(cudaStreamCreateWithFlags(&c->part_stream[i], cudaStreamNonBlocking));
            MM_CUDA(cudaEventCreateWithFlags(&c->part_join[i], cudaEventDisableTiming));
        }
    }
    const size_t HW = (size_t)c->H * c->W, V3 = (size_t)c->V * 3, F3 = (size_t)c->F * 3;
    const size_t texn = (size_t)3 * (c->tex_mirror ? Ht / 2 : Ht) * Wt;
    char* ws = (char*)workspace;
    // workspace: [part 0 | part 1 | ... | part losses (parts x 4 floats)]
    size_t ws_off[MM_MAX_PARTS + 1];
    ws_off[0] = 0;
    for (int i = 0; i < parts; ++i) ws_off[i + 1] = ws_off[i] + mm_align_up(mm_ws_make(c, part_images(B, parts, i)).total, 256);
    float* part_loss = (float*)(ws + ws_off[parts]);
    MM_CUDA(cudaEventRecord(c->part_fork, s));
    float wgt[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    int b0 = 0;
    for (int i = 0; i < parts; ++i) {
        const int Bi = part_images(B, parts, i);
        cudaStream_t si = (i == 0) ? s : c->part_stream[i - 1];
        if (i > 0) MM_CUDA(cudaStreamWaitEvent(si, c->part_fork, 0));
        wgt[i] = (float)Bi / (float)B;
        const size_t o = (size_t)b0;
        const int r = fused_step(c, Bi, vertices + o * V3, azim + o, elev + o, dist + o, bias + o * 2, tex + o * texn, Ht, Wt,
                                 lights + o * 9, bg ? bg + o * 3 * HW : nullptr, no_mask, gt + o * 4 * HW, image_weight, contour,
                                 loss_scale * wgt[i], g_rgba_extra ? g_rgba_extra + o * 4 * HW : nullptr,
                                 g_face_normals ? g_face_normals + o * F3 : nullptr, rgba + o * 4 * HW,
                                 face_normals ? face_normals + o * F3 : nullptr, part_loss + i * 4, g_vertices + o * V3,
                                 g_azim + o, g_elev + o, g_dist + o, g_bias + o * 2, g_tex + o * texn, g_lights + o * 9,
                                 g_bg ? g_bg + o * 3 * HW : nullptr, ws + ws_off[i], si);
        if (i > 0) MM_CUDA(cudaEventRecord(c->part_join[i - 1], si));
        if (r) {                                           // keep the caller's stream ordered after whatever was enqueued
            for (int j = 1; j <= i; ++j) cudaStreamWaitEvent(s, c->part_join[j - 1], 0);
            return r;
        }
        b0 += Bi;
    }
    for (int i = 1; i < parts; ++i) MM_CUDA(cudaStreamWaitEvent(s, c->part_join[i - 1], 0));
    mm_launch(k_combine_loss, dim3(1), dim3(32), 0, s, false, (const float*)part_loss, parts,
              make_float4(wgt[0], wgt[1], wgt[2], wgt[3]), loss);
    return check_launch("combine_loss");
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
    MM_CUDA(cudaSetDevice(c->device));
    cudaFree(c->d_edges); cudaFree(c->d_edge2faces); cudaFree(c->d_flip); cudaFree(c->d_sign_init);
    cudaFree(c->d_lap_off); cudaFree(c->d_lap_col); cudaFree(c->d_lap_val); cudaFree(c->d_reg_ticket);
    cudaFree(c->d_lapT_off); cudaFree(c->d_lapT_row); cudaFree(c->d_lapT_val);
    c->d_lapT_off = c->d_lapT_row = nullptr; c->d_lapT_val = nullptr;
    c->d_edges = c->d_edge2faces = c->d_flip = c->d_lap_off = c->d_lap_col = nullptr;
    c->d_sign_init = c->d_lap_val = nullptr; c->d_reg_ticket = nullptr;
    MM_CUDA(cudaMalloc(&c->d_edges, (size_t)E * 2 * 4));
    MM_CUDA(cudaMalloc(&c->d_edge2faces, (size_t)E * 2 * 4));
    MM_CUDA(cudaMalloc(&c->d_flip, (size_t)c->V * 4));
    MM_CUDA(cudaMalloc(&c->d_sign_init, (size_t)c->V * 4));
    MM_CUDA(cudaMalloc(&c->d_lap_off, (size_t)(c->V + 1) * 4));
    MM_CUDA(cudaMalloc(&c->d_lap_col, (size_t)nnz * 4));
    MM_CUDA(cudaMalloc(&c->d_lap_val, (size_t)nnz * 4));
    MM_CUDA(cudaMalloc(&c->d_reg_ticket, 4));
    MM_CUDA(cudaMemcpy(c->d_edges, edges_host, (size_t)E * 2 * 4, cudaMemcpyHostToDevice));
    MM_CUDA(cudaMemcpy(c->d_edge2faces, edge2faces_host, (size_t)E * 2 * 4, cudaMemcpyHostToDevice));
    MM_CUDA(cudaMemcpy(c->d_flip, flip_index_host, (size_t)c->V * 4, cudaMemcpyHostToDevice));
    MM_CUDA(cudaMemcpy(c->d_sign_init, sign_init_host, (size_t)c->V * 4, cudaMemcpyHostToDevice));
    MM_CUDA(cudaMemcpy(c->d_lap_off, lap_row_off_host, (size_t)(c->V + 1) * 4, cudaMemcpyHostToDevice));
    MM_CUDA(cudaMemcpy(c->d_lap_col, lap_col_host, (size_t)nnz * 4, cudaMemcpyHostToDevice));
    MM_CUDA(cudaMemcpy(c->d_lap_val, lap_val_host, (size_t)nnz * 4, cudaMemcpyHostToDevice));
    MM_CUDA(cudaMemset(c->d_reg_ticket, 0, 4));
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
    MM_REQUIRE(c->d_edges, "mm_ctx_set_regularizer_topology has not been called");
    MM_REQUIRE(terms && scratch, "terms / scratch");
    MM_REQUIRE(!(term_mask & (1u | 64u | 128u)) || delta_vertices, "laplacian / deform / flip terms need delta_vertices");
    MM_REQUIRE(!(term_mask & (4u | 8u | 16u | 32u)) || vertices, "edge / depth terms need vertices");
    MM_REQUIRE(!(term_mask & 2u) || face_normals, "flat term needs face_normals");
    mm_launch_meshreg_fwd(c, B, delta_vertices, vertices, face_normals, temp, eps, flip_l1, term_mask, scratch, terms,
                          (cudaStream_t)stream);
    return check_launch("meshreg_fwd");
}

int mm_mesh_reg_backward(mm_ctx* c, int B, const float* delta_vertices, const float* vertices, const float* face_normals,
                         float temp, float eps, int flip_l1, unsigned term_mask, const float* g_terms, float* g_delta,
                         float* g_vertices, float* g_face_normals, void* stream)
{
    MM_REQUIRE(c && B > 0, "ctx / B");
    MM_REQUIRE(c->d_edges, "mm_ctx_set_regularizer_topology has not been called");
    MM_REQUIRE(g_terms, "g_terms");
    MM_REQUIRE(!(term_mask & (1u | 64u | 128u)) || (delta_vertices && g_delta), "laplacian / deform / flip terms need delta_vertices and g_delta");
    MM_REQUIRE(!(term_mask & (4u | 8u | 16u | 32u)) || (vertices && g_vertices), "edge / depth terms need vertices and g_vertices");
    MM_REQUIRE(!(term_mask & 2u) || (face_normals && g_face_normals), "flat term needs face_normals and g_face_normals");
    mm_launch_meshreg_bwd(c, B, delta_vertices, vertices, face_normals, temp, eps, flip_l1, term_mask, g_terms, g_delta,
                          g_vertices, g_face_normals, (cudaStream_t)stream);
    return check_launch("meshreg_bwd");
}

int mm_ctx_set_texture_mirror(mm_ctx* c, int enable) {
    MM_REQUIRE(c, "ctx");
    c->tex_mirror = enable ? 1 : 0;
    return MM_OK;
}

int mm_template_features_forward(mm_ctx* c, int N, int h, int w, const float* x, const float* template_xyz, float* local,
                                 float* neighbor_diff, void* stream)
{
    MM_REQUIRE(c && N > 0 && h > 0 && w > 0, "ctx / N / h / w");
    MM_REQUIRE(x && template_xyz && local, "NULL argument");
    MM_REQUIRE(!neighbor_diff || c->d_lapT_off, "neighbor_diff needs mm_ctx_set_regularizer_topology (the Laplacian)");
    if (cudaError_t e = mm_launch_template_fwd(c, N, h, w, x, template_xyz, local, neighbor_diff, (cudaStream_t)stream))
        return fail(MM_E_CUDA, "template_features_fwd: %s", cudaGetErrorString(e));
    return check_launch("template_features_fwd");
}

int mm_template_features_backward(mm_ctx* c, int N, int h, int w, const float* template_xyz, const float* g_local,
                                  const float* g_neighbor_diff, float* g_x, void* stream)
{
    MM_REQUIRE(c && N > 0 && h > 0 && w > 0, "ctx / N / h / w");
    MM_REQUIRE(template_xyz && g_x && (g_local || g_neighbor_diff), "NULL argument");
    MM_REQUIRE(!g_neighbor_diff || c->d_lap_off, "g_neighbor_diff needs mm_ctx_set_regularizer_topology (the Laplacian)");
    if ((long long)h * w > mm_template_max_plane())
        return fail(MM_E_UNSUPPORTED, "feature plane %dx%d exceeds %d elements (shared-memory accumulator of the backward)", h, w,
                    mm_template_max_plane());
    if (cudaError_t e = mm_launch_template_bwd(c, N, h, w, template_xyz, g_local, g_neighbor_diff, g_x, (cudaStream_t)stream))
        return fail(MM_E_CUDA, "template_features_bwd: %s", cudaGetErrorString(e));
    return check_launch("template_features_bwd");
}

int mm_ctx_set_parts(mm_ctx* c, int parts) {
    MM_REQUIRE(c, "ctx");
    MM_REQUIRE(parts >= 1 && parts <= MM_MAX_PARTS, "parts out of range");
    c->parts = parts;
    return MM_OK;
}

int mm_ctx_get_parts(const mm_ctx* c) { return c ? c->parts : 0; }

int mm_ctx_set_timing(mm_ctx* c, int enable) {
    MM_REQUIRE(c, "ctx");
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

int mm_debug_export_faces(mm_ctx* c, int B, const void* workspace, float* fvi, float* fvz, float* fnz, void* stream)
{
    MM_REQUIRE(c && B > 0 && workspace, "ctx / B / workspace");
    const mm_ws_layout L = mm_ws_make(c, B);
    const char* ws = (const char*)workspace;
    mm_launch_export_faces(c, B, (const float*)(ws + L.frec), (const float*)(ws + L.vimg), fvi, fvz, fnz,
                           (cudaStream_t)stream);
    return check_launch("export_faces");
}

}  // extern "C"
