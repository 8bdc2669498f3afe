/*
 * magicmirror.h -- C ABI of libmagicmirror.so, the sm_100a differentiable
 * render-and-compare hot path of layumi/3D-Magic-Mirror.
 *
 * The reference has no FFI of its own for this path: it reaches NVIDIA Kaolin's
 * torch C++ extension from Python.  Each entry point below names the reference
 * call(s) it replaces (paths relative to /root/reference).
 *
 * Conventions
 *   - every tensor argument is a DEVICE pointer to a contiguous fp32 array unless
 *     stated otherwise; the caller owns all buffers, including `workspace`.
 *   - the library owns only `mm_ctx` (mesh topology, per-face UVs, raster params).
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *     never synchronises and never allocates device memory: every call can be
 *     captured into a CUDA graph.
 *   - return value 0 = ok, < 0 = error (MM_E_*); mm_last_error() returns a
 *     thread-local message for the last failing call.
 *   - a ctx is bound to one device, which must be the CURRENT device of the calling
 *     thread (checked); calls on one ctx may come from any thread; two calls that
 *     share a workspace must be stream-ordered.  The ctx holds no per-call state.
 *   - `workspace` / `workspace_bytes`: caller-owned scratch of at least
 *     mm_workspace_bytes(ctx, B) bytes, 256-byte aligned (checked).  A workspace
 *     written by a forward call must be passed unmodified to the matching backward
 *     call (and to mm_recon_data_forward for lazy fusion, below).
 *   - gradient outputs are OVERWRITTEN (the library zero-fills), never accumulated.
 *   - image-sized planes (rgba, gt, bg, g_rgba, g_bg, imnormal, face_idx) may have any
 *     alignment; 16-byte aligned planes with W % 4 == 0 take the vectorised kernels.
 *   - `tex_mirror`: the atlas the renderer reads is produced as
 *     cat([t, t.flip(2)], dim=2) (TextureEncoder.forward, network/model_res.py:609-610):
 *     its lower half is the upper half upside down.  With tex_mirror = 1, `tex` / `g_tex`
 *     are the UPPER HALF alone, [B,3,Ht/2,Wt], while `Ht` stays the logical atlas height
 *     (even): logical row r >= Ht/2 is read from (and its gradient accumulated into)
 *     physical row Ht-1-r.  Same texels, same weights: the image is bit-identical, g_tex
 *     equals the sum of the two halves' gradients, half the texture bytes move.
 *
 * Environment switches read at mm_ctx_create (diagnostics; defaults are the measured best):
 *   MM_PDL=0        no programmatic dependent launch between the library's kernels
 *   MM_PDL_LATE=m   bit k set: raster kernel k releases its dependent launch at CTA exit instead of at its first instruction
 *                   (0 hard, 1 soft forward, 2 overflow, 3 shading, 4 soft backward; default 15)
 *   MM_VCHUNKS=n    CTAs per image of the vertex forward kernel (default 8)
 *   MM_PLIST_CAP=n  test hook: caps the forward's candidate list so the backward's fallback path runs
 */
#ifndef MAGICMIRROR_H_
#define MAGICMIRROR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MM_ABI_VERSION 3     /* 3: mm_render_backward no longer reads `rgba`; mm_debug_workspace_offset; larger workspace */

#define MM_OK            0
#define MM_E_INVALID    -1   /* bad argument (NULL pointer, non-positive size, undersized workspace, wrong device, ...) */
#define MM_E_CUDA       -2   /* a CUDA runtime call failed; message has the cudaError string */
#define MM_E_UNSUPPORTED -3  /* configuration outside what the kernels support */

typedef struct mm_ctx mm_ctx;

int         mm_abi_version(void);
const char* mm_last_error(void);

/* Replaces DiffRender.__init__'s device-side state (networks.py:165-256: faces,
 * face_uvs, cam_proj = generate_perspective_projection(fovy, 1/ratio) -> proj_x =
 * 2.5*ratio, proj_y = 2.5) and the defaults of kaolin dibr_rasterization
 * (sigmainv=7000, boxlen=0.02, knum=30, multiplier=1000, eps=1e-8; call site
 * networks.py:297-299).  faces_host / face_uvs_host are HOST pointers. */
int mm_ctx_create(mm_ctx** out, int device, int V, int F,
                  const int32_t* faces_host /* F*3 */, const float* face_uvs_host /* F*3*2 */,
                  int H, int W, float proj_x, float proj_y,
                  float sigmainv, float boxlen, int knum, float multiplier, float eps);
int mm_ctx_destroy(mm_ctx* ctx);

/* Introspection (tests, bench): "fused_kernels" (kernel launches of one mm_render_compare_fwd_bwd for H, W multiples of 4),
 * "api_kernels" (render_forward + recon_data_forward + render_backward).  Returns -1 for an unknown key. */
int mm_ctx_get_int(const mm_ctx* ctx, const char* key);

/* Bytes of caller-owned scratch needed by any call on `ctx` with batch B. */
size_t mm_workspace_bytes(const mm_ctx* ctx, int B);

/* Replaces DiffRender.render (networks.py:258-324): camera_position_from_spherical_angles
 * + generate_transformation_matrix (smr_utils.py:257-311), kaolin prepare_vertices
 * (:284), face_normals (:289), dibr_rasterization (:297), texture_mapping (:305),
 * spherical_harmonic_lighting (:306), composite + clamp + pack (:307-317).
 *   rgba          [B,4,H,W]   out
 *   face_normals  [B,F,3]     out  (attributes['face_normals'], networks.py:319), may be NULL
 *   imnormal      [B,H,W,3]   out, may be NULL (attributes['imnormal'], :320)
 *   face_idx      [B,H,W]     out int32, may be NULL (-1 = no face) */
int mm_render_forward(mm_ctx* ctx, int B,
                      const float* vertices /* B,V,3 */, const float* azim /* B */,
                      const float* elev /* B */, const float* dist /* B */, const float* bias /* B,2 */,
                      const float* tex /* B,3,Ht,Wt */, int Ht, int Wt, int tex_mirror,
                      const float* lights /* B,9 */, const float* bg /* B,3,H,W or NULL */, int no_mask,
                      float* rgba, float* face_normals, float* imnormal, int32_t* face_idx,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Backward of mm_render_forward (replaces autograd through the same Kaolin calls:
 * rasterize_backward + dibr_soft_mask_backward + grid_sample backward + ...).
 * The upstream gradient of the image is the SUM of two optional parts:
 *   g_rgba          [B,4,H,W]  a materialised upstream gradient, or NULL
 *   recon_gt        [B,4,H,W]  or NULL.  LAZY FUSION of trainer.py:441 + :509: when DiffRender.recon_data was evaluated on
 *                   this render's output with mm_recon_data_forward ON THIS WORKSPACE, its backward does not materialise
 *                   d(loss)/d(rgba): it hands over (recon_gt, image_weight, contour, loss_scale, loss_scale_dev) and the
 *                   gradient loss_scale * [*loss_scale_dev] * d(recon_data)/d(rgba) is formed inside the shading kernel
 *                   (the per-image IoU sums are in the workspace).  loss_scale_dev: DEVICE scalar or NULL (= 1).
 *   g_face_normals  [B,F,3]    upstream gradient of the face_normals output, may be NULL
 * All g_* outputs are overwritten; g_bg may be NULL (and must be when bg is NULL). */
int mm_render_backward(mm_ctx* ctx, int B,
                       const float* vertices, const float* azim, const float* elev, const float* dist,
                       const float* bias, const float* tex, int Ht, int Wt, int tex_mirror, const float* lights,
                       const float* bg, int no_mask,
                       const float* rgba /* unused, may be NULL: the backward reads the silhouette from the workspace, so the
                                            caller is free to edit the image in place after the forward */,
                       const float* g_rgba, const float* g_face_normals,
                       const float* recon_gt, float image_weight, float contour, float loss_scale,
                       const float* loss_scale_dev,
                       float* g_vertices /* B,V,3 */, float* g_azim, float* g_elev, float* g_dist,
                       float* g_bias /* B,2 */, float* g_tex /* B,3,Ht,Wt */, float* g_lights /* B,9 */,
                       float* g_bg /* B,3,H,W or NULL */,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Replaces DiffRender.recon_data (networks.py:364-390) incl. kaolin mask_iou (:377).  One kernel (+ a 1 KB memset).
 *   loss      [4] out: data (= image_weight*image + mask + contour*contour_term), image, mask(1-IoU), contour_term
 *   iou_sums  [B,2] out: N_b = sum(p*g), D_b = sum(p+g-p*g); may be NULL
 * The per-image sums are left in `workspace` (see mm_render_backward's recon_gt). */
int mm_recon_data_forward(mm_ctx* ctx, int B, const float* pred /* B,4,H,W */, const float* gt /* B,4,H,W */,
                          float image_weight, float contour,
                          float* loss, float* iou_sums, void* workspace, size_t workspace_bytes, void* stream);

/* d(loss_scale * [*loss_scale_dev] * loss_data)/d(pred) -> g_pred [B,4,H,W] (overwritten), for a `pred` that is not a
 * render output of this library; needs the per-image sums mm_recon_data_forward left in `workspace`. */
int mm_recon_data_backward(mm_ctx* ctx, int B, const float* pred, const float* gt,
                           float image_weight, float contour, float loss_scale, const float* loss_scale_dev,
                           float* g_pred, void* workspace, size_t workspace_bytes, void* stream);

/* Fused path: render -> recon_data -> backward of (loss_scale*loss_data + <g_rgba_extra, rgba>) in one call; the loss
 * gradient is formed in-kernel and never materialised.  Replaces trainer.py:276 + :441 + the autograd walk at :509.
 *   gt            [B,4,H,W]
 *   g_rgba_extra  [B,4,H,W] or NULL  (upstream gradient from another consumer, e.g. the GAN)
 *   g_face_normals [B,F,3] or NULL   (upstream gradient of face_normals, e.g. from calc_reg_loss, networks.py:422)
 *   rgba          [B,4,H,W] out
 *   loss          [4] out, as mm_recon_data_forward */
int mm_render_compare_fwd_bwd(mm_ctx* ctx, int B,
                              const float* vertices, const float* azim, const float* elev, const float* dist,
                              const float* bias, const float* tex, int Ht, int Wt, int tex_mirror, const float* lights,
                              const float* bg, int no_mask,
                              const float* gt, float image_weight, float contour, float loss_scale,
                              const float* g_rgba_extra, const float* g_face_normals,
                              float* rgba, float* face_normals, float* loss,
                              float* g_vertices, float* g_azim, float* g_elev, float* g_dist, float* g_bias,
                              float* g_tex, float* g_lights, float* g_bg,
                              void* workspace, size_t workspace_bytes, void* stream);

/* ---- render de-duplication.  trainer.py:367 renders only to refresh attributes['face_normals'] (the image
 * is discarded): these two calls run the vertex stage alone (prepare_vertices + face_normals, networks.py:284-290) and its
 * backward from an upstream gradient of the face normals.  Same argument conventions as mm_render_forward / _backward. */
int mm_face_normals_forward(mm_ctx* ctx, int B, const float* vertices, const float* azim, const float* elev, const float* dist,
                            const float* bias, float* face_normals /* B,F,3 out */, void* workspace, size_t workspace_bytes,
                            void* stream);
int mm_face_normals_backward(mm_ctx* ctx, int B, const float* vertices, const float* azim, const float* elev, const float* dist,
                             const float* bias, const float* g_face_normals /* B,F,3 */, float* g_vertices, float* g_azim,
                             float* g_elev, float* g_dist, float* g_bias, void* workspace, size_t workspace_bytes, void* stream);

/* ---- the mesh regularisers next to the render path (networks.py:392-491), one launch per direction.
 * Topology they need beyond mm_ctx_create's (DiffRender.__init__, networks.py:197-252): edges [E,2], edge2faces [E,2]
 * (faces sharing each edge), flip_index [V] (z-mirror partner), sign_init [V] (sign of the template depth) and the uniform
 * Laplacian (kaolin uniform_laplacian, networks.py:249) in CSR form.  All HOST pointers; copied into the ctx. */
int mm_ctx_set_regularizer_topology(mm_ctx* ctx, int E, const int32_t* edges_host, const int32_t* edge2faces_host,
                                    const int32_t* flip_index_host, const float* sign_init_host, int nnz,
                                    const int32_t* lap_row_off_host /* V+1 */, const int32_t* lap_col_host,
                                    const float* lap_val_host, float ratio);

/* terms[8] (device, out) = laplacian, flat (the two summands of calc_reg_loss :412-451, before lambda_lpl / lambda_flat),
 * calc_reg_edge (:453), calc_reg_depth (:463), calc_reg_depthR (:468), calc_reg_depthC (:477), calc_reg_deform (:487),
 * recon_flip (:392; flip_l1 selects its L1 form).  term_mask bit k selects term k; an input a selected term does not need may
 * be NULL (delta_vertices [B,V,3]: terms 0,6,7; vertices [B,V,3]: 2-5; face_normals [B,F,3]: 1).
 * scratch: caller-owned device floats [B*8 + 8] (per-image partial sums, reduced over the batch in image order, + the
 * last-CTA ticket; per call, so concurrent calls on different streams do not share state). */
int mm_mesh_reg_forward(mm_ctx* ctx, int B, const float* delta_vertices, const float* vertices, const float* face_normals,
                        float temp, float eps, int flip_l1, unsigned term_mask, float* terms, float* scratch, void* stream);

/* Gradient of sum_k g_terms[k] * term_k (g_terms: device [8]) w.r.t. the three inputs; outputs overwritten; a NULL input
 * goes with a NULL output. */
int mm_mesh_reg_backward(mm_ctx* ctx, int B, const float* delta_vertices, const float* vertices, const float* face_normals,
                         float temp, float eps, int flip_l1, unsigned term_mask, const float* g_terms,
                         float* g_delta_vertices, float* g_vertices, float* g_face_normals, void* stream);

/* ---- SURVEY 8(f)-3, encoder side: ShapeEncoder.forward's template conditioning (network/model_res.py:317-325):
 *   local         = F.grid_sample(x, template[..., 0:2], 'bilinear', align_corners=True, padding_mode='zeros')   [N,V]
 *   neighbor_diff = torch.mm(local.view(-1, V), lpl)   with lpl = vertices_laplacian_matrix (trainer.py:91)       [N,V]
 * for the N = B*C feature planes x [N,h,w]; one kernel, the dense V x V GEMM replaced by the sparse Laplacian set with
 * mm_ctx_set_regularizer_topology.  template_xyz: device [V,3] (x, y used; the reference detaches them).  neighbor_diff may
 * be NULL.  Backward: g_x [N,h,w] (overwritten) from g_local and / or g_neighbor_diff (either may be NULL); h*w <= 2048. */
int mm_template_features_forward(mm_ctx* ctx, int N, int h, int w, const float* x, const float* template_xyz,
                                 float* local, float* neighbor_diff, void* stream);
int mm_template_features_backward(mm_ctx* ctx, int N, int h, int w, const float* template_xyz, const float* g_local,
                                  const float* g_neighbor_diff, float* g_x, void* stream);

/* ---- SURVEY 8(f)-3, texture side: TextureEncoder.forward's tail (network/model_res.py:598-611):
 *   textures = F.grid_sample(img, flow.permute(0,2,3,1), mode='bicubic', align_corners=True)      (padding_mode 'zeros')
 *   [textures = cat([textures, textures.flip(2)], dim=2)   when concat != 0; the reference's optional `makeup` network sits
 *    between the two lines, so the concat is a switch]
 * img [B,C,Hi,Wi], flow [B,2,Ho,Wo] (channel 0 = x, 1 = y in [-1,1], exactly the tensor the reference permutes);
 * out [B,C,Ho*(concat?2:1),Wo].  Backward: g_img (overwritten) and g_flow (overwritten) from g_out (the two halves of a
 * concatenated g_out are summed on the fly). */
int mm_texture_flow_forward(mm_ctx* ctx, int B, int C, int Hi, int Wi, int Ho, int Wo, int concat,
                            const float* img, const float* flow, float* out, void* stream);
int mm_texture_flow_backward(mm_ctx* ctx, int B, int C, int Hi, int Wi, int Ho, int Wo, int concat,
                             const float* img, const float* flow, const float* g_out, float* g_img, float* g_flow,
                             void* stream);

/* Test hook: copies the vertex-stage products of the last forward on `workspace`
 * (what kaolin prepare_vertices returns, networks.py:284-287) so that the oracle's
 * rasteriser can be run on bit-identical inputs.  Any pointer may be NULL.
 *   fvi [B,F,3,2] image-plane xy (unscaled), fvz [B,F,3] camera z, fnz [B,F] unit-normal z */
int mm_debug_export_faces(mm_ctx* ctx, int B, const void* workspace, size_t workspace_bytes,
                          float* fvi, float* fvz, float* fnz, void* stream);

/* Test / probe hook: byte offset of a named block of the workspace for batch size B ("frec", "frect", "zbuf", "lacc", "cov",
 * "ovf_count", "sched_n", "ovf_list", "sched_list", "plist", "gsoft", "gfacc", "img_fwd", "img_bwd"; layout: csrc/mm_common.cuh),
 * or (size_t)-1 for an unknown name.  Lets tests look at the counters and the shading schedule without restating the layout. */
size_t mm_debug_workspace_offset(const mm_ctx* ctx, int B, const char* block);

/* Measurement hook (bench.py): when enabled, mm_render_compare_fwd_bwd records a CUDA event on
 * `stream` around each of its launch groups.  mm_ctx_get_timing waits for the last call's final event
 * and writes the group durations in milliseconds, in launch order (vertex_fwd, geometry forward [hard + soft],
 * fused shading, d/d-silhouette pass [only for H or W not a multiple of 4], geometry backward,
 * vertex_bwd + loss finalisation, -; capacity >= 7); returns the count written.  Event records switch
 * programmatic dependent launch off across them, so the figures are slightly above the in-step cost.
 * (The one piece of per-call state in a ctx: do not enable it on a ctx shared between threads.) */
int mm_ctx_set_timing(mm_ctx* ctx, int enable);
int mm_ctx_get_timing(mm_ctx* ctx, float* ms_host, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* MAGICMIRROR_H_ */
