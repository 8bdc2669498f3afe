/*
 * dibr_oracle_impl.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the four Kaolin DIB-R CUDA kernels that the reference's
 * hot path reaches through `kaolin.render.mesh.dibr_rasterization`
 * (call site: /root/reference/networks.py:297-299).  Kaolin is a third-party
 * dependency (pinned v0.12.0 / 0.18.0 in /root/reference/INSTALL.md:31-34,45)
 * that is NOT vendored in /root/reference and is not installed here, so this
 * file restates its published algorithm from memory: **parity unpinned**
 * (see docs/DIBR_SPEC.md for every assumption).
 *
 * The file is a "template": it is included twice by dibr_oracle.c, once with
 * REAL=float (the oracle proper: same fp32 operation order the CUDA product
 * implements, compiled with -ffp-contract=off) and once with REAL=double
 * (used only to validate the analytic backward against finite differences).
 *
 * Plain per-pixel x per-face loops, exactly the structure of the Kaolin
 * kernels: no tiling, no binning, no early outs other than the ones Kaolin has.
 */

#ifndef REAL
#error "define REAL and FN() before including"
#endif

/* pixel centre in multiplier-scaled NDC (DIBR_SPEC A.1).  MMO_V_PIXEL_ORDER: the other plausible expression order. */
static inline REAL FN(px_x)(int ix, int W, REAL mult) {
    if (g_mmo_variant & MMO_V_PIXEL_ORDER) return ((REAL)(2 * ix + 1 - W) / (REAL)W) * mult;
    return (mult / (REAL)W) * (REAL)(2 * ix + 1 - W);
}
static inline REAL FN(px_y)(int iy, int H, REAL mult) {
    if (g_mmo_variant & MMO_V_PIXEL_ORDER) return ((REAL)(H - 2 * iy - 1) / (REAL)H) * mult;
    return (mult / (REAL)H) * (REAL)(H - 2 * iy - 1);
}

/* barycentric weights of (x0,y0) in triangle a,b,c.  Default: DIB-R's k1/k2/k3 form with k/(k3+eps) (DIBR_SPEC A.2).
 * MMO_V_EDGE_BARY: the edge-function form (w0, w1 from the sub-triangle areas opposite a and b, w2 = 1-w0-w1), which a newer
 * Kaolin may use; MMO_V_EPS_ZERO: no eps in the denominator. */
static inline void FN(bary)(REAL ax, REAL ay, REAL bx, REAL by, REAL cx, REAL cy, REAL x0, REAL y0, REAL eps,
                            REAL* w0, REAL* w1, REAL* w2)
{
    if (g_mmo_variant & MMO_V_EPS_ZERO) eps = 0;
    if (g_mmo_variant & MMO_V_EDGE_BARY) {
        const REAL det = (by - cy) * (ax - cx) + (cx - bx) * (ay - cy);
        *w0 = ((by - cy) * (x0 - cx) + (cx - bx) * (y0 - cy)) / (det + eps);
        *w1 = ((cy - ay) * (x0 - cx) + (ax - cx) * (y0 - cy)) / (det + eps);
        *w2 = (REAL)1 - *w0 - *w1;
        return;
    }
    const REAL m = bx - ax, p = by - ay;
    const REAL n = cx - ax, q = cy - ay;
    const REAL s = x0 - ax, t = y0 - ay;
    const REAL k1 = s * q - n * t;
    const REAL k2 = m * t - s * p;
    const REAL k3 = m * q - n * p;
    *w1 = k1 / (k3 + eps);
    *w2 = k2 / (k3 + eps);
    *w0 = (REAL)1 - *w1 - *w2;
}

/* ------------------------------------------------------------------------
 * Hard rasterisation forward  (Kaolin packed_rasterize_forward_cuda_kernel)
 *   fvz   [B,F,3]    camera-space z of the three corners
 *   fvi   [B,F,3,2]  image-plane xy (NDC, y up), NOT yet multiplied
 *   feat  [B,F,3,D]  per-corner features
 *   valid [B,F]      u8, face takes part iff != 0 (front-facing: normal_z >= 0)
 * out:
 *   face_idx [B,H,W] int64 (-1 = none), weights [B,H,W,3], interp [B,H,W,D]
 * ---------------------------------------------------------------------- */
void FN(mmo_rasterize_forward)(int B, int H, int W, int F, int D,
                               const REAL* fvz, const REAL* fvi, const REAL* feat,
                               const unsigned char* valid, REAL multiplier, REAL eps,
                               long long* face_idx, REAL* weights, REAL* interp)
{
    #pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int b = 0; b < B; ++b) {
        for (int iy = 0; iy < H; ++iy) {
            const REAL* z_b = fvz + (size_t)b * F * 3;
            const REAL* p_b = fvi + (size_t)b * F * 6;
            const REAL* f_b = feat + (size_t)b * F * 3 * D;
            const unsigned char* v_b = valid + (size_t)b * F;
            const REAL y0 = FN(px_y)(iy, H, multiplier);
            for (int ix = 0; ix < W; ++ix) {
                const REAL x0 = FN(px_x)(ix, W, multiplier);
                REAL best_z = -INFINITY;
                int best_f = -1;
                REAL bw0 = 0, bw1 = 0, bw2 = 0;
                for (int f = 0; f < F; ++f) {
                    if (!v_b[f]) continue;
                    const REAL ax = p_b[f * 6 + 0] * multiplier, ay = p_b[f * 6 + 1] * multiplier;
                    const REAL bx = p_b[f * 6 + 2] * multiplier, by = p_b[f * 6 + 3] * multiplier;
                    const REAL cx = p_b[f * 6 + 4] * multiplier, cy = p_b[f * 6 + 5] * multiplier;
                    const REAL xmin = fmin(fmin(ax, bx), cx), xmax = fmax(fmax(ax, bx), cx);
                    const REAL ymin = fmin(fmin(ay, by), cy), ymax = fmax(fmax(ay, by), cy);
                    if (g_mmo_variant & MMO_V_BBOX_CLOSED) {         /* closed instead of half-open bbox test */
                        if (x0 < xmin || x0 > xmax || y0 < ymin || y0 > ymax) continue;
                    } else if (x0 < xmin || x0 >= xmax || y0 < ymin || y0 >= ymax) continue;
                    REAL w0, w1, w2;
                    FN(bary)(ax, ay, bx, by, cx, cy, x0, y0, eps, &w0, &w1, &w2);
                    if (g_mmo_variant & MMO_V_INSIDE_STRICT) { if (w0 <= 0 || w1 <= 0 || w2 <= 0) continue; }
                    else if (w0 < 0 || w1 < 0 || w2 < 0) continue;
                    const REAL zz = w0 * z_b[f * 3 + 0] + w1 * z_b[f * 3 + 1] + w2 * z_b[f * 3 + 2];
                    if (g_mmo_variant & MMO_V_DEPTH_GE) { if (zz < best_z) continue; }     /* >=: LAST face wins exact ties */
                    else if (zz <= best_z) continue;     /* strict >: first face wins exact ties */
                    best_z = zz; best_f = f; bw0 = w0; bw1 = w1; bw2 = w2;
                }
                const size_t pix = ((size_t)b * H + iy) * W + ix;
                face_idx[pix] = best_f;
                weights[pix * 3 + 0] = bw0; weights[pix * 3 + 1] = bw1; weights[pix * 3 + 2] = bw2;
                for (int d = 0; d < D; ++d) {
                    REAL v = 0;
                    if (best_f >= 0) {
                        const REAL* c = f_b + (size_t)best_f * 3 * D;
                        v = bw0 * c[d] + bw1 * c[D + d] + bw2 * c[2 * D + d];
                    }
                    interp[pix * D + d] = v;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------
 * Hard rasterisation backward (Kaolin rasterize_backward_cuda_kernel)
 *   grad_interp [B,H,W,D] -> grad_fvi [B,F,3,2], grad_feat [B,F,3,D]
 *   (both ACCUMULATED into; caller zero-fills).  No gradient to z or to the
 *   visibility decision.
 * ---------------------------------------------------------------------- */
void FN(mmo_rasterize_backward)(int B, int H, int W, int F, int D,
                                const REAL* grad_interp, const long long* face_idx,
                                const REAL* weights, const REAL* fvi, const REAL* feat,
                                REAL multiplier, REAL eps,
                                REAL* grad_fvi, REAL* grad_feat)
{
    #pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const REAL* p_b = fvi + (size_t)b * F * 6;
        const REAL* f_b = feat + (size_t)b * F * 3 * D;
        REAL* gp_b = grad_fvi + (size_t)b * F * 6;
        REAL* gf_b = grad_feat + (size_t)b * F * 3 * D;
        for (int iy = 0; iy < H; ++iy) {
            const REAL y0 = FN(px_y)(iy, H, multiplier);
            for (int ix = 0; ix < W; ++ix) {
                const size_t pix = ((size_t)b * H + iy) * W + ix;
                const long long f = face_idx[pix];
                if (f < 0) continue;
                const REAL x0 = FN(px_x)(ix, W, multiplier);
                const REAL* g = grad_interp + pix * D;
                /* features: dL/dc_i = w_i * g */
                for (int i = 0; i < 3; ++i) {
                    const REAL w = weights[pix * 3 + i];
                    for (int d = 0; d < D; ++d) gf_b[(f * 3 + i) * D + d] += g[d] * w;
                }
                const REAL ax = p_b[f * 6 + 0] * multiplier, ay = p_b[f * 6 + 1] * multiplier;
                const REAL bx = p_b[f * 6 + 2] * multiplier, by = p_b[f * 6 + 3] * multiplier;
                const REAL cx = p_b[f * 6 + 4] * multiplier, cy = p_b[f * 6 + 5] * multiplier;
                const REAL m = bx - ax, p = by - ay;
                const REAL n = cx - ax, q = cy - ay;
                const REAL s = x0 - ax, t = y0 - ay;
                const REAL k1 = s * q - n * t;
                const REAL k2 = m * t - s * p;
                const REAL k3 = m * q - n * p;
                /* d(k1)/d(m,n,p,q,s,t) etc. */
                const REAL dk1dm = 0, dk1dn = -t, dk1dp = 0, dk1dq = s, dk1ds = q, dk1dt = -n;
                const REAL dk2dm = t, dk2dn = 0, dk2dp = -s, dk2dq = 0, dk2ds = -p, dk2dt = m;
                const REAL dk3dm = q, dk3dn = -p, dk3dp = -n, dk3dq = m, dk3ds = 0, dk3dt = 0;
                /* numerators of d(w)/d(.) ; the common 1/k3^2 is applied below */
                const REAL dw1dm = dk1dm * k3 - dk3dm * k1, dw1dn = dk1dn * k3 - dk3dn * k1;
                const REAL dw1dp = dk1dp * k3 - dk3dp * k1, dw1dq = dk1dq * k3 - dk3dq * k1;
                const REAL dw1ds = dk1ds * k3 - dk3ds * k1, dw1dt = dk1dt * k3 - dk3dt * k1;
                const REAL dw2dm = dk2dm * k3 - dk3dm * k2, dw2dn = dk2dn * k3 - dk3dn * k2;
                const REAL dw2dp = dk2dp * k3 - dk3dp * k2, dw2dq = dk2dq * k3 - dk3dq * k2;
                const REAL dw2ds = dk2ds * k3 - dk3ds * k2, dw2dt = dk2dt * k3 - dk3dt * k2;
                const REAL dw1dax = -(dw1dm + dw1dn + dw1ds), dw1day = -(dw1dp + dw1dq + dw1dt);
                const REAL dw1dbx = dw1dm, dw1dby = dw1dp, dw1dcx = dw1dn, dw1dcy = dw1dq;
                const REAL dw2dax = -(dw2dm + dw2dn + dw2ds), dw2day = -(dw2dp + dw2dq + dw2dt);
                const REAL dw2dbx = dw2dm, dw2dby = dw2dp, dw2dcx = dw2dn, dw2dcy = dw2dq;
                const REAL* c = f_b + (size_t)f * 3 * D;
                for (int d = 0; d < D; ++d) {
                    const REAL c0 = c[d], c1 = c[D + d], c2 = c[2 * D + d];
                    const REAL dIdax = (c1 - c0) * dw1dax + (c2 - c0) * dw2dax;
                    const REAL dIday = (c1 - c0) * dw1day + (c2 - c0) * dw2day;
                    const REAL dIdbx = (c1 - c0) * dw1dbx + (c2 - c0) * dw2dbx;
                    const REAL dIdby = (c1 - c0) * dw1dby + (c2 - c0) * dw2dby;
                    const REAL dIdcx = (c1 - c0) * dw1dcx + (c2 - c0) * dw2dcx;
                    const REAL dIdcy = (c1 - c0) * dw1dcy + (c2 - c0) * dw2dcy;
                    const REAL dldI = multiplier * g[d] / (k3 * k3 + eps);
                    gp_b[f * 6 + 0] += dldI * dIdax;
                    gp_b[f * 6 + 1] += dldI * dIday;
                    gp_b[f * 6 + 2] += dldI * dIdbx;
                    gp_b[f * 6 + 3] += dldI * dIdby;
                    gp_b[f * 6 + 4] += dldI * dIdcx;
                    gp_b[f * 6 + 5] += dldI * dIdcy;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------
 * DIB-R soft silhouette forward (Kaolin dibr_soft_mask_forward_cuda_kernel)
 *   face_idx [B,H,W] from the hard pass; covered pixels get soft = 1.
 *   Uncovered pixels: faces in INDEX ORDER (all faces, no back-face test),
 *   enlarged half-open bbox test, min over 3 edge + 3 vertex squared
 *   distances, p = exp(-sigmainv * d2 / mult^2), first `knum` hits kept.
 * out: soft [B,H,W]; close_prob [B,H,W,knum]; close_idx [B,H,W,knum] int64
 *      (-1 = unused); close_type [B,H,W,knum] u8 (1..3 edge, 4..6 vertex, 0 unused)
 * ---------------------------------------------------------------------- */
void FN(mmo_soft_mask_forward)(int B, int H, int W, int F, int knum,
                               const REAL* fvi, const long long* face_idx,
                               REAL sigmainv, REAL boxlen, REAL multiplier,
                               REAL* soft, REAL* close_prob, long long* close_idx,
                               unsigned char* close_type)
{
    const REAL blen = boxlen * multiplier;
    #pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int b = 0; b < B; ++b) {
        for (int iy = 0; iy < H; ++iy) {
            const REAL* p_b = fvi + (size_t)b * F * 6;
            const REAL y0 = FN(px_y)(iy, H, multiplier);
            for (int ix = 0; ix < W; ++ix) {
                const size_t pix = ((size_t)b * H + iy) * W + ix;
                REAL* cp = close_prob + pix * knum;
                long long* ci = close_idx + pix * knum;
                unsigned char* ct = close_type + pix * knum;
                for (int k = 0; k < knum; ++k) { cp[k] = 0; ci[k] = -1; ct[k] = 0; }
                if (face_idx[pix] >= 0) { soft[pix] = (REAL)1; continue; }
                const REAL x0 = FN(px_x)(ix, W, multiplier);
                int kid = 0;
                for (int f = 0; f < F && kid < knum; ++f) {
                    REAL X[3], Y[3];
                    for (int i = 0; i < 3; ++i) {
                        X[i] = p_b[f * 6 + 2 * i] * multiplier;
                        Y[i] = p_b[f * 6 + 2 * i + 1] * multiplier;
                    }
                    const REAL xmin = fmin(fmin(X[0], X[1]), X[2]) - blen;
                    const REAL xmax = fmax(fmax(X[0], X[1]), X[2]) + blen;
                    const REAL ymin = fmin(fmin(Y[0], Y[1]), Y[2]) - blen;
                    const REAL ymax = fmax(fmax(Y[0], Y[1]), Y[2]) + blen;
                    if (g_mmo_variant & MMO_V_SOFT_BBOX_CLOSED) {
                        if (x0 < xmin || x0 > xmax || y0 < ymin || y0 > ymax) continue;
                    } else if (x0 < xmin || x0 >= xmax || y0 < ymin || y0 >= ymax) continue;
                    REAL pdis[6];
                    for (int i = 0; i < 3; ++i) {
                        const REAL x1 = X[i], y1 = Y[i];
                        const REAL x2 = X[(i + 1) % 3], y2 = Y[(i + 1) % 3];
                        const REAL A = y2 - y1, Bc = x1 - x2, C = x2 * y1 - x1 * y2;
                        const REAL up = A * x0 + Bc * y0 + C;
                        const REAL down = A * A + Bc * Bc;
                        REAL x3 = Bc * Bc * x0 - A * Bc * y0 - A * C;
                        REAL y3 = A * A * y0 - A * Bc * x0 - Bc * C;
                        x3 = x3 / (down + (REAL)1e-10);
                        y3 = y3 / (down + (REAL)1e-10);
                        const REAL direct = (x3 - x1) * (x3 - x2) + (y3 - y1) * (y3 - y2);
                        if (direct > 0) pdis[i] = (REAL)4 * multiplier * multiplier;   /* foot outside the segment */
                        else            pdis[i] = up * up / (down + (REAL)1e-10);
                    }
                    for (int i = 0; i < 3; ++i)
                        pdis[i + 3] = (x0 - X[i]) * (x0 - X[i]) + (y0 - Y[i]) * (y0 - Y[i]);
                    int edgeid = 0;
                    REAL d2 = pdis[0];
                    for (int i = 1; i < 6; ++i) if (d2 > pdis[i]) { d2 = pdis[i]; edgeid = i; }
                    const REAL z = sigmainv * d2 / multiplier / multiplier;
                    const REAL prob = EXPFN(-z);
                    cp[kid] = prob; ci[kid] = f; ct[kid] = (unsigned char)(edgeid + 1);
                    ++kid;
                }
                REAL allprob = (REAL)1;
                for (int k = 0; k < kid; ++k) allprob *= ((REAL)1 - cp[k]);
                soft[pix] = (REAL)1 - allprob;
            }
        }
    }
}

/* ------------------------------------------------------------------------
 * DIB-R soft silhouette backward (Kaolin dibr_soft_mask_backward_cuda_kernel)
 *   grad_soft [B,H,W] -> grad_fvi [B,F,3,2] (ACCUMULATED into)
 * ---------------------------------------------------------------------- */
void FN(mmo_soft_mask_backward)(int B, int H, int W, int F, int knum,
                                const REAL* grad_soft, const REAL* soft,
                                const long long* face_idx, const REAL* close_prob,
                                const long long* close_idx, const unsigned char* close_type,
                                const REAL* fvi, REAL sigmainv, REAL multiplier,
                                REAL* grad_fvi)
{
    #pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const REAL* p_b = fvi + (size_t)b * F * 6;
        REAL* gp_b = grad_fvi + (size_t)b * F * 6;
        for (int iy = 0; iy < H; ++iy) {
            const REAL y0 = FN(px_y)(iy, H, multiplier);
            for (int ix = 0; ix < W; ++ix) {
                const size_t pix = ((size_t)b * H + iy) * W + ix;
                if (face_idx[pix] >= 0) continue;
                const REAL x0 = FN(px_x)(ix, W, multiplier);
                const REAL dLdp = grad_soft[pix];
                const REAL allprob = soft[pix];
                for (int k = 0; k < knum; ++k) {
                    const long long f = close_idx[pix * knum + k];
                    if (f < 0) break;
                    const REAL prob = close_prob[pix * knum + k];
                    const REAL dLdz = (REAL)-1.0 * sigmainv * dLdp * ((REAL)1 - allprob) /
                                      ((REAL)1 - prob + (REAL)1e-6) * prob;
                    const int edgeid = (int)close_type[pix * knum + k] - 1;
                    if (edgeid >= 3) {
                        const int i = edgeid - 3;
                        const REAL x1 = p_b[f * 6 + 2 * i] * multiplier;
                        const REAL y1 = p_b[f * 6 + 2 * i + 1] * multiplier;
                        gp_b[f * 6 + 2 * i]     += dLdz * 2 * (x1 - x0) / multiplier;
                        gp_b[f * 6 + 2 * i + 1] += dLdz * 2 * (y1 - y0) / multiplier;
                    } else {
                        const int i = edgeid, j = (edgeid + 1) % 3;
                        const REAL x1 = p_b[f * 6 + 2 * i] * multiplier, y1 = p_b[f * 6 + 2 * i + 1] * multiplier;
                        const REAL x2 = p_b[f * 6 + 2 * j] * multiplier, y2 = p_b[f * 6 + 2 * j + 1] * multiplier;
                        const REAL A = y2 - y1, Bc = x1 - x2, C = x2 * y1 - x1 * y2;
                        const REAL up = A * x0 + Bc * y0 + C;
                        const REAL down = A * A + Bc * Bc;
                        const REAL d2 = up * up / (down + (REAL)1e-10);
                        const REAL dzdA = 2 * (x0 * up - d2 * A) / (down + (REAL)1e-10);
                        const REAL dzdB = 2 * (y0 * up - d2 * Bc) / (down + (REAL)1e-10);
                        const REAL dzdC = 2 * up / (down + (REAL)1e-10);
                        const REAL dLdx1 = dLdz * (dzdB - y2 * dzdC);
                        const REAL dLdy1 = dLdz * (x2 * dzdC - dzdA);
                        const REAL dLdx2 = dLdz * (y1 * dzdC - dzdB);
                        const REAL dLdy2 = dLdz * (dzdA - x1 * dzdC);
                        gp_b[f * 6 + 2 * i]     += dLdx1 / multiplier;
                        gp_b[f * 6 + 2 * i + 1] += dLdy1 / multiplier;
                        gp_b[f * 6 + 2 * j]     += dLdx2 / multiplier;
                        gp_b[f * 6 + 2 * j + 1] += dLdy2 / multiplier;
                    }
                }
            }
        }
    }
}
