/*
 * dibr_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU oracle for the Kaolin DIB-R
 * kernels reached from /root/reference/networks.py:297-299.  parity unpinned
 * (Kaolin is un-vendored and not installable here; see docs/DIBR_SPEC.md).
 *
 * Build: `make -C oracle`  ->  oracle/_build/libdibr_oracle.so
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#include <math.h>
#include <stddef.h>

/* Variant switches, ORACLE ONLY (tools/dibr_sensitivity.py -> docs/DIBR_SENSITIVITY.md): each bit replaces one
 * medium-confidence assumption of docs/DIBR_SPEC.md by its plausible alternative, to measure how much of the output
 * would change if the recollection of Kaolin were wrong there.  0 = the spec the CUDA product implements. */
#define MMO_V_BBOX_CLOSED      1     /* hard pass: closed bbox test (<=) instead of half-open (<) on the max side */
#define MMO_V_SOFT_BBOX_CLOSED 2     /* soft pass: the same for the enlarged bbox */
#define MMO_V_DEPTH_GE         4     /* depth test >= (last face wins exact ties) instead of > */
#define MMO_V_PIXEL_ORDER      8     /* pixel centre ((2ix+1-W)/W)*mult instead of (mult/W)*(2ix+1-W) */
#define MMO_V_EDGE_BARY        16    /* edge-function barycentrics instead of the k1/k2/k3 form */
#define MMO_V_EPS_ZERO         32    /* no eps in the barycentric denominator */
#define MMO_V_INSIDE_STRICT    64    /* inside test w > 0 instead of w >= 0 */
static int g_mmo_variant = 0;
void mmo_set_variant(int v) { g_mmo_variant = v; }
int mmo_get_variant(void) { return g_mmo_variant; }

#define REAL float
#define FN(name) name##_f32
#define EXPFN expf
#include "dibr_oracle_impl.h"
#undef REAL
#undef FN
#undef EXPFN

#define REAL double
#define FN(name) name##_f64
#define EXPFN exp
#include "dibr_oracle_impl.h"
#undef REAL
#undef FN
#undef EXPFN

int mmo_abi_version(void) { return 1; }
