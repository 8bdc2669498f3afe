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
