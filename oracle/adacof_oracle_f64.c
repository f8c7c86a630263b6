/* oracle/adacof_oracle_f64.c -- TEST INFRASTRUCTURE.  The fp64 instantiation of adacof_oracle.c (same expression trees,
 * double arithmetic): the high-precision arbiter for the per-stage error budget tests.  Not a restatement of anything the
 * reference runs -- the reference computes in fp32. */
#define REAL double
#define SUF(x) x##_f64
#include "adacof_oracle.c"
