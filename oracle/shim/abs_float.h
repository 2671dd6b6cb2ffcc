/* Force-included (gcc -include) when building activ_functions.c for the "serial" oracle
 * variant: makes every abs() in that file the float overload, i.e. the semantics of the
 * reference's CUDA kernels (fabsf) instead of C's integer abs (SURVEY.md 8c caveat viii).
 * TEST INFRASTRUCTURE ONLY. */
#include <stdlib.h>
#include <math.h>
#include <tgmath.h>
#define abs(x) fabsf(x)
