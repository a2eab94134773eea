/* oracle/ref_d.c -- TEST INFRASTRUCTURE ONLY (see ref_arpack.h). Double-precision instantiation. */
#include "ref_ctx.h"
#define R double
#define FN(x) ref_d##x
#define BL(x) scipy_d##x##_
#define IS_DOUBLE 1
#include "ref_impl_common.inc"
#include "ref_impl_sym.inc"
#include "ref_impl_nonsym.inc"
