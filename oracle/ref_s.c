/* oracle/ref_s.c -- TEST INFRASTRUCTURE ONLY (see ref_arpack.h). Single-precision instantiation. */
#include "ref_ctx.h"
#define R float
#define FN(x) ref_s##x
#define BL(x) scipy_s##x##_
#define IS_DOUBLE 0
#include "ref_impl_common.inc"
#include "ref_impl_sym.inc"
#include "ref_impl_nonsym.inc"
