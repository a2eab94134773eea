/* oracle/ref_c.c -- TEST INFRASTRUCTURE ONLY (see ref_arpack.h). Single-complex instantiation (cnaupd/cneupd). */
#define REF_COMPLEX_IMPL
#include "ref_ctx.h"
#define R float
#define CX float _Complex
#define ZF(x) ref_c##x
#define ZB(x) scipy_c##x##_
#define ZDB(x) scipy_cs##x##_
#define NRM2 scipy_scnrm2_
#define RB(x) scipy_s##x##_
#define IS_DOUBLE 0
#include "ref_impl_complex.inc"
