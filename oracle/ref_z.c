/* oracle/ref_z.c -- TEST INFRASTRUCTURE ONLY (see ref_arpack.h). Double-complex instantiation (znaupd/zneupd). */
#define REF_COMPLEX_IMPL
#include "ref_ctx.h"
#define R double
#define CX double _Complex
#define ZF(x) ref_z##x
#define ZB(x) scipy_z##x##_
#define ZDB(x) scipy_zd##x##_
#define NRM2 scipy_dznrm2_
#define RB(x) scipy_d##x##_
#define IS_DOUBLE 1
#include "ref_impl_complex.inc"
