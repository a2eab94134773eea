/* oracle/ref_ctx.h -- TEST INFRASTRUCTURE ONLY (see ref_arpack.h). Internal context layout. */
#ifndef REF_CTX_H
#define REF_CTX_H
#include "ref_arpack.h"
struct ref_ctx {
  /* PARPACK semantics when par != 0 (communicator replaced by a callback) */
  int par, rank, nranks;
  ref_allreduce_fn ar;
  void* ar_user;
  /* COMMON /timing/ counters (stat.h:11) */
  int nopx, nbx, nrorth, nitref, nrstrt;
  /* SAVE'd locals of every routine, one block per precision, allocated lazily */
  void* dstate;
  void* sstate;
  void* zstate; /* complex paths (ref_impl_complex.inc) */
  void* cstate;
};
#endif
