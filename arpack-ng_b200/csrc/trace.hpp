// trace.hpp -- the reference's run-time tracing, COMMON /debug/ (debug.h:8-16) + UTIL/ivout.f / dvout.f, in small.
//
// debug_c(logfil, ndigit, mgetv0, msaupd, ... mceupd) (ICB/debug_c.h) stores one verbosity level per routine; a
// routine prints through ivout/dvout when its level exceeds a threshold (e.g. dsaupd.f:632-680 at msaupd > 0,
// dsaup2.f:408-412,498-508 at msaup2 > 0, dsaitr.f:398-403 at msaitr > 0, dgetv0.f:398-401 at mgetv0 > 0).  The
// UTIL printers themselves are out of scope as a port; this is the tiny printer that honours
// logfil/ndigit and the per-routine levels for the messages that describe the PATH of a solve: iteration starts, the
// NEV/NP/converged triple, restarts, the start-vector norm, and the exit summary with the COMMON /timing/ counters.
// The line formats are those of ivout.f:2000,1000-1003 and dvout.f:9999,9995-9998 (E exponent letter written as D).
// logfil = 6 is stdout (Fortran unit 6); any other unit goes to stderr.  PARPACK prints on rank 0 only
// (PARPACK/UTIL/MPI/pivout.f, pdvout.f:53) -- callers pass rank.
#pragma once
#include <complex>
#include <cstdio>
#include <cstring>

namespace ab200 {

struct TraceLevels {
  int logfil = 6, ndigit = -3, mgetv0 = 0;
  int msaupd = 0, msaup2 = 0, msaitr = 0, mseigt = 0, msapps = 0, msgets = 0, mseupd = 0;
  int mnaupd = 0, mnaup2 = 0, mnaitr = 0, mneigh = 0, mnapps = 0, mngets = 0, mneupd = 0;
  int mcaupd = 0, mcaup2 = 0, mcaitr = 0, mceigh = 0, mcapps = 0, mcgets = 0, mceupd = 0;
};
inline TraceLevels& trace_levels() {
  static TraceLevels t;  // one per loaded library, like the COMMON block
  return t;
}

namespace trace {

inline FILE* unit() { return trace_levels().logfil == 6 ? stdout : stderr; }

inline void title(const char* t) {  // FORMAT ( /1X, A /1X, A )
  const int len = (int)std::strlen(t) < 80 ? (int)std::strlen(t) : 80;
  char dash[81];
  std::memset(dash, '-', (size_t)len);
  dash[len] = 0;
  std::fprintf(unit(), "\n %s\n %s\n", t, dash);
}
// ivout.f:26-81
inline void ivout(int n, const int* ix, const char* t) {
  title(t);
  if (n <= 0) return;
  const int idigit = trace_levels().ndigit;
  int nd = idigit == 0 ? 4 : (idigit < 0 ? -idigit : idigit);
  int per, width;
  if (nd <= 4) { per = idigit < 0 ? 10 : 20; width = 5; }
  else if (nd <= 6) { per = idigit < 0 ? 7 : 15; width = 7; }
  else if (nd <= 10) { per = idigit < 0 ? 5 : 10; width = 11; }
  else { per = idigit < 0 ? 3 : 7; width = 15; }
  for (int k1 = 1; k1 <= n; k1 += per) {
    const int k2 = (n < k1 + per - 1) ? n : k1 + per - 1;
    std::fprintf(unit(), " %4d - %4d:", k1, k2);
    for (int i = k1; i <= k2; ++i) std::fprintf(unit(), " %*d", width, ix[i - 1]);
    std::fprintf(unit(), "\n");
  }
  std::fprintf(unit(), "  \n");
  std::fflush(unit());
}
inline void ivout1(int v, const char* t) { ivout(1, &v, t); }
// dvout.f:52-121 (1P, nD w.d)
template <typename R>
inline void dvout(int n, const R* x, const char* t) {
  title(t);
  if (n <= 0) return;
  const int idigit = trace_levels().ndigit;
  int nd = idigit == 0 ? 4 : (idigit < 0 ? -idigit : idigit);
  int per, width, prec;
  if (nd <= 4) { per = idigit < 0 ? 5 : 10; width = 12; prec = 3; }
  else if (nd <= 6) { per = idigit < 0 ? 4 : 8; width = 14; prec = 5; }
  else if (nd <= 10) { per = idigit < 0 ? 3 : 6; width = 18; prec = 9; }
  else { per = idigit < 0 ? 2 : 5; width = 24; prec = 13; }
  for (int k1 = 1; k1 <= n; k1 += per) {
    const int k2 = (n < k1 + per - 1) ? n : k1 + per - 1;
    std::fprintf(unit(), " %4d - %4d:%s", k1, k2, nd <= 4 ? "" : " ");
    for (int i = k1; i <= k2; ++i) {
      char buf[64];
      std::snprintf(buf, sizeof(buf), "%*.*E", width, prec, (double)x[i - 1]);
      for (char* p = buf; *p; ++p)
        if (*p == 'E') *p = 'D';
      std::fputs(buf, unit());
    }
    std::fprintf(unit(), "\n");
  }
  std::fprintf(unit(), "  \n");
  std::fflush(unit());
}
template <typename R>
inline void dvout1(R v, const char* t) { dvout(1, &v, t); }
// complex vectors (zvout.f): printed as (re, im) pairs, two reals per entry
template <typename R>
inline void zvout(int n, const std::complex<R>* x, const char* t) {
  title(t);
  const int idigit = trace_levels().ndigit;
  const int nd = idigit == 0 ? 4 : (idigit < 0 ? -idigit : idigit);
  const int prec = nd <= 4 ? 3 : (nd <= 6 ? 5 : (nd <= 10 ? 9 : 13));
  const int per = nd <= 4 ? 2 : 1;
  for (int k1 = 1; k1 <= n; k1 += per) {
    const int k2 = (n < k1 + per - 1) ? n : k1 + per - 1;
    std::fprintf(unit(), " %4d - %4d:", k1, k2);
    for (int i = k1; i <= k2; ++i) {
      char buf[96];
      std::snprintf(buf, sizeof(buf), "  (%*.*E,%*.*E)", prec + 7, prec, (double)x[i - 1].real(), prec + 7, prec,
                    (double)x[i - 1].imag());
      for (char* p = buf; *p; ++p)
        if (*p == 'E') *p = 'D';
      std::fputs(buf, unit());
    }
    std::fprintf(unit(), "\n");
  }
  std::fprintf(unit(), "  \n");
  std::fflush(unit());
}

// the exit banner of [ds]saupd / [ds]naupd / [cz]naupd (dsaupd.f:650-680): counters live, timers 0 as in every
// arpack-ng build (UTIL/second_NONE.f:29-31)
inline void summary(const char* what, int mxiter, int nopx, int nbx, int nrorth, int nitref, int nrstrt) {
  FILE* f = unit();
  std::fprintf(f, "\n\n     =============================================\n     = %-41s =\n"
                  "     = %-41s =\n"
                  "     =============================================\n"
                  "     = Summary of timing statistics              =\n"
                  "     =============================================\n\n\n", what, "Version: arpack_b200 0.1 (sm_100a)");
  std::fprintf(f, "     Total number update iterations             = %5d\n", mxiter);
  std::fprintf(f, "     Total number of OP*x operations            = %5d\n", nopx);
  std::fprintf(f, "     Total number of B*x operations             = %5d\n", nbx);
  std::fprintf(f, "     Total number of reorthogonalization steps  = %5d\n", nrorth);
  std::fprintf(f, "     Total number of iterative refinement steps = %5d\n", nitref);
  std::fprintf(f, "     Total number of restart steps              = %5d\n", nrstrt);
  static const char* timers[] = {"user OP*x operation", "user B*x operation", "Arnoldi update routine",
                                 "the update (aup2) routine", "basic Arnoldi iteration loop",
                                 "reorthogonalization phase", "(re)start vector generation",
                                 "projected eigen-subproblem", "getting the shifts", "applying the shifts",
                                 "convergence testing", "computing final Ritz vectors"};
  for (const char* tname : timers) std::fprintf(f, "     Total time in %-28s = %12.6f\n", tname, 0.0);
  std::fprintf(f, "\n");
  std::fflush(f);
}

}  // namespace trace
}  // namespace ab200
