// Host-only check of the pieces of arpackmm_b200's ILU preconditioner that run on the CPU: the dual-threshold
// factorisation (ilut) and the level sets its triangular-solve kernel walks (level_sets).  The tool's translation unit
// is included as is (its main renamed); nothing here touches a GPU.
#define main arpackmm_b200_main
#include "../../arpack-ng_b200/csrc/arpackmm_b200.cu"
#undef main
#include <random>

template <typename H>
int check(int n, double droptol, int fillfactor, bool expect_exact) {
  // banded, diagonally dominant, non-symmetric, with holes in the band
  std::mt19937 g(5);
  std::uniform_real_distribution<double> u(-1, 1);
  Csr<H> A;
  A.n = A.m = n;
  A.rowptr.assign(n + 1, 0);
  for (int r = 0; r < n; ++r) {
    for (int c = std::max(0, r - 3); c <= std::min(n - 1, r + 3); ++c) {
      if (c != r && ((r * 7 + c * 3) % 4 == 0)) continue;
      A.col.push_back(c);
      H v = make_host<H>(u(g), std::is_same<H, double>::value ? 0.0 : u(g));
      if (c == r) v += H(8.0);
      A.val.push_back(v);
    }
    A.rowptr[r + 1] = (int)A.col.size();
  }
  auto f = ilut(A, droptol, fillfactor);
  const long long fill = (long long)A.val.size() * fillfactor / n + 1;
  std::vector<H> L((size_t)n * n, H(0.0)), U((size_t)n * n, H(0.0)), D((size_t)n * n, H(0.0));
  for (int r = 0; r < n; ++r) {
    L[(size_t)r * n + r] = H(1.0);
    if (f.L.rowptr[r + 1] - f.L.rowptr[r] > fill || f.U.rowptr[r + 1] - f.U.rowptr[r] > fill) return 12;  // fill bound
    for (int p = f.L.rowptr[r]; p < f.L.rowptr[r + 1]; ++p) {
      if (f.L.col[p] >= r || (p > f.L.rowptr[r] && f.L.col[p] <= f.L.col[p - 1])) return 10;  // strictly lower, sorted
      L[(size_t)r * n + f.L.col[p]] = f.L.val[p];
    }
    U[(size_t)r * n + r] = H(1.0) / f.dinv[r];
    for (int p = f.U.rowptr[r]; p < f.U.rowptr[r + 1]; ++p) {
      if (f.U.col[p] <= r || (p > f.U.rowptr[r] && f.U.col[p] <= f.U.col[p - 1])) return 11;  // strictly upper, sorted
      U[(size_t)r * n + f.U.col[p]] = f.U.val[p];
    }
    for (int p = A.rowptr[r]; p < A.rowptr[r + 1]; ++p) D[(size_t)r * n + A.col[p]] = A.val[p];
  }
  double err = 0, nrm = 0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      H s = H(0.0);
      for (int k = 0; k < n; ++k) s += L[(size_t)i * n + k] * U[(size_t)k * n + j];
      err = std::max(err, (double)std::abs(s - D[(size_t)i * n + j]));
      nrm = std::max(nrm, (double)std::abs(D[(size_t)i * n + j]));
    }
  // the two triangular solves, level by level, in the order k_sptrsv_levels processes them
  std::vector<int> lp, lr, up, ur;
  level_sets(f.L, true, lp, lr);
  level_sets(f.U, false, up, ur);
  std::vector<H> b(n), x(n, H(0.0)), y(n, H(0.0));
  for (int i = 0; i < n; ++i) b[i] = make_host<H>(u(g), 0.0);
  std::vector<char> done(n, 0);
  for (size_t l = 0; l + 1 < lp.size(); ++l) {
    for (int k = lp[l]; k < lp[l + 1]; ++k) {
      const int r = lr[k];
      H s = b[r];
      for (int p = f.L.rowptr[r]; p < f.L.rowptr[r + 1]; ++p) {
        if (!done[f.L.col[p]]) return 20;  // a row may only read rows of EARLIER levels
        s -= f.L.val[p] * x[f.L.col[p]];
      }
      x[r] = s;
    }
    for (int k = lp[l]; k < lp[l + 1]; ++k) done[lr[k]] = 1;
  }
  std::fill(done.begin(), done.end(), 0);
  for (size_t l = 0; l + 1 < up.size(); ++l) {
    for (int k = up[l]; k < up[l + 1]; ++k) {
      const int r = ur[k];
      H s = x[r];
      for (int p = f.U.rowptr[r]; p < f.U.rowptr[r + 1]; ++p) {
        if (!done[f.U.col[p]]) return 21;
        s -= f.U.val[p] * y[f.U.col[p]];
      }
      y[r] = s * f.dinv[r];
    }
    for (int k = up[l]; k < up[l + 1]; ++k) done[ur[k]] = 1;
  }
  double res = 0;
  std::vector<H> uy(n, H(0.0));
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < n; ++k) uy[i] += U[(size_t)i * n + k] * y[k];
  for (int i = 0; i < n; ++i) {
    H s = H(0.0);
    for (int k = 0; k < n; ++k) s += L[(size_t)i * n + k] * uy[k];
    res = std::max(res, (double)std::abs(s - b[i]));
  }
  std::printf("n=%d droptol=%g fill=%d: |LU-A|=%.3e nnz(L)=%zu nnz(U)=%zu levels %zu/%zu solve residual %.3e\n", n, droptol,
              fillfactor, err, f.L.val.size(), f.U.val.size(), lp.size() - 1, up.size() - 1, res);
  if (expect_exact && err > 1e-12 * nrm) return 1;
  if (!expect_exact && droptol >= 1.0 && (f.L.val.size() != 0 || f.U.val.size() != 0)) return 3;  // everything dropped: Jacobi
  if (res > 1e-10) return 2;
  return 0;
}

int main() {
  int rc = 0;
  rc = rc ? rc : check<double>(40, 0.0, 50, true);      // nothing dropped: the exact LU of a banded matrix
  rc = rc ? rc : check<hcomplex>(40, 0.0, 50, true);
  rc = rc ? rc : check<double>(40, 1e-2, 2, false);
  rc = rc ? rc : check<double>(40, 1.0, 2, false);       // arpackmm's default for a bare "ILU"
  rc = rc ? rc : check<hcomplex>(33, 1e-3, 1, false);
  rc = rc ? rc : check<double>(57, 0.0, 1, false);       // smallest fill factor: the band still fits
  std::printf(rc == 0 ? "ILUT OK\n" : "ILUT FAILED %d\n", rc);
  return rc;
}
