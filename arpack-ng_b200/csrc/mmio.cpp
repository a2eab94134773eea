// mmio.cpp -- on-disk formats of the reference's tool layer, host side only.
//
//  * Matrix-Market style coordinate reader with the semantics of
//    EXAMPLES/MATRIX_MARKET/arpackSolver.hpp:361-416 (readMatrixMarket): '%' comment lines and blank lines are
//    skipped; the first data line is the header "n m [nnz]" (nnz optional); every other line is "i j value";
//    indices are taken as 1-based when the largest row index equals n or the largest column index equals m,
//    0-based otherwise (:405-413); a malformed line is an error.  The triplets then become a CSR matrix with
//    duplicates summed, as Eigen's setFromTriplets does in createMatrix (:417-424).
//  * The "--restart" dump files (arpackSolver.hpp:664-704, 770-772, 871-872): first line = element count, then
//    one value per line; on load a count mismatch is an error and |value| < 1e-6 may be replaced by machine
//    epsilon (restartSolve's allowZero = false, used for resid so that info = -9 cannot happen).
//
// Plain C++ (no CUDA): these run on any host; the arpackmm_b200 tool and the tests use them through the C-ABI.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/arpack_b200.h"

namespace {

template <typename V>
struct TripletsT {
  long long n = 0, m = 0;
  std::vector<long long> i, j;
  std::vector<V> v;
};
using Triplets = TripletsT<double>;

// returns 0, or 1 (cannot open), 2 (bad header), 3 (bad body line), 4 (index out of range).  V = double reads
// "i j value"; V = std::complex<double> reads "i j (re, im)" -- the stream extraction the reference's reader uses for
// its complex instantiation (arpackSolver.hpp:398-400), which also accepts "(re)" and a bare real.
template <typename V>
int read_triplets(const char* path, TripletsT<V>& t, long long* bad_line) {
  std::ifstream inp(path);
  if (!inp) return 1;
  std::string line;
  long long lineno = 0;
  bool have_header = false;
  while (std::getline(inp, line)) {
    ++lineno;
    size_t p = 0;
    while (p < line.size() && std::isspace((unsigned char)line[p])) ++p;
    if (p == line.size()) continue;   // empty line
    if (line[p] == '%') continue;     // comment (the %%MatrixMarket banner included)
    std::stringstream ss(line.substr(p));
    if (!have_header) {
      ss >> t.n >> t.m;
      if (!ss || t.n < 0 || t.m < 0) { if (bad_line) *bad_line = lineno; return 2; }
      long long nnz = 0;
      ss >> nnz;
      if (ss && nnz > 0) { t.i.reserve((size_t)nnz); t.j.reserve((size_t)nnz); t.v.reserve((size_t)nnz); }
      have_header = true;
    } else {
      long long k = 0, l = 0;
      V val = V(0);
      ss >> k >> l >> val;
      if (!ss) { if (bad_line) *bad_line = lineno; return 3; }
      t.i.push_back(k);
      t.j.push_back(l);
      t.v.push_back(val);
    }
  }
  if (!have_header) { if (bad_line) *bad_line = lineno; return 2; }
  if (!t.i.empty()) {
    const long long imax = *std::max_element(t.i.begin(), t.i.end());
    const long long jmax = *std::max_element(t.j.begin(), t.j.end());
    if (imax == t.n || jmax == t.m) {  // 1-based -> 0-based
      for (auto& x : t.i) x -= 1;
      for (auto& x : t.j) x -= 1;
    }
    for (size_t k = 0; k < t.i.size(); ++k)
      if (t.i[k] < 0 || t.i[k] >= t.n || t.j[k] < 0 || t.j[k] >= t.m) { if (bad_line) *bad_line = (long long)k; return 4; }
  }
  return 0;
}

template <typename V>
int mm_read_csr_t(const char* path, int* nrows, int* ncols, long long* nnz, int** rowptr_host, int** col_host,
                  V** val_host) {
  if (!path || !nrows || !ncols || !nnz || !rowptr_host || !col_host || !val_host) return -1;
  TripletsT<V> t;
  long long bad = 0;
  const int rc = read_triplets(path, t, &bad);
  if (rc != 0) {
    const char* what[] = {"", "can not open", "bad header (n, m)", "bad line", "index out of range"};
    std::fprintf(stderr, "arpack_b200: %s: %s (at %lld)\n", path, what[rc], bad);
    return rc;
  }
  if (t.n > 2147483647LL || t.m > 2147483647LL || (long long)t.i.size() > 2147483647LL) return 5;
  const size_t nz = t.i.size();
  // counting sort by row, then order each row by column and sum duplicates
  std::vector<long long> start((size_t)t.n + 1, 0);
  for (size_t k = 0; k < nz; ++k) start[(size_t)t.i[k] + 1]++;
  for (long long r = 0; r < t.n; ++r) start[(size_t)r + 1] += start[(size_t)r];
  std::vector<int> cj(nz);
  std::vector<V> cv(nz);
  {
    std::vector<long long> fill(start.begin(), start.end() - 1);
    for (size_t k = 0; k < nz; ++k) {
      const long long p = fill[(size_t)t.i[k]]++;
      cj[(size_t)p] = (int)t.j[k];
      cv[(size_t)p] = t.v[k];
    }
  }
  int* rp = (int*)std::malloc(sizeof(int) * ((size_t)t.n + 1));
  int* co = (int*)std::malloc(sizeof(int) * (nz ? nz : 1));
  V* va = (V*)std::malloc(sizeof(V) * (nz ? nz : 1));
  if (!rp || !co || !va) { std::free(rp); std::free(co); std::free(va); return 6; }
  long long out = 0;
  std::vector<std::pair<int, V>> row;
  for (long long r = 0; r < t.n; ++r) {
    rp[r] = (int)out;
    row.clear();
    for (long long p = start[(size_t)r]; p < start[(size_t)r + 1]; ++p) row.emplace_back(cj[(size_t)p], cv[(size_t)p]);
    std::stable_sort(row.begin(), row.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    for (size_t q = 0; q < row.size(); ++q) {
      if (q > 0 && row[q].first == co[out - 1]) {
        va[out - 1] += row[q].second;  // duplicates are summed (Eigen setFromTriplets)
      } else {
        co[out] = row[q].first;
        va[out] = row[q].second;
        ++out;
      }
    }
  }
  rp[t.n] = (int)out;
  *nrows = (int)t.n;
  *ncols = (int)t.m;
  *nnz = out;
  *rowptr_host = rp;
  *col_host = co;
  *val_host = va;
  return 0;
}
}  // namespace

extern "C" {

int ab200_mm_read_csr(const char* path, int* nrows, int* ncols, long long* nnz, int** rowptr_host, int** col_host,
                      double** val_host) {
  return mm_read_csr_t<double>(path, nrows, ncols, nnz, rowptr_host, col_host, val_host);
}
// complex coordinate files (EXAMPLES/MATRIX_MARKET/Az.mtx, Bz.mtx: "i j (re, im)"); val_host is interleaved (re, im)
int ab200_mm_read_csr_z(const char* path, int* nrows, int* ncols, long long* nnz, int** rowptr_host, int** col_host,
                        double** val_host) {
  std::complex<double>* zv = nullptr;
  const int rc = mm_read_csr_t<std::complex<double>>(path, nrows, ncols, nnz, rowptr_host, col_host, &zv);
  if (val_host) *val_host = reinterpret_cast<double*>(zv);
  return rc;
}

void ab200_mm_free(void* p) { std::free(p); }

int ab200_restart_save_f64(const char* path, long long count, const double* values) {
  std::ofstream ofs(path, std::ofstream::trunc);
  if (!ofs.is_open()) return 1;
  ofs.precision(17);
  ofs << count << "\n";
  for (long long k = 0; values && k < count; ++k) ofs << values[k] << "\n";
  return ofs.good() ? 0 : 2;
}

int ab200_restart_load_f64(const char* path, long long count, double* values, int allow_zero) {
  std::ifstream ifs(path);
  if (!ifs.is_open()) return 1;
  long long have = 0;
  ifs >> have;
  if (!ifs || have != count) {
    std::fprintf(stderr, "arpack_b200: %s: bad dim - restart KO\n", path);
    return 2;
  }
  const double eps = std::numeric_limits<double>::epsilon();
  for (long long k = 0; values && k < count; ++k) {
    double v = 0.0;
    ifs >> v;
    if (!ifs) return 3;
    if (std::fabs(v) < 1.e-6 && !allow_zero) v = eps;  // never hand dsaupd a zero residual (info = -9)
    values[k] = v;
  }
  return 0;
}

}  // extern "C"
