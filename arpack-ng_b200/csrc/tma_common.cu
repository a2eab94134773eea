// tma_common.cu -- host side of tma_common.cuh: cached cuTensorMapEncodeTiled descriptors.
#include "tma_common.cuh"

#include <mutex>

namespace ab200 {
namespace tma {

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  });
  return fn;
}

struct Entry {
  const void* base;
  int64_t n, ldv;
  int ncols, esize, br, bc;
  CUtensorMap map;
};
std::vector<Entry>& cache() {
  static std::vector<Entry> c;
  return c;
}
std::mutex g_mu;
}  // namespace

bool tensor_maps_available() { return encode_fn() != nullptr; }

bool get_tensor_map(CUtensorMap* map, const void* base, int esize, int64_t n, int64_t ldv, int ncols, int box_rows,
                    int box_cols) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto& c = cache();
  for (const Entry& e : c) {
    if (e.base == base && e.n == n && e.ldv == ldv && e.ncols == ncols && e.esize == esize && e.br == box_rows &&
        e.bc == box_cols) {
      *map = e.map;
      return true;
    }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const CUtensorMapDataType dt = esize == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r;
  if (ncols > 0) {
    const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)ncols};
    const cuuint64_t gstride[1] = {(cuuint64_t)ldv * (cuuint64_t)esize};
    const cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)box_cols};
    const cuuint32_t estr[2] = {1, 1};
    r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t gdim[1] = {(cuuint64_t)n};
    const cuuint64_t gstride[1] = {0};
    const cuuint32_t box[1] = {(cuuint32_t)box_rows};
    const cuuint32_t estr[1] = {1};
    r = fn(map, dt, 1, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) return false;
  if (c.size() >= 1024) c.clear();
  c.push_back(Entry{base, n, ldv, ncols, esize, box_rows, box_cols, *map});
  return true;
}

}  // namespace tma
}  // namespace ab200
