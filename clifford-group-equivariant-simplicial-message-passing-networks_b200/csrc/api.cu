// Host-only entry points: version, status strings, algebra tables.
#include "common.cuh"

namespace csmpn {
unsigned long long& launch_counter() {
  static unsigned long long n = 0;
  return n;
}
}  // namespace csmpn

using namespace csmpn;

extern "C" {

int csmpn_version(void) { return 100; }

const char* csmpn_status_string(int status) {
  switch (status) {
    case CSMPN_OK: return "ok";
    case CSMPN_ERR_BAD_DIM: return "algebra dimension outside 1..5";
    case CSMPN_ERR_BAD_ARG: return "bad argument (null pointer, negative or inconsistent size)";
    case CSMPN_ERR_UNSUPPORTED: return "configuration not supported by this build";
    case CSMPN_ERR_CUDA: return "CUDA runtime error";
    case CSMPN_ERR_WORKSPACE: return "workspace missing or too small";
    default: return "unknown status";
  }
}

const char* csmpn_last_cuda_error(void) { return last_error_buf(); }

int csmpn_sm_count(void) { return sm_count_cached(); }

int64_t csmpn_launch_count(void) { return (int64_t)launch_counter(); }

// metric.py:18-120 / cliffordalgebra.py:27-42,238-252 restated with bit arithmetic.
int csmpn_algebra_tables(int dim, const float* metric, int32_t* out_idx, float* coef, int32_t* grades, uint8_t* paths) {
  if (dim < 1 || dim > 5) return CSMPN_ERR_BAD_DIM;
  if (!metric) return CSMPN_ERR_BAD_ARG;
  const int B = 1 << dim, G = dim + 1;
  int bitmap[32], index_of[32], grade[32];
  host_blade_bitmaps(dim, bitmap);
  for (int i = 0; i < B; ++i) {
    index_of[bitmap[i]] = i;
    grade[i] = __builtin_popcount(bitmap[i]);
    if (grades) grades[i] = grade[i];
  }
  if (paths) memset(paths, 0, (size_t)G * G * G);
  for (int i = 0; i < B; ++i)
    for (int k = 0; k < B; ++k) {
      const int a = bitmap[i], b = bitmap[k];
      int swaps = 0;
      for (int j = 0; j < dim; ++j)
        if (b >> j & 1) swaps += __builtin_popcount(a >> (j + 1));
      float c = (swaps & 1) ? -1.f : 1.f;
      for (int v = 0; v < dim; ++v)
        if ((a & b) >> v & 1) c *= metric[v];
      const int j = index_of[a ^ b];
      if (out_idx) out_idx[i * B + k] = j;
      if (coef) coef[i * B + k] = c;
      if (paths && c != 0.f) paths[(grade[i] * G + grade[j]) * G + grade[k]] = 1;
    }
  return CSMPN_OK;
}

}  // extern "C"
