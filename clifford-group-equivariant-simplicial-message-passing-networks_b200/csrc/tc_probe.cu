// Single-tile tcgen05 probe: pins the descriptor encodings of tc_common.cuh on real hardware (tests/test_tc_probe.py).
// One CTA stages two small matrices into shared memory in the plane layout, issues kind::tf32 MMAs and dumps the
// accumulator lane by lane, so a test can check operand majors, LBO/SBO roles, the M=64 lane mapping and the accuracy
// of the error-compensated split.
#include "tc_common.cuh"
#include "csmpn_debug.h"

namespace csmpn {
using namespace tc;

// mode 0: D[m][n] = sum_k A[m][k]  * Bm[n][k]    A: [128 x K] K-major plane,   Bm: [N x K] K-major plane
// mode 1: D[m][n] = sum_k A[m][k]  * Bt[k][n]    A as above,                   Bt: [K x N] plane used MN-major
// mode 2: D[m][n] = sum_k At[k][m] * Bt[k][n]    At: [K x M] plane MN-major,   Bt as above        (M = 64 or 128)
// flags: bit0 swap LBO/SBO of A, bit1 swap LBO/SBO of B, bit2 three-MMA hi/lo split
__global__ void __launch_bounds__(128, 1) tc_probe_kernel(const float* __restrict__ A, const float* __restrict__ Bsrc,
                                                         float* __restrict__ dump, int mode, int M, int N, int K, int flags) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int a_rows = (mode == 2) ? K : 128, a_ch = (mode == 2) ? M : K;
  const int b_rows = (mode == 0) ? N : K, b_ch = (mode == 0) ? K : N;
  const uint32_t a_bytes = (uint32_t)a_rows * a_ch * 4, b_bytes = (uint32_t)b_rows * b_ch * 4;
  uint8_t* a_hi = smem;
  uint8_t* a_lo = a_hi + a_bytes;
  uint8_t* b_hi = a_lo + a_bytes;
  uint8_t* b_lo = b_hi + b_bytes;
  const bool split = flags & 4;
  for (int idx = tid; idx < a_rows * a_ch; idx += 128) {
    const int r = idx / a_ch, c = idx % a_ch;
    const float x = A[idx];
    const float hi = split ? tf32_hi(x) : x;
    *reinterpret_cast<float*>(a_hi + plane_off(a_rows, r, c)) = hi;
    *reinterpret_cast<float*>(a_lo + plane_off(a_rows, r, c)) = x - hi;
  }
  for (int idx = tid; idx < b_rows * b_ch; idx += 128) {
    const int r = idx / b_ch, c = idx % b_ch;
    const float x = Bsrc[idx];
    const float hi = split ? tf32_hi(x) : x;
    *reinterpret_cast<float*>(b_hi + plane_off(b_rows, r, c)) = hi;
    *reinterpret_cast<float*>(b_lo + plane_off(b_rows, r, c)) = x - hi;
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(M, N, mode == 2, mode != 0);
    const int ksteps = K / 8;
    uint32_t acc = 0;
    for (int ks = 0; ks < ksteps; ++ks) {
      for (int pass = 0; pass < (split ? 3 : 1); ++pass) {
        const uint8_t* ap = (pass == 2) ? a_lo : a_hi;
        const uint8_t* bp = (pass == 1) ? b_lo : b_hi;
        uint64_t ad = desc_kmajor(smem_addr(ap), a_rows, ks);
        uint64_t bd = desc_kmajor(smem_addr(bp), b_rows, ks);
        if (flags & 1) ad = (ad & ~0x3FFF3FFF0000ull) | (((ad >> 16) & 0x3FFF) << 32) | (((ad >> 32) & 0x3FFF) << 16);
        if (flags & 2) bd = (bd & ~0x3FFF3FFF0000ull) | (((bd >> 16) & 0x3FFF) << 32) | (((bd >> 32) & 0x3FFF) << 16);
        mma_tf32(tbase, ad, bd, idesc, acc);
        acc = 1;
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (int n = 0; n < N; ++n) {
    float v;
    tmem_ld1(tmem_at(tbase, warp * 32, n), v);
    tmem_wait_ld();
    dump[(warp * 32 + lane) * N + n] = v;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 64);
}

}  // namespace csmpn

using namespace csmpn;

extern "C" int csmpn_tc_probe(int mode, int M, int N, int K, int flags, const float* A, const float* Bsrc, float* dump,
                              csmpn_stream_t stream) {
  if (mode != 0 || (M != 64 && M != 128) || N < 8 || N > 64 || N % 8 || K < 8 || K > 128 || K % 8)
    return CSMPN_ERR_BAD_ARG;
  if (mode != 2 && M != 128) return CSMPN_ERR_BAD_ARG;
  if (!A || !Bsrc || !dump) return CSMPN_ERR_BAD_ARG;
  const int a_elems = (mode == 2) ? K * M : 128 * K, b_elems = N * K;
  const size_t smem = (size_t)(a_elems + b_elems) * 8;
  CSMPN_CUDA_TRY(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, Bsrc, dump, mode, M, N, K, flags);
  CSMPN_LAUNCH_CHECK("tc_probe_kernel");
  return CSMPN_OK;
}

namespace csmpn {
// Raw probe: the host supplies byte-exact shared-memory images of A and B and every descriptor field, so operand
// layouts can be explored from Python without recompiling.  prm: 0 M, 1 N, 2 a_mn, 3 b_mn, 4 a_lbo, 5 a_sbo, 6 a_layout,
// 7 b_lbo, 8 b_sbo, 9 b_layout, 10 ksteps, 11 a_kinc (bytes), 12 b_kinc (bytes), 13 a_off, 14 b_off (bytes into each image)
struct ProbeRaw { uint32_t v[16]; };
__global__ void __launch_bounds__(128, 1) tc_probe_raw_kernel(const float* __restrict__ a_img, int a_words,
                                                             const float* __restrict__ b_img, int b_words, ProbeRaw prm,
                                                             float* __restrict__ dump) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* a_s = reinterpret_cast<float*>(smem_raw);
  const int a_pad = ((a_words * 4 + 1023) / 1024) * 1024 / 4;
  float* b_s = a_s + a_pad;
  for (int i = tid; i < a_words; i += 128) a_s[i] = a_img[i];
  for (int i = tid; i < b_words; i += 128) b_s[i] = b_img[i];
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  const uint32_t M = prm.v[0], N = prm.v[1];
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(M, N, prm.v[2] != 0, prm.v[3] != 0);
    for (uint32_t ks = 0; ks < prm.v[10]; ++ks) {
      uint64_t ad = smem_desc(smem_addr(a_s) + prm.v[13] + ks * prm.v[11], prm.v[4], prm.v[5]) | ((uint64_t)prm.v[6] << 61);
      uint64_t bd = smem_desc(smem_addr(b_s) + prm.v[14] + ks * prm.v[12], prm.v[7], prm.v[8]) | ((uint64_t)prm.v[9] << 61);
      mma_tf32(tbase, ad, bd, idesc, ks > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (uint32_t n = 0; n < N; ++n) {
    float v;
    tmem_ld1(tmem_at(tbase, warp * 32, n), v);
    tmem_wait_ld();
    dump[(warp * 32 + lane) * N + n] = v;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 64);
}
}  // namespace csmpn

extern "C" int csmpn_tc_probe_raw(const float* a_img, int a_words, const float* b_img, int b_words, const uint32_t* prm16,
                                  float* dump, csmpn_stream_t stream) {
  if (!a_img || !b_img || !prm16 || !dump || a_words <= 0 || b_words <= 0) return CSMPN_ERR_BAD_ARG;
  csmpn::ProbeRaw prm;
  for (int i = 0; i < 16; ++i) prm.v[i] = prm16[i];
  if ((prm.v[0] != 64 && prm.v[0] != 128) || prm.v[1] < 8 || prm.v[1] > 64 || prm.v[1] % 8) return CSMPN_ERR_BAD_ARG;
  const size_t a_pad = ((size_t)a_words * 4 + 1023) / 1024 * 1024;
  const size_t smem = a_pad + (size_t)b_words * 4 + 1024;
  if (smem > 200 * 1024) return CSMPN_ERR_BAD_ARG;
  CSMPN_CUDA_TRY(cudaFuncSetAttribute(csmpn::tc_probe_raw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  csmpn::tc_probe_raw_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a_img, a_words, b_img, b_words, prm, dump);
  CSMPN_LAUNCH_CHECK("tc_probe_raw_kernel");
  return CSMPN_OK;
}
