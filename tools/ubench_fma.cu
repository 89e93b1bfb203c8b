// Micro-benchmark: FP32 FFMA (3-register) vs packed FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
// Decides the inner-loop form of the per-grade channel GEMMs (DESIGN.md "FP32 pipe").
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a0, float b0) {
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x * 0.001f + i;
  float a[4] = {a0, a0 + 1.f, a0 + 2.f, a0 + 3.f};
  float b[4] = {b0, b0 + .5f, b0 + .25f, b0 + .125f};
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = fmaf(a[(i + r) & 3], b[(i >> 2) & 3], acc[i]);
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          unsigned long long d, x, y, c;
          asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[(i + r) & 3]), "f"(a[(i + r + 1) & 3]));
          asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b[(i >> 2) & 3]), "f"(b[((i >> 2) + 1) & 3]));
          asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc[i]), "f"(acc[i + 1]));
          asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(y), "l"(c));
          asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i]), "=f"(acc[i + 1]) : "l"(d));
        }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name) {
  float* out;
  int blocks = 148 * 8, threads = 256, iters = 4096;
  cudaMalloc(&out, blocks * threads * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0001f, 0.9999f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double fma = double(blocks) * threads * iters * 128.0;
    printf("%s rep%d: %.3f ms  %.2f TFLOP/s (%.1f FMA/clk/SM @1.965GHz)\n", name, rep, ms, 2 * fma / ms * 1e-9,
           fma / (ms * 1e-3) / 148 / 1.965e9);
  }
  cudaFree(out);
}

int main() {
  run<0>("ffma_scalar");
  run<1>("ffma2_packed");
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
