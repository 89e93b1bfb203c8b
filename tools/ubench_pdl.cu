// Kernel-to-kernel dependency latency inside a CUDA graph, with and without programmatic dependent launch (PDL):
// a chain of N kernels shaped like the tensor-core kernels (148 CTAs x 512 threads, 200 KB dynamic shared memory, a
// ~4 us "prologue" that does not depend on the predecessor, then a short dependent body).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench_pdl.cu -o /tmp/ubench_pdl && /tmp/ubench_pdl
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 1) body_kernel(float* buf, int spin_pro, int spin_body, int pdl) {
  extern __shared__ float sm[];
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // prologue: independent of the predecessor
  long long t0 = clock64();
  while (clock64() - t0 < spin_pro) {}
  sm[threadIdx.x] = (float)threadIdx.x;
  __syncthreads();
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  // dependent body
  float v = buf[blockIdx.x * 512 + threadIdx.x];
  t0 = clock64();
  while (clock64() - t0 < spin_body) {}
  buf[blockIdx.x * 512 + threadIdx.x] = v + sm[threadIdx.x ^ 1];
}

static float run(int n, int pdl, int spin_pro, int spin_body, float* buf) {
  cudaStream_t s;
  cudaStreamCreate(&s);
  cudaFuncSetAttribute(body_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < n; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 200 * 1024; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, body_kernel, buf, spin_pro, spin_body, pdl);
  }
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, s);
  cudaStreamSynchronize(s);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, s);
  for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s);
  cudaStreamSynchronize(s);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(err));
  return ms * 1e3f / (10 * n);
}

int main() {
  float* buf;
  cudaMalloc(&buf, 148 * 512 * sizeof(float));
  cudaMemset(buf, 0, 148 * 512 * sizeof(float));
  const int n = 40;
  for (int body_us : {0, 20}) {
    for (int pro_us : {0, 4}) {
      const int sp = (int)(pro_us * 1965), sb = (int)(body_us * 1965);
      const float a = run(n, 0, sp, sb, buf), b = run(n, 1, sp, sb, buf);
      printf("body %2d us, prologue %d us: plain %.2f us per kernel, PDL %.2f us per kernel  (saves %.2f us)\n", body_us, pro_us, a, b, a - b);
    }
  }
  return 0;
}
