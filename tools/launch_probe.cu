// Micro-benchmark: per-kernel cost of dependent kernel chains inside a CUDA graph on this GPU (empty kernels, big dynamic
// shared memory, TMEM allocation, alternating configurations).  nvcc -arch=sm_100a -o launch_probe launch_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_empty(float* p) { if (p && threadIdx.x == 9999) p[0] = 1.f; }
__global__ void k_smem(float* p) { extern __shared__ float s[]; if (p && threadIdx.x == 9999) p[0] = s[0]; }
__global__ void __launch_bounds__(192, 1) k_tmem(float* p) {
  extern __shared__ float s[];
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(32u) : "memory");
  if (p && threadIdx.x == 9999) p[0] = s[0];
}
__global__ void k_work(float* p, int n) {   // streams n floats per block
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += p[(size_t)blockIdx.x * n + i];
  if (acc == 12345.f) p[0] = acc;
}

template <typename F> float time_graph(cudaStream_t st, int reps, F enqueue) {
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < reps; ++i) enqueue(i);
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
  cudaEventRecord(a, st);
  for (int r = 0; r < 5; ++r) cudaGraphLaunch(ge, st);
  cudaEventRecord(b, st); cudaStreamSynchronize(st);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return ms * 1000.f / (5 * reps);
}

int main() {
  cudaStream_t st; CK(cudaStreamCreate(&st));
  float* buf; CK(cudaMalloc(&buf, 256 << 20)); CK(cudaMemset(buf, 0, 256 << 20));
  CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024));
  CK(cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024));
  const int R = 200;
  printf("empty <<<32,256>>>                 : %.2f us/kernel\n", time_graph(st, R, [&](int) { k_empty<<<32, 256, 0, st>>>(buf); }));
  printf("empty <<<148,192>>>                : %.2f us/kernel\n", time_graph(st, R, [&](int) { k_empty<<<148, 192, 0, st>>>(buf); }));
  printf("smem164K <<<148,192>>>             : %.2f us/kernel\n", time_graph(st, R, [&](int) { k_smem<<<148, 192, 164 * 1024, st>>>(buf); }));
  printf("smem164K+tmem <<<148,192>>>        : %.2f us/kernel\n", time_graph(st, R, [&](int) { k_tmem<<<148, 192, 164 * 1024, st>>>(buf); }));
  printf("alternate empty32 / smem164K+tmem  : %.2f us/kernel\n", time_graph(st, R, [&](int i) { if (i & 1) k_tmem<<<148, 192, 164 * 1024, st>>>(buf); else k_empty<<<32, 256, 0, st>>>(buf); }));
  CK(cudaFuncSetAttribute(k_empty, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  printf("alternate (empty carveout=max)     : %.2f us/kernel\n", time_graph(st, R, [&](int i) { if (i & 1) k_tmem<<<148, 192, 164 * 1024, st>>>(buf); else k_empty<<<32, 256, 0, st>>>(buf); }));
  printf("work 32 blocks x 68KB              : %.2f us/kernel\n", time_graph(st, R, [&](int) { k_work<<<32, 256, 0, st>>>(buf, 17000); }));
  printf("work 148 blocks x 120KB (17.7MB)   : %.2f us/kernel\n", time_graph(st, R, [&](int i) { k_work<<<148, 256, 0, st>>>(buf + (size_t)(i % 3) * 5000000, 30000); }));
  // PDL-style: same but eager back-to-back launches (no graph)
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a, st);
  for (int i = 0; i < 2000; ++i) k_empty<<<32, 256, 0, st>>>(buf);
  cudaEventRecord(b, st); cudaStreamSynchronize(st);
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("eager empty <<<32,256>>>           : %.2f us/kernel\n", ms * 1000 / 2000);
  return 0;
}
