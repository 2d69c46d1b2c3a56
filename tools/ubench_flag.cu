// tools/ubench_flag.cu -- cross-warp hand-off through a shared-memory flag: warp A sets flag[i], warp B spins on it and
// answers through flag2[i]; cycles per round trip, with and without a CTA fence before the flag store, with other
// warps idle / polling / doing fp64 work (the situations of the data-flow impulse kernel).
#include <cstdio>
#include <cuda_runtime.h>

template <int FENCE, int BUSY>
__global__ void k_pingpong(long long* cyc, int n, double* sink) {
  __shared__ volatile unsigned char f1[4096], f2[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) { f1[i] = 0; f2[i] = 0; }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    long long c0 = clock64();
    for (int i = 0; i < n; i++) {
      if (FENCE) asm volatile("fence.acq_rel.cta;" ::: "memory");
      f1[i] = 1;
      while (!f2[i]) { }
    }
    cyc[0] = clock64() - c0;
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < n; i++) {
      while (!f1[i]) { }
      if (FENCE) asm volatile("fence.acq_rel.cta;" ::: "memory");
      f2[i] = 1;
    }
  } else if (BUSY == 1 && lane == 0) {  // other warps: one lane spinning on a flag that never comes
    while (!f1[4095]) { if (f2[n - 1]) break; }
  } else if (BUSY == 2 && lane == 0) {  // other warps: fp64 work
    double t = threadIdx.x;
    while (!f2[n - 1]) { for (int k = 0; k < 64; k++) t = t * 1.0000001 + 0.5; }
    sink[threadIdx.x] = t;
  }
}

int main() {
  long long* cyc; double* sink;
  cudaMalloc(&cyc, 64); cudaMalloc(&sink, 8192);
  const int n = 4000;
  long long h;
#define RUN(name, F, B, T)                                                   \
  for (int r = 0; r < 2; r++) { k_pingpong<F, B><<<1, T>>>(cyc, n, sink); cudaDeviceSynchronize(); } \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                          \
  printf("%-58s %7.1f cycles per round trip (%s)\n", name, (double)h / n, cudaGetErrorString(cudaGetLastError()));
  RUN("2 warps, no fence", 0, 0, 64)
  RUN("2 warps, fence.acq_rel.cta before each flag store", 1, 0, 64)
  RUN("8 warps (6 idle), no fence", 0, 0, 256)
  RUN("8 warps (6 spinning lanes), no fence", 0, 1, 256)
  RUN("8 warps (6 spinning lanes), fence", 1, 1, 256)
  RUN("8 warps (6 doing fp64 chains), no fence", 0, 2, 256)
  RUN("8 warps (6 doing fp64 chains), fence", 1, 2, 256)
  return 0;
}
