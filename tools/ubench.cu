// tools/ubench.cu -- single-warp latency / issue-rate probes for the fp64 pipe, shared memory and shuffles on the
// target GPU (numbers quoted in DESIGN.md).  nvcc -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dadd_chain(double* out, long long* cyc, int n, double a) {
  double t = out[threadIdx.x];
  long long c0 = clock64();
  for (int i = 0; i < n; i++) { t = t + a; t = t - a * 0.5; t = t + a; t = t - a * 0.25; }
  long long c1 = clock64();
  out[threadIdx.x] = t;
  if (threadIdx.x == 0) cyc[0] = c1 - c0;
}
__global__ void k_dmul_chain(double* out, long long* cyc, int n, double a) {
  double t = out[threadIdx.x];
  long long c0 = clock64();
  for (int i = 0; i < n; i++) { t = t * a; t = t * a; t = t * a; t = t * a; }
  long long c1 = clock64();
  out[threadIdx.x] = t;
  if (threadIdx.x == 0) cyc[0] = c1 - c0;
}
__global__ void k_dmul_indep(double* out, long long* cyc, int n, double a) {
  double t0 = out[threadIdx.x], t1 = t0 + 1, t2 = t0 + 2, t3 = t0 + 3, t4 = t0 + 4, t5 = t0 + 5, t6 = t0 + 6, t7 = t0 + 7;
  long long c0 = clock64();
  for (int i = 0; i < n; i++) { t0 *= a; t1 *= a; t2 *= a; t3 *= a; t4 *= a; t5 *= a; t6 *= a; t7 *= a; }
  long long c1 = clock64();
  out[threadIdx.x] = t0 + t1 + t2 + t3 + t4 + t5 + t6 + t7;
  if (threadIdx.x == 0) cyc[0] = c1 - c0;
}
__global__ void k_ffma_indep(float* out, long long* cyc, int n, float a) {
  float t0 = out[threadIdx.x], t1 = t0 + 1, t2 = t0 + 2, t3 = t0 + 3, t4 = t0 + 4, t5 = t0 + 5, t6 = t0 + 6, t7 = t0 + 7;
  long long c0 = clock64();
  for (int i = 0; i < n; i++) { t0 *= a; t1 *= a; t2 *= a; t3 *= a; t4 *= a; t5 *= a; t6 *= a; t7 *= a; }
  long long c1 = clock64();
  out[threadIdx.x] = t0 + t1 + t2 + t3 + t4 + t5 + t6 + t7;
  if (threadIdx.x == 0) cyc[0] = c1 - c0;
}
__global__ void k_lds_chain(int* out, long long* cyc, int n) {
  __shared__ int sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 37 + 11) & 1023;
  __syncthreads();
  int idx = threadIdx.x;
  long long c0 = clock64();
  for (int i = 0; i < n; i++) { idx = sm[idx]; idx = sm[idx]; idx = sm[idx]; idx = sm[idx]; }
  long long c1 = clock64();
  out[threadIdx.x] = idx;
  if (threadIdx.x == 0) cyc[0] = c1 - c0;
}
__global__ void k_shfl_chain(double* out, long long* cyc, int n) {
  double t = out[threadIdx.x];
  long long c0 = clock64();
  for (int i = 0; i < n; i++) { t = __shfl_sync(0xffffffffu, t, (threadIdx.x + 1) & 31); t = __shfl_sync(0xffffffffu, t, (threadIdx.x + 3) & 31);
                                t = __shfl_sync(0xffffffffu, t, (threadIdx.x + 5) & 31); t = __shfl_sync(0xffffffffu, t, (threadIdx.x + 7) & 31); }
  long long c1 = clock64();
  out[threadIdx.x] = t;
  if (threadIdx.x == 0) cyc[0] = c1 - c0;
}
__global__ void k_syncwarp_sts(double* out, long long* cyc, int n) {
  __shared__ double sm[64];
  sm[threadIdx.x] = out[threadIdx.x];
  __syncwarp();
  double t = 0;
  long long c0 = clock64();
  for (int i = 0; i < n; i++) {
    t += sm[(threadIdx.x + 1) & 31]; __syncwarp(); sm[threadIdx.x] = t; __syncwarp();
    t += sm[(threadIdx.x + 1) & 31]; __syncwarp(); sm[threadIdx.x] = t; __syncwarp();
  }
  long long c1 = clock64();
  out[threadIdx.x] = t;
  if (threadIdx.x == 0) cyc[0] = c1 - c0;
}

int main() {
  double* d; float* f; int* ii; long long* cyc;
  cudaMalloc(&d, 4096); cudaMalloc(&f, 4096); cudaMalloc(&ii, 4096); cudaMalloc(&cyc, 64);
  cudaMemset(d, 0, 4096); cudaMemset(f, 0, 4096);
  const int n = 4096;
  long long h;
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  printf("device %s sm_%d%d, %d SMs\n", pr.name, pr.major, pr.minor, pr.multiProcessorCount);
#define RUN(name, launch, per)                                                     \
  for (int rep = 0; rep < 2; rep++) { launch; cudaDeviceSynchronize(); }           \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                  \
  printf("%-34s %8.2f cycles per op (%s)\n", name, (double)h / ((double)n * per), cudaGetErrorString(cudaGetLastError()));
  RUN("DADD dependent chain, 1 warp", (k_dadd_chain<<<1, 32>>>(d, cyc, n, 1.0000001)), 4)
  RUN("DMUL dependent chain, 1 warp", (k_dmul_chain<<<1, 32>>>(d, cyc, n, 1.0000001)), 4)
  RUN("DMUL independent x8, 1 warp", (k_dmul_indep<<<1, 32>>>(d, cyc, n, 1.0000001)), 8)
  RUN("DMUL independent x8, 4 warps", (k_dmul_indep<<<1, 128>>>(d, cyc, n, 1.0000001)), 8)
  RUN("DMUL independent x8, 8 warps", (k_dmul_indep<<<1, 256>>>(d, cyc, n, 1.0000001)), 8)
  RUN("DMUL independent x8, 1 warp 3 lanes", (k_dmul_indep<<<1, 3>>>(d, cyc, n, 1.0000001)), 8)
  RUN("FMUL independent x8, 1 warp", (k_ffma_indep<<<1, 32>>>(f, cyc, n, 1.0000001f)), 8)
  RUN("LDS dependent chain", (k_lds_chain<<<1, 32>>>(ii, cyc, n)), 4)
  RUN("SHFL.64 dependent chain", (k_shfl_chain<<<1, 32>>>(d, cyc, n)), 4)
  RUN("LDS+syncwarp+STS+syncwarp round", (k_syncwarp_sts<<<1, 32>>>(d, cyc, n)), 2)
  return 0;
}
