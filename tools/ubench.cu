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

// independent conflict-free LDS.64 x8 per iteration (addresses do not depend on loaded data)
__global__ void k_lds64_indep(double* out, long long* cyc, int n, int stride) {
  __shared__ double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int base = (threadIdx.x & 31) * stride;
  double t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0, t7 = 0;
  long long c0 = clock64();
  for (int i = 0; i < n; i++) {
    const volatile double* p = sm + base + (i & 7) * 256;
    t0 += p[0]; t1 += p[32 * stride]; t2 += p[64 * stride]; t3 += p[96 * stride];
    t4 += p[1]; t5 += p[32 * stride + 1]; t6 += p[64 * stride + 1]; t7 += p[96 * stride + 1];
  }
  long long c1 = clock64();
  out[threadIdx.x] = t0 + t1 + t2 + t3 + t4 + t5 + t6 + t7;
  if (threadIdx.x == 0) cyc[0] = c1 - c0;
}
// the shape of one k-group of the ILU factorisation: 24 LDS.64 (dynamic addresses), 18 DMUL + 18 DADD, 6 STS.64, syncwarp
__global__ void k_kgroup_shape(double* out, long long* cyc, int n) {
  __shared__ double sm[4096];
  __shared__ int off[64];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + i * 1e-6;
  if (threadIdx.x < 64) off[threadIdx.x] = (threadIdx.x * 97) & 1023;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double acc = 0;
  long long c0 = clock64();
  for (int i = 0; i < n; i++) {
    const double a0 = sm[lane], a1 = sm[lane + 32], a2 = sm[lane + 64];
    double b0[6], b1[6], b2[6], t[6];
    int pt[6];
#pragma unroll
    for (int j = 0; j < 6; j++) {
      const int o = off[(i + j) & 63];
      pt[j] = 2048 + ((o + lane) & 1023);
      b0[j] = sm[o + lane]; b1[j] = sm[o + lane + 3]; b2[j] = sm[o + lane + 6];
      t[j] = sm[pt[j]];
    }
#pragma unroll
    for (int j = 0; j < 6; j++) t[j] = t[j] - (a0 * b0[j] + a1 * b1[j] + a2 * b2[j]);
#pragma unroll
    for (int j = 0; j < 6; j++) sm[pt[j]] = t[j];
    __syncwarp();
    acc += t[0];
  }
  long long c1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[0] = c1 - c0;
}
// division and square root latency (dependent chains)
__global__ void k_ddiv_chain(double* out, long long* cyc, int n, double a) {
  double t = out[threadIdx.x] + 3.0;
  long long c0 = clock64();
  for (int i = 0; i < n; i++) { t = a / t; t = a / t; t = a / t; t = a / t; }
  long long c1 = clock64();
  out[threadIdx.x] = t;
  if (threadIdx.x == 0) cyc[0] = c1 - c0;
}
__global__ void k_dsqrt_chain(double* out, long long* cyc, int n, double a) {
  double t = out[threadIdx.x] + 3.0;
  long long c0 = clock64();
  for (int i = 0; i < n; i++) { t = sqrt(t) + a; t = sqrt(t) + a; t = sqrt(t) + a; t = sqrt(t) + a; }
  long long c1 = clock64();
  out[threadIdx.x] = t;
  if (threadIdx.x == 0) cyc[0] = c1 - c0;
}
// division by a divisor whose correctly rounded reciprocal is known (Markstein): q0 = a*y; r = fma(-b,q0,a); q1 = fma(r,y,q0); again
__global__ void k_ddiv_markstein(double* out, long long* cyc, int n, double b) {
  double t = out[threadIdx.x] + 3.0;
  const double y = 1.0 / b;
  long long c0 = clock64();
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double q0 = t * y;
      const double r0 = __fma_rn(-b, q0, t);
      const double q1 = __fma_rn(r0, y, q0);
      const double r1 = __fma_rn(-b, q1, t);
      t = __fma_rn(r1, y, q1) + 1.0;
    }
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
  RUN("LDS.64 independent x8, 1 warp", (k_lds64_indep<<<1, 32>>>(d, cyc, n, 1)), 8)
  RUN("LDS.64 independent x8, 4 warps", (k_lds64_indep<<<1, 128>>>(d, cyc, n, 1)), 8)
  RUN("LDS.64 independent x8, 8 warps", (k_lds64_indep<<<1, 256>>>(d, cyc, n, 1)), 8)
  RUN("k-group shape (24 LDS.64, 36 DP, 6 STS), 1 warp", (k_kgroup_shape<<<1, 32>>>(d, cyc, n)), 1)
  RUN("k-group shape, 4 warps", (k_kgroup_shape<<<1, 128>>>(d, cyc, n)), 1)
  RUN("k-group shape, 8 warps", (k_kgroup_shape<<<1, 256>>>(d, cyc, n)), 1)
  RUN("DDIV dependent chain", (k_ddiv_chain<<<1, 32>>>(d, cyc, n, 1.7)), 4)
  RUN("DSQRT(+DADD) dependent chain", (k_dsqrt_chain<<<1, 32>>>(d, cyc, n, 1.7)), 4)
  RUN("Markstein division (5 FMA + DADD) chain", (k_ddiv_markstein<<<1, 32>>>(d, cyc, n, 1.7)), 4)
  return 0;
}
