// tools/ubench_icache.cu -- instruction-fetch probe: one warp runs a straight-line body of N fp64 adds (8 independent
// accumulators, so the body is issue bound at ~2.3 cycles per instruction when it hits the instruction caches) in a
// loop; cycles per instruction against the body's code footprint.
#include <cstdio>
#include <cuda_runtime.h>

template <int N>
__global__ void k_body(double* out, long long* cyc, int iters, double a) {
  double t0 = out[0], t1 = t0 + 1, t2 = t0 + 2, t3 = t0 + 3, t4 = t0 + 4, t5 = t0 + 5, t6 = t0 + 6, t7 = t0 + 7;
  long long c0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < N / 8; k++) { t0 += a; t1 += a; t2 += a; t3 += a; t4 += a; t5 += a; t6 += a; t7 += a; }
    a = a * 1.0000001;  // keeps the compiler from collapsing iterations
  }
  long long c1 = clock64();
  out[threadIdx.x] = t0 + t1 + t2 + t3 + t4 + t5 + t6 + t7;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = c1 - c0;
}

template <int N>
void run(double* d, long long* cyc, int warps) {
  const int iters = (1 << 22) / N;
  long long h;
  for (int rep = 0; rep < 2; rep++) { k_body<N><<<1, 32 * warps>>>(d, cyc, iters, 1.0000001); cudaDeviceSynchronize(); }
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("body %6d instr (%4d KB) x %2d warps: %6.2f cycles per instruction\n", N, N * 16 / 1024, warps, (double)h / ((double)iters * N));
}

int main() {
  double* d; long long* cyc;
  cudaMalloc(&d, 4096); cudaMalloc(&cyc, 64); cudaMemset(d, 0, 4096);
  run<256>(d, cyc, 1); run<1024>(d, cyc, 1); run<2048>(d, cyc, 1); run<4096>(d, cyc, 1); run<8192>(d, cyc, 1);
  run<16384>(d, cyc, 1); run<32768>(d, cyc, 1);
  run<2048>(d, cyc, 8); run<8192>(d, cyc, 8); run<16384>(d, cyc, 8);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
