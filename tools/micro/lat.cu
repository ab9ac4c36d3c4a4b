// Dependent-issue latencies of the operations on the pivot chain (one warp): DFMA, SHFL (64-bit), rsqrt(double),
// __drcp_rn, LDS.64, DMMA m8n8k4.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 lat.cu -o lat
#include <cstdio>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ double sm[64];
  const int lane = threadIdx.x;
  sm[lane] = (double)((lane + 1) & 31); sm[lane + 32] = 0;
  __syncwarp();
  const int N = 512;
  double x = seed + lane * 1e-3, y = 1.0000001, acc = 0;
  long long t0, t1;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
  t1 = clock64(); if (lane == 0) cyc[0] = (t1 - t0) / N; acc += x;
  x = seed + lane;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
  t1 = clock64(); if (lane == 0) cyc[1] = (t1 - t0) / N; acc += x;
  x = seed + 2.0 + lane;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) x = rsqrt(x) + 3.0;
  t1 = clock64(); if (lane == 0) cyc[2] = (t1 - t0) / N; acc += x;
  x = seed + 2.0 + lane;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) x = __drcp_rn(x) + 3.0;
  t1 = clock64(); if (lane == 0) cyc[3] = (t1 - t0) / N; acc += x;
  int idx = lane;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) idx = (int)sm[idx & 31];
  t1 = clock64(); if (lane == 0) cyc[4] = (t1 - t0) / N; acc += idx;
  double c0 = 0, c1 = 0; x = 1e-3 * lane;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) dmma(c0, c1, x, y);
  t1 = clock64(); if (lane == 0) cyc[5] = (t1 - t0) / N; acc += c0 + c1;
  // independent DMMAs (4 accumulator pairs): issue interval
  double d0 = 0, d1 = 0, e0 = 0, e1 = 0, f0 = 0, f1 = 0;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) { dmma(c0, c1, x, y); dmma(d0, d1, x, y); dmma(e0, e1, x, y); dmma(f0, f1, x, y); }
  t1 = clock64(); if (lane == 0) cyc[6] = (t1 - t0) / (4 * N); acc += c0 + d0 + e0 + f0 + c1 + d1 + e1 + f1;
  // independent DFMAs (4 chains)
  double x1 = x + 1, x2 = x + 2, x3 = x + 3;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) { x = fma(x, y, 1e-9); x1 = fma(x1, y, 1e-9); x2 = fma(x2, y, 1e-9); x3 = fma(x3, y, 1e-9); }
  t1 = clock64(); if (lane == 0) cyc[7] = (t1 - t0) / (4 * N); acc += x + x1 + x2 + x3;
  // sqrt + division
  x = seed + 2.0 + lane;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) x = sqrt(x) + 3.0;
  t1 = clock64(); if (lane == 0) cyc[8] = (t1 - t0) / N; acc += x;
  x = seed + 2.0 + lane;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) x = 1.0 / x + 3.0;
  t1 = clock64(); if (lane == 0) cyc[9] = (t1 - t0) / N; acc += x;
  out[lane] = acc;
}
int main() {
  double* o; long long* c; cudaMalloc(&o, 32 * 8); cudaMalloc(&c, 16 * 8);
  k<<<1, 32>>>(o, c, 1.5); k<<<1, 32>>>(o, c, 1.5);
  long long h[16]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
  printf("dependent latency (cycles): DFMA %lld  SHFL64 %lld  rsqrt+add %lld  drcp+add %lld  LDS64+cvt %lld  DMMA(dep) %lld | issue interval: DMMA %lld  DFMA(4 chains) %lld | sqrt+add %lld  div+add %lld\n",
         h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9]);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
