// Micro-benchmark of the CTA-wide tiled Cholesky (bvg_tiled.cuh): cycles per phase for one CTA, and throughput with
// the SMs filled.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../bvartools_b200/csrc
//                        -DBVG_TILED_PROF chol_prof.cu -o chol_prof
#include <cstdio>
#include <vector>
#include <cmath>
#include "bvg_tiled.cuh"
using namespace bvg;

__global__ void __launch_bounds__(512, 1) k_chol(const double* A, const double* b0, double* out, int n, int reps, long long* prof) {
  extern __shared__ double sm[];
  const int tl = tiled_len(n), npad = tiled_nt(n) * 8;
  double* L = sm; double* b = L + tl; double* a = b + npad;
  int fail = 0;
  long long t_ch = 0, t_bs = 0, t_ld = 0;
  for (int r = 0; r < reps; ++r) {
    long long c0 = clock64();
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) { int i = e % n, j = e / n; if (j <= i) L[tiled_idx(i, j)] = A[e]; }
    for (int e = threadIdx.x; e < n; e += blockDim.x) b[e] = b0[e];
    __syncthreads();
    tiled_finish_build(L, b, n);
    long long c1 = clock64();
#ifdef BVG_TILED_PROF
    tiled_chol(L, b, n, fail, (blockIdx.x == 0 && r == reps - 1) ? prof + 8 : nullptr);
#else
    tiled_chol(L, b, n, fail);
#endif
    long long c2 = clock64();
    tiled_back_solve(L, b, a, n);
    long long c3 = clock64();
    t_ld += c1 - c0; t_ch += c2 - c1; t_bs += c3 - c2;
  }
  if (blockIdx.x == 0) {
    for (int e = threadIdx.x; e < n; e += blockDim.x) out[e] = a[e];
    if (threadIdx.x == 0) { prof[0] = t_ld / reps; prof[1] = t_ch / reps; prof[2] = t_bs / reps; prof[3] = fail; }
  }
}

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 150, threads = argc > 2 ? atoi(argv[2]) : 256, grid = argc > 3 ? atoi(argv[3]) : 1;
  int reps = 50;
  std::vector<double> A(n * n), b(n), x(n);
  // SPD: A = B B' + n I
  std::vector<double> B(n * n);
  unsigned s = 12345;
  for (auto& v : B) { s = s * 1664525u + 1013904223u; v = ((s >> 8) & 0xffff) / 65536.0 - 0.5; }
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double t = 0; for (int k = 0; k < n; ++k) t += B[i + n * k] * B[j + n * k]; A[i + n * j] = t + (i == j ? n * 0.05 : 0); }
  for (int i = 0; i < n; ++i) b[i] = std::sin(i + 1.0);
  double *dA, *db, *dout; long long* dprof;
  cudaMalloc(&dA, A.size() * 8); cudaMalloc(&db, n * 8); cudaMalloc(&dout, n * 8); cudaMalloc(&dprof, 64 * 8);
  cudaMemset(dprof, 0, 64 * 8);
  cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), n * 8, cudaMemcpyHostToDevice);
  size_t smem = (tiled_len(n) + 2 * (tiled_nt(n) * 8) + 64) * 8;
  cudaFuncSetAttribute(k_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_chol<<<grid, threads, smem>>>(dA, db, dout, n, 2, dprof);
  cudaEventRecord(e0);
  k_chol<<<grid, threads, smem>>>(dA, db, dout, n, reps, dprof);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long prof[64]; cudaMemcpy(prof, dprof, sizeof(prof), cudaMemcpyDeviceToHost); cudaMemcpy(x.data(), dout, n * 8, cudaMemcpyDeviceToHost);
  // check: x solves (L L') ... a = L^-T L^-1 b = A^-1 b
  double rmax = 0, bn = 0;
  for (int i = 0; i < n; ++i) { double t = 0; for (int j = 0; j < n; ++j) t += A[i + n * j] * x[j]; rmax = fmax(rmax, fabs(t - b[i])); bn = fmax(bn, fabs(b[i])); }
  printf("n=%d threads=%d grid=%d err=%s fail=%lld resid=%.2e | cycles: load %lld chol %lld backsolve %lld | kernel %.3f ms -> %.1f us per factor+solve per CTA-slot\n",
         n, threads, grid, cudaGetErrorString(err), prof[3], rmax / bn, prof[0], prof[1], prof[2], ms, ms * 1e3 / reps);
  printf("  phase cycles (thread 0 / warp 1): panel %lld  rhs+tile0 %lld  factor %lld  wait %lld | w1 trailing %lld  w1 wait %lld  first factor %lld\n",
         prof[8], prof[9], prof[10], prof[11], prof[12], prof[13], prof[14]);
  return 0;
}
