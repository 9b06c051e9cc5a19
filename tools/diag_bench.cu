// Stand-alone timing + correctness harness for diag_block_kernel (includes the TU to reach the anonymous namespace).
#define DIAG_DBG 1
#include "../gaussdca.jl_b200/csrc/chol.cu"
#include <vector>
#include <cmath>
#include <cstdlib>
int main() {
  const int n = 128;
  std::vector<double> A(n * n), R(n * n), X(n * n);
  srand(1);
  for (auto& r : R) r = (rand() / (double)RAND_MAX) - 0.5;
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += R[i*n+k]*R[j*n+k]; A[i*n+j] = s / n + (i == j ? 0.5 : 0.0); }
  double *dA, *dX; int* dInfo;
  cudaMalloc(&dA, n*n*8); cudaMalloc(&dX, n*n*8); cudaMalloc(&dInfo, 4); cudaMemset(dInfo, 0, 4);
  cudaMemcpy(dA, A.data(), n*n*8, cudaMemcpyHostToDevice);
  const size_t dsmem = (size_t)NB * DLD * sizeof(double);
  cudaFuncSetAttribute(diag_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    for (int i = 0; i < 100; ++i) diag_block_kernel<<<1, DT, dsmem>>>(dA, n, dX, n, 0, n, dInfo);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("diag_block_kernel: %.1f us per launch (100 back-to-back)\n", ms * 10);
  }
  cudaFuncSetAttribute(diag_block_kernel2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D2_SMEM);
  long long clk[8]; cudaMemcpyFromSymbol(clk, g_diag_clk, sizeof clk);
  printf("v1 cycles: load %lld  phase1 %lld  phase2 %lld  store %lld\n", clk[1]-clk[0], clk[2]-clk[1], clk[3]-clk[2], clk[4]-clk[3]);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    for (int i = 0; i < 100; ++i) diag_block_kernel2<<<1, DT, D2_SMEM>>>(dA, n, dX, n, 0, n, dInfo);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("diag_block_kernel2: %.1f us per launch (100 back-to-back)\n", ms * 10);
  }
  cudaMemcpyFromSymbol(clk, g_diag_clk, sizeof clk);
  printf("v2 first 32x32 block (factor + inverse): %lld cycles\n", clk[5]-clk[1]);
  printf("cycles: load %lld  phase1 %lld  phase2 %lld  store %lld\n", clk[1]-clk[0], clk[2]-clk[1], clk[3]-clk[2], clk[4]-clk[3]);
  cudaMemcpy(X.data(), dX, n*n*8, cudaMemcpyDeviceToHost);
  // check: X * A * X' == I  (X = L^-1)
  double err = 0;
  std::vector<double> T(n*n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += X[i*n+k]*A[k*n+j]; T[i*n+j] = s; }
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += T[i*n+k]*X[j*n+k]; err = fmax(err, fabs(s - (i == j))); }
  printf("max |X A X' - I| = %.3e   %s\n", err, cudaGetErrorString(cudaGetLastError()));
}
