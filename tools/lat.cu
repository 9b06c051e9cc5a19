// Latency microbenchmarks (single warp unless noted): dependent DFMA chain, rsqrt, __drcp_rn, shfl(double), LDS->DFMA->STS, __syncthreads(512)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ double sh[256];
  double x = seed + threadIdx.x * 1e-3, y = 1.0000001;
  long long t0, t1;
  int slot = 0;
  // 1. dependent DFMA chain
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) x = fma(x, y, 1e-9);
  t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = (t1 - t0) / 256; slot++;
  // 2. dependent rsqrt
  double r = fabs(x) + 2.0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) r = rsqrt(r) + 2.0;
  t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = (t1 - t0) / 64; slot++;
  // 3. dependent __drcp_rn
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) r = __drcp_rn(r) + 2.0;
  t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = (t1 - t0) / 64; slot++;
  // 4. dependent shfl of a double
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) r = __shfl_sync(0xffffffffu, r, (i * 7) & 31) + 1.0;
  t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = (t1 - t0) / 64; slot++;
  // 5. LDS -> DFMA -> STS dependent through shared memory
  sh[threadIdx.x & 255] = r;
  __syncthreads();
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) { double v = sh[(threadIdx.x + i) & 255]; v = fma(v, y, 1e-9); sh[(threadIdx.x + i + 1) & 255] = v; __syncwarp(); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = (t1 - t0) / 64; slot++;
  // 6. __syncthreads with the whole block
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) __syncthreads();
  t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = (t1 - t0) / 64; slot++;
  // 7. independent DFMA throughput per warp (8 chains)
  double a[8]; for (int q = 0; q < 8; ++q) a[q] = x + q;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = fma(a[q], y, 1e-9);
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = (t1 - t0) / 64; slot++;
  double s = r + x; for (int q = 0; q < 8; ++q) s += a[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + sh[threadIdx.x & 255];
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMallocManaged(&cyc, 64);
  const char* names[] = {"dependent DFMA", "dependent rsqrt(+add)", "dependent __drcp_rn(+add)", "dependent shfl double(+add)",
                         "LDS->DFMA->STS->syncwarp", "__syncthreads", "8 independent DFMA (per iteration)"};
  for (int threads : {32, 512}) {
    k<<<1, threads>>>(out, cyc, 1.5); cudaDeviceSynchronize();
    k<<<1, threads>>>(out, cyc, 1.5); cudaDeviceSynchronize();
    printf("block of %d threads:\n", threads);
    for (int i = 0; i < 7; ++i) printf("  %-36s %lld cycles\n", names[i], cyc[i]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
