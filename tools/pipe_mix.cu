// Microbenchmark: how do LOP3, POPC and IMAD share issue slots on sm_100a?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int IT = 2048;
template <int NL, int NP, int MODE>
__global__ void __launch_bounds__(256) mix(uint32_t* out, uint32_t seed) {
  uint32_t x[16], acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { x[i] = seed * (threadIdx.x + 1) + i; acc[i] = 0; }
  uint32_t a = seed ^ 0x9e3779b9u, b = threadIdx.x * 0x85ebca6bu;
  for (int it = 0; it < IT; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
#pragma unroll
      for (int l = 0; l < NL; ++l) asm volatile("lop3.b32 %0, %0, %1, %2, 0xf6;" : "+r"(x[i]) : "r"(a), "r"(b));
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        uint32_t pc;
        asm volatile("popc.b32 %0, %1;" : "=r"(pc) : "r"(x[i]));
        if (MODE == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(acc[i]) : "r"(pc));
        if (MODE == 1) asm volatile("mad.lo.u32 %0, %1, 3, %0;" : "+r"(acc[i]) : "r"(pc));
      }
    }
    a += it; b ^= a;
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) r ^= x[i] + acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int NL, int NP, int MODE>
void run(const char* name, uint32_t* buf, int warps_per_sm) {
  int blocks = 148 * warps_per_sm / 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  mix<NL, NP, MODE><<<blocks, 256>>>(buf, 123);
  cudaEventRecord(e0);
  mix<NL, NP, MODE><<<blocks, 256>>>(buf, 123);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double groups = (double)blocks * 256 * IT * 16;  // thread-level (NL lop3 + NP popc+add) groups
  // cycles per warp-group per SMSP at 1.9 GHz (approx): time * clk * (148*4 SMSP) / (groups/32)
  double cyc = ms * 1e-3 * 1.9e9 * 148 * 4 / (groups / 32);
  printf("%-28s warps/SM=%2d  %.3f ms  ~%.2f cyc per warp-group per SMSP\n", name, warps_per_sm, ms, cyc);
}
int main() {
  uint32_t* buf; cudaMalloc(&buf, 148 * 64 * 256 * 4);
  for (int w : {8, 32}) {
    run<5, 0, 0>("5 LOP3", buf, w);
    run<0, 1, 2>("1 POPC", buf, w);
    run<5, 1, 2>("5 LOP3 + 1 POPC", buf, w);
    run<5, 1, 0>("5 LOP3 + 1 POPC + 1 IADD", buf, w);
    run<5, 1, 1>("5 LOP3 + 1 POPC + 1 IMAD", buf, w);
    run<10, 1, 1>("10 LOP3 + 1 POPC + 1 IMAD", buf, w);
    run<0, 1, 1>("1 POPC + 1 IMAD", buf, w);
  }
  cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
