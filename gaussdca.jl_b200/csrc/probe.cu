// probe.cu -- raw pipe-throughput probes (bench only): the roofline denominators that
// MEASURED_PEAKS.json does not carry (INT32 ALU issue, POPC, FP64 DMMA, FP64 FMA).
#include "gdca_internal.cuh"

namespace {
constexpr int PITER = 4096;

__global__ void __launch_bounds__(256) lop3_probe(uint32_t *out, uint32_t seed) {
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = seed * (threadIdx.x + 1) + i;
  uint32_t a = seed ^ 0x9e3779b9u, b = threadIdx.x * 0x85ebca6bu;
  for (int it = 0; it < PITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = x[i] | (a ^ (b + x[(i + 1) & 15]));  // 1 LOP3 + 1 IADD per element
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) r ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void __launch_bounds__(256) lop3_only_probe(uint32_t *out, uint32_t seed) {
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = seed * (threadIdx.x + 1) + i;
  for (int it = 0; it < PITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      // x_i = x_i | (x_{i+1} ^ x_{i+2}) : exactly one LOP3 (lut 0xf6)
      asm volatile("lop3.b32 %0, %0, %1, %2, 0xf6;" : "+r"(x[i]) : "r"(x[(i + 1) & 15]), "r"(x[(i + 2) & 15]));
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) r ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void __launch_bounds__(256) popc_probe(uint32_t *out, uint32_t seed) {
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = seed * (threadIdx.x + 1) + i * 0x01000193u;
  for (int it = 0; it < PITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("popc.b32 %0, %1;" : "=r"(x[i]) : "r"(x[i] | 0x10000u));
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) r ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void __launch_bounds__(256) dmma_probe(double *out) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < PITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void __launch_bounds__(256) dfma_probe(double *out) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  for (int it = 0; it < PITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = __fma_rn(c[i], a, b);
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
}  // namespace

int32_t gdca_k_probe(gdca_ctx *ctx, double *lop3, double *popc, double *dmma, double *dfma) {
  const int blocks = ctx->num_sms * 8, threads = 256;
  void *buf = nullptr;
  GDCA_CUDA(ctx, cudaMalloc(&buf, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  GDCA_CUDA(ctx, cudaEventCreate(&e0));
  GDCA_CUDA(ctx, cudaEventCreate(&e1));
  float ms = 0;
  const double nthreads = (double)blocks * threads;
  auto timeit = [&](auto launch) -> float {
    launch();  // warm-up
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
      cudaEventRecord(e0, ctx->stream);
      launch();
      cudaEventRecord(e1, ctx->stream);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
    return best;
  };
  float t;
  t = timeit([&] { lop3_only_probe<<<blocks, threads, 0, ctx->stream>>>((uint32_t *)buf, 12345u); });
  if (lop3) *lop3 = nthreads * PITER * 16.0 / (t * 1e-3) / 1e12;
  t = timeit([&] { popc_probe<<<blocks, threads, 0, ctx->stream>>>((uint32_t *)buf, 12345u); });
  if (popc) *popc = nthreads * PITER * 16.0 / (t * 1e-3) / 1e12;
  t = timeit([&] { dmma_probe<<<blocks, threads, 0, ctx->stream>>>((double *)buf); });
  if (dmma) *dmma = (nthreads / 32.0) * PITER * 16.0 * 512.0 / (t * 1e-3) / 1e12;
  t = timeit([&] { dfma_probe<<<blocks, threads, 0, ctx->stream>>>((double *)buf); });
  if (dfma) *dfma = nthreads * PITER * 16.0 * 2.0 / (t * 1e-3) / 1e12;
  ctx->launches += 16;
  (void)lop3_probe;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  GDCA_CUDA(ctx, cudaFree(buf));
  GDCA_CUDA(ctx, cudaGetLastError());
  return GDCA_OK;
}
