// ozaki_gemm.cu -- ROUND-2 DRAFT (not part of the library, not built by build(), no test depends on it).
//
// STATUS: compiles for sm_100a (nvcc -gencode arch=compute_100a,code=sm_100a); it has NOT run on a GPU yet -- the round's GPU
// budget was spent when it was written.  It is the worked design of DESIGN.md section 8.1 as a self-checking stand-alone
// program, so that the next round can start with `nvcc ... && ./ozaki_gemm` under gpurun instead of with a blank page.
//
// What it does:  C[m x n] += A[m x k] * B[n x k]^T  in FP64-equivalent arithmetic on the INT8 tensor cores:
//   1. slice_rows_kernel: per row the exponent e (frexp of the row maximum, +1) and S = 7 signed 7-bit digits
//      a / 2^e ~ sum_t d_t 2^(-7 (t+1)), |d_t| <= 64, written as S K-major int8 matrices [S][rows][k]
//      (tools/ozaki_numerics.py: this arithmetic keeps the blocked Cholesky + inverse of chol.cu at 3.8e-13 normwise);
//   2. ozaki_gemm_kernel: D_d = sum_{t+u=d} A_t B_u^T for d = 0..6 with tcgen05.mma kind::i8 (exact S32 accumulation),
//      tile 128 x 64 with all seven diagonal accumulators resident in TMEM (7 x 64 = 448 columns); same warp roles, TMA ring,
//      mbarrier protocol and lean issue loops as csrc/tcfilter.cu, but 64-byte k-blocks (SWIZZLE_64B) so that two stages of the
//      7 + 7 digit tiles (84 KB each) fit in shared memory;
//   3. epilogue: C += 2^(e_i + f_j) * sum_d 2^(-7 (d+2)) D_d, summed in FP64 in ascending d (fma with exact products).
// main() checks the result BIT FOR BIT against the same digit arithmetic on the CPU (int64 sums, same FP64 recombination order)
// and reports the error against a plain FP64 product and the achieved INT8 rate.
//
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a tools/drafts/ozaki_gemm.cu -o ozaki_gemm && ./ozaki_gemm 4096 4096 4096
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CHECK(x)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (x);                                                                         \
    if (e_ != cudaSuccess) {                                                                      \
      fprintf(stderr, "%s:%d: %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));      \
      exit(2);                                                                                    \
    }                                                                                             \
  } while (0)

namespace {

constexpr int S = 7;          // digits per operand
constexpr int W = 7;          // bits per digit
constexpr int BM = 128;       // tile rows
constexpr int BN = 64;        // tile columns (7 x 64 accumulator columns = 448 <= 512)
constexpr int BKB = 64;       // k bytes (= int8 elements) per stage, the SWIZZLE_64B atom width
constexpr int NSTAGE = 2;
constexpr int A_TILE = BM * BKB;                       // 8 KB per digit
constexpr int B_TILE = BN * BKB;                       // 4 KB per digit
constexpr int STAGE_BYTES = S * (A_TILE + B_TILE);     // 84 KB
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr size_t SMEM = (size_t)NSTAGE * STAGE_BYTES + 1024;
// kind::i8: D = S32 (2 at bit 4), A = B = signed int8 (1 at bits 7, 10), K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// ---------------------------------------------------------------------------------------------- slicing
// one warp per row: exponent from the row maximum, then the digits
__global__ void slice_rows_kernel(const double *__restrict__ A, long long rows, long long k, long long lda, int8_t *__restrict__ SA,
                                  int *__restrict__ expo) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const double *a = A + r * lda;
  double mx = 0.0;
  for (long long j = lane; j < k; j += 32) mx = fmax(mx, fabs(a[j]));
  for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  int e0 = 0;
  frexp(mx, &e0);                       // mx = f 2^e0, f in [1/2, 1)
  const int e = (mx > 0.0) ? e0 + 1 : 0;  // |a / 2^e| < 1/2
  if (lane == 0) expo[r] = e;
  for (long long j = lane; j < k; j += 32) {
    double rem = ldexp(a[j], -e);
#pragma unroll
    for (int t = 0; t < S; ++t) {
      const double scale = ldexp(1.0, W * (t + 1));
      const double d = rint(rem * scale);  // |d| <= 64
      rem -= d / scale;                    // exact: d / scale is a dyadic rational inside rem's precision window
      SA[((long long)t * rows + r) * k + j] = (int8_t)(int)d;
    }
  }
}

// ---------------------------------------------------------------------------------------------- PTX helpers (as in csrc/tcfilter.cu)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct GemmParams {
  long long m, n, k, ldc;
  double *C;
  const int *ea, *eb;
  double alpha;  // +1 or -1
};

// ---------------------------------------------------------------------------------------------- the GEMM
// tmapA: [S][m][k] int8 as a 3-D tensor (k fastest), box {64, 128, 1};  tmapB: [S][n][k], box {64, 64, 1};  SWIZZLE_64B.
__global__ void __launch_bounds__(THREADS, 1)
    ozaki_gemm_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, GemmParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bars[2 * NSTAGE + 2];
  __shared__ uint32_t s_tmem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_u32(s_bars);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (NSTAGE + s); };
  const uint32_t tfull_bar = bars + 8u * (2 * NSTAGE), tempty_bar = bars + 8u * (2 * NSTAGE + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapB) : "memory");
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *(volatile uint32_t *)&s_tmem;

  const int tiles_n = (int)(P.n / BN), tiles_m = (int)(P.m / BM);
  const int ntiles = tiles_m * tiles_n;
  const int KB = (int)(P.k / BKB);

  if (warp == 0) {
    // ===== TMA producer: 7 + 7 digit tiles per stage =====
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        if (elect_one()) {
          const uint32_t sa = base + (uint32_t)s * (uint32_t)STAGE_BYTES;
          mbar_expect_tx(full_bar(s), STAGE_BYTES);
#pragma unroll
          for (int t = 0; t < S; ++t) tma_load_3d(sa + t * A_TILE, &tmapA, full_bar(s), kb * BKB, m0, t);
#pragma unroll
          for (int u = 0; u < S; ++u) tma_load_3d(sa + S * A_TILE + u * B_TILE, &tmapB, full_bar(s), kb * BKB, n0, u);
        }
        __syncwarp();
        if (++s == NSTAGE) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: D[d] += A_t B_u^T for all t + u = d < S, two K = 32 steps per 64-byte k-block =====
    int s = 0;
    uint32_t ph = 0, nt = 0;
    // SWIZZLE_64B K-major descriptor: rows of 64 bytes, 8-row groups 512 bytes apart; hi = SBO | version | layout 4
    constexpr uint32_t DESC_HI = (uint32_t)(512 >> 4) | (1u << 14) | (4u << 29);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++nt) {
      mbar_wait(tempty_bar, (nt & 1u) ^ 1u);  // the epilogue has drained the accumulators of the previous tile
      tc_fence_after();
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = base + (uint32_t)s * (uint32_t)STAGE_BYTES;
#pragma unroll
          for (int t = 0; t < S; ++t) {
#pragma unroll
            for (int u = 0; u < S - t; ++u) {
              const uint32_t lo_a = (((sa + t * A_TILE) >> 4) & 0x3FFFu) | (1u << 16);
              const uint32_t lo_b = (((sa + S * A_TILE + u * B_TILE) >> 4) & 0x3FFFu) | (1u << 16);
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                const uint64_t da = ((uint64_t)DESC_HI << 32) | (uint64_t)(lo_a + 2u * kk);
                const uint64_t db = ((uint64_t)DESC_HI << 32) | (uint64_t)(lo_b + 2u * kk);
                // the first product of a diagonal in this tile overwrites: (t == 0, first k-block, first K step)
                const uint32_t accumulate = (t == 0 && kk == 0) ? (uint32_t)(kb != 0) : 1u;
                umma_i8(tmem_base + (uint32_t)((t + u) * BN), da, db, accumulate);
              }
            }
          }
          umma_commit(empty_bar(s));
          if (kb == KB - 1) umma_commit(tfull_bar);
        }
        __syncwarp();
        if (++s == NSTAGE) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else {
    // ===== epilogue: recombine the seven integer accumulators in FP64 and add into C =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    uint32_t nt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++nt) {
      const long long m0 = (long long)(tile / tiles_n) * BM, n0 = (long long)(tile % tiles_n) * BN;
      mbar_wait(tfull_bar, nt & 1u);
      tc_fence_after();
      const int ea = P.ea[m0 + row];
      double *crow = P.C + (m0 + row) * P.ldc + n0;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        double acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.0;
#pragma unroll 1
        for (int d = 0; d < S; ++d) {
          uint32_t v[32];
          tmem_ld32(taddr + (uint32_t)(d * BN + c * 32), v);
          tmem_ld_wait();
          const double scale = ldexp(1.0, -W * (d + 2));
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = fma((double)(int)v[j], scale, acc[j]);  // product exact: same as the CPU model
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int eb = P.eb[n0 + c * 32 + j];
          crow[c * 32 + j] += P.alpha * ldexp(acc[j], ea + eb);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

CUtensorMap make_map(void *base, long long rows, long long k, int box_rows) {
  static encode_tiled_fn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    fn = (encode_tiled_fn)p;
  }
  CUtensorMap map;
  const cuuint64_t gdim[3] = {(cuuint64_t)k, (cuuint64_t)rows, (cuuint64_t)S};
  const cuuint64_t gstride[2] = {(cuuint64_t)k, (cuuint64_t)(k * rows)};
  const cuuint32_t box[3] = {(cuuint32_t)BKB, (cuuint32_t)box_rows, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r);
    exit(2);
  }
  return map;
}

}  // namespace

int main(int argc, char **argv) {
  const long long m = argc > 1 ? atoll(argv[1]) : 1024, n = argc > 2 ? atoll(argv[2]) : 1024, k = argc > 3 ? atoll(argv[3]) : 1024;
  if (m % BM || n % BN || k % BKB) {
    fprintf(stderr, "m, n, k must be multiples of %d, %d, %d\n", BM, BN, BKB);
    return 1;
  }
  std::vector<double> A((size_t)m * k), B((size_t)n * k), C((size_t)m * n, 0.0);
  uint64_t st = 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() {
    st ^= st << 13;
    st ^= st >> 7;
    st ^= st << 17;
    return (double)(st >> 11) * (1.0 / 9007199254740992.0) - 0.5;
  };
  for (auto &x : A) x = rnd() * exp2(8.0 * rnd());  // rows with a spread of magnitudes
  for (auto &x : B) x = rnd() * exp2(8.0 * rnd());

  double *dA, *dB, *dC;
  int8_t *dSA, *dSB;
  int *dEa, *dEb;
  CHECK(cudaMalloc(&dA, A.size() * 8));
  CHECK(cudaMalloc(&dB, B.size() * 8));
  CHECK(cudaMalloc(&dC, C.size() * 8));
  CHECK(cudaMalloc(&dSA, (size_t)S * m * k));
  CHECK(cudaMalloc(&dSB, (size_t)S * n * k));
  CHECK(cudaMalloc(&dEa, m * sizeof(int)));
  CHECK(cudaMalloc(&dEb, n * sizeof(int)));
  CHECK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
  CHECK(cudaMemset(dC, 0, C.size() * 8));

  cudaEvent_t e0, e1, e2;
  CHECK(cudaEventCreate(&e0));
  CHECK(cudaEventCreate(&e1));
  CHECK(cudaEventCreate(&e2));
  int sms = 148;
  CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const CUtensorMap mapA = make_map(dSA, m, k, BM), mapB = make_map(dSB, n, k, BN);
  GemmParams P{m, n, k, n, dC, dEa, dEb, 1.0};
  CHECK(cudaFuncSetAttribute(ozaki_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  for (int rep = 0; rep < 2; ++rep) {  // rep 0 warms up; C accumulates twice, the check below accounts for it
    CHECK(cudaEventRecord(e0));
    slice_rows_kernel<<<(unsigned)((m + 7) / 8), 256>>>(dA, m, k, k, dSA, dEa);
    slice_rows_kernel<<<(unsigned)((n + 7) / 8), 256>>>(dB, n, k, k, dSB, dEb);
    CHECK(cudaEventRecord(e1));
    ozaki_gemm_kernel<<<sms, THREADS, SMEM>>>(mapA, mapB, P);
    CHECK(cudaEventRecord(e2));
    CHECK(cudaDeviceSynchronize());
  }
  float ms_slice = 0, ms_gemm = 0;
  CHECK(cudaEventElapsedTime(&ms_slice, e0, e1));
  CHECK(cudaEventElapsedTime(&ms_gemm, e1, e2));
  CHECK(cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost));

  // ---- CPU model of the same arithmetic on a sample of rows (bit-for-bit), and the plain FP64 product
  auto slice_row = [&](const double *a, std::vector<int8_t> &dig, int &e) {
    double mx = 0;
    for (long long j = 0; j < k; ++j) mx = fmax(mx, fabs(a[j]));
    int e0 = 0;
    frexp(mx, &e0);
    e = mx > 0 ? e0 + 1 : 0;
    dig.assign((size_t)S * k, 0);
    for (long long j = 0; j < k; ++j) {
      double rem = ldexp(a[j], -e);
      for (int t = 0; t < S; ++t) {
        const double scale = ldexp(1.0, W * (t + 1));
        const double d = rint(rem * scale);
        rem -= d / scale;
        dig[(size_t)t * k + j] = (int8_t)(int)d;
      }
    }
  };
  std::vector<std::vector<int8_t>> digB(n);
  std::vector<int> eB(n);
  for (long long j = 0; j < n; ++j) slice_row(&B[(size_t)j * k], digB[j], eB[j]);
  long long mismatches = 0, checked = 0;
  double max_rel = 0, max_ref = 0;
  std::vector<int8_t> digA;
  const long long nsample = (m * n * k > (1ll << 31)) ? 8 : 64;
  for (long long i = 0; i < m; i += (m > nsample ? m / nsample : 1)) {
    int eA;
    slice_row(&A[(size_t)i * k], digA, eA);
    for (long long j = 0; j < n; ++j) {
      double acc = 0.0;
      for (int d = 0; d < S; ++d) {
        long long Dd = 0;
        for (int t = 0; t <= d; ++t) {
          const int8_t *x = &digA[(size_t)t * k], *y = &digB[j][(size_t)(d - t) * k];
          for (long long kk = 0; kk < k; ++kk) Dd += (long long)x[kk] * y[kk];
        }
        acc = fma((double)Dd, ldexp(1.0, -W * (d + 2)), acc);
      }
      const double one = ldexp(acc, eA + eB[j]);
      const double model = one + one;  // two accumulating launches: 0 + x, then x + x (exact doubling)
      double ref = 0;
      for (long long kk = 0; kk < k; ++kk) ref += A[(size_t)i * k + kk] * B[(size_t)j * k + kk];
      const double got = C[(size_t)i * n + j];
      ++checked;
      if (got != model) ++mismatches;
      max_rel = fmax(max_rel, fabs(got - 2 * ref));
      max_ref = fmax(max_ref, fabs(2 * ref));
    }
  }
  const double ops = 2.0 * m * n * k * (S * (S + 1) / 2);
  printf("m=%lld n=%lld k=%lld  slice %.3f ms  gemm %.3f ms  = %.1f INT8 TOP/s = %.1f FP64-equivalent TFLOP/s\n", m, n, k, ms_slice,
         ms_gemm, ops / ms_gemm * 1e-9, 2.0 * m * n * k / ms_gemm * 1e-9);
  printf("checked %lld entries: %lld differ from the CPU digit model (bit for bit); max |err| / max |ref| vs FP64 = %.3e\n", checked,
         mismatches, max_rel / max_ref);
  return mismatches ? 3 : 0;
}
