// rank.cu -- K8 correct_APC (reference src/GaussDCA.jl:78-86) and K9 compute_ranking (:88-99).
//
//   APC:      Si = sum(S, dims=1), Sj = sum(S, dims=2), Sa = sum(S) * (1 - 1/N);  S -= (Sj*Si)/Sa
//             (the zero diagonal takes part in the sums, exactly as in the reference).
//   ranking:  for i = 1:N-ms, j = i+ms:N  ->  (i, j, S[j,i]);  stable sort, descending score.
//             Stability = exact ties keep enumeration order (i-major, j-minor).  The device sort
//             orders the unique composite key (score descending, i*N+j ascending), which is the
//             same total order, with a bitonic network (shared-memory stages fused per 2048 keys).
#include "gdca_internal.cuh"

namespace {

// deterministic row sums (warps 0..L-1) and column sums (warps L..2L-1): one warp per line, the SAME association of the
// additions for both, so a symmetric S gives bit-identical row and column sums and the corrected matrix stays symmetric
__global__ void line_sums_kernel(const double *__restrict__ S, int L, double *__restrict__ rows, double *__restrict__ cols) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= 2 * L) return;
  const bool col = warp >= L;
  const int line = col ? warp - L : warp;
  double acc = 0.0;
  for (int c = lane; c < L; c += 32) acc += col ? S[(long long)c * L + line] : S[(long long)line * L + c];
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) (col ? cols : rows)[line] = acc;
}

__global__ void __launch_bounds__(1024) total_kernel(const double *__restrict__ rows, int L, double *__restrict__ out) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int r = threadIdx.x; r < L; r += 1024) acc += rows[r];
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = sh[threadIdx.x];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) out[0] = acc;
  }
}

// out[r][c] = S[r][c] - Sj[r] * Si[c] / Sa with Sj = sum(S, dims=2) (row sums) and Si = sum(S, dims=1) (column sums),
// src/GaussDCA.jl:80-84 -- correct for a non-symmetric S too (the formula reads the same in Julia's column-major view)
__global__ void apc_kernel(const double *__restrict__ S, const double *__restrict__ rows, const double *__restrict__ cols,
                           const double *__restrict__ tot, int L, double *__restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)L * L) return;
  const int r = (int)(e / L), c = (int)(e - (long long)r * L);
  const double Sa = tot[0] * (1.0 - 1.0 / L);
  out[e] = S[e] - (rows[r] * cols[c]) / Sa;
}

__device__ __forceinline__ unsigned long long desc_key(double x) {
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
  if (x != x) b = 0x7FF8000000000000ull;  // Julia's isless: every NaN, whatever its sign bit, sorts above +Inf and NaNs tie
  const unsigned long long asc = b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull);
  return ~asc;  // ascending sort of this key == descending score
}

// enumerate i = 0..L-ms-1, j = i+ms..L-1 (0-based) in i-major order
__global__ void enumerate_kernel(const double *__restrict__ S, int L, int ms, long long npairs, long long P,
                                 unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P) return;
  if (e >= npairs) {
    keys[e] = ~0ull;
    vals[e] = 0xffffffffu;
    return;
  }
  // row i holds (L - ms - i) pairs; offset(i) = i*(L-ms) - i*(i-1)/2
  const double T = (double)(L - ms) + 0.5;
  int i = (int)(T - sqrt(T * T - 2.0 * (double)e));
  if (i < 0) i = 0;
  if (i > L - ms - 1) i = L - ms - 1;
  while (true) {
    const long long off = (long long)i * (L - ms) - (long long)i * (i - 1) / 2;
    if (off > e) { --i; continue; }
    if (e >= off + (L - ms - i)) { ++i; continue; }
    const int j = i + ms + (int)(e - off);
    keys[e] = desc_key(S[(long long)j * L + i]);
    vals[e] = (uint32_t)((long long)i * L + j);
    return;
  }
}

__device__ __forceinline__ bool kv_less(unsigned long long ka, uint32_t va, unsigned long long kb, uint32_t vb) {
  return (ka < kb) || (ka == kb && va < vb);
}

// one global compare-exchange stage (k, jj)
__global__ void bitonic_global_kernel(unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals, long long P,
                                      long long k, long long jj) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per pair
  if (t >= P / 2) return;
  const long long i = ((t / jj) * 2 * jj) + (t % jj);
  const long long l = i + jj;
  const bool up = ((i & k) == 0);
  const unsigned long long ki = keys[i], kl = keys[l];
  const uint32_t vi = vals[i], vl = vals[l];
  const bool swap = up ? kv_less(kl, vl, ki, vi) : kv_less(ki, vi, kl, vl);
  if (swap) {
    keys[i] = kl; keys[l] = ki;
    vals[i] = vl; vals[l] = vi;
  }
}

constexpr int SORT_CHUNK = 2048;  // keys per CTA in the shared-memory stages

// all stages with jj < SORT_CHUNK for a given k (or the full local sort when k_lo == 2)
__global__ void __launch_bounds__(SORT_CHUNK / 2) bitonic_shared_kernel(unsigned long long *__restrict__ keys,
                                                                         uint32_t *__restrict__ vals, long long k_lo,
                                                                         long long k_hi) {
  __shared__ unsigned long long sk[SORT_CHUNK];
  __shared__ uint32_t sv[SORT_CHUNK];
  const long long base = (long long)blockIdx.x * SORT_CHUNK;
  const int t = threadIdx.x;
  sk[t] = keys[base + t];
  sv[t] = vals[base + t];
  sk[t + SORT_CHUNK / 2] = keys[base + t + SORT_CHUNK / 2];
  sv[t + SORT_CHUNK / 2] = vals[base + t + SORT_CHUNK / 2];
  __syncthreads();
  for (long long k = k_lo; k <= k_hi; k <<= 1) {
    for (int jj = (int)((k >> 1) < SORT_CHUNK / 2 ? (k >> 1) : SORT_CHUNK / 2); jj > 0; jj >>= 1) {
      const int i = ((t / jj) * 2 * jj) + (t % jj);
      const int l = i + jj;
      const bool up = (((base + i) & k) == 0);
      const unsigned long long ki = sk[i], kl = sk[l];
      const uint32_t vi = sv[i], vl = sv[l];
      const bool swap = up ? kv_less(kl, vl, ki, vi) : kv_less(ki, vi, kl, vl);
      if (swap) {
        sk[i] = kl; sk[l] = ki;
        sv[i] = vl; sv[l] = vi;
      }
      __syncthreads();
    }
  }
  keys[base + t] = sk[t];
  vals[base + t] = sv[t];
  keys[base + t + SORT_CHUNK / 2] = sk[t + SORT_CHUNK / 2];
  vals[base + t + SORT_CHUNK / 2] = sv[t + SORT_CHUNK / 2];
}

__global__ void emit_kernel(const uint32_t *__restrict__ vals, const double *__restrict__ S, int L, long long npairs,
                            gdca_rank_t *__restrict__ R) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= npairs) return;
  const uint32_t v = vals[e];
  const int i = (int)(v / (uint32_t)L), j = (int)(v - (uint32_t)i * (uint32_t)L);
  gdca_rank_t r;
  r.i = i + 1;
  r.j = j + 1;
  r.score = S[(long long)j * L + i];
  R[e] = r;
}

}  // namespace

int32_t gdca_k_apc(gdca_ctx *ctx) {
  const int L = (int)ctx->L;
  GDCA_TRY(gdca_reserve(ctx, ctx->dS2, ctx->capS2, (size_t)L * L));
  GDCA_TRY(gdca_reserve(ctx, ctx->dRed, ctx->capRed, (size_t)2 * L + 4096));
  double *rows = ctx->dRed, *cols = ctx->dRed + L, *tot = ctx->dRed + 2 * L;
  line_sums_kernel<<<(unsigned)((2 * L * 32 + 255) / 256), 256, 0, ctx->stream>>>(ctx->dS, L, rows, cols);
  GDCA_LAUNCH_CHECK(ctx);
  total_kernel<<<1, 1024, 0, ctx->stream>>>(rows, L, tot);
  GDCA_LAUNCH_CHECK(ctx);
  const long long ne = (long long)L * L;
  apc_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, ctx->stream>>>(ctx->dS, rows, cols, tot, L, ctx->dS2);
  GDCA_LAUNCH_CHECK(ctx);
  return GDCA_OK;
}

int32_t gdca_k_rank(gdca_ctx *ctx, int64_t min_sep, int64_t R_len) {
  const int L = (int)ctx->L;
  if (min_sep < 1) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "invalid min_separation value (must be >= 1)");
  if (L >= 65536) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "ranking: L must be < 65536");
  const long long npairs = gdca_ranking_length(L, min_sep);
  if (npairs != R_len) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "ranking: R_len != (L-ms)*(L-ms+1)/2");
  if (npairs == 0) return GDCA_OK;
  long long P = SORT_CHUNK;
  while (P < npairs) P <<= 1;
  GDCA_TRY(gdca_reserve(ctx, ctx->dKeys, ctx->capKeys, (size_t)P));
  GDCA_TRY(gdca_reserve(ctx, ctx->dVals, ctx->capVals, (size_t)P));
  GDCA_TRY(gdca_reserve(ctx, ctx->dR, ctx->capR, (size_t)npairs));
  enumerate_kernel<<<(unsigned)((P + 255) / 256), 256, 0, ctx->stream>>>(ctx->dS2, L, (int)min_sep, npairs, P,
                                                                         ctx->dKeys, ctx->dVals);
  GDCA_LAUNCH_CHECK(ctx);
  const unsigned nchunks = (unsigned)(P / SORT_CHUNK);
  // local sort of every 2048-key chunk (alternating directions come from the global index)
  bitonic_shared_kernel<<<nchunks, SORT_CHUNK / 2, 0, ctx->stream>>>(ctx->dKeys, ctx->dVals, 2, SORT_CHUNK);
  GDCA_LAUNCH_CHECK(ctx);
  for (long long k = 2 * SORT_CHUNK; k <= P; k <<= 1) {
    for (long long jj = k >> 1; jj >= SORT_CHUNK; jj >>= 1) {
      bitonic_global_kernel<<<(unsigned)((P / 2 + 255) / 256), 256, 0, ctx->stream>>>(ctx->dKeys, ctx->dVals, P, k, jj);
      GDCA_LAUNCH_CHECK(ctx);
    }
    bitonic_shared_kernel<<<nchunks, SORT_CHUNK / 2, 0, ctx->stream>>>(ctx->dKeys, ctx->dVals, k, k);
    GDCA_LAUNCH_CHECK(ctx);
  }
  emit_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, ctx->stream>>>(ctx->dVals, ctx->dS2, L, npairs, ctx->dR);
  GDCA_LAUNCH_CHECK(ctx);
  return GDCA_OK;
}
