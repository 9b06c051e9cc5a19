// weights.cu -- neighbour counts -> W = 1/count and Meff = sum W.
//
// Tail of DCAUtils compute_weights (un-vendored; reference call site src/GaussDCA.jl:28).
// W[k] = 1.0/count[k] is a single IEEE division, bit-exact by construction.  Meff = sum_k W[k] is
// accumulated in double-double so the returned double is the correctly rounded value of the exact
// rational sum -- independent of summation order, GPU count or grid shape.
#include "gdca_internal.cuh"

namespace {

struct dd {
  double hi, lo;
};

__device__ __forceinline__ dd dd_add(dd a, dd b) {
  // two_sum of the high parts, then fold the low parts in
  double s = __dadd_rn(a.hi, b.hi);
  double bb = __dadd_rn(s, -a.hi);
  double e = __dadd_rn(__dadd_rn(a.hi, -__dadd_rn(s, -bb)), __dadd_rn(b.hi, -bb));
  e = __dadd_rn(e, __dadd_rn(a.lo, b.lo));
  double hi = __dadd_rn(s, e);
  double lo = __dadd_rn(e, -__dadd_rn(hi, -s));
  return dd{hi, lo};
}

__device__ __forceinline__ dd dd_recip(double c) {
  double q1 = __ddiv_rn(1.0, c);
  double r = __fma_rn(-q1, c, 1.0);  // exact remainder
  double q2 = __ddiv_rn(r, c);
  return dd{q1, q2};
}

constexpr int WB = 256;   // threads per block
constexpr int WG = 128;   // blocks (fixed: the reduction tree must not depend on the device)

__device__ __forceinline__ dd block_reduce_dd(dd v, dd *sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o; o >>= 1) {
    dd other{__shfl_xor_sync(0xffffffffu, v.hi, o), __shfl_xor_sync(0xffffffffu, v.lo, o)};
    v = dd_add(v, other);
  }
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : dd{0.0, 0.0};
    for (int o = 16; o; o >>= 1) {
      dd other{__shfl_xor_sync(0xffffffffu, v.hi, o), __shfl_xor_sync(0xffffffffu, v.lo, o)};
      v = dd_add(v, other);
    }
  }
  return v;  // valid in warp 0
}

// counts == nullptr  ->  theta == 0 path: W = 1, Meff = M
__global__ void __launch_bounds__(WB) weights_kernel(const int32_t *__restrict__ counts, long long M,
                                                     double *__restrict__ W, dd *__restrict__ partial) {
  __shared__ dd sh[WB / 32];
  dd acc{0.0, 0.0};
  for (long long k = (long long)blockIdx.x * WB + threadIdx.x; k < M; k += (long long)WG * WB) {
    const double c = counts ? (double)(counts[k] + 1) : 1.0;  // + the sequence itself
    W[k] = __ddiv_rn(1.0, c);
    acc = dd_add(acc, dd_recip(c));
  }
  acc = block_reduce_dd(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(WG) meff_final_kernel(const dd *__restrict__ partial, double *__restrict__ out) {
  __shared__ dd sh[WG / 32];
  dd v = partial[threadIdx.x];
  v = block_reduce_dd(v, sh);
  if (threadIdx.x == 0) {
    out[0] = __dadd_rn(v.hi, v.lo);
    out[1] = v.lo;
  }
}

}  // namespace

// which: row of the counts buffer (0..2), or -1 for the theta == 0 path (no pair sweep at all).
int32_t gdca_k_finish_weights(gdca_ctx *ctx, int which) {
  if (!ctx->have_alignment) return gdca_fail(ctx, GDCA_ERR_STATE, "finish_weights: no alignment loaded");
  if (which < -1 || which > 2) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "finish_weights: which out of range");
  GDCA_TRY(gdca_reserve(ctx, ctx->dW, ctx->capW, (size_t)ctx->Mpad));
  GDCA_TRY(gdca_reserve(ctx, ctx->dRed, ctx->capRed, (size_t)4096));
  const int32_t *cnt = (which < 0) ? nullptr : ctx->dCounts + (size_t)which * ctx->Mpad;
  weights_kernel<<<WG, WB, 0, ctx->stream>>>(cnt, ctx->M, ctx->dW, reinterpret_cast<dd *>(ctx->dRed));
  GDCA_LAUNCH_CHECK(ctx);
  meff_final_kernel<<<1, WG, 0, ctx->stream>>>(reinterpret_cast<const dd *>(ctx->dRed), ctx->dMeff);
  GDCA_LAUNCH_CHECK(ctx);
  GDCA_CUDA(ctx, cudaMemcpyAsync(&ctx->meff, ctx->dMeff, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->have_weights = true;
  ctx->counts_row = which;
  ctx->weights_from_counts = true;
  return GDCA_OK;
}
