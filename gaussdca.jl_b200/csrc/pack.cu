// pack.cu -- K1: encoded alignment -> bit-plane ("bit-sliced") layout for the pair sweep.
//
// Replaces DCAUtils compress_Z (un-vendored; reached from compute_weighted_frequencies, reference call
// site src/GaussDCA.jl:28), which packs 12 residues x 5 bits per UInt64.  That layout costs ~5 ALU
// ops per 6 residues when two sequences are compared.  Here the 5 bits of every residue code are
// stored in 5 separate planes, 32 sites per 32-bit word:
//
//      planes[w][p][k]  bit b  =  bit p of Z[site 32w+b, sequence k]          (k fastest)
//
// so comparing 32 sites of two sequences is 5 LOP3 (x |= a_p ^ b_p) + 1 POPC: 0.19 op / residue
// instead of 0.83.  Sites beyond L and sequences beyond M are zero in every plane (never differ).
//
// Also here: q = max(Z) (src/GaussDCA.jl:25).
#include "gdca_internal.cuh"

__global__ void maxq_kernel(const int8_t *__restrict__ Z, size_t nbytes, int *__restrict__ out) {
  int m = 0;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t nvec = nbytes / 16;
  const int4 *Z4 = reinterpret_cast<const int4 *>(Z);
  for (size_t v = i; v < nvec; v += stride) {
    int4 x = Z4[v];
    unsigned w[4] = {(unsigned)x.x, (unsigned)x.y, (unsigned)x.z, (unsigned)x.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      // per-byte max via __vmaxu4 against the running max replicated in all bytes
      unsigned mm = __vmaxu4(w[t], (unsigned)m * 0x01010101u);
      mm = max(max(mm & 0xff, (mm >> 8) & 0xff), max((mm >> 16) & 0xff, mm >> 24));
      m = (int)mm;
    }
  }
  for (size_t b = nvec * 16 + i; b < nbytes; b += stride) m = max(m, (int)(uint8_t)Z[b]);
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

int32_t gdca_k_maxq(gdca_ctx *ctx) {
  GDCA_CUDA(ctx, cudaMemsetAsync(ctx->dQ, 0, sizeof(int), ctx->stream));
  const size_t nbytes = (size_t)ctx->L * ctx->M;
  maxq_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(ctx->dZ, nbytes, ctx->dQ);
  GDCA_LAUNCH_CHECK(ctx);
  return GDCA_OK;
}

// One CTA = 32 sequences (one per warp).  Each warp walks its sequence 32 sites at a time, turns
// the 5 code bits into 5 ballots, parks them in shared memory; then the CTA writes 128-byte rows.
__global__ void __launch_bounds__(1024) pack_planes_kernel(const int8_t *__restrict__ Z, int64_t L, int64_t M,
                                                           int64_t Mpad, int nwords, int nplanes,
                                                           uint32_t *__restrict__ planes) {
  extern __shared__ uint32_t tile[];  // [nwords][nplanes][32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * 32 + warp;
  for (int w = 0; w < nwords; ++w) {
    const int64_t site = (int64_t)w * 32 + lane;
    unsigned v = 0;
    if (k < M && site < L) v = (unsigned)(uint8_t)Z[k * L + site] & 31u;
    unsigned mine = 0;
#pragma unroll
    for (int p = 0; p < GDCA_MAX_PLANES; ++p) {
      unsigned bal = __ballot_sync(0xffffffffu, (v >> p) & 1u);
      if (lane == p) mine = bal;
    }
    if (lane < nplanes) tile[(w * nplanes + lane) * 32 + warp] = mine;
  }
  __syncthreads();
  const int rows = nwords * nplanes;
  for (int r = warp; r < rows; r += 32) {
    planes[(int64_t)r * Mpad + (int64_t)blockIdx.x * 32 + lane] = tile[r * 32 + lane];
  }
}

int32_t gdca_k_pack(gdca_ctx *ctx) {
  const int64_t Mpad = ctx->Mpad;
  const size_t words = (size_t)ctx->nwords * ctx->nplanes * Mpad;
  GDCA_TRY(gdca_reserve(ctx, ctx->dPlanes, ctx->capPlanes, words));
  const size_t smem = (size_t)ctx->nwords * ctx->nplanes * 32 * sizeof(uint32_t);
  if (smem > 48 * 1024) {
    GDCA_CUDA(ctx, cudaFuncSetAttribute(pack_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  pack_planes_kernel<<<(unsigned)(Mpad / 32), 1024, smem, ctx->stream>>>(ctx->dZ, ctx->L, ctx->M, Mpad,
                                                                          (int)ctx->nwords, ctx->nplanes, ctx->dPlanes);
  GDCA_LAUNCH_CHECK(ctx);
  return GDCA_OK;
}
