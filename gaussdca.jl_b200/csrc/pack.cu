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
// "site 32w+b" is a position of the PACKED order (most variable columns first, see site_order_kernel).
//
// Also here: q = max(Z) (src/GaussDCA.jl:25).
#include "gdca_internal.cuh"

// out[0] = max(Z), out[1] = min(Z) as SIGNED bytes (Julia's maximum(Z) on a Matrix{Int8}); the caller's pointer may sit
// at any byte offset (gdca_run_resident takes device views), so a scalar head runs up to the first 16-byte boundary.
__global__ void maxq_kernel(const int8_t *__restrict__ Z, size_t nbytes, int *__restrict__ out) {
  int mx = -128, mn = 127;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t head = (16 - ((size_t)Z & 15)) & 15;
  if (head > nbytes) head = nbytes;
  const size_t nvec = (nbytes - head) / 16;
  const int4 *Z4 = reinterpret_cast<const int4 *>(Z + head);
  unsigned vmx = 0x80808080u, vmn = 0x7f7f7f7fu;  // per-byte signed running max / min
  for (size_t v = i; v < nvec; v += stride) {
    const int4 x = Z4[v];
    vmx = __vmaxs4(__vmaxs4(vmx, (unsigned)x.x), __vmaxs4(__vmaxs4((unsigned)x.y, (unsigned)x.z), (unsigned)x.w));
    vmn = __vmins4(__vmins4(vmn, (unsigned)x.x), __vmins4(__vmins4((unsigned)x.y, (unsigned)x.z), (unsigned)x.w));
  }
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    mx = max(mx, (int)(int8_t)(vmx >> (8 * b)));
    mn = min(mn, (int)(int8_t)(vmn >> (8 * b)));
  }
  for (size_t b = i; b < head; b += stride) mx = max(mx, (int)Z[b]), mn = min(mn, (int)Z[b]);
  for (size_t b = head + nvec * 16 + i; b < nbytes; b += stride) mx = max(mx, (int)Z[b]), mn = min(mn, (int)Z[b]);
  for (int o = 16; o; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out, mx);
    atomicMin(out + 1, mn);
  }
}

int32_t gdca_k_maxq(gdca_ctx *ctx) {
  const int init[2] = {-128, 127};
  GDCA_CUDA(ctx, cudaMemcpyAsync(ctx->dQ, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
  const size_t nbytes = (size_t)ctx->L * ctx->M;
  maxq_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(ctx->dZ, nbytes, ctx->dQ);
  GDCA_LAUNCH_CHECK(ctx);
  return GDCA_OK;
}

// ---- site order of the packed planes ---------------------------------------------------------------------
// Hamming distance does not depend on the order of the sites, but the early exit of the count sweep does: a
// pair leaves the sweep as soon as its PARTIAL distance reaches thresh.  Sites are therefore packed in order of
// decreasing mismatch probability  1 - sum_v (n_iv / M)^2  (n_iv = per-site state counts = bucket sizes of the
// per-site lists): conserved and gap-dominated columns -- common in Pfam alignments -- go last.  Ties keep
// site order, so the permutation is deterministic.  (A no-op for i.i.d. columns such as the synthetic configs.)
__global__ void site_order_kernel(const int32_t *__restrict__ listoff, int L, int nstate_p1, int32_t *__restrict__ perm) {
  extern __shared__ unsigned long long key[];  // sum_v n_iv^2 per site (smaller = more variable)
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    unsigned long long a = 0;
    for (int v = 0; v + 1 < nstate_p1; ++v) {
      const unsigned long long nv = (unsigned long long)(listoff[i * nstate_p1 + v + 1] - listoff[i * nstate_p1 + v]);
      a += nv * nv;
    }
    key[i] = a;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const unsigned long long ki = key[i];
    int rank = 0;
    for (int j = 0; j < L; ++j) rank += (key[j] < ki) || (key[j] == ki && j < i);
    perm[rank] = i;  // position rank of the packed order holds site i
  }
}

__global__ void identity_order_kernel(int L, int32_t *__restrict__ perm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < L) perm[i] = i;
}

// One CTA = 32 sequences (one per warp).  Each warp walks its sequence 32 packed positions at a time, turns
// the 5 code bits into 5 ballots, parks them in shared memory; then the CTA writes 128-byte rows.
__global__ void __launch_bounds__(1024) pack_planes_kernel(const int8_t *__restrict__ Z, int64_t L, int64_t M,
                                                           int64_t Mpad, int nwords, int nplanes,
                                                           const int32_t *__restrict__ perm,
                                                           uint32_t *__restrict__ planes) {
  extern __shared__ uint32_t tile[];  // [nwords][nplanes][32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * 32 + warp;
  for (int w = 0; w < nwords; ++w) {
    const int64_t pos = (int64_t)w * 32 + lane;
    unsigned v = 0;
    if (k < M && pos < L) v = (unsigned)(uint8_t)Z[k * L + perm[pos]] & 31u;
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
  // site order from the per-site state histograms (built here if they are not there yet)
  GDCA_TRY(gdca_k_site_hist(ctx));
  GDCA_TRY(gdca_reserve(ctx, ctx->dPerm, ctx->capPerm, (size_t)ctx->L));
  const size_t osmem = (size_t)ctx->L * sizeof(unsigned long long);
  if (osmem <= 64 * 1024) {
    if (osmem > 48 * 1024)
      GDCA_CUDA(ctx, cudaFuncSetAttribute(site_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)osmem));
    site_order_kernel<<<1, 1024, osmem, ctx->stream>>>(ctx->dListOff, (int)ctx->L, 33, ctx->dPerm);
  } else {  // L > 8192: keep the natural order (the O(L^2) ranking kernel is meant for alignment-sized L)
    identity_order_kernel<<<(unsigned)((ctx->L + 255) / 256), 256, 0, ctx->stream>>>((int)ctx->L, ctx->dPerm);
  }
  GDCA_LAUNCH_CHECK(ctx);
  const size_t smem = (size_t)ctx->nwords * ctx->nplanes * 32 * sizeof(uint32_t);
  if (smem > 48 * 1024) {
    GDCA_CUDA(ctx, cudaFuncSetAttribute(pack_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  pack_planes_kernel<<<(unsigned)(Mpad / 32), 1024, smem, ctx->stream>>>(ctx->dZ, ctx->L, ctx->M, Mpad,
                                                                          (int)ctx->nwords, ctx->nplanes, ctx->dPerm,
                                                                          ctx->dPlanes);
  GDCA_LAUNCH_CHECK(ctx);
  return GDCA_OK;
}
