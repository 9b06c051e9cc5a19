// cov.cu -- K4: weighted two-point frequencies + pseudocount + covariance, fused.
//
// Replaces, in one pass over the alignment:
//   DCAUtils compute_freqs   (Pi_true, Pij_true; un-vendored, reference call site src/GaussDCA.jl:28)
//   DCAUtils add_pseudocount (call site src/GaussDCA.jl:30)
//   compute_C(Pi, Pij) = Pij - Pi*Pi'            (src/GaussDCA.jl:32,76)
//
// Pij_true = X' W X / Meff with X the M x n one-hot matrix.  As a dense FP64 contraction that is
// M*n^2 = 2e13 flop at L=500, M=200k (0.5 s at DMMA peak) -- but X has exactly one 1 per site, so only
// 1/400 of the products are non-zero.  This kernel does the M*L^2/2 = 2.5e10 useful additions directly:
//
//   1. per site i, sequence ids are grouped by the state at i (stable counting sort -> list(i,a));
//   2. one CTA owns output row (i,a) x a chunk of 512 sites starting at the diagonal block; thread t owns 4
//      adjacent sites.  For every k in list(i,a) it adds W[k] to PRIVATE shared-memory accumulators
//      acc[Z[j,k]][u][t] (list entries are staged 128 at a time in shared memory; the four states come from
//      one 32-bit load of a recoded copy of Z).  No atomics, no
//      bank conflicts (thread t always hits bank pair t mod 16), every element is summed in ascending
//      sequence order, so the result is deterministic and bit-identical for (r,c) and (c,r);
//   3. the epilogue applies 1/Meff, the pseudocount mix and "- Pi Pi'" and writes the row chunk.
//
// Only site blocks j >= i are computed; symmetrize_C mirrors them.  The bound is shared-memory
// bandwidth (one 8-byte read-modify-write per addition), not the FP64 pipe and not HBM.
#include "gdca_internal.cuh"

namespace {

constexpr int JT = 128;      // sites (threads) per covariance CTA
constexpr int LB = 1024;     // threads of the list-building CTA (32 warps, one contiguous chunk of the sequences each)
constexpr int NSTATE = 32;   // states are 1..31 (q <= 31)

// ---- Z [M][L]  ->  Zt [L][M]  (site-major) so one site's column can be streamed coalesced ----
__global__ void transpose_Z_kernel(const int8_t *__restrict__ Z, long long L, long long M, int8_t *__restrict__ Zt) {
  __shared__ int8_t tile[64][65];
  const long long k0 = (long long)blockIdx.x * 64, i0 = (long long)blockIdx.y * 64;
  for (int e = threadIdx.x; e < 64 * 64; e += blockDim.x) {
    const int kk = e >> 6, ii = e & 63;
    if (k0 + kk < M && i0 + ii < L) tile[kk][ii] = Z[(k0 + kk) * L + i0 + ii];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 64 * 64; e += blockDim.x) {
    const int ii = e >> 6, kk = e & 63;
    if (k0 + kk < M && i0 + ii < L) Zt[(i0 + ii) * M + k0 + kk] = tile[kk][ii];
  }
}

// ---- per-site state histograms, no sort: listoff[i][v] = #{k : Z[i,k] < v} (the bucket offsets of the per-site lists) ----
// All that theta = :auto (ident_sum_kernel), the site order of the bit planes and the tensor-core covariance need of the lists.
// One CTA per site; a thread counts four sequences per instruction triple (byte-wise compare of a 32-bit word against the
// replicated state, popc) into NS private registers; warp shuffles + one shared-memory pass finish the site.
template <int NS>
__global__ void __launch_bounds__(256) site_hist_kernel(const int8_t *__restrict__ Zt, long long M, int32_t *__restrict__ listoff) {
  __shared__ int part[8][NSTATE];
  const long long i = blockIdx.x;
  const uint8_t *z = reinterpret_cast<const uint8_t *>(Zt) + i * M;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int acc[NS];
#pragma unroll
  for (int v = 0; v < NS; ++v) acc[v] = 0;
  // bytes up to the first 4-byte boundary and behind the last whole word: one thread each, into the same counters
  const long long head = (4 - (reinterpret_cast<uintptr_t>(z) & 3)) & 3;
  const long long h = head < M ? head : M;
  const long long nwords = (M - h) / 4;
  const uint32_t *zw = reinterpret_cast<const uint32_t *>(z + h);
  for (long long w = tid; w < nwords; w += 256) {
    const uint32_t x = zw[w];
#pragma unroll
    for (int v = 0; v < NS; ++v) acc[v] += __popc(__vcmpeq4(x, 0x01010101u * (uint32_t)v)) >> 3;
  }
  auto one = [&](long long k) {
    const int st = (int)z[k];
#pragma unroll
    for (int v = 0; v < NS; ++v) acc[v] += (st == v);
  };
  if (tid < h) one(tid);
  const long long tail0 = h + 4 * nwords;
  if (tail0 + tid < M && tid < 4) one(tail0 + tid);
#pragma unroll
  for (int v = 0; v < NS; ++v) {
    int a = acc[v];
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) part[warp][v] = a;
  }
  if (NS < NSTATE && tid < 8 * (NSTATE - NS)) part[tid / (NSTATE - NS)][NS + tid % (NSTATE - NS)] = 0;
  __syncthreads();
  if (warp == 0) {
    int tot = 0;
    for (int w = 0; w < 8; ++w) tot += part[w][lane];
    int incl = tot;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    listoff[i * (NSTATE + 1) + lane] = incl - tot;
    if (lane == 31) listoff[i * (NSTATE + 1) + NSTATE] = incl;
  }
}

// ---- per-site stable counting sort of sequence ids by state ----
// One CTA of 32 warps per site; warp w owns the contiguous chunk [w c, (w+1) c) of the sequences.  Pass 1: per-warp state
// histograms (match_any, no CTA barrier in the loop); one prefix over (state, warp) turns them into every warp's first output
// slot per state; pass 2: each warp scatters its chunk in order.  Ascending sequence ids inside every bucket (stable), and only
// three CTA barriers per site instead of four per 256 sequences (0.90 -> 0.1 ms at config C).
__global__ void __launch_bounds__(LB) build_lists_kernel(const int8_t *__restrict__ Zt, long long M,
                                                         int32_t *__restrict__ list, int32_t *__restrict__ listoff) {
  __shared__ int cnt[LB / 32][NSTATE];
  const long long i = blockIdx.x;
  const int8_t *z = Zt + i * M;
  int32_t *out = list + i * M;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long chunk = ((M + (LB / 32) - 1) / (LB / 32) + 127) / 128 * 128;
  const long long k0 = (long long)warp * chunk, k1 = k0 + chunk < M ? k0 + chunk : M;
  cnt[warp][lane] = 0;
  __syncwarp();
  // four 32-sequence steps per round: their loads (one 128-byte line) are in flight together
  for (long long base = k0; base < k1; base += 128) {
    unsigned vv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long k = base + 32 * u + lane;
      vv[u] = k < k1 ? ((unsigned)z[k] & 31u) : 32u + (unsigned)lane;  // invalid lanes match nobody
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned m = __match_any_sync(0xffffffffu, vv[u]);
      if (vv[u] < 32u && (m & ((1u << lane) - 1u)) == 0) cnt[warp][vv[u]] += __popc(m);   // one lane per state present
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == 0) {
    // lane v: bucket sizes -> exclusive prefix over the states, then over the warps inside the bucket
    int tot = 0;
    for (int w = 0; w < LB / 32; ++w) tot += cnt[w][lane];
    int incl = tot;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    int run = incl - tot;
    listoff[i * (NSTATE + 1) + lane] = run;
    if (lane == 31) listoff[i * (NSTATE + 1) + NSTATE] = incl;
    for (int w = 0; w < LB / 32; ++w) {
      const int c = cnt[w][lane];
      cnt[w][lane] = run;
      run += c;
    }
  }
  __syncthreads();
  for (long long base = k0; base < k1; base += 128) {
    unsigned vv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long k = base + 32 * u + lane;
      vv[u] = k < k1 ? ((unsigned)z[k] & 31u) : 32u + (unsigned)lane;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned m = __match_any_sync(0xffffffffu, vv[u]);
      const int rank = __popc(m & ((1u << lane) - 1u));
      const bool valid = vv[u] < 32u;
      if (valid) out[cnt[warp][vv[u]] + rank] = (int32_t)(base + 32 * u + lane);
      __syncwarp();
      if (valid && rank == 0) cnt[warp][vv[u]] += __popc(m);
      __syncwarp();
    }
  }
}

// ---- sum over pairs k<l of #identical positions, WITHOUT visiting any pair ---------------------------------
// A position i contributes one identity for every pair of sequences that carry the same state there:
//     sum_{k<l} ident(k,l) = sum_i sum_v n_iv (n_iv - 1) / 2,     n_iv = #{k : Z[i,k] = v}  (gap state included).
// The n_iv are the bucket sizes of the per-site lists, so theta = :auto costs O(M L), not O(M^2 L).
__global__ void __launch_bounds__(256) ident_sum_kernel(const int32_t *__restrict__ listoff, long long L,
                                                        unsigned long long *__restrict__ out) {
  unsigned long long acc = 0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < L * NSTATE; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / NSTATE;
    const int v = (int)(e - i * NSTATE);
    const unsigned long long nv = (unsigned long long)(listoff[i * (NSTATE + 1) + v + 1] - listoff[i * (NSTATE + 1) + v]);
    acc += nv * (nv - 1) / 2;
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// ---- Pi[(i,a)] = (1-pc) * sum_{k in list(i,a)} W[k] / Meff + pc/q  (deterministic tree) ----
__global__ void __launch_bounds__(256) pi_kernel(const int32_t *__restrict__ list, const int32_t *__restrict__ listoff,
                                                 const double *__restrict__ W, const double *__restrict__ meff,
                                                 long long M, int q, double pc, double *__restrict__ Pi) {
  const long long i = blockIdx.x;
  const int s = q - 1, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double Meff = meff[0];
  for (int a = 1 + warp; a <= s; a += (int)(blockDim.x >> 5)) {
    const int beg = listoff[i * (NSTATE + 1) + a], end = listoff[i * (NSTATE + 1) + a + 1];
    const int32_t *l = list + i * M;
    double acc = 0.0;
    for (int e = beg + lane; e < end; e += 32) acc += W[l[e]];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) Pi[i * s + (a - 1)] = (1.0 - pc) * (acc / Meff) + pc / q;
  }
}

// ---- Zq: recoded copy of the alignment for the covariance kernel --------------------------------------
// Zq[k][j] = accumulator SLOT of sequence k at site j: state-1, or s (the dump slot) for the gap state and for
// the padding sites j >= L.  Row stride Lq is a multiple of 128, so a thread reads its SPT adjacent sites with
// one aligned load and the slot needs no decoding in the inner loop.
__global__ void build_zq_kernel(const int8_t *__restrict__ Z, long long L, long long M, long long Lq, int s,
                                uint8_t *__restrict__ Zq) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M * Lq) return;
  const long long k = e / Lq;
  const long long j = e - k * Lq;
  int slot = s;
  if (j < L) {
    const int st = (int)Z[k * L + j];
    if (st >= 1 && st <= s) slot = st - 1;
  }
  Zq[e] = (uint8_t)slot;
}

struct CovParams {
  const uint8_t *Zq;  // [M][Lq]
  const int32_t *list, *listoff;
  const double *W, *meff, *Pi;
  double *C;  // [npad][npad], leading dimension ld
  long long L, M, n, ld, Lq;
  int q, s, nchunks;
  int rank, world;
  double pc;
};

constexpr int UNR = 8;
constexpr int STG = 128;  // list entries staged per round (one per thread)

template <int SPT> struct ZLoad;
template <> struct ZLoad<2> { typedef unsigned short type; };
template <> struct ZLoad<4> { typedef unsigned int type; };

// CTA (row r = (i,a), chunk c): sites [start + c*CH, start + (c+1)*CH), CH = JT*SPT, start = i rounded down to a
// warp's worth of sites (32*SPT).  Thread t owns the SPT adjacent sites start + c*CH + SPT*t + u.
// GRID2D: (chunks, rows) grid (rows <= 65535); otherwise a 1-D grid, row-major over (row, chunk), for L > 3276 at q = 21.
// (A compile-time switch on purpose: the common instantiation keeps the exact code -- 48 registers, 19.3 ms at config C -- that a
// run-time selection of the two index computations lost to a 40-register schedule, 21.4 ms.)
template <int SPT, bool RAW, bool GRID2D>
__global__ void __launch_bounds__(JT) cov_rows_kernel(CovParams P) {
  constexpr int CH = JT * SPT;
  extern __shared__ double acc[];      // [q][SPT][JT]; slot s is the dump for the gap state / padding
  __shared__ unsigned stage_off[STG];  // row offset k*Lq of the staged sequences
  __shared__ double stage_w[STG];      // W[k]
  const int r = GRID2D ? (int)blockIdx.y : (int)(blockIdx.x / (unsigned)P.nchunks);  // output row (i, a)
  const int chunk = GRID2D ? (int)blockIdx.x : (int)(blockIdx.x - (unsigned)r * (unsigned)P.nchunks);
  const int i = r / P.s, a = r - i * P.s + 1;
  if (P.world > 1 && (i % P.world) != P.rank) return;  // rows are dealt to ranks by site
  const int start = (i / (32 * SPT)) * (32 * SPT) + chunk * CH;
  if (start >= P.L) return;
  const int t = threadIdx.x;
  const int j0 = start + SPT * t;                 // first site of this thread
  const bool active = ((t >> 5) * 32 * SPT + start) < (int)P.Lq;   // warp-uniform: warp has sites inside the row
  for (int b = 0; b < SPT * P.q; ++b) acc[b * JT + t] = 0.0;
  // (private columns: no barrier needed before the accumulation loop)

  const int beg = P.listoff[(long long)i * (NSTATE + 1) + a], end = P.listoff[(long long)i * (NSTATE + 1) + a + 1];
  const int32_t *l = P.list + (long long)i * P.M;
  const uint8_t *Zt = P.Zq + j0;
  const unsigned Lq = (unsigned)P.Lq;
  double *acct = acc + t;
  constexpr int slot_stride = SPT * JT;
  typedef typename ZLoad<SPT>::type zword;

  for (int base = beg; base < end; base += STG) {
    const int nst = min(STG, end - base);
    __syncthreads();  // previous round fully consumed
    {
      // pad the round with no-op entries (weight 0, row 0) so the inner loop has no tail
      const bool real = t < nst;
      const int k = real ? l[base + t] : 0;
      stage_off[t] = (unsigned)k * Lq;  // < 2^32 (checked on the host)
      stage_w[t] = real ? P.W[k] : 0.0;
    }
    __syncthreads();
    if (!active) continue;
    // software pipeline: the Zq words of group g+1 are in flight while group g is accumulated
    zword cur[UNR], nxt[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) cur[u] = *reinterpret_cast<const zword *>(Zt + stage_off[u]);
    for (int u0 = 0; u0 < nst; u0 += UNR) {
      if (u0 + UNR < nst) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) nxt[u] = *reinterpret_cast<const zword *>(Zt + stage_off[u0 + UNR + u]);
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const double wv = stage_w[u0 + u];  // broadcast
#pragma unroll
        for (int sidx = 0; sidx < SPT; ++sidx) acct[((cur[u] >> (8 * sidx)) & 0xff) * slot_stride + sidx * JT] += wv;
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) cur[u] = nxt[u];
    }
  }
  __syncthreads();

  // ---- epilogue: row r, columns (j, b) of this chunk, coalesced over (j, b) ----
  const double Meff = P.meff[0];
  const double pir = P.Pi[r];
  const double pcq = P.pc / P.q, pcqq = pcq / P.q, omp = 1.0 - P.pc;
  double *Crow = P.C + (long long)r * P.ld;
  for (int e = t; e < CH * P.s; e += JT) {
    const int jl = e / P.s, b = e - jl * P.s;       // site within the chunk, state
    const long long jj = (long long)start + jl;
    if (jj >= P.L || jj < i) continue;
    const long long c = jj * P.s + b;
    const double ptrue = acc[b * slot_stride + (jl % SPT) * JT + jl / SPT] / Meff;
    double pij;
    if (jj == i)
      pij = omp * ptrue + ((b == a - 1) ? pcq : 0.0);
    else
      pij = omp * ptrue + pcqq;
    Crow[c] = RAW ? ptrue : pij - pir * P.Pi[c];
  }
}

// lower site blocks <- transpose of the upper ones (diagonal blocks are already complete)
__global__ void symmetrize_kernel(double *__restrict__ C, long long n, long long ld, int s) {
  __shared__ double tile[32][33];
  const int tr = blockIdx.y, tc = blockIdx.x;
  if (tc > tr) return;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;  // 32 x 8
  // load tile (tc, tr): rows tc*32.., cols tr*32..
  for (int yy = ly; yy < 32; yy += 8) {
    const long long rr = (long long)tc * 32 + yy, cc = (long long)tr * 32 + lx;
    tile[yy][lx] = (rr < n && cc < n) ? C[rr * ld + cc] : 0.0;
  }
  __syncthreads();
  for (int yy = ly; yy < 32; yy += 8) {
    const long long rr = (long long)tr * 32 + yy, cc = (long long)tc * 32 + lx;
    if (rr < n && cc < n && (rr / s) > (cc / s)) C[rr * ld + cc] = tile[lx][yy];
  }
}

__global__ void extract_diag_blocks_kernel(const double *__restrict__ C, long long ld, int s, double *__restrict__ out) {
  const long long i = blockIdx.x;
  for (int e = threadIdx.x; e < s * s; e += blockDim.x) {
    const int a = e / s, b = e - a * s;
    out[i * s * s + e] = C[(i * s + a) * ld + i * s + b];
  }
}

// DCAUtils add_pseudocount (call site src/GaussDCA.jl:30) on contiguous n x n / n buffers:
//   Pi = (1-pc) Pi_true + pc/q;  off-diagonal site blocks: (1-pc) Pij_true + pc/q^2;  diagonal site blocks: (1-pc) Pij_true + delta_ab pc/q
__global__ void add_pseudocount_kernel(const double *__restrict__ Pi_true, const double *__restrict__ Pij_true, long long n, int s,
                                       int q, double pc, double *__restrict__ Pi, double *__restrict__ Pij) {
  const double pcq = pc / q, pcqq = pcq / q, omp = 1.0 - pc;
  const long long total = n * n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / n, c = e - r * n;
    const double base = omp * Pij_true[e];
    Pij[e] = (r / s == c / s) ? base + (r == c ? pcq : 0.0) : base + pcqq;
    if (e < n) Pi[e] = omp * Pi_true[e] + pcq;
  }
}

// compute_C(Pi, Pij) = Pij - Pi * Pi'  (src/GaussDCA.jl:32,76)
__global__ void compute_C_kernel(const double *__restrict__ Pi, const double *__restrict__ Pij, long long n, double *__restrict__ C) {
  const long long total = n * n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / n, c = e - r * n;
    C[e] = Pij[e] - Pi[r] * Pi[c];
  }
}

}  // namespace

int32_t gdca_k_add_pseudocount(gdca_ctx *ctx, const double *Pi_true, const double *Pij_true, long long n, int q, double pc, double *Pi,
                               double *Pij) {
  const long long total = n * n;
  const int grid = (int)((total + 255) / 256 < (long long)ctx->num_sms * 16 ? (total + 255) / 256 : (long long)ctx->num_sms * 16);
  add_pseudocount_kernel<<<grid, 256, 0, ctx->stream>>>(Pi_true, Pij_true, n, q - 1, q, pc, Pi, Pij);
  GDCA_LAUNCH_CHECK(ctx);
  return GDCA_OK;
}

int32_t gdca_k_compute_C(gdca_ctx *ctx, const double *Pi, const double *Pij, long long n, double *C) {
  const long long total = n * n;
  const int grid = (int)((total + 255) / 256 < (long long)ctx->num_sms * 16 ? (total + 255) / 256 : (long long)ctx->num_sms * 16);
  compute_C_kernel<<<grid, 256, 0, ctx->stream>>>(Pi, Pij, n, C);
  GDCA_LAUNCH_CHECK(ctx);
  return GDCA_OK;
}

int32_t gdca_k_symmetrize_C(gdca_ctx *ctx) {
  const long long n = ctx->n;
  if (!ctx->cov_full) {  // the tensor-core engine writes every tile together with its mirror image
    const unsigned nt = (unsigned)((n + 31) / 32);
    symmetrize_kernel<<<dim3(nt, nt), 256, 0, ctx->stream>>>(ctx->dC, n, ctx->npad, ctx->s);
    GDCA_LAUNCH_CHECK(ctx);
  }
  return gdca_k_extract_diag(ctx);
}

int32_t gdca_k_extract_diag(gdca_ctx *ctx) {
  GDCA_TRY(gdca_reserve(ctx, ctx->dCdiag, ctx->capCdiag, (size_t)ctx->L * ctx->s * ctx->s));
  extract_diag_blocks_kernel<<<(unsigned)ctx->L, 128, 0, ctx->stream>>>(ctx->dC, ctx->npad, ctx->s, ctx->dCdiag);
  GDCA_LAUNCH_CHECK(ctx);
  return GDCA_OK;
}

// site-major copy of the alignment + per-site state histograms as bucket offsets (once per loaded alignment; theta = :auto, the
// site order of the bit planes and the tensor-core covariance need nothing else of the per-site lists)
int32_t gdca_k_site_hist(gdca_ctx *ctx) {
  if (!ctx->have_alignment) return gdca_fail(ctx, GDCA_ERR_STATE, "site_hist: no alignment loaded");
  if (ctx->have_hist) return GDCA_OK;
  const long long L = ctx->L, M = ctx->M;
  GDCA_TRY(gdca_reserve(ctx, ctx->dListOff, ctx->capListOff, (size_t)L * (NSTATE + 1)));
  GDCA_TRY(gdca_reserve(ctx, ctx->dZt, ctx->capZt, (size_t)L * M));
  transpose_Z_kernel<<<dim3((unsigned)((M + 63) / 64), (unsigned)((L + 63) / 64)), 256, 0, ctx->stream>>>(ctx->dZ, L, M, ctx->dZt);
  GDCA_LAUNCH_CHECK(ctx);
  if (ctx->q < 24)
    site_hist_kernel<24><<<(unsigned)L, 256, 0, ctx->stream>>>(ctx->dZt, M, ctx->dListOff);
  else
    site_hist_kernel<32><<<(unsigned)L, 256, 0, ctx->stream>>>(ctx->dZt, M, ctx->dListOff);
  GDCA_LAUNCH_CHECK(ctx);
  ctx->have_hist = true;
  return GDCA_OK;
}

// per-site lists of sequence ids grouped by state (built on demand: the scatter-add covariance engine and its Pi)
int32_t gdca_k_build_lists(gdca_ctx *ctx) {
  if (!ctx->have_alignment) return gdca_fail(ctx, GDCA_ERR_STATE, "build_lists: no alignment loaded");
  if (ctx->have_lists) return GDCA_OK;
  GDCA_TRY(gdca_k_site_hist(ctx));
  const long long L = ctx->L, M = ctx->M;
  GDCA_TRY(gdca_reserve(ctx, ctx->dList, ctx->capList, (size_t)L * M));
  build_lists_kernel<<<(unsigned)L, LB, 0, ctx->stream>>>(ctx->dZt, M, ctx->dList, ctx->dListOff);
  GDCA_LAUNCH_CHECK(ctx);
  ctx->have_lists = true;
  return GDCA_OK;
}

// dHam[0] <- sum over pairs k<l of the number of identical positions (exact)
int32_t gdca_k_ident_sum(gdca_ctx *ctx, unsigned long long *ident_out) {
  GDCA_TRY(gdca_k_site_hist(ctx));
  GDCA_CUDA(ctx, cudaMemsetAsync(ctx->dHam, 0, 2 * sizeof(unsigned long long), ctx->stream));
  ident_sum_kernel<<<64, 256, 0, ctx->stream>>>(ctx->dListOff, ctx->L, ctx->dHam);
  GDCA_LAUNCH_CHECK(ctx);
  unsigned long long ident = 0;
  GDCA_CUDA(ctx, cudaMemcpyAsync(&ident, ctx->dHam, sizeof ident, cudaMemcpyDeviceToHost, ctx->stream));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ident_out) *ident_out = ident;
  return GDCA_OK;
}

int32_t gdca_k_covariance(gdca_ctx *ctx, double pc, bool raw) {
  if (!ctx->have_alignment) return gdca_fail(ctx, GDCA_ERR_STATE, "covariance: no alignment loaded");
  if (!ctx->have_weights) return gdca_fail(ctx, GDCA_ERR_STATE, "covariance: weights not computed");
  const long long L = ctx->L, M = ctx->M, n = ctx->n;
  GDCA_TRY(gdca_k_site_hist(ctx));
  GDCA_TRY(gdca_reserve(ctx, ctx->dPi, ctx->capPi, (size_t)n));
  const long long npad = ctx->npad;
  GDCA_TRY(gdca_reserve(ctx, ctx->dC, ctx->capC, (size_t)npad * npad));
  ctx->cov_full = false;
  ctx->last_cov_engine = 1;
  if (ctx->cov_engine != 1 && ctx->weights_from_counts) {
    // weights that are 1/(integer count): exact co-occurrence counts per weight class on the tensor cores (covtc.cu) when the
    // classes are few and large enough to pay (it computes Pi from the same class counts: no per-site lists at all);
    // otherwise, and for arbitrary weights, the scatter-add engine below
    bool done = false;
    GDCA_TRY(gdca_k_covariance_tc(ctx, pc, raw, &done));
    if (done) {
      ctx->pseudocount = pc;
      ctx->have_cov = true;
      ctx->have_inv = false;
      ctx->cov_full = true;
      ctx->last_cov_engine = 2;
      return GDCA_OK;
    }
  }
  GDCA_TRY(gdca_k_build_lists(ctx));
  pi_kernel<<<(unsigned)L, 256, 0, ctx->stream>>>(ctx->dList, ctx->dListOff, ctx->dW, ctx->dMeff, M, ctx->q, pc, ctx->dPi);
  GDCA_LAUNCH_CHECK(ctx);
  const long long Lq = (L + 127) / 128 * 128;
  if ((unsigned long long)M * (unsigned long long)Lq >= (1ull << 32))
    return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "covariance: M * roundup(L,128) must be < 2^32");
  GDCA_TRY(gdca_reserve(ctx, ctx->dZq, ctx->capZq, (size_t)M * Lq));
  build_zq_kernel<<<(unsigned)(((size_t)M * Lq + 255) / 256), 256, 0, ctx->stream>>>(ctx->dZ, L, M, Lq, ctx->s, ctx->dZq);
  GDCA_LAUNCH_CHECK(ctx);

  // zero everything: padding rows/cols, the not-yet-mirrored lower part, and other shards' rows.  In peer mode the
  // host zeroes rank 0's buffer (gdca_dev_zero_C) and barriers BEFORE any rank launches this stage.
  if (!(ctx->peers_ready && ctx->shard_world > 1))
    GDCA_CUDA(ctx, cudaMemsetAsync(ctx->dC, 0, (size_t)npad * npad * sizeof(double), ctx->stream));

  CovParams P;
  P.Zq = ctx->dZq;
  P.Lq = Lq;
  P.list = ctx->dList;
  P.listoff = ctx->dListOff;
  P.W = ctx->dW;
  P.meff = ctx->dMeff;
  P.Pi = ctx->dPi;
  // one process per GPU with peer buffers imported: every rank stores its rows straight into rank 0's C over
  // NVLink (rows are dealt by site, so the writes are disjoint: the reduce is fused into the kernel's epilogue)
  const bool peer_out = ctx->peers_ready && ctx->shard_world > 1;
  P.C = peer_out ? ctx->peer_C[0] : ctx->dC;
  P.L = L;
  P.M = M;
  P.n = n;
  P.ld = npad;
  P.q = ctx->q;
  P.s = ctx->s;
  constexpr int SPT = 2;
  P.nchunks = (int)((L + JT * SPT - 1) / (JT * SPT));
  P.rank = ctx->shard_rank;
  P.world = ctx->shard_world;
  P.pc = pc;
  const size_t smem = (size_t)ctx->q * SPT * JT * sizeof(double);
  // RAW (compile time): write Pij_true (no pseudocount, no - Pi Pi') for DCAUtils compute_weighted_frequencies
  const bool g2 = n <= 65535;
  const dim3 cgrid = g2 ? dim3((unsigned)P.nchunks, (unsigned)n) : dim3((unsigned)((long long)P.nchunks * n));
  auto launch = [&](auto kern) -> int32_t {
    GDCA_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_cov0, ctx->stream));
    kern<<<cgrid, JT, smem, ctx->stream>>>(P);
    return GDCA_OK;
  };
  if (raw)
    GDCA_TRY(g2 ? launch(cov_rows_kernel<SPT, true, true>) : launch(cov_rows_kernel<SPT, true, false>));
  else
    GDCA_TRY(g2 ? launch(cov_rows_kernel<SPT, false, true>) : launch(cov_rows_kernel<SPT, false, false>));
  GDCA_LAUNCH_CHECK(ctx);
  GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_cov1, ctx->stream));
  ctx->pseudocount = pc;
  ctx->have_cov = true;
  ctx->have_inv = false;
  return GDCA_OK;
}
