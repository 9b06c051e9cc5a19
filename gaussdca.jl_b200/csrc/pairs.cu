// pairs.cu -- K2/K3: the M x M pairwise-identity sweep (theta :auto and sequence weights).
//
// Replaces DCAUtils compute_theta and compute_weights (un-vendored; both reached from
// compute_weighted_frequencies, reference call site src/GaussDCA.jl:28):
//   theta  = min(0.5, 0.38*0.32 / meanfracid),  meanfracid = mean over k<l of ident(k,l)/L
//   count[k] = 1 + #{l != k : hamming(k,l) < floor(theta*L)},  W[k] = 1/count[k]
// ident = L - hamming, gap == gap counts as identical.  Everything here is exact integer work.
//
// Design (B200): INT32-ALU bound, not HBM bound -- the packed alignment (<= 64 MB at L=500, M=200k)
// lives in L2.  The bit-plane layout of pack.cu makes 32 sites of one pair cost 5 LOP3 + 1 POPC + 1 ADD.
//   * CTA tile 128 x 128 sequences, 512 threads (4 warps per scheduler), 4 x 8 register tile of pairs
//     per thread, <= 128 registers; the popcount accumulation runs on the otherwise idle FMA pipe.
//   * operands staged through shared memory with cp.async: a 2-slot ring of 16-word stages (a whole
//     tile for L <= 512 -> one barrier per tile); the ring runs across tile boundaries, so the next
//     tile is in flight while this one is computed.
//   * only tiles bi <= bj of the symmetric pair matrix are visited; a hit credits both sequences.
//   * production (mode 1, M >= 16384): the work list is NOT all tiles but the blocks the tensor-core prefilter
//     (tcfilter.cu) could not prove neighbour-free, each with a 16-bit mask of its 32 x 32 cells; a warp whose sub-tile
//     holds no flagged cell skips the block.  Counts are identical with and without the prefilter.
//   * persistent grid (one CTA per SM), items strided over (rank, world) for multi-GPU sharding; with peer
//     buffers imported (gdca_dev_peer_import) the epilogue adds the (rare) hits into EVERY rank's counters with
//     peer atomics over NVLink -- the all-reduce of the counts is fused into the sweep, no collective follows.
//   * mode 1 (neighbour counts, the production path) exits a warp's 16 x 64 sub-tile as soon as all of its
//     partial hamming distances have reached thresh -- exact, and ~40 % fewer words on typical alignments.
//     theta = :auto no longer needs a sweep at all (cov.cu:ident_sum_kernel); modes 0 and 2 (hamming sum,
//     three thresholds at once) remain for cross-checks and for hosts that want the sum from the sweep.
#include "gdca_internal.cuh"

namespace {

constexpr int TILE = GDCA_TILE;  // 128
constexpr int WC = 16;           // 32-site words per pipeline stage (L <= 512: one stage per tile, one barrier per tile)
constexpr int STAGES = 2;
constexpr int NTHREADS = 512;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// linear item t in [0, T(T+1)/2)  ->  (bi, bj), bi <= bj, rows enumerated bi = 0..T-1
__device__ __forceinline__ void item_to_tile(long long t, int T, int &bi, int &bj) {
  // row bi starts at off(bi) = bi*T - bi*(bi-1)/2
  double Td = (double)T + 0.5;
  int b = (int)(Td - sqrt(Td * Td - 2.0 * (double)t));
  if (b < 0) b = 0;
  if (b > T - 1) b = T - 1;
  while (true) {
    long long off = (long long)b * T - (long long)b * (b - 1) / 2;
    if (off > t) {
      --b;
      continue;
    }
    long long nxt = off + (T - b);
    if (t >= nxt) {
      ++b;
      continue;
    }
    bi = b;
    bj = b + (int)(t - off);
    return;
  }
}

struct PairParams {
  const uint32_t *planes;
  long long Mpad, M;
  int nwords, nchunks, T;
  long long n_items;   // T(T+1)/2
  const int2 *items;      // optional: explicit list of (bi, bj) blocks (tensor-core prefilter, tcfilter.cu) ...
  const int *n_items_dev; // ... and its length, known only on the device
  const uint32_t *item_mask;  // ... and per block the 32x32 cells that may hold a neighbour pair (bit 4*(r/32) + c/32)
  int rank, world;
  int thresh;
  int32_t *counts[GDCA_MAX_PEERS];  // [3][Mpad] of every rank that must see the hits (fused all-reduce over peer memory)
  int npeers;
  unsigned long long *ham_sum;  // [0] sum of hamming distances, [1] pairs visited (mode 1: pair-words executed)
};

// acc += pc on the FMA pipe: a 3-register IMAD (multiplier held in a register so ptxas cannot
// fold it into an ALU-pipe IADD3).  The ALU pipe is the bound of this kernel; FMA is idle.
__device__ __forceinline__ void add_on_fma_pipe(unsigned &acc, unsigned pc, unsigned one) {
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(pc), "r"(one));
}

template <int NPL, int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) pair_sweep_kernel(PairParams P) {
  extern __shared__ __align__(16) uint32_t smem[];
  // ring: [STAGES][2 operands][WC][NPL][TILE]
  constexpr int OP_WORDS = WC * NPL * TILE;
  constexpr int STAGE_WORDS = 2 * OP_WORDS;
  constexpr int CHUNKS = 2 * WC * NPL * (TILE / 4);          // 16-byte chunks per stage
  constexpr int CPT = (CHUNKS + NTHREADS - 1) / NTHREADS;    // chunks per thread
  __shared__ int s_row[3][TILE], s_col[3][TILE];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  // 512 threads = 32 (ty) x 16 (tx); thread tile 4 rows x 8 cols; a warp is 4 (ty) x 8 (tx)
  const int ty = (warp >> 1) * 4 + (lane >> 3);  // 0..31 -> rows ty*4 .. ty*4+3
  const int tx = (warp & 1) * 8 + (lane & 7);    // 0..15 -> cols tx*4.. and 64+tx*4..
  const unsigned one = (unsigned)(P.T > 0);      // == 1 at run time, unknown at compile time

  // items of this CTA: local index it -> global item (blockIdx.x + it*gridDim.x)*world + rank
  const long long my_first = (long long)blockIdx.x;
  const long long n_items = P.items ? (long long)*P.n_items_dev : P.n_items;
  auto tile_of = [&](long long t, int &bi, int &bj) {
    if (P.items) {
      const int2 v = P.items[t];
      bi = v.x;
      bj = v.y;
    } else {
      item_to_tile(t, P.T, bi, bj);
    }
  };
  const long long per_rank_items = (n_items - P.rank + P.world - 1) / P.world;  // items t with t%world==rank
  long long n_my = 0;
  if (my_first < per_rank_items) n_my = (per_rank_items - my_first + gridDim.x - 1) / gridDim.x;
  const long long n_flat = n_my * P.nchunks;

  // per-thread copy plan, fixed for the whole kernel: chunk ch -> (operand, row = w*NPL+p, 4-sequence column)
  int cp_dst[CPT], cp_wl[CPT], cp_op[CPT];
  long long cp_src[CPT];
#pragma unroll
  for (int u = 0; u < CPT; ++u) {
    const int ch = tid + u * NTHREADS;
    const int op = ch / (WC * NPL * (TILE / 4));
    const int rem = ch - op * (WC * NPL * (TILE / 4));
    const int row = rem / (TILE / 4), col = rem - row * (TILE / 4);
    const int wl = row / NPL, p = row - wl * NPL;
    cp_op[u] = (ch < CHUNKS) ? op : -1;
    cp_wl[u] = wl;
    cp_dst[u] = op * OP_WORDS + row * TILE + col * 4;
    cp_src[u] = (long long)p * P.Mpad + col * 4;  // + (w*NPL)*Mpad + tile*TILE at issue time
  }

  // flat pipeline position of the loader: tile coordinates are advanced incrementally
  long long ld_f = 0;
  int ld_c = 0, ld_bi = 0, ld_bj = 0;
  if (n_flat > 0) tile_of(my_first * P.world + P.rank, ld_bi, ld_bj);
  const long long item_step = (long long)gridDim.x * P.world;

  auto issue_load = [&]() {
    if (ld_f < n_flat) {
      uint32_t *dst = smem + (size_t)(ld_f % STAGES) * STAGE_WORDS;
#pragma unroll
      for (int u = 0; u < CPT; ++u) {
        if (cp_op[u] >= 0) {
          const int w = ld_c * WC + cp_wl[u];
          const bool valid = w < P.nwords;
          const long long seq0 = (long long)(cp_op[u] == 0 ? ld_bi : ld_bj) * TILE;
          const uint32_t *src = P.planes + (long long)(valid ? w : 0) * NPL * P.Mpad + cp_src[u] + seq0;
          cp_async16(dst + cp_dst[u], src, valid);
        }
      }
      ++ld_f;
      if (++ld_c == P.nchunks) {
        ld_c = 0;
        if (ld_f < n_flat) {
          const long long it = ld_f / P.nchunks;
          tile_of((my_first + it * gridDim.x) * P.world + P.rank, ld_bi, ld_bj);
        }
      }
    }
    cp_async_commit();
  };
  (void)item_step;

  unsigned acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0;
  unsigned long long ham_local = 0, pairs_local = 0;

  for (int i = tid; i < 3 * TILE; i += NTHREADS) {
    (&s_row[0][0])[i] = 0;
    (&s_col[0][0])[i] = 0;
  }

  issue_load();

  int cur_c = 0;
  long long cur_it = 0;
  bool warp_done = false;  // MODE 1: this warp's sub-tile can no longer contain a neighbour pair
  bool warp_skip = false;  // MODE 1 with cell masks: the prefilter cleared every cell this warp owns
  // this warp owns rows 16 (warp >> 1) .. +15 and columns 32 (warp & 1) .. +31 and 64 + 32 (warp & 1) .. +31
  const uint32_t my_cells = (1u << (4 * (warp >> 2) + (warp & 1))) | (1u << (4 * (warp >> 2) + 2 + (warp & 1)));
  unsigned long long words_done = 0;  // MODE 1: 32-site words this warp really processed (x 1024 pairs each)
  for (long long f = 0; f < n_flat; ++f) {
    cp_async_wait<0>();
    __syncthreads();  // stage f landed for everyone; everyone is done with stage f-1
    issue_load();     // prefetch stage f+1 into the slot freed by f-1 while f is being computed

    if (MODE == 1 && cur_c == 0 && P.item_mask) {
      warp_skip = (P.item_mask[(my_first + cur_it * gridDim.x) * P.world + P.rank] & my_cells) == 0;
      warp_done = warp_skip;
    }
    const uint32_t *A = smem + (size_t)(f % STAGES) * STAGE_WORDS;
    const uint32_t *B = A + OP_WORDS;
    const int wcount = min(WC, P.nwords - cur_c * WC);  // words really present in this stage
    auto process_word = [&](int w) {
      unsigned x[4][8];
#pragma unroll
      for (int p = 0; p < NPL; ++p) {
        const uint4 a0 = *reinterpret_cast<const uint4 *>(A + (w * NPL + p) * TILE + ty * 4);
        const uint4 b0 = *reinterpret_cast<const uint4 *>(B + (w * NPL + p) * TILE + tx * 4);
        const uint4 b1 = *reinterpret_cast<const uint4 *>(B + (w * NPL + p) * TILE + 64 + tx * 4);
        const unsigned a[4] = {a0.x, a0.y, a0.z, a0.w};
        const unsigned b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) x[i][j] = (p == 0) ? (a[i] ^ b[j]) : (x[i][j] | (a[i] ^ b[j]));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) add_on_fma_pipe(acc[i][j], (unsigned)__popc(x[i][j]), one);
    };
    if (MODE == 1) {
      // Early exit (exact): hamming only grows with more sites, so once every pair of this WARP's 16 x 64
      // sub-tile has reached thresh none of them can be a neighbour and the remaining words are skipped.
      // Each warp decides alone (no CTA barrier); the vote costs ~20 instructions per word (224 instructions).
      // no pair can reach thresh before ceil(thresh/32) words: run those without votes, fully pipelined,
      // vote once right after them, then after every word
      auto vote = [&]() {
        unsigned dmin = acc[0][0];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) dmin = min(dmin, acc[i][j]);
        warp_done = __all_sync(0xffffffffu, (int)dmin >= P.thresh) != 0;
      };
      int n_free = (P.thresh + 31) / 32 - cur_c * WC;
      n_free = max(0, min(n_free, wcount));
      if (!warp_done && n_free > 0) {
#pragma unroll 4
        for (int w = 0; w < n_free; ++w) process_word(w);
        words_done += n_free;
        vote();
      }
      for (int w0 = n_free; w0 < wcount && !warp_done; ++w0) {
        process_word(w0);
        words_done += 1;
        vote();
      }
    } else {
#pragma unroll 4
      for (int w = 0; w < wcount; ++w) process_word(w);
    }

    if (++cur_c == P.nchunks) {
      cur_c = 0;
      // ---- tile epilogue: acc[i][j] = hamming(row r_i, col c_j) ----
      const long long t = (my_first + cur_it * gridDim.x) * P.world + P.rank;
      ++cur_it;
      int bi, bj;
      tile_of(t, bi, bj);
      const bool diag = (bi == bj);
      if (MODE == 1 && warp_skip) {  // nothing was accumulated: none of these pairs is a neighbour (proved by the prefilter)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = 0x7fffffffu;
      }
      const int hi_thresh = (MODE == 2) ? P.thresh + 1 : P.thresh;
      unsigned tile_ham = 0, tile_pairs = 0, any = 0;
      const bool interior = !diag && ((long long)(bj + 1) * TILE <= P.M);  // bi < bj: rows are in range too
      if (interior) {
        // fast path (all but O(T) of the T^2/2 tiles): every pair is valid
        unsigned dmin = 0x7fffffffu;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (MODE != 1) add_on_fma_pipe(tile_ham, acc[i][j], one);
            if (MODE != 0) dmin = min(dmin, acc[i][j]);
          }
        tile_pairs = 32;
        any = ((int)dmin < hi_thresh) ? 1u : 0u;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rl = ty * 4 + i;
          const bool rvalid = (long long)bi * TILE + rl < P.M;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int cl = (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4);
            const bool valid = rvalid && ((long long)bj * TILE + cl < P.M) && (!diag || rl < cl);
            const unsigned d = valid ? acc[i][j] : 0x7fffffffu;  // invalid pairs can never be neighbours
            acc[i][j] = d;
            if (MODE != 1) {
              tile_ham += valid ? d : 0u;
              tile_pairs += valid ? 1u : 0u;
            }
            if (MODE != 0) any |= ((int)d < hi_thresh) ? 1u : 0u;
          }
        }
      }
      ham_local += tile_ham;
      pairs_local += tile_pairs;
      if (MODE != 0) {
        // most tiles contain no neighbour pair at all: one vote, then skip the counting entirely
        if (__syncthreads_or((int)any)) {
          constexpr int NK = (MODE == 2) ? 3 : 1;
#pragma unroll
          for (int k = 0; k < NK; ++k) {
            const int th = (MODE == 2) ? P.thresh - 1 + k : P.thresh;
            int rh[4], chh[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) rh[i] = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) chh[j] = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int h = ((int)acc[i][j] < th) ? 1 : 0;
                rh[i] += h;
                chh[j] += h;
              }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              int v = rh[i];  // same rows across the 8 lanes sharing lane>>3
              v += __shfl_xor_sync(0xffffffffu, v, 1);
              v += __shfl_xor_sync(0xffffffffu, v, 2);
              v += __shfl_xor_sync(0xffffffffu, v, 4);
              if ((lane & 7) == 0 && v) atomicAdd(&s_row[k][ty * 4 + i], v);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              int v = chh[j];  // same cols across the 4 lanes sharing lane&7
              v += __shfl_xor_sync(0xffffffffu, v, 8);
              v += __shfl_xor_sync(0xffffffffu, v, 16);
              const int cl = (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4);
              if ((lane >> 3) == 0 && v) atomicAdd(&s_col[k][cl], v);
            }
          }
          __syncthreads();
          for (int e = tid; e < NK * 2 * TILE; e += NTHREADS) {
            const int k = e / (2 * TILE);
            const int r = e - k * 2 * TILE;
            if (r < TILE) {
              const int v = s_row[k][r];
              if (v) {
                for (int pr = 0; pr < P.npeers; ++pr)  // own buffer and, over NVLink, every peer's
                  atomicAdd(P.counts[pr] + (long long)k * P.Mpad + (long long)bi * TILE + r, v);
                s_row[k][r] = 0;
              }
            } else {
              const int v = s_col[k][r - TILE];
              if (v) {
                for (int pr = 0; pr < P.npeers; ++pr)
                  atomicAdd(P.counts[pr] + (long long)k * P.Mpad + (long long)bj * TILE + (r - TILE), v);
                s_col[k][r - TILE] = 0;
              }
            }
          }
          // the __syncthreads at the top of the next iteration orders these resets before reuse
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0;
      warp_done = false;
    }
  }
  cp_async_wait<0>();

  if (MODE == 1) {
    if (lane == 0 && words_done) atomicAdd(P.ham_sum + 1, words_done * 1024ull);  // executed pair-words (roofline evidence)
  }
  if (MODE == 0 || MODE == 2) {
    for (int o = 16; o; o >>= 1) {
      ham_local += __shfl_xor_sync(0xffffffffu, ham_local, o);
      pairs_local += __shfl_xor_sync(0xffffffffu, pairs_local, o);
    }
    if (lane == 0 && pairs_local) {
      atomicAdd(P.ham_sum, ham_local);
      atomicAdd(P.ham_sum + 1, pairs_local);
    }
  }
}

// ---- exact sweep of the flagged 32 x 32 CELLS (production path behind the tensor-core prefilter) -------------------------
// The prefilter leaves a list of blocks with a 16-bit mask of the cells it could not clear.  On a favourable sequence order few
// blocks are listed; on a random order almost every block holds a few flagged cells (~22 % of all cells at config C), and a
// block-per-CTA sweep then costs as much as no filter at all.  Here the unit of work is the flagged cell: ONE WARP per cell,
// no shared memory, no CTA barrier -- lane (ry, cx) owns the 4 x 8 pairs of rows 4 ry.. and columns 8 cx.. and reads its
// operands (16-byte words of the bit planes) straight through L1 from the L2-resident planes; the per-warp early exit is the
// one of the block kernel.  Every warp walks a contiguous range of the cell list (cells of one block are adjacent), found by
// one binary search on the running cell counts of the blocks.
struct CellParams {
  const uint32_t *planes;
  long long Mpad, M;
  int nwords, thresh;
  const int2 *items;
  const uint32_t *item_mask, *cellbase;
  const unsigned long long *n_packed;  // low word: listed blocks, high word: flagged cells
  int32_t *counts[GDCA_MAX_PEERS];
  int npeers;
  unsigned long long *ham_sum;
  const unsigned long long *npairs;    // candidate pairs the prefilter listed; when they fit their list (<= pair_cap) the pair
  unsigned long long pair_cap;         // kernel below has done the exact stage and this kernel returns at once
};

// ---- exact stage on the candidate PAIRS of the prefilter: one warp per pair, on the alignment itself ----
// The prefilter knows the projected distance of every pair exactly, so it can list the pairs that may be neighbours instead of
// flagging 32 x 32 cells of them.  A warp reads the two sequences (L bytes each, coalesced), counts the differing positions
// (byte-wise compare, popc) and credits both sequences when the distance is below thresh.  The work is proportional to the
// number of candidates -- true neighbour pairs plus the few the 4-class projection cannot tell apart -- and no longer to how
// they are spread over the pair matrix: config C 2.0 -> 0.2 ms, the same sequences in random order 27 -> 0.2 ms.
struct PairListParams {
  const int8_t *Z;        // [M][L]
  long long L, M;
  int thresh, nwords;
  const int2 *pairs;
  const unsigned long long *npairs;
  unsigned long long pair_cap;
  int32_t *counts[GDCA_MAX_PEERS];
  int npeers;
  unsigned long long *ham_sum;
};

__global__ void __launch_bounds__(256) pair_list_kernel(PairListParams P) {
  const unsigned long long n = *P.npairs;
  if (n > P.pair_cap) return;  // the list overflowed: the cell sweep runs instead
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const bool words = (P.L & 3) == 0 && (reinterpret_cast<uintptr_t>(P.Z) & 3) == 0;
  // A pair costs two dependent trips to L2 (its list entry, then the two sequences): the entry of the warp's NEXT pair is fetched
  // before the current pair is compared, so only the sequence reads remain on the critical path of the loop.
  unsigned long long p = (unsigned long long)gw;
  int2 kl = p < n ? P.pairs[p] : make_int2(-1, -1);
  for (; p < n; p += (unsigned long long)nw) {
    const unsigned long long pn = p + (unsigned long long)nw;
    const int2 kl_next = pn < n ? P.pairs[pn] : make_int2(-1, -1);
    if (kl.x >= 0) {  // else: unused slot of a warp's chunk
      const int8_t *za = P.Z + (long long)kl.x * P.L, *zb = P.Z + (long long)kl.y * P.L;
      int ham = 0;
      if (words) {
        const uint32_t *wa = reinterpret_cast<const uint32_t *>(za), *wb = reinterpret_cast<const uint32_t *>(zb);
#pragma unroll 4
        for (int w = lane; w < (int)(P.L >> 2); w += 32) ham += __popc(__vcmpne4(__ldg(wa + w), __ldg(wb + w))) >> 3;
      } else {
#pragma unroll 4
        for (int i = lane; i < (int)P.L; i += 32) ham += za[i] != zb[i];
      }
      ham = __reduce_add_sync(0xffffffffu, ham);
      if (lane == 0 && ham < P.thresh)
        for (int pr = 0; pr < P.npeers; ++pr) {  // own buffer and, over NVLink, the peers'
          atomicAdd(P.counts[pr] + kl.x, 1);
          atomicAdd(P.counts[pr] + kl.y, 1);
        }
    }
    kl = kl_next;
  }
  if (gw == 0 && lane == 0) atomicAdd(P.ham_sum + 1, n * (unsigned long long)P.nwords);  // executed pair-words (roofline evidence)
}

template <int NPL>
__global__ void __launch_bounds__(128, 3) cell_sweep_kernel(CellParams P) {
  if (P.pair_cap && *P.npairs <= P.pair_cap) return;  // the candidate pairs fitted their list: pair_list_kernel did the exact stage
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const unsigned long long np = *P.n_packed;
  const int n_items = (int)(np & 0xffffffffull);
  const long long n_cells = (long long)(np >> 32);
  const int ry = lane >> 2, cx = lane & 3;
  const unsigned one = (unsigned)(P.nwords > 0);  // == 1 at run time, unknown at compile time
  const long long c_beg = n_cells * gw / nw, c_end = n_cells * (gw + 1) / nw;
  if (c_beg >= c_end) return;
  // block that holds cell c_beg: the last slot with cellbase <= c_beg
  int slot = 0;
  {
    int lo = 0, hi = n_items - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if ((long long)P.cellbase[mid] <= c_beg) lo = mid; else hi = mid - 1;
    }
    slot = lo;
  }
  int2 it = P.items[slot];
  uint32_t mask = P.item_mask[slot];
  // drop the cells of this block that belong to the warp in front
  for (long long skip = c_beg - (long long)P.cellbase[slot]; skip > 0; --skip) mask &= mask - 1;
  unsigned long long words_done = 0;
  const int n_free = min((P.thresh + 31) / 32, P.nwords);  // no pair can reach thresh before this many words
  for (long long c = c_beg; c < c_end; ++c) {
    while (mask == 0) {  // next listed block
      ++slot;
      it = P.items[slot];
      mask = P.item_mask[slot];
    }
    const int bit = __ffs(mask) - 1;
    mask &= mask - 1;
    const int cr = bit >> 2, cc = bit & 3;
    const long long row0 = (long long)it.x * TILE + 32 * cr, col0 = (long long)it.y * TILE + 32 * cc;
    const uint32_t *pa = P.planes + row0 + 4 * ry, *pb = P.planes + col0 + 8 * cx;
    unsigned acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0;
    auto process_word = [&](int w) {
      unsigned x[4][8];
#pragma unroll
      for (int p = 0; p < NPL; ++p) {
        const long long o = (long long)(w * NPL + p) * P.Mpad;
        const uint4 a0 = __ldg(reinterpret_cast<const uint4 *>(pa + o));
        const uint4 b0 = __ldg(reinterpret_cast<const uint4 *>(pb + o));
        const uint4 b1 = __ldg(reinterpret_cast<const uint4 *>(pb + o + 4));
        const unsigned a[4] = {a0.x, a0.y, a0.z, a0.w};
        const unsigned b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) x[i][j] = (p == 0) ? (a[i] ^ b[j]) : (x[i][j] | (a[i] ^ b[j]));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) add_on_fma_pipe(acc[i][j], (unsigned)__popc(x[i][j]), one);
    };
    // exact early exit: hamming only grows with more sites; once all 1024 pairs of the cell have reached thresh none of them
    // can be a neighbour.  The first ceil(thresh/32) words run without votes.  (Measured: 12 warps per SM with the loads issued
    // where they are used beat 8 warps per SM with the operands of the next word prefetched into registers, 27.1 vs 28.7 ms on
    // the shuffled config C.)
    bool done = false;
    auto vote = [&]() {
      unsigned dmin = acc[0][0];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmin = min(dmin, acc[i][j]);
      done = __all_sync(0xffffffffu, (int)dmin >= P.thresh) != 0;
    };
#pragma unroll 2
    for (int w = 0; w < n_free; ++w) process_word(w);
    vote();
    int w = n_free;
    for (; w < P.nwords && !done; ++w) {
      process_word(w);
      vote();
    }
    words_done += (unsigned long long)w;
    if (done) continue;  // every pair is at or beyond thresh
    // ---- hits: a pair credits both sequences; diagonal cells of diagonal blocks count each unordered pair once ----
    const bool diag_cell = (it.x == it.y) && (cr == cc);
    int rh[4] = {0, 0, 0, 0}, ch[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned any = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long r = row0 + 4 * ry + i;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const long long cl = col0 + 8 * cx + j;
        const bool valid = r < P.M && cl < P.M && (!diag_cell || r < cl);
        const int h = (valid && (int)acc[i][j] < P.thresh) ? 1 : 0;
        rh[i] += h;
        ch[j] += h;
        any |= (unsigned)h;
      }
    }
    if (!__any_sync(0xffffffffu, any)) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // rows: the 4 lanes that share ry
      int v = rh[i];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (cx == 0 && v)
        for (int pr = 0; pr < P.npeers; ++pr) atomicAdd(P.counts[pr] + row0 + 4 * ry + i, v);  // own buffer and, over NVLink, the peers'
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // columns: the 8 lanes that share cx
      int v = ch[j];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (ry == 0 && v)
        for (int pr = 0; pr < P.npeers; ++pr) atomicAdd(P.counts[pr] + col0 + 8 * cx + j, v);
    }
  }
  if (lane == 0 && words_done) atomicAdd(P.ham_sum + 1, words_done * 1024ull);  // executed pair-words (roofline evidence)
}

template <int NPL>
int32_t launch_cells(gdca_ctx *ctx, const CellParams &P) {
  cell_sweep_kernel<NPL><<<ctx->num_sms * 3, 128, 0, ctx->stream>>>(P);
  GDCA_LAUNCH_CHECK(ctx);
  return GDCA_OK;
}

template <int NPL>
int32_t launch_pairs(gdca_ctx *ctx, int mode, const PairParams &P) {
  const size_t smem = (size_t)STAGES * 2 * WC * NPL * TILE * sizeof(uint32_t);
  const int grid = ctx->num_sms;
#define GDCA_PAIR_LAUNCH(MODE)                                                                                      \
  do {                                                                                                              \
    GDCA_CUDA(ctx, cudaFuncSetAttribute(pair_sweep_kernel<NPL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                        (int)smem));                                                                \
    pair_sweep_kernel<NPL, MODE><<<grid, NTHREADS, smem, ctx->stream>>>(P);                                         \
  } while (0)
  if (mode == 0)
    GDCA_PAIR_LAUNCH(0);
  else if (mode == 1)
    GDCA_PAIR_LAUNCH(1);
  else
    GDCA_PAIR_LAUNCH(2);
#undef GDCA_PAIR_LAUNCH
  GDCA_LAUNCH_CHECK(ctx);
  return GDCA_OK;
}

}  // namespace

int32_t gdca_k_pair_pass(gdca_ctx *ctx, int mode, int thresh, int sample_stride) {
  if (!ctx->have_alignment) return gdca_fail(ctx, GDCA_ERR_STATE, "pair_pass: no alignment loaded");
  if (mode < 0 || mode > 2) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "pair_pass: mode must be 0, 1 or 2");
  GDCA_TRY(gdca_reserve(ctx, ctx->dCounts, ctx->capCounts, (size_t)3 * ctx->Mpad));
  // in peer mode the host zeroes every rank's counters (gdca_dev_zero_counts) and barriers before the sweep
  if (mode != 0 && !(ctx->peers_ready && ctx->shard_world > 1))
    GDCA_CUDA(ctx, cudaMemsetAsync(ctx->dCounts, 0, (size_t)3 * ctx->Mpad * sizeof(int32_t), ctx->stream));
  GDCA_CUDA(ctx, cudaMemsetAsync(ctx->dHam, 0, 2 * sizeof(unsigned long long), ctx->stream));
  if (sample_stride < 1) sample_stride = 1;

  // production sweep: clear whole blocks on the tensor cores first, sweep only what is left (exact, tcfilter.cu)
  const bool filtered = mode == 1 && sample_stride == 1 && gdca_tc_filter_wanted(ctx);
  if (mode == 1) GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_sweep0, ctx->stream));
  if (filtered) {
    GDCA_TRY(gdca_k_tc_filter(ctx, thresh, nullptr, 0));
    GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_filter, ctx->stream));
  }

  PairParams P;
  P.planes = ctx->dPlanes;
  P.Mpad = ctx->Mpad;
  P.M = ctx->M;
  P.nwords = (int)ctx->nwords;
  P.nchunks = (int)((ctx->nwords + WC - 1) / WC);
  P.T = (int)(ctx->Mpad / TILE);
  P.n_items = (long long)P.T * (P.T + 1) / 2;
  // a sampled sweep visits every sample_stride-th item of this shard: items t = rank (mod world*stride)
  P.rank = ctx->shard_rank;
  P.world = ctx->shard_world * sample_stride;
  P.items = nullptr;
  P.n_items_dev = nullptr;
  P.item_mask = nullptr;
  if (filtered) {  // the list is already this rank's share
    P.items = ctx->dItems;
    P.n_items_dev = reinterpret_cast<const int *>(ctx->dNItems);  // low word of the packed counter
    P.item_mask = ctx->dItemMask;
    P.rank = 0;
    P.world = 1;
  }
  ctx->last_sweep_filtered = filtered;
  P.thresh = thresh;
  if (ctx->peers_ready && ctx->shard_world > 1) {
    P.npeers = ctx->shard_world;
    for (int r = 0; r < P.npeers; ++r) P.counts[r] = ctx->peer_counts[r];
  } else {
    P.npeers = 1;
    P.counts[0] = ctx->dCounts;
  }
  P.ham_sum = ctx->dHam;
  int32_t st;
  if (filtered && ctx->cell_sweep) {
    // behind the prefilter the unit of work is the flagged 32 x 32 cell (one warp each), not the 128 x 128 block
    CellParams Q;
    Q.planes = ctx->dPlanes;
    Q.Mpad = ctx->Mpad;
    Q.M = ctx->M;
    Q.nwords = (int)ctx->nwords;
    Q.thresh = thresh;
    Q.items = ctx->dItems;
    Q.item_mask = ctx->dItemMask;
    Q.cellbase = ctx->dCellBase;
    Q.n_packed = ctx->dNItems;
    Q.npeers = P.npeers;
    for (int r = 0; r < GDCA_MAX_PEERS; ++r) Q.counts[r] = r < P.npeers ? P.counts[r] : nullptr;
    Q.ham_sum = ctx->dHam;
    Q.npairs = ctx->dNPairs;
    Q.pair_cap = ctx->pair_cap;
    if (ctx->pair_cap) {
      PairListParams R;
      R.Z = ctx->dZ;
      R.L = ctx->L;
      R.M = ctx->M;
      R.thresh = thresh;
      R.nwords = (int)ctx->nwords;
      R.pairs = ctx->dPairs;
      R.npairs = ctx->dNPairs;
      R.pair_cap = ctx->pair_cap;
      R.npeers = P.npeers;
      for (int r = 0; r < GDCA_MAX_PEERS; ++r) R.counts[r] = r < P.npeers ? P.counts[r] : nullptr;
      R.ham_sum = ctx->dHam;
      pair_list_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(R);
      GDCA_LAUNCH_CHECK(ctx);
    }
    switch (ctx->nplanes) {
      case 1: st = launch_cells<1>(ctx, Q); break;
      case 2: st = launch_cells<2>(ctx, Q); break;
      case 3: st = launch_cells<3>(ctx, Q); break;
      case 4: st = launch_cells<4>(ctx, Q); break;
      default: st = launch_cells<5>(ctx, Q); break;
    }
    if (st == GDCA_OK) GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_sweep1, ctx->stream));
    return st;
  }
  switch (ctx->nplanes) {
    case 1: st = launch_pairs<1>(ctx, mode, P); break;
    case 2: st = launch_pairs<2>(ctx, mode, P); break;
    case 3: st = launch_pairs<3>(ctx, mode, P); break;
    case 4: st = launch_pairs<4>(ctx, mode, P); break;
    default: st = launch_pairs<5>(ctx, mode, P); break;
  }
  if (st == GDCA_OK && mode == 1) GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_sweep1, ctx->stream));
  return st;
}
