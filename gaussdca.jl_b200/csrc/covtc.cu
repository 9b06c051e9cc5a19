// covtc.cu -- K4 on the tensor cores: the weighted two-point frequencies as exact co-occurrence COUNTS per weight class
// (tcgen05 kind::mxf4 / TMEM / TMA multicast, sm_100a), combined in FP64.
//
// Replaces, like cov.cu (the scatter-add engine that remains for arbitrary weights):
//   DCAUtils compute_freqs (Pij_true; un-vendored, reference call site src/GaussDCA.jl:28), add_pseudocount (:30), compute_C (:32,:76).
//
//   Pij_true = X' W X / Meff,  X the M x n one-hot matrix.  The reference's weights are W[k] = 1/count[k] with
//   INTEGER neighbour counts, so the sequences fall into a few classes of equal weight (46 at config C) and
//
//       X' W X = sum_c (1/c) * N_c,      N_c = X_c' X_c  = integer co-occurrence counts among the sequences with count c.
//
//   N_c is a 0/1 GEMM: exact on the FP4 tensor cores (e2m1 holds 0 and 1, FP32 accumulation of at most 2^23 ones is exact),
//   2 n^2 M = 4e13 flop at config C -- 3 ms at the FP4 rate, where the scatter-add engine's 2.5e10 shared-memory
//   read-modify-writes take 18.4 ms.  The FP64 result is sum_c w_c N_c in ascending class order with one FMA per class:
//   deterministic, bit-identical for (r,c) and (c,r), for any tile order and any number of GPUs.
//
// Pieces:
//   * classes: histogram of the counts, compacted on the device, planned on the host (a few hundred bytes cross PCIe): class c
//     gets a run of whole 256-sequence k-blocks, padded with empty columns; `perm` lists the sequences in that order;
//   * operand: Xt[(i,a)][kpos] = [Z[i, perm[kpos]] == a] as packed e2m1 (1.0 = 0x2), K-major, n x Mk/2 bytes (1.03 GB at C).
//     While a warp encodes a k-block it also counts, per state, the sequences of that class segment carrying the state at the
//     site: Pi = sum_c w_c n_c(i,a) / Meff comes from the same exact integers (no per-site lists on this path at all);
//   * GEMM: Xt Xt' over the lower triangle of 128 x 128 output tiles.  CTA PAIRS (clusters of 2) issue ONE
//     tcgen05.mma.cta_group::2 of 256 x 128: the even CTA issues for both, every CTA keeps its own A tile (its 128 output rows)
//     and only HALF of the B tile -- the tensor cores read the other half from the peer's shared memory -- so 24 KB instead of
//     32 KB of operands arrive in every SM per k-block (the L2 -> SM path delivers ~64-70 B/clk/SM, which capped the first,
//     2 x 2-cluster multicast version of this kernel at 0.42 of the FP4 rate).  The barriers of the pair live in the even CTA:
//     both producers' TMA bytes are counted there (cp.async.bulk.tensor.cta_group::2), one tcgen05.commit.cta_group::2 frees
//     a stage in both CTAs, the epilogue warps of both CTAs release the accumulator stage there.
//     Per CTA: warp 0 = TMA producer (9-stage ring of 24 KB), warp 1 = MMA issuer (even CTA), warps 2..9 = epilogue.  Two
//     accumulator stages of 128 TMEM columns: while the MMAs of class c+1 run, the epilogue drains class c
//     (tcgen05.ld -> FP64 -> acc += w_c * N_c, 64 FP64 accumulators per thread in registers).  After the last class the fused
//     epilogue of cov.cu (1/Meff, pseudocount, - Pi Pi') writes the tile and its mirror image.
//   * device groups: pair tiles dealt round-robin, stored straight into the leader's C (disjoint, no reduction).
#include <algorithm>
#include <string.h>
#include <vector>

#include "gdca_internal.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace tcptx;

constexpr int CT = 128;            // output tile edge
constexpr int CBK = 128;           // K bytes per stage = 256 packed e2m1 = 256 sequences
constexpr int CSEQ = 2 * CBK;      // sequences per k-block
constexpr int HALF_BYTES = 64 * CBK;     // one TMA box: 64 rows x 128 B
constexpr int TILE_BYTES = CT * CBK;     // 16 KB
constexpr int CSTAGE_BYTES = TILE_BYTES + HALF_BYTES;  // the CTA's own A tile + its half of the B tile: 24 KB
constexpr int CNSTAGE = 9;
constexpr int C_THREADS = 320;     // warp 0 producer, warp 1 MMA issuer, warps 2..9 epilogue
constexpr int C_TMEM_COLS = 512;   // 2 x 128 accumulator columns + 64 scale-factor columns -> next power of two
constexpr int C_SF_COL = 2 * CT;
constexpr int MAXSEG = GDCA_COV_MAXCLS;
constexpr size_t C_SMEM = (size_t)CNSTAGE * CSTAGE_BYTES + 1024;
// kind::mxf4.block_scale.block32: A = B = E2M1 (1 at bits 7 and 10), scale format UE8M0 (bit 23), N >> 3 at bit 17, M >> 4 at bit 24;
// cta_group::2: M = 256 rows across the CTA pair
constexpr uint32_t C_IDESC = (1u << 7) | (1u << 10) | ((uint32_t)(CT >> 3) << 17) | (1u << 23) | ((uint32_t)(2 * CT >> 4) << 24);
constexpr long long SEG_MAX_SEQ = (1ll << 23) - 256;  // FP32 accumulation of ones stays exact, and count + 2^23 keeps the count in its mantissa

struct CovTcParams {
  const int2 *tiles;     // [ntiles] pair tiles (RB2, CB): row blocks 2 RB2 and 2 RB2 + 1 x column block CB <= 2 RB2 + 1, this rank's share
  int ntiles;
  int nblk;              // 128-row blocks of the output
  int nseg;
  const int *seg_end;    // [nseg] k-block index one past the segment
  const double *seg_w;   // [nseg] weight of the class
  const double *meff, *Pi;
  double *C;
  long long n, ld;
  int s, q, raw;
  double pc;
  unsigned int *round_sync;  // producers that have started their j-th tile (nullptr: free running)
};

// FP32 accumulator -> FP64 without a conversion instruction (F2F.F64.F32 issues at 16 lanes per clock and SM: 64 of them per
// thread and class made the drain of a class ~1000 clocks, the floor of the plan's cost model).  The accumulator holds an exact
// integer count < 2^23: count + 2^23 (one FADD, exact) carries it in its mantissa field, and 2^52 + count - 2^52 (one DADD on the
// full-rate FP64 pipe, exact) is the same double the conversion gives.
__device__ __forceinline__ double count_to_double(uint32_t fbits) {
  const uint32_t u = __float_as_uint(__uint_as_float(fbits) + 8388608.0f) & 0x007fffffu;
  return __hiloint2double(0x43300000, (int)u) - 4503599627370496.0;
}

__global__ void __launch_bounds__(C_THREADS, 1) cov_tc_kernel(const __grid_constant__ CUtensorMap tmap, CovTcParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bars[2 * CNSTAGE + 4];
  __shared__ uint32_t s_tmem;
  __shared__ int s_seg_end[MAXSEG];
  __shared__ double s_seg_w[MAXSEG];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_u32(s_bars);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (CNSTAGE + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * CNSTAGE + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * CNSTAGE + 2 + a); };
  const uint32_t tmem_slot = smem_u32(&s_tmem);
  volatile uint32_t *tmem_slot_ptr = &s_tmem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();   // y: this CTA computes the tile (2 RB2 + y, CB); the even CTA issues the MMAs of the pair
  const int cy = (int)crank;
  const int first = (int)cluster_id_x(), step = (int)cluster_nid_x();

  for (int i = threadIdx.x; i < P.nseg; i += C_THREADS) {
    s_seg_end[i] = P.seg_end[i];
    s_seg_w[i] = P.seg_w[i];
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    for (int s = 0; s < CNSTAGE; ++s) {
      mbar_init(full_bar(s), 1);   // even CTA: its producer's expect_tx; the bytes of BOTH CTAs are counted here
      mbar_init(empty_bar(s), 1);  // one cta_group::2 commit reaches both CTAs
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 16);  // even CTA: one arrival per epilogue warp of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();  // both CTAs are resident, their barriers initialised, before the pair allocates tensor memory
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(C_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp >= 2 && warp < 6) {
    // UE8M0 scale factors, all 2^0: every byte of the 64 scale columns of all 128 lanes is 0x7F
    const uint32_t t = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)C_SF_COL;
    tmem_st32_fill(t, 0x7F7F7F7Fu);
    tmem_st32_fill(t + 32u, 0x7F7F7F7Fu);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int nseg = P.nseg;
  const int KB = nseg > 0 ? s_seg_end[nseg - 1] : 0;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): my A tile (two 64-row boxes) and MY half of the B tile; the even CTA's barrier counts it all
    int s = 0;
    uint32_t ph = 0;
    unsigned round = 0;
    for (int t = first; t < P.ntiles; t += step) {
      if (P.round_sync) {
        // Long operand rows (K in the millions): the pairs running at one time share a few row and column blocks, but only pairs
        // that walk K TOGETHER find each other's rows in L2.  Every producer checks in when it starts its j-th tile and waits
        // -- for a bounded time: this is a hint, never a dependency -- until the others have started theirs.
        if (lane == 0) {
          atomicAdd(P.round_sync, 1u);
          const unsigned target = (round + 1) * gridDim.x;
          const long long t0 = clock64();
          while (*reinterpret_cast<volatile unsigned int *>(P.round_sync) < target && clock64() - t0 < 200000) __nanosleep(200);
        }
        __syncwarp();
        ++round;
      }
      const int2 st = P.tiles[t];
      const int row_a = (2 * st.x + cy) * CT;
      const int row_b = st.y * CT + cy * 64;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        if (elect_one()) {
          const uint32_t sa = base + (uint32_t)s * (uint32_t)CSTAGE_BYTES;
          if (cy == 0) mbar_expect_tx(full_bar(s), 2 * CSTAGE_BYTES);
          tma_load_2d_cg2(sa, &tmap, full_bar(s), kb * CBK, row_a);  // rows >= n are zero-filled
          tma_load_2d_cg2(sa + HALF_BYTES, &tmap, full_bar(s), kb * CBK, row_a + 64);
          tma_load_2d_cg2(sa + TILE_BYTES, &tmap, full_bar(s), kb * CBK, row_b);
        }
        __syncwarp();
        if (++s == CNSTAGE) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1 && cy == 0) {
    // ===== MMA issuer (the even CTA, for the pair): one accumulator stage per (tile, class) =====
    int s = 0;
    uint32_t ph = 0, nacc = 0;
    const uint32_t sfa = tmem_base + (uint32_t)C_SF_COL, sfb = sfa + 32u;
    for (int t = first; t < P.ntiles; t += step) {
      int kb = 0;
      for (int g = 0; g < nseg; ++g) {
        const int kend = s_seg_end[g];
        const uint32_t as = nacc & 1u, aph = (nacc >> 1) & 1u;
        ++nacc;
        mbar_wait(tempty_bar(as), aph ^ 1u);  // the epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * (uint32_t)CT;
        const int kb0 = kb;
        for (; kb < kend; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = base + (uint32_t)s * (uint32_t)CSTAGE_BYTES;
            const uint32_t lo_a = desc_lo(sa), lo_b = desc_lo(sa + TILE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // 32 bytes = 64 e2m1 along K per instruction = +2 in the address field
              const uint64_t da = ((uint64_t)DESC_HI_SW128 << 32) | (uint64_t)(lo_a + 2u * k);
              const uint64_t db = ((uint64_t)DESC_HI_SW128 << 32) | (uint64_t)(lo_b + 2u * k);
              umma_mxf4_cg2(tmem_d, da, db, C_IDESC, (k > 0) ? 1u : (uint32_t)(kb != kb0), sfa, sfb);
            }
            umma_commit_cg2(empty_bar(s));                       // the stage is free again in both CTAs
            if (kb == kend - 1) umma_commit_cg2(tfull_bar(as));  // both halves of the accumulator are complete
          }
          __syncwarp();
          if (++s == CNSTAGE) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp >= 2) {
    // ===== epilogue: warps 2..9; warp w reads TMEM lanes 32 (w & 3) .. +31 (tile rows), columns 64 h .. 64 h + 63 =====
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const double Meff = P.meff[0];
    const double pcq = P.pc / P.q, pcqq = pcq / P.q, omp = 1.0 - P.pc;
    uint32_t nacc = 0;
    for (int t = first; t < P.ntiles; t += step) {
      const int2 st = P.tiles[t];
      const int rb = 2 * st.x + cy, cb = st.y;
      double acc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] = 0.0;
      for (int g = 0; g < nseg; ++g) {
        const double w = s_seg_w[g];
        const uint32_t as = nacc & 1u, aph = (nacc >> 1) & 1u;
        ++nacc;
        mbar_wait(tfull_bar(as), aph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * (uint32_t)CT + (uint32_t)(half * 64);
        uint32_t v[32];
        tmem_ld32(taddr, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = fma(w, count_to_double(v[j]), acc[j]);
        tmem_ld32(taddr + 32u, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_even(tempty_bar(as));  // the counts are in registers: the next class may overwrite the stage
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[32 + j] = fma(w, count_to_double(v[j]), acc[32 + j]);
      }
      // ---- fused epilogue of the tile: 1/Meff, pseudocount, - Pi Pi'; the tile and (off the diagonal) its mirror image
      if (rb >= P.nblk || cb > rb) continue;  // computed for the neighbours' sake only
      const int r = rb * CT + quarter * 32 + lane;
      const int c0 = cb * CT + half * 64;
      const bool rok = r < (int)P.n;
      const int site_r = r / P.s;
      const double pir = rok ? P.Pi[r] : 0.0;
      const bool mirror = rb != cb;
      int site_c = c0 / P.s, rem = c0 - site_c * P.s;  // site of column c0 + j, advanced without divisions
      double *Crow = P.C + (long long)r * P.ld + c0;
      double *Ccol = P.C + (long long)c0 * P.ld + r;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const int c = c0 + j;
        if (c < (int)P.n) {  // warp-uniform
          const double ptrue = acc[j] / Meff;
          double out;
          if (P.raw) {
            out = ptrue;
          } else {
            const double pij = (site_c == site_r) ? omp * ptrue + ((c == r) ? pcq : 0.0) : omp * ptrue + pcqq;
            out = pij - pir * P.Pi[c];
          }
          if (rok) {
            Crow[j] = out;
            if (mirror) Ccol[(long long)j * P.ld] = out;
          }
        }
        if (++rem == P.s) {
          rem = 0;
          ++site_c;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still arrive on this CTA's barriers / read its shared memory until it is done too
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C_TMEM_COLS) : "memory");
  }
}

// Device groups: a member computes its pair tiles into its OWN copy of C and then pushes them to the leader in whole 1 KB row
// segments (the tile epilogue writes one orientation 8 bytes per lane and row: fine for local memory, where L2 merges the
// sectors, but a stream of 8-byte packets over NVLink -- 8 GPUs spent 5.7 ms in the stage that way).
__global__ void __launch_bounds__(256) push_tiles_kernel(const int2 *__restrict__ tiles, int nblk, long long n, long long ld,
                                                         const double *__restrict__ src, double *__restrict__ dst) {
  const int2 st = tiles[blockIdx.x >> 1];
  const int rb = 2 * st.x + (int)(blockIdx.x & 1), cb = st.y;
  if (rb >= nblk || cb > rb) return;
  for (int pass = 0; pass < (rb != cb ? 2 : 1); ++pass) {
    const long long r0 = (long long)(pass ? cb : rb) * CT, c0 = (long long)(pass ? rb : cb) * CT;
    for (int e = threadIdx.x; e < CT * CT / 2; e += 256) {
      const long long r = r0 + (e >> 6), c = c0 + (e & 63) * 2;
      if (r >= n || c >= n) continue;
      if (c + 1 < n)
        *reinterpret_cast<double2 *>(dst + r * ld + c) = *reinterpret_cast<const double2 *>(src + r * ld + c);
      else
        dst[r * ld + c] = src[r * ld + c];
    }
  }
}

// ---- weight classes ---------------------------------------------------------------------------------------------------
constexpr int HSM = 4096;  // count values below this are histogrammed in shared memory first (almost all of them)
// hist[v] = #{k : count[k] == v}, hist[M] = largest count value present
__global__ void __launch_bounds__(256) cls_hist_kernel(const int32_t *__restrict__ cnt, long long M, int32_t *__restrict__ hist) {
  __shared__ int sh[HSM];
  __shared__ int smax;
  for (int v = threadIdx.x; v < HSM; v += 256) sh[v] = 0;
  if (threadIdx.x == 0) smax = 0;
  __syncthreads();
  int mx = 0;
  for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < M; k += (long long)gridDim.x * 256) {
    const int v = cnt ? cnt[k] : 0;
    mx = max(mx, v);
    if (v < HSM) atomicAdd(&sh[v], 1); else atomicAdd(&hist[v], 1);
  }
  for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(&smax, mx);
  __syncthreads();
  for (int v = threadIdx.x; v < HSM && v < M; v += 256)
    if (sh[v]) atomicAdd(&hist[v], sh[v]);
  if (threadIdx.x == 0) atomicMax(&hist[M], smax);
}

// distinct count values in ascending order: out[0] = number of classes, out[1 + 2 c] = value, out[2 + 2 c] = size (first cap
// classes); only [0, largest value present] is scanned
__global__ void __launch_bounds__(1024) cls_compact_kernel(const int32_t *__restrict__ hist, long long M, int32_t *__restrict__ out, int cap) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int vmax = hist[M];
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int v0 = 0; v0 <= vmax; v0 += 1024) {
    const int v = v0 + tid;
    const int h = v <= vmax ? hist[v] : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, h != 0);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const int pos = before + __popc(bal & ((1u << lane) - 1u));
    if (h && pos < cap) {
      out[1 + 2 * pos] = v;
      out[2 + 2 * pos] = h;
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < 32; ++w) t += s_warp[w];
      s_base += t;
    }
    __syncthreads();
  }
  if (tid == 0) out[0] = s_base;
}

// perm[base(class of k) + running index] = k   (the order inside a class does not matter: the counts are exact integers).
// A block owns a contiguous range of the sequences: it counts its members per class in shared memory, reserves one range per
// class with ONE global atomic each, and hands out the slots from shared-memory cursors.
__global__ void __launch_bounds__(256) cls_assign_kernel(const int32_t *__restrict__ cnt, long long M, const int32_t *__restrict__ cls_val,
                                                         const long long *__restrict__ cls_base, int ncls, int32_t *__restrict__ cursor,
                                                         int32_t *__restrict__ perm) {
  __shared__ int s_val[MAXSEG], s_cnt[MAXSEG], s_off[MAXSEG];
  for (int c = threadIdx.x; c < ncls; c += 256) {
    s_val[c] = cls_val[c];
    s_cnt[c] = 0;
  }
  __syncthreads();
  const long long per = (M + gridDim.x - 1) / gridDim.x, k0 = per * blockIdx.x, k1 = k0 + per < M ? k0 + per : M;
  auto cls_of = [&](int v) {
    int lo = 0, hi = ncls - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_val[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  for (long long k = k0 + threadIdx.x; k < k1; k += 256) atomicAdd(&s_cnt[cls_of(cnt ? cnt[k] : 0)], 1);
  __syncthreads();
  for (int c = threadIdx.x; c < ncls; c += 256) {
    s_off[c] = s_cnt[c] ? atomicAdd(&cursor[c], s_cnt[c]) : 0;
    s_cnt[c] = 0;
  }
  __syncthreads();
  for (long long k = k0 + threadIdx.x; k < k1; k += 256) {
    const int c = cls_of(cnt ? cnt[k] : 0);
    perm[cls_base[c] + s_off[c] + atomicAdd(&s_cnt[c], 1)] = (int32_t)k;
  }
}

// Xt[(i, a)][kpos] = [Zt[i][perm[kpos]] == a + 1] as packed e2m1 (1.0 = 0x2; element 2b in the low nibble of byte b), one
// 32-bit word = 8 consecutive positions; perm < 0 (padding of a class to whole k-blocks) encodes zeros.
// A warp covers exactly one k-block (32 words = 256 positions), i.e. one class segment: while it encodes it also counts, per
// state, the sequences of that segment carrying the state at this site (segcnt[i][segment][a], the numerators of Pi).
__global__ void __launch_bounds__(256) encode_onehot4_kernel(const int8_t *__restrict__ Zt, long long M, const int32_t *__restrict__ perm,
                                                             long long words_per_row, int s, const int *__restrict__ seg_end, int nseg,
                                                             uint32_t *__restrict__ Xt, int32_t *__restrict__ segcnt) {
  const long long w = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long i = blockIdx.y;
  const bool live = w < words_per_row;
  int zz[8];
  if (live) {
    const int4 p0 = reinterpret_cast<const int4 *>(perm)[2 * w], p1 = reinterpret_cast<const int4 *>(perm)[2 * w + 1];
    const int p[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    const int8_t *z = Zt + i * M;
#pragma unroll
    for (int e = 0; e < 8; ++e) zz[e] = p[e] >= 0 ? (int)z[p[e]] : 0;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) zz[e] = 0;
  }
  // the segment of this warp's k-block (warp-uniform)
  const int kb = (int)(w >> 5);
  int lo = 0, hi = nseg - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (seg_end[mid] <= kb) lo = mid + 1; else hi = mid;
  }
  uint32_t *out = Xt + (i * s) * words_per_row + w;
  int32_t *sc = segcnt + (i * nseg + lo) * 32;
  const int lane = threadIdx.x & 31;
  for (int a = 1; a <= s; ++a) {
    uint32_t word = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) word |= (zz[e] == a) ? (0x2u << (4 * e)) : 0u;
    if (live) out[(long long)(a - 1) * words_per_row] = word;
    const int tot = __reduce_add_sync(0xffffffffu, __popc(word));
    if (lane == 0 && tot) atomicAdd(sc + a, tot);
  }
}

// Pi[(i,a)] = (1 - pc) * (sum_segments w_seg * segcnt[i][seg][a]) / Meff + pc / q, segments in ascending order (deterministic)
__global__ void __launch_bounds__(128) pi_from_segments_kernel(const int32_t *__restrict__ segcnt, const double *__restrict__ seg_w, int nseg,
                                                               const double *__restrict__ meff, int s, int q, double pc,
                                                               double *__restrict__ Pi) {
  const long long i = blockIdx.x;
  const int a = threadIdx.x + 1;
  if (a > s) return;
  double acc = 0.0;
  for (int g = 0; g < nseg; ++g) acc = fma(seg_w[g], (double)segcnt[(i * nseg + g) * 32 + a], acc);
  Pi[i * s + (a - 1)] = (1.0 - pc) * (acc / meff[0]) + pc / q;
}

}  // namespace

// plan of the class segments (host): -> false when the tensor-core engine should not run (too many / too small classes)
struct CovTcPlan {
  std::vector<int> seg_end;
  std::vector<double> seg_w;
  std::vector<int> cls_val;
  std::vector<long long> cls_base;
  long long kblocks = 0;
  double cost_clk_per_tile = 0.0;  // model: MMA clocks, or the FP64 drain of a class when that is longer
};

static bool plan_classes(const int32_t *cls, int ncls, CovTcPlan &pl) {
  pl = CovTcPlan{};
  long long kb = 0;
  for (int c = 0; c < ncls; ++c) {
    const int val = cls[2 * c];
    long long left = cls[2 * c + 1];
    pl.cls_val.push_back(val);
    pl.cls_base.push_back(kb * CSEQ);
    const double w = 1.0 / (double)(val + 1);  // the same IEEE division as weights_kernel: W[k] of the class, bit for bit
    while (left > 0) {  // a class larger than 2^23 sequences is cut into several segments (FP32 accumulation stays exact)
      const long long take = left < SEG_MAX_SEQ ? left : SEG_MAX_SEQ;
      const long long nb = (take + CSEQ - 1) / CSEQ;
      kb += nb;
      if ((int)pl.seg_end.size() >= MAXSEG || kb > 0x7fffffffll / CBK) return false;
      pl.seg_end.push_back((int)kb);
      pl.seg_w.push_back(w);
      pl.cost_clk_per_tile += std::max(272.0 * (double)nb, 900.0);
      left -= take;
    }
  }
  pl.kblocks = kb;
  return true;
}

// distinct neighbour counts of the loaded alignment and how many sequences carry each: host_out[0] = number of classes,
// host_out[1 + 2 c] = count value (without the sequence itself), host_out[2 + 2 c] = size, ascending, first GDCA_COV_MAXCLS classes
int32_t gdca_k_cov_classes(gdca_ctx *ctx, int32_t *host_out) {
  const long long M = ctx->M;
  constexpr int CAP = MAXSEG;
  const int32_t *cnt = ctx->counts_row < 0 ? nullptr : ctx->dCounts + (size_t)ctx->counts_row * ctx->Mpad;
  GDCA_TRY(gdca_reserve(ctx, ctx->dClsHist, ctx->capClsHist, (size_t)M + 4 * CAP + 16));
  int32_t *hist = ctx->dClsHist, *compact = ctx->dClsHist + M + 1;  // hist[M] = largest value; compact [1 + 2 CAP]
  GDCA_CUDA(ctx, cudaMemsetAsync(hist, 0, (size_t)(M + 1) * sizeof(int32_t), ctx->stream));
  cls_hist_kernel<<<ctx->num_sms, 256, 0, ctx->stream>>>(cnt, M, hist);
  GDCA_LAUNCH_CHECK(ctx);
  cls_compact_kernel<<<1, 1024, 0, ctx->stream>>>(hist, M, compact, CAP);
  GDCA_LAUNCH_CHECK(ctx);
  GDCA_CUDA(ctx, cudaMemcpyAsync(host_out, compact, (size_t)(1 + 2 * CAP) * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GDCA_OK;
}

int32_t gdca_k_covariance_tc(gdca_ctx *ctx, double pc, bool raw, bool *done) {
  *done = false;
  const long long L = ctx->L, M = ctx->M, n = ctx->n, npad = ctx->npad;
  const int32_t *cnt = ctx->counts_row < 0 ? nullptr : ctx->dCounts + (size_t)ctx->counts_row * ctx->Mpad;
  // ---- classes of equal weight (device groups: the leader's copy, handed over by run_group)
  constexpr int CAP = MAXSEG;
  std::vector<int32_t> h(1 + 2 * CAP);
  if (ctx->cov_cls_host)
    std::copy(ctx->cov_cls_host, ctx->cov_cls_host + h.size(), h.begin());
  else
    GDCA_TRY(gdca_k_cov_classes(ctx, h.data()));
  const int ncls = h[0];
  ctx->cov_tc_classes = ncls;
  if (ncls < 1 || ncls > CAP) return GDCA_OK;
  CovTcPlan pl;
  if (!plan_classes(h.data() + 1, ncls, pl)) return GDCA_OK;
  const int nblk = (int)(npad / CT), ns = (nblk + 1) / 2;
  if (ctx->cov_engine == 0) {
    // auto: run on the tensor cores when the model says they win.  Scatter-add engine: M L^2 / 2 additions at the measured
    // 1.36e12 additions/s (18.4 ms at config C); this engine: super-tile rounds x clocks per tile at 1.9 GHz, plus the encode.
    const double t_sparse = 0.5 * (double)M * (double)L * (double)L / 1.36e12;
    const double rounds = ceil((double)ns * (ns + 1) / (double)(ctx->num_sms / 2));
    const double t_tc = rounds * pl.cost_clk_per_tile * 1.5 / 1.9e9 + (double)n * (double)pl.kblocks * CBK / 3.0e12 + 60e-6;
    if (t_tc >= t_sparse) return GDCA_OK;
  }
  const long long Kbytes = pl.kblocks * CBK, Mk = pl.kblocks * CSEQ;
  const int nseg = (int)pl.seg_end.size();
  // ---- pair tiles of the lower triangle: (RB2, CB) = row blocks 2 RB2, 2 RB2 + 1 x column block CB, in bands of 8 pair rows with the
  // row varying fastest (the pairs running at one time share a few row and column blocks: Xt streams from HBM about once per band);
  // dealt round-robin to the members of a device group
  std::vector<int2> tiles;
  {
    constexpr int BH = 8;
    long long idx = 0;
    for (int r0 = 0; r0 < ns; r0 += BH) {
      const int r1 = std::min(ns, r0 + BH);
      for (int c = 0; c < std::min(nblk, 2 * r1); ++c)
        for (int r = std::max(r0, c / 2); r < r1; ++r, ++idx)
          if (idx % ctx->shard_world == ctx->shard_rank) tiles.push_back(make_int2(r, c));
    }
  }
  const int ntiles = (int)tiles.size();
  GDCA_TRY(gdca_reserve(ctx, ctx->dCovTiles, ctx->capCovTiles, (size_t)std::max(ntiles, 1)));
  // ---- sequences in class order, operand matrix
  GDCA_TRY(gdca_reserve(ctx, ctx->dClsPerm, ctx->capClsPerm, (size_t)Mk + 4 * CAP));
  GDCA_TRY(gdca_reserve(ctx, ctx->dClsTab, ctx->capClsTab, (size_t)8 * CAP));
  GDCA_TRY(gdca_reserve(ctx, ctx->dXt, ctx->capXt, (size_t)n * Kbytes));
  int32_t *perm = ctx->dClsPerm, *cursor = ctx->dClsPerm + Mk, *d_val = cursor + CAP, *d_segend = d_val + CAP;
  long long *d_base = reinterpret_cast<long long *>(ctx->dClsTab);
  double *d_segw = reinterpret_cast<double *>(ctx->dClsTab) + CAP;
  // The small tables go through PINNED host memory: a copy from pageable memory makes the host wait for the stream, which in a
  // device group serialises the members' covariance stages behind one another (8 GPUs: 5.7 ms instead of 1.3 ms).
  {
    const size_t need = (size_t)ntiles * sizeof(int2) + (size_t)ncls * (sizeof(int) + sizeof(long long)) +
                        (size_t)nseg * (sizeof(int) + sizeof(double)) + 64;
    if (need > ctx->capHostTab) {
      if (ctx->hostTab) cudaFreeHost(ctx->hostTab);
      ctx->hostTab = nullptr;
      ctx->capHostTab = 0;
      GDCA_CUDA(ctx, cudaHostAlloc(&ctx->hostTab, 2 * need, cudaHostAllocDefault));
      ctx->capHostTab = 2 * need;
    }
    char *h = static_cast<char *>(ctx->hostTab);
    auto put = [&](void *dst, const void *src, size_t bytes) -> int32_t {
      if (!bytes) return GDCA_OK;
      memcpy(h, src, bytes);
      GDCA_CUDA(ctx, cudaMemcpyAsync(dst, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
      h += (bytes + 15) / 16 * 16;
      return GDCA_OK;
    };
    GDCA_TRY(put(ctx->dCovTiles, tiles.data(), (size_t)ntiles * sizeof(int2)));
    GDCA_TRY(put(d_val, pl.cls_val.data(), (size_t)ncls * sizeof(int)));
    GDCA_TRY(put(d_base, pl.cls_base.data(), (size_t)ncls * sizeof(long long)));
    GDCA_TRY(put(d_segend, pl.seg_end.data(), (size_t)nseg * sizeof(int)));
    GDCA_TRY(put(d_segw, pl.seg_w.data(), (size_t)nseg * sizeof(double)));
  }
  GDCA_CUDA(ctx, cudaMemsetAsync(perm, 0xFF, (size_t)Mk * sizeof(int32_t), ctx->stream));
  GDCA_CUDA(ctx, cudaMemsetAsync(cursor, 0, (size_t)CAP * sizeof(int32_t), ctx->stream));
  cls_assign_kernel<<<ctx->num_sms, 256, 0, ctx->stream>>>(cnt, M, d_val, d_base, ncls, cursor, perm);
  GDCA_LAUNCH_CHECK(ctx);
  const long long wpr = Kbytes / 4;
  GDCA_TRY(gdca_reserve(ctx, ctx->dSegCnt, ctx->capSegCnt, (size_t)L * nseg * 32));
  GDCA_CUDA(ctx, cudaMemsetAsync(ctx->dSegCnt, 0, (size_t)L * nseg * 32 * sizeof(int32_t), ctx->stream));
  encode_onehot4_kernel<<<dim3((unsigned)((wpr + 255) / 256), (unsigned)L), 256, 0, ctx->stream>>>(
      ctx->dZt, M, perm, wpr, ctx->s, d_segend, nseg, reinterpret_cast<uint32_t *>(ctx->dXt), ctx->dSegCnt);
  GDCA_LAUNCH_CHECK(ctx);
  // Pi from the class counts: the same exact integers the tiles accumulate (Pi_true is the diagonal of Pij_true)
  pi_from_segments_kernel<<<(unsigned)L, 128, 0, ctx->stream>>>(ctx->dSegCnt, d_segw, nseg, ctx->dMeff, ctx->s, ctx->q, pc, ctx->dPi);
  GDCA_LAUNCH_CHECK(ctx);
  // the padding strips of C (rows / columns n .. npad-1) must be zero; everything else is written by the tiles
  const bool peer_out = ctx->peers_ready && ctx->shard_world > 1;
  if (!peer_out && npad > n) {
    GDCA_CUDA(ctx, cudaMemsetAsync(ctx->dC + n * npad, 0, (size_t)(npad - n) * npad * sizeof(double), ctx->stream));
    GDCA_CUDA(ctx, cudaMemset2DAsync(ctx->dC + n, (size_t)npad * sizeof(double), 0, (size_t)(npad - n) * sizeof(double), (size_t)n, ctx->stream));
  }
  CUtensorMap map;
  GDCA_TRY(gdca_make_tensor_map_2d(ctx, &map, ctx->dXt, n, Kbytes, 64));
  CovTcParams P;
  P.tiles = ctx->dCovTiles;
  P.ntiles = ntiles;
  P.nblk = nblk;
  P.nseg = nseg;
  P.seg_end = d_segend;
  P.seg_w = d_segw;
  P.meff = ctx->dMeff;
  P.Pi = ctx->dPi;
  const bool push = peer_out && ctx->peer_C[0] != ctx->dC;  // a member other than the leader: compute locally, then push
  P.C = ctx->dC;
  P.n = n;
  P.ld = npad;
  P.s = ctx->s;
  P.q = ctx->q;
  P.raw = raw ? 1 : 0;
  P.pc = pc;
  // round hint of the producers: when the operand is several times the size of the L2 cache (1 GB at config C, 15 GB at E: kernel 350 -> 146 ms there)
  P.round_sync = nullptr;
  if (ctx->cov_round_sync == 1 || (ctx->cov_round_sync < 0 && (double)n * (double)Kbytes > 5.0e8)) {
    P.round_sync = ctx->dCovSync;
    GDCA_CUDA(ctx, cudaMemsetAsync(P.round_sync, 0, sizeof(unsigned int), ctx->stream));
  }
  GDCA_CUDA(ctx, cudaFuncSetAttribute(cov_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_SMEM));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(C_THREADS);
  cfg.dynamicSmemBytes = C_SMEM;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3((unsigned)(ctx->num_sms / 2 * 2));
  int max_clusters = ctx->cov_max_clusters;  // per device: asked once
  if (max_clusters <= 0) {
    GDCA_CUDA(ctx, cudaOccupancyMaxActiveClusters(&max_clusters, cov_tc_kernel, &cfg));
    ctx->cov_max_clusters = max_clusters;
  }
  int nclusters = std::min(std::min(max_clusters, ctx->num_sms / 2), std::max(ntiles, 1));
  if (nclusters < 1) return gdca_fail(ctx, GDCA_ERR_CUDA, "covariance (tensor cores): no CTA pair fits on this device");
  cfg.gridDim = dim3((unsigned)(2 * nclusters));
  GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_cov0, ctx->stream));
  GDCA_CUDA(ctx, cudaLaunchKernelEx(&cfg, cov_tc_kernel, map, P));
  GDCA_LAUNCH_CHECK(ctx);
  GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_cov1, ctx->stream));
  if (push && ntiles > 0) {
    push_tiles_kernel<<<(unsigned)(2 * ntiles), 256, 0, ctx->stream>>>(ctx->dCovTiles, nblk, n, npad, ctx->dC, ctx->peer_C[0]);
    GDCA_LAUNCH_CHECK(ctx);
  }
  ctx->cov_tc_kblocks = pl.kblocks;
  ctx->cov_tc_segments = nseg;
  ctx->cov_tc_clusters = nclusters;
  // 128 x 128 x 256 MMA work of every CTA tile visited (incl. the padded / mirrored tiles of diagonal super-tiles)
  ctx->cov_tc_tflop = 2.0 * CT * CT * (double)Mk * 2.0 * (double)ntiles * 1e-12;
  ctx->cov_tc_l2_bytes = (double)ntiles * 2.0 * (double)pl.kblocks * CSTAGE_BYTES;
  *done = true;
  return GDCA_OK;
}
