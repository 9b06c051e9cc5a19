// tcfilter.cu -- K3a: tensor-core prefilter of the neighbour-count sweep (tcgen05 / TMEM / TMA, sm_100a).
//
// Part of the replacement of DCAUtils compute_weights (un-vendored; reference call site src/GaussDCA.jl:28):
//   count[k] = 1 + #{l != k : hamming(k,l) < thresh}.
// On real and synthetic alignments almost every pair is far beyond thresh.  This kernel PROVES that for whole 32 x 32 cells
// of sequence pairs on the tensor cores; only the 128 x 128 blocks with a cell left go to the exact bit-plane sweep
// (pairs.cu), and there only the warps that own such a cell work.  The result is unchanged by construction:
//
//   * every residue state is mapped to one of 4 classes (state & 3) and each class to a vertex of the regular
//     simplex in {-1,+1}^3:  c0=(+,+,+) c1=(+,-,-) c2=(-,+,-) c3=(-,-,+);  v_a . v_b = 3 if a == b else -1.
//   * for two sequences  S = sum_i v(Z_ik) . v(Z_il) = 4 * ident_proj - L,  ident_proj = #sites with equal CLASS
//     >= ident (equal states have equal classes), so  hamming_proj = (3L - S) / 4  <=  hamming.
//   * a pair with hamming_proj >= thresh cannot be a neighbour.  A cell is cleared iff that holds for all its pairs, i.e.
//     iff max S <= 3L - 4 thresh.  S is an exact integer (|S| <= 3L < 2^24: FP32 accumulation of +-1 products, or S32
//     for the INT8 variant), so the test is exact and conservative: flagged cells are a superset of the cells that contain
//     a neighbour pair; the sweep counts exactly in those.
//
// The M x M x 3L contraction runs as  V V^T  with V = [rows][3L padded to whole 128-byte k-blocks], K-major:
//   * one persistent CTA per SM, 192 threads: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one elected lane),
//     warps 2..9 = epilogue (two per TMEM lane quarter, alternate 32 x 32 cells);
//   * operand variants of one kernel template:
//       FP4 (default)  packed e2m1 +-1.0, kind::mxf4.block_scale.block32, UMMA 128 x 224 x 64, two accumulator stages of
//                      224 TMEM columns + 64 columns of UE8M0 scale factors that are all 1.0 (written once with tcgen05.st:
//                      every byte is 0x7F, so the scale-factor layout never matters); 5-stage TMA ring of 44 KB;
//       FP8            e4m3 +-1.0, kind::f8f6f4, UMMA 128 x 256 x 32, two stages of 256 columns, 4-stage ring of 48 KB;
//       INT8           +-1 bytes, kind::i8, S32 accumulators, same shape as FP8;
//     K is streamed in 128-byte blocks (SWIZZLE_128B), full/empty mbarriers, tcgen05.commit releases the stages;
//   * default launch: clusters of 2 CTAs on neighbouring row blocks of the same column tile; each CTA fetches its A tile and
//     half of the B tile, the half is TMA-multicast into both (-32 % L2->SM operand bytes);
//   * the epilogue of tile t (tcgen05.ld, max per 32 x 32 cell, one vote and at most one atomicOr per cell) overlaps
//     the MMAs of tile t+1.  No C matrix is ever written: the output is a 16-bit cell mask per 128 x 128 block.
//   * only tiles that touch the upper triangle are visited, in bands of 16 row blocks with the row block varying
//     fastest: the CTAs running at one time share a few column tiles, so V streams from HBM once per band and the
//     operands come out of L2; multi-GPU: rank r owns the row blocks bi = r (mod world).
#include <cuda.h>  // CUtensorMap (types only: cuTensorMapEncodeTiled is fetched through the runtime)

#include "gdca_internal.cuh"

namespace {

constexpr int BM = 128;              // tile rows (sequences)
constexpr int BK = 128;              // K bytes per stage (= SWIZZLE_128B atom width)
constexpr int MAX_STAGE = 7;
constexpr int TC_THREADS = 320;       // warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue (two per TMEM lane quarter: alternate cells)
constexpr int TMEM_COLS = 512;
constexpr int BAND = 16;             // row blocks per band of the tile order

enum Operand { OP_FP8 = 0, OP_FP4 = 1, OP_I8 = 2 };  // e4m3 / packed e2m1 / signed int8 (exact S32 accumulation)

template <int OP>
struct Cfg {
  static constexpr bool FP4 = OP == OP_FP4;
  static constexpr int BN = FP4 ? 224 : 256;        // tile columns (sequences)
  static constexpr int A_BYTES = BM * BK;           // 16 KB
  static constexpr int B_BYTES = BN * BK;           // 28 / 32 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NSTAGE = FP4 ? 5 : 4;        // 5 x 44 KB / 4 x 48 KB of operands in flight
  static constexpr int SF_COL = 2 * BN;             // FP4: scale factors behind the two accumulator stages (448..511)
  static constexpr size_t SMEM = (size_t)NSTAGE * STAGE_BYTES + 1024 /*alignment slack*/;
  // instruction descriptors (cute::UMMA::InstrDescriptor / InstrDescriptorBlockScaled bit layout), both operands K-major:
  //   f8f6f4:          D=F32 (bit 4), A=B=E4M3 (0), N>>3 at bit 17, M>>4 at bit 24
  //   mxf4 block32:    A=B=E2M1 (1 at bits 7 and 10), scale format UE8M0 (bit 23), N>>3 at bit 17, M>>4 at bit 24, K=64
  //   i8:              D=S32 (2 at bit 4), A=B=signed INT8 (1 at bits 7 and 10)
  static constexpr uint32_t IDESC =
      FP4 ? ((1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | (1u << 23) | ((uint32_t)(BM >> 4) << 24))
      : OP == OP_I8 ? ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24))
                    : ((1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24));
};

// cta_group::2 (FP4 only): the two CTAs of a cluster form ONE 256 x 224 MMA -- each keeps its own 128 rows of A and only HALF of
// the B tile (the tensor cores read the other half from the peer's shared memory): 30 KB instead of 44 KB of operands arrive in
// every SM per k-block (the L2 -> SM path, ~70 B/clk/SM, is what bounds the multicast variant at 0.75 of the FP4 rate), 7 stages
constexpr int CG2_STAGE_BYTES = BM * BK + (224 / 2) * BK;  // 30 KB
constexpr int CG2_NSTAGE = 7;
constexpr size_t CG2_SMEM = (size_t)CG2_NSTAGE * CG2_STAGE_BYTES + 1024;
constexpr int PAIR_CHUNK = 256;                            // candidate pairs an epilogue warp reserves at a time
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;                // shared::cluster address of the same offset in the even CTA of the pair

struct FilterParams {
  int T, NT, NB, KB;    // 128-row blocks, column tiles, bands of this rank's rows, 128-byte k-blocks
  int rank, world;
  float bound;          // a 32 x 32 cell is flagged iff max S > bound,  bound = 3L - 4 thresh
  uint32_t *flags;      // [T][T] cell masks: bit 4*(row/32) + (col/32) of block (bi, bj)
  float *dump;          // tests only: S of every visited tile, [T*128][dump_ld]
  long long dump_ld;
  // candidate PAIRS: every (k, l), k < l < M, whose projected distance is below thresh (S > bound) is appended here; the exact
  // stage then checks these pairs alone (pairs.cu: pair_list_kernel).  The counter keeps counting past the capacity: a list
  // that overflowed is ignored and the flagged cells are swept instead.
  int2 *pairs;
  unsigned long long *npairs;
  unsigned long long pair_cap;
  long long M;
};

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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// the same box into the same shared-memory offset (and onto the same mbarrier offset) of every CTA of the cluster in `mask`
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(dst),
      "l"(map), "r"(bar), "h"(mask), "r"(c0), "r"(c1)
      : "memory");
}
// tcgen05.commit that arrives on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one lane of the (converged) warp; the compiler keeps warp-uniform operands in uniform registers around it
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);  // start address
  d |= (uint64_t)1 << 16;                  // leading byte offset (unused for swizzled K-major; CUTLASS writes 1)
  d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                  // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_mxf4(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate,
                                          uint32_t tmem_sfa, uint32_t tmem_sfb) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}

// ---- cta_group::2 forms: one MMA across the CTA pair, issued by the even CTA ----
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  // executed by both CTAs: the bytes land in the issuing CTA's shared memory, the transaction count on the EVEN CTA's barrier
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_mxf4_cg2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate,
                                              uint32_t tmem_sfa, uint32_t tmem_sfb) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}
// arrives on the mbarrier at this offset in both CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// plain arrival on the EVEN CTA's barrier at this offset (from either CTA)
__device__ __forceinline__ void mbar_arrive_even(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_MASK) : "memory");
}

// 32 consecutive 32-bit columns of this thread's TMEM lane
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
// the same 32-bit word into 32 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st32_fill(uint32_t taddr, uint32_t v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(v)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Tile order: this rank owns the row blocks bi = rank + world * i; bands of BAND of ITS rows; inside band b the column
// tile cj runs from the first one that reaches the band's diagonal to NT-1, and the row varies FASTEST -- so the ~148 tiles
// in flight at one time use 16 A tiles and a few B tiles out of L2.  (row, cj) pairs of a band that lie wholly below the
// diagonal are skipped (valid() == false).  Whole rows per rank: every 128 x 128 block is flagged by exactly one rank.
struct TileIter {
  int T, NT, NB, colw, rank, world;
  int b, u, cmin_b, tiles_b;  // band, index inside the band, first column tile and tile count of the band (cached)
  __host__ __device__ __forceinline__ int row_of(int i) const { return rank + world * i; }
  __host__ __device__ __forceinline__ void set_band() {
    cmin_b = (int)(((long long)BM * row_of(b * BAND)) / colw);
    tiles_b = (NT - cmin_b) * BAND;
  }
  __host__ __device__ __forceinline__ void start(const FilterParams &P, int colw_, int first) {
    T = P.T;
    NT = P.NT;
    NB = P.NB;
    colw = colw_;
    rank = P.rank;
    world = P.world;
    b = 0;
    u = 0;
    if (NB > 0) set_band();
    advance(first);
  }
  __host__ __device__ __forceinline__ void advance(int d) {
    u += d;
    while (b < NB && u >= tiles_b) {
      u -= tiles_b;
      if (++b < NB) set_band();
    }
  }
  __host__ __device__ __forceinline__ bool done() const { return b >= NB; }
  __host__ __device__ __forceinline__ int bi() const { return row_of(b * BAND + (u & (BAND - 1))); }
  __host__ __device__ __forceinline__ int cj() const { return cmin_b + (u / BAND); }
  __host__ __device__ __forceinline__ bool valid_row(int row) const {
    return row < T && (long long)colw * (cj() + 1) > (long long)BM * row;
  }
  __host__ __device__ __forceinline__ bool valid() const { return valid_row(bi()); }
  // the tile of the other CTA of a 2-CTA cluster: the neighbouring row of the same band, same column tile (BAND is even)
  __host__ __device__ __forceinline__ bool peer_valid() const { return valid_row(row_of(b * BAND + ((u & (BAND - 1)) ^ 1))); }
};
static_assert((BAND & (BAND - 1)) == 0, "BAND must be a power of two");

// MC: launched as clusters of 2 CTAs that work on neighbouring row blocks of the SAME column tile.  Each CTA loads its own A
// tile and one half of the B tile; the half is TMA-multicast into both CTAs (-32 % L2->SM operand traffic at 128 x 224).  A
// shared-memory stage is therefore written by both CTAs' producers: its empty barrier counts the tcgen05.commit of BOTH
// CTAs (multicast commit), its full barrier the bytes of all three boxes.
// MC = 2: the cta_group::2 pair (FP4): same tile pairing as MC = 1, no multicast -- each CTA loads its A tile and its half of B.
template <int OP, int MC>
__global__ void __launch_bounds__(TC_THREADS, 1)
    tc_filter_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, FilterParams P) {
  using C = Cfg<OP>;
  constexpr bool FP4 = C::FP4;
  constexpr bool CG2 = MC == 2;
  static_assert(!CG2 || FP4, "cta_group::2 variant: FP4 operands only");
  const uint32_t crank = MC ? cluster_ctarank() : 0u;
  constexpr int BN = C::BN;
  constexpr int NSTAGE = CG2 ? CG2_NSTAGE : C::NSTAGE;
  constexpr int STAGE_BYTES = CG2 ? CG2_STAGE_BYTES : C::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bars[2 * MAX_STAGE + 4];
  __shared__ uint32_t s_tmem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B operands need 1024-byte alignment
  const uint32_t bars = smem_u32(s_bars);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (NSTAGE + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * NSTAGE + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * NSTAGE + 2 + a); };
  const uint32_t tmem_slot = smem_u32(&s_tmem);
  volatile uint32_t *tmem_slot_ptr = &s_tmem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapB) : "memory");
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), MC == 1 ? 2 : 1);  // multicast pair: both producers write the stage; cta_group::2: one commit reaches both CTAs
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), CG2 ? 16 : 8);  // one arrival per epilogue warp (cta_group::2: of both CTAs, on the even CTA's barrier)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (CG2) {
    __syncthreads();
    cluster_sync_all();  // both CTAs are resident and their barriers initialised before the pair allocates tensor memory
  }
  if (warp == 1) {
    if (CG2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();  // the peer's barriers are initialised before anything of this CTA can arrive on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (FP4 && warp >= 2 && warp < 6) {
    // UE8M0 scale factors, all 2^0: every byte of the 64 scale columns of all 128 lanes is 0x7F
    const uint32_t t = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)C::SF_COL;
    tmem_st32_fill(t, 0x7F7F7F7Fu);
    tmem_st32_fill(t + 32u, 0x7F7F7F7Fu);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int first = (int)blockIdx.x;
  const int step = (int)gridDim.x;

  // The producer and the MMA issuer are ONE thread's instruction stream each and the tensor pipe can only be as fast as
  // they are: the whole warp runs the loops (warp-uniform control flow, operands in uniform registers), one elected lane
  // issues, slot / phase advance by increments (no divisions), descriptors are a constant plus the stage offset.
  if (warp == 0) {
    // ===== TMA producer =====
    TileIter it;
    int s = 0;
    uint32_t ph = 0;  // ring slot and phase, carried across tiles
    for (it.start(P, BN, first); !it.done(); it.advance(step)) {
      if (!(it.valid() || (MC && it.peer_valid()))) continue;
      const int row_a = it.bi() * BM;
      const int row_b = it.cj() * BN + (MC ? (int)crank * (BN / 2) : 0);
      for (int kb = 0; kb < P.KB; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        if (elect_one()) {
          const uint32_t sa = base + (uint32_t)s * (uint32_t)STAGE_BYTES;
          if (CG2) {
            // the even CTA's barrier collects the bytes of both CTAs; every CTA keeps its A tile and ITS half of the B tile
            if (crank == 0) mbar_expect_tx(full_bar(s), 2 * STAGE_BYTES);
            tma_load_2d_cg2(sa, &tmapA, full_bar(s), kb * BK, row_a);
            tma_load_2d_cg2(sa + C::A_BYTES, &tmapB, full_bar(s), kb * BK, row_b);
          } else {
            mbar_expect_tx(full_bar(s), STAGE_BYTES);
            tma_load_2d(sa, &tmapA, full_bar(s), kb * BK, row_a);  // rows beyond the matrix are zero-filled
            if (MC)  // tmapB boxes are BN/2 rows here: my half of the B tile, delivered to both CTAs
              tma_load_2d_mc(sa + C::A_BYTES + crank * (C::B_BYTES / 2), &tmapB, full_bar(s), kb * BK, row_b, (uint16_t)3);
            else
              tma_load_2d(sa + C::A_BYTES, &tmapB, full_bar(s), kb * BK, row_b);
          }
        }
        __syncwarp();
        if (++s == NSTAGE) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1 && !(CG2 && crank != 0)) {
    // ===== MMA issuer (cta_group::2: the even CTA issues for the pair) =====
    TileIter it;
    int s = 0;
    uint32_t ph = 0, n = 0;
    // descriptor halves: hi = SBO 1024 B | version 1 | SWIZZLE_128B, lo = start address >> 4 | LBO field 1
    constexpr uint32_t DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    const uint32_t sfa = tmem_base + (uint32_t)C::SF_COL, sfb = sfa + 32u;
    for (it.start(P, BN, first); !it.done(); it.advance(step)) {
      if (!(it.valid() || (MC && it.peer_valid()))) continue;
      const uint32_t as = n & 1u, aph = (n >> 1) & 1u;
      ++n;
      mbar_wait(tempty_bar(as), aph ^ 1u);  // the epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * (uint32_t)BN;
      for (int kb = 0; kb < P.KB; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = base + (uint32_t)s * (uint32_t)STAGE_BYTES;
          const uint32_t lo_a = ((sa >> 4) & 0x3FFFu) | (1u << 16), lo_b = (((sa + C::A_BYTES) >> 4) & 0x3FFFu) | (1u << 16);
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 32 bytes along K per instruction (32 e4m3 / 64 e2m1) = +2 in the address field
            const uint64_t da = ((uint64_t)DESC_HI << 32) | (uint64_t)(lo_a + 2u * k);
            const uint64_t db = ((uint64_t)DESC_HI << 32) | (uint64_t)(lo_b + 2u * k);
            const uint32_t accumulate = (k > 0) ? 1u : (uint32_t)(kb != 0);
            if (CG2)  // M = 256 across the pair
              umma_mxf4_cg2(tmem_d, da, db, (C::IDESC & ~(0x1Fu << 24)) | ((uint32_t)(2 * BM >> 4) << 24), accumulate, sfa, sfb);
            else if (FP4)
              umma_mxf4(tmem_d, da, db, C::IDESC, accumulate, sfa, sfb);
            else if (OP == OP_I8)
              umma_i8(tmem_d, da, db, C::IDESC, accumulate);
            else
              umma_f8(tmem_d, da, db, C::IDESC, accumulate);
          }
          if (CG2) {
            umma_commit_cg2(empty_bar(s));                          // the stage of BOTH CTAs is free again
            if (kb == P.KB - 1) umma_commit_cg2(tfull_bar(as));     // both halves of the accumulator are complete
          } else {
            if (MC)
              umma_commit_mc(empty_bar(s), (uint16_t)3);  // both producers write this stage: tell both
            else
              umma_commit(empty_bar(s));  // frees the smem stage once these MMAs have read it
            if (kb == P.KB - 1) umma_commit(tfull_bar(as));  // accumulator complete
          }
        }
        __syncwarp();
        if (++s == NSTAGE) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp >= 2) {
    // ===== epilogue: 4 warps, warp w owns TMEM lanes 32 (w & 3) .. +31 = tile rows =====
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;  // this warp reduces the cells half, half + 2, ... of its lane quarter
    const int row = quarter * 32 + lane;
    TileIter it;
    uint32_t n = 0;
    // this warp's chunk of the candidate-pair list (warp-uniform)
    unsigned long long chunk_base = 0;
    int chunk_used = PAIR_CHUNK;   // "full": the first candidates reserve a chunk
    bool chunk_ok = false;
    for (it.start(P, BN, first); !it.done(); it.advance(step)) {
      if (!(it.valid() || (MC && it.peer_valid()))) continue;
      const int bi = it.bi(), cj = it.cj();
      const bool mine = it.valid();  // MC: a tile that only exists because the peer's does is computed but never flagged
      const uint32_t as = n & 1u, aph = (n >> 1) & 1u;
      ++n;
      mbar_wait(tfull_bar(as), aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * (uint32_t)BN;
      // one 32 x 32 cell per warp and step.  Measured: neither a software-pipelined tcgen05.ld, nor 8 epilogue warps, nor
      // skipping the reads altogether changes the kernel time -- the epilogue is hidden behind the MMAs of the next tile.
      auto reduce_cell = [&](const uint32_t (&v)[32], int c) {
        const long long col0 = (long long)cj * BN + c * 32;
        if (P.dump && mine) {
          float *d = P.dump + ((long long)bi * BM + row) * P.dump_ld + col0;
#pragma unroll
          for (int j = 0; j < 32; ++j) d[j] = OP == OP_I8 ? (float)(int)v[j] : __uint_as_float(v[j]);
        }
        float mx = -3.0e38f;
        if (OP == OP_I8) {  // S32 accumulators: exact integers, |S| <= 3L < 2^24 converts exactly
          int mi = (int)0x80000000;
#pragma unroll
          for (int j = 0; j < 32; ++j) mi = max(mi, (int)v[j]);
          mx = (float)mi;
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
        const bool hit = __any_sync(0xffffffffu, mx > P.bound);
        const int cb = (int)(col0 >> 7);
        if (hit && lane == 0 && mine && cb >= bi && cb < P.T)
          atomicOr(P.flags + (long long)bi * P.T + cb, 1u << (quarter * 4 + (int)((col0 >> 5) & 3)));
        if (hit && mine && P.pairs) {  // rare: one warp-aggregated reservation, then every lane writes its own candidates
          const int k = bi * BM + row;
          unsigned m = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float sv = OP == OP_I8 ? (float)(int)v[j] : __uint_as_float(v[j]);
            m |= (unsigned)(sv > P.bound) << j;
          }
          // columns l = col0 + j with k < l < M: the bits [lo, hi)
          const int c0i = (int)col0;
          const int lo = max(0, k + 1 - c0i), hi = min(32, (int)P.M - c0i);
          m = (lo < hi) ? (m >> lo << lo) & (hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u)) : 0u;
          const int mine_n = __popc(m);
          // exclusive prefix over the lanes that hold candidates (usually one or two of them)
          unsigned holders = __ballot_sync(0xffffffffu, m != 0);
          int total = 0, before = 0;
          while (holders) {
            const int src = __ffs(holders) - 1;
            holders &= holders - 1;
            const int cnt = __shfl_sync(0xffffffffu, mine_n, src);
            if (lane == src) before = total;
            total += cnt;
          }
          const int incl = before + mine_n;
          if (total) {
            // slots come out of this warp's current CHUNK of the list; a new chunk costs one global atomic (a single counter cannot
            // take one atomic per flagged cell: 4.3 M of them on the shuffled config C cost 9 ms), the rest of the old chunk is
            // marked empty (k = -1)
            if (total > PAIR_CHUNK / 2) {
              // a dense cell (up to 1024 candidates): its own run of the list, exactly as long as needed
              unsigned long long nb = 0;
              if (lane == 0) nb = atomicAdd(P.npairs, (unsigned long long)total);
              nb = __shfl_sync(0xffffffffu, nb, 0);
              if (nb + (unsigned long long)total <= P.pair_cap) {
                unsigned long long at = nb + (unsigned long long)(incl - mine_n);
                while (m) {
                  const int j = __ffs(m) - 1;
                  m &= m - 1;
                  P.pairs[at++] = make_int2(k, c0i + j);
                }
              }
              return;
            }
            if (chunk_used + total > PAIR_CHUNK) {
              if (chunk_ok)
                for (int i = chunk_used + lane; i < PAIR_CHUNK; i += 32) P.pairs[chunk_base + i] = make_int2(-1, -1);
              unsigned long long nb = 0;
              if (lane == 0) nb = atomicAdd(P.npairs, (unsigned long long)PAIR_CHUNK);
              chunk_base = __shfl_sync(0xffffffffu, nb, 0);
              chunk_ok = chunk_base + PAIR_CHUNK <= P.pair_cap;
              chunk_used = 0;
            }
            if (chunk_ok) {
              unsigned long long at = chunk_base + (unsigned long long)(chunk_used + incl - mine_n);
              while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                P.pairs[at++] = make_int2(k, c0i + j);
              }
            }
            chunk_used += total;
          }
        }
      };
      // two register buffers: the tcgen05.ld of the next cell is in flight while this one is reduced (with candidate pairs to
      // emit the epilogue of a tile no longer hides behind the MMAs of the next by itself when the flags are dense)
      constexpr int NC = BN / 32;
      uint32_t va[32], vb[32];
      tmem_ld32(taddr + (uint32_t)(half * 32), va);
#pragma unroll
      for (int i = 0; i < (NC + 1) / 2; i += 2) {
        const int c0 = half + 2 * i, c1 = c0 + 2, c2 = c0 + 4;  // warp-uniform
        if (c0 < NC) {
          tmem_ld_wait();
          if (c1 < NC) tmem_ld32(taddr + (uint32_t)(c1 * 32), vb);
          reduce_cell(va, c0);
        }
        if (c1 < NC) {
          tmem_ld_wait();
          if (c2 < NC) tmem_ld32(taddr + (uint32_t)(c2 * 32), va);
          reduce_cell(vb, c1);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG2)
          mbar_arrive_even(tempty_bar(as));  // the issuing CTA waits for the epilogues of both CTAs
        else
          mbar_arrive(tempty_bar(as));
      }
    }
    if (chunk_ok)  // the unused rest of this warp's last chunk
      for (int i = chunk_used + lane; i < PAIR_CHUNK; i += 32) P.pairs[chunk_base + i] = make_int2(-1, -1);
  }

  tc_fence_before();
  __syncthreads();
  if (MC) cluster_sync_all();  // the peer may still multicast into / arrive on this CTA's shared memory until it is done too
  if (warp == 1) {
    tc_fence_after();
    if (CG2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// simplex coordinate j of the class of state z: c0=(+,+,+) c1=(+,-,-) c2=(-,+,-) c3=(-,-,+)
__device__ __forceinline__ bool simplex_neg(int z, int j) {
  const int c = z & 3;
  return (c != 0) && (j != c - 1);
}

// FP8 / INT8: V[k][3 i + j] as e4m3 (+1.0 = 0x38, -1.0 = 0xB8) or int8 (+1 = 0x01, -1 = 0xFF), one byte per element; zero padding
__global__ void encode_simplex_kernel(const int8_t *__restrict__ Z, long long L, long long M, long long VM, long long Kpad,
                                      uint32_t *__restrict__ V, uint32_t plus, uint32_t minus) {
  const long long words_per_row = Kpad / 4;
  const long long total = VM * words_per_row;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long k = t / words_per_row;
    const int w = (int)(t - k * words_per_row);
    uint32_t out = 0;
    if (k < M) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int b = 4 * w + e;
        const int site = b / 3, j = b - 3 * site;
        if (site < L) out |= (simplex_neg((int)Z[k * L + site], j) ? minus : plus) << (8 * e);
      }
    }
    V[t] = out;
  }
}

// FP4: the same elements as packed e2m1 (+1.0 = 0x2, -1.0 = 0xA), element 2b in the low nibble of byte b; zero padding
__global__ void encode_simplex4_kernel(const int8_t *__restrict__ Z, long long L, long long M, long long VM, long long Kbytes,
                                       uint32_t *__restrict__ V) {
  const long long words_per_row = Kbytes / 4;
  const long long total = VM * words_per_row;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long k = t / words_per_row;
    const int w = (int)(t - k * words_per_row);
    uint32_t out = 0;
    if (k < M) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int b = 8 * w + e;
        const int site = b / 3, j = b - 3 * site;
        if (site < L) out |= (simplex_neg((int)Z[k * L + site], j) ? 0xAu : 0x2u) << (4 * e);
      }
    }
    V[t] = out;
  }
}

// cell masks [T][T] -> list of blocks (bi <= bj) with their masks for the exact sweep, plus the running number of flagged
// cells in front of every block (cellbase), so the sweep can deal CELLS, not blocks, to its warps.  One 64-bit atomic hands out
// the block slot (low word) and the cell offset (high word) together: cellbase grows with the slot.
__global__ void compact_flags_kernel(const uint32_t *__restrict__ flags, int T, int2 *__restrict__ items,
                                     uint32_t *__restrict__ masks, uint32_t *__restrict__ cellbase,
                                     unsigned long long *__restrict__ n_packed) {
  const long long total = (long long)T * T;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    uint32_t m = flags[t];
    if (m) {
      const int bi = (int)(t / T), bj = (int)(t - (long long)bi * T);
      if (bj >= bi) {
        // diagonal blocks: cell (r, c) with r > c only repeats the pairs of cell (c, r)
        if (bi == bj) m &= 0x8CEFu;  // bits 4r + c with c >= r
        if (!m) continue;
        const unsigned long long old = atomicAdd(n_packed, ((unsigned long long)__popc(m) << 32) | 1ull);
        const uint32_t slot = (uint32_t)old;
        items[slot] = make_int2(bi, bj);
        masks[slot] = m;
        cellbase[slot] = (uint32_t)(old >> 32);
      }
    }
  }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int32_t make_tensor_map(gdca_ctx *ctx, CUtensorMap *map, void *V, long long VM, long long Kbytes, int box_rows) {
  static encode_tiled_fn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    GDCA_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (!p || qres != cudaDriverEntryPointSuccess)
      return gdca_fail(ctx, GDCA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    fn = (encode_tiled_fn)p;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)Kbytes, (cuuint64_t)VM};
  const cuuint64_t gstride[1] = {(cuuint64_t)Kbytes};  // bytes between rows
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, V, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[96];
    snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    ctx->err = b;
    return GDCA_ERR_CUDA;
  }
  return GDCA_OK;
}

template <int OP>
int32_t run_filter(gdca_ctx *ctx, int thresh, float *dump, long long dump_ld) {
  using C = Cfg<OP>;
  constexpr bool FP4 = C::FP4;
  constexpr int TAG = FP4 ? 4 : (OP == OP_I8 ? 80 : 8);  // what ctx->dV currently holds
  const long long T = ctx->Mpad / GDCA_TILE;
  const long long NT = (T * BM + C::BN - 1) / C::BN;
  const long long world = ctx->shard_world, rank = ctx->shard_rank;
  const long long my_rows = T > rank ? (T - rank + world - 1) / world : 0;   // row blocks bi = rank + world * i < T
  const long long NB = (my_rows + BAND - 1) / BAND;
  const long long VM = NT * C::BN > T * BM ? NT * C::BN : T * BM;
  const long long Kel = 3 * ctx->L;                                             // elements per row
  const long long Kbytes = (((FP4 ? (Kel + 1) / 2 : Kel) + BK - 1) / BK) * BK;  // bytes per row, whole k-blocks
  GDCA_TRY(gdca_reserve(ctx, ctx->dV, ctx->capV, (size_t)(VM * Kbytes)));
  GDCA_TRY(gdca_reserve(ctx, ctx->dFlags, ctx->capFlags, (size_t)(T * T)));
  GDCA_TRY(gdca_reserve(ctx, ctx->dItems, ctx->capItems, (size_t)(T * (T + 1) / 2)));
  GDCA_TRY(gdca_reserve(ctx, ctx->dItemMask, ctx->capItemMask, (size_t)(T * (T + 1) / 2)));
  GDCA_TRY(gdca_reserve(ctx, ctx->dCellBase, ctx->capCellBase, (size_t)(T * (T + 1) / 2)));

  if (ctx->have_V != TAG) {
    const long long words = VM * Kbytes / 4;
    const int grid = (int)((words + 255) / 256 < (long long)ctx->num_sms * 16 ? (words + 255) / 256 : (long long)ctx->num_sms * 16);
    if (FP4)
      encode_simplex4_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->dZ, ctx->L, ctx->M, VM, Kbytes, reinterpret_cast<uint32_t *>(ctx->dV));
    else
      encode_simplex_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->dZ, ctx->L, ctx->M, VM, Kbytes, reinterpret_cast<uint32_t *>(ctx->dV),
                                                           OP == OP_I8 ? 0x01u : 0x38u, OP == OP_I8 ? 0xFFu : 0xB8u);
    GDCA_LAUNCH_CHECK(ctx);
    ctx->have_V = TAG;
  }
  // 2-CTA clusters with TMA multicast of the B tile unless switched off (gdca_set_tc_filter_multicast) or the grid is odd
  const bool mc = ctx->tc_filter_want_multicast && (ctx->num_sms % 2 == 0);
  const bool cg2 = mc && FP4 && ctx->tc_filter_want_multicast == 2;   // one cta_group::2 MMA per CTA pair
  CUtensorMap mapA, mapB;
  GDCA_TRY(make_tensor_map(ctx, &mapA, ctx->dV, VM, Kbytes, BM));
  GDCA_TRY(make_tensor_map(ctx, &mapB, ctx->dV, VM, Kbytes, mc ? C::BN / 2 : C::BN));

  GDCA_CUDA(ctx, cudaMemsetAsync(ctx->dFlags, 0, (size_t)(T * T) * sizeof(uint32_t), ctx->stream));
  GDCA_CUDA(ctx, cudaMemsetAsync(ctx->dNItems, 0, sizeof(unsigned long long), ctx->stream));

  FilterParams P;
  P.T = (int)T;
  P.NT = (int)NT;
  P.NB = (int)NB;
  P.KB = (int)(Kbytes / BK);
  P.rank = ctx->shard_rank;
  P.world = ctx->shard_world;
  P.bound = (float)(3 * ctx->L - 4 * (long long)thresh);
  P.flags = ctx->dFlags;
  P.dump = dump;
  P.dump_ld = dump_ld;
  P.M = ctx->M;
  P.pairs = nullptr;
  P.npairs = ctx->dNPairs;
  P.pair_cap = 0;
  if (ctx->pair_list) {
    // capacity: 32 candidates per sequence (16 M at least, 256 M at most): far above what a weighted alignment produces; an
    // alignment of near-duplicates overflows it and takes the cell sweep
    unsigned long long cap = 32ull * (unsigned long long)ctx->M;
    if (cap < (1ull << 24)) cap = 1ull << 24;
    if (cap > (1ull << 28)) cap = 1ull << 28;
    GDCA_TRY(gdca_reserve(ctx, ctx->dPairs, ctx->capPairs, (size_t)cap));
    P.pairs = ctx->dPairs;
    P.pair_cap = cap;
  }
  ctx->pair_cap = P.pair_cap;
  GDCA_CUDA(ctx, cudaMemsetAsync(ctx->dNPairs, 0, sizeof(unsigned long long), ctx->stream));
  auto launch = [&](auto kern, int threads, bool cluster) -> int32_t {
    const size_t smem = cg2 ? CG2_SMEM : C::SMEM;
    GDCA_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctx->num_sms);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cluster ? 1 : 0;
    GDCA_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, mapA, mapB, P));
    return GDCA_OK;
  };
  if constexpr (FP4) {
    if (cg2) GDCA_TRY(launch(tc_filter_kernel<OP, 2>, TC_THREADS, true));
  }
  if (mc && !cg2)
    GDCA_TRY(launch(tc_filter_kernel<OP, 1>, TC_THREADS, true));
  else if (!mc)
    GDCA_TRY(launch(tc_filter_kernel<OP, 0>, TC_THREADS, false));
  GDCA_LAUNCH_CHECK(ctx);
  ctx->tc_filter_multicast = mc ? (cg2 ? 2 : 1) : 0;

  const long long tt = T * T;
  const int cgrid = (int)((tt + 255) / 256 < (long long)ctx->num_sms * 8 ? (tt + 255) / 256 : (long long)ctx->num_sms * 8);
  compact_flags_kernel<<<cgrid, 256, 0, ctx->stream>>>(ctx->dFlags, (int)T, ctx->dItems, ctx->dItemMask, ctx->dCellBase, ctx->dNItems);
  GDCA_LAUNCH_CHECK(ctx);
  // tiles this rank visits (host arithmetic, same rule as TileIter::valid)
  long long tiles = 0;
  for (long long bi = rank; bi < T; bi += world) tiles += NT - (BM * bi) / C::BN;
  ctx->tc_filter_tiles = tiles;
  ctx->tc_filter_tflop = 2.0 * (double)BM * C::BN * (double)(Kbytes * (FP4 ? 2 : 1)) * 1e-12 * (double)tiles;
  // TMA operand bytes requested from L2 (multicast: each CTA fetches its A tile and half a B tile; the few tiles computed only
  // because the cluster peer's tile is valid are not counted)
  ctx->tc_filter_l2_bytes = (double)(BM + (mc ? C::BN / 2 : C::BN)) * (double)Kbytes * (double)tiles;
  return GDCA_OK;
}

}  // namespace

int32_t gdca_make_tensor_map_2d(gdca_ctx *ctx, CUtensorMap_st *map, void *base, long long rows, long long row_bytes, int box_rows) {
  return make_tensor_map(ctx, map, base, rows, row_bytes, box_rows);
}

// Host-side replay of the kernel's tile order for one CTA (the same TileIter code, compiled for the host): rows of
// {bi, cj, valid, peer_valid} in visiting order.  Lets CPU tests check coverage, disjointness across ranks and the
// lock-step of cluster pairs without a GPU.
extern "C" int32_t gdca_tc_filter_tile_order(int32_t T, int32_t bits, int32_t rank, int32_t world, int32_t grid, int32_t cta,
                                             int32_t *out, int64_t cap_rows, int64_t *n_rows) {
  if (T < 1 || (bits != 4 && bits != 8) || world < 1 || rank < 0 || rank >= world || grid < 1 || cta < 0 || cta >= grid || !n_rows)
    return GDCA_ERR_INVALID_ARG;
  const int colw = bits == 4 ? Cfg<OP_FP4>::BN : Cfg<OP_FP8>::BN;
  FilterParams P{};
  P.T = T;
  P.NT = (int)(((long long)T * BM + colw - 1) / colw);
  const long long my_rows = T > rank ? ((long long)T - rank + world - 1) / world : 0;
  P.NB = (int)((my_rows + BAND - 1) / BAND);
  P.rank = rank;
  P.world = world;
  TileIter it;
  int64_t n = 0;
  for (it.start(P, colw, cta); !it.done(); it.advance(grid)) {
    if (out && n < cap_rows) {
      out[4 * n + 0] = it.bi();
      out[4 * n + 1] = it.cj();
      out[4 * n + 2] = it.valid() ? 1 : 0;
      out[4 * n + 3] = it.peer_valid() ? 1 : 0;
    }
    ++n;
  }
  *n_rows = n;
  return GDCA_OK;
}

// Flags the 32 x 32 cells of the 128 x 128 blocks (bi <= bj, this rank's share of the tiles) that may contain a pair
// with hamming < thresh and compacts the blocks with a non-empty mask into ctx->dItems / dItemMask / dNItems.
// dump (device, optional): S of every visited tile.
int32_t gdca_k_tc_filter(gdca_ctx *ctx, int thresh, float *dump, long long dump_ld) {
  if (!ctx->have_alignment) return gdca_fail(ctx, GDCA_ERR_STATE, "tc_filter: no alignment loaded");
  switch (ctx->tc_filter_bits) {
    case 8: return run_filter<OP_FP8>(ctx, thresh, dump, dump_ld);
    case 80: return run_filter<OP_I8>(ctx, thresh, dump, dump_ld);
    default: return run_filter<OP_FP4>(ctx, thresh, dump, dump_ld);
  }
}
