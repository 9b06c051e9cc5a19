// gdca_internal.cuh -- shared declarations of the gDCA B200 library (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/gdca_b200.h"

#define GDCA_NUM_SMS_DEFAULT 148
#define GDCA_TILE 128          // sequences per pair-sweep tile edge
#define GDCA_MAX_PLANES 5      // q <= 31  ->  5 bit planes
#define GDCA_NB 128            // Cholesky / GEMM block
#define GDCA_MAX_PEERS 16
#define GDCA_STAGE_SLOTS 4       // pinned ring of the pageable host-to-device copy (stage_copy.cpp)
#define GDCA_COV_MAXCLS 512      // weight classes / class segments the tensor-core covariance handles
#define GDCA_EV_POTRF 15        // ctx->ev[15]: recorded by chol.cu between the factorisation and the inversion

struct gdca_ctx {
  int device = 0;
  int num_sms = GDCA_NUM_SMS_DEFAULT;
  cudaStream_t stream = nullptr;   // main stream (created with the highest priority: it carries every critical path)
  cudaStream_t stream2 = nullptr;  // helper stream: bulk trailing updates of the Cholesky (look-ahead)
  cudaStream_t stream3 = nullptr;  // panel stream of the Cholesky: rest of the panel + inner updates, beside the diagonal chain
  cudaEvent_t ev_fact = nullptr, ev_trail = nullptr, ev_trail_a = nullptr;  // look-ahead hand-shakes (ev_trail_a: first part of a split bulk update)
  cudaEvent_t ev_diag = nullptr, ev_p1 = nullptr, ev_u2a = nullptr, ev_u2b = nullptr;  // inner look-ahead hand-shakes
  int inv_graph_mode = 1;          // single GPU: the inversion's launch sequence is captured once per shape and replayed as CUDA graphs (env GDCA_INV_GRAPH=0: direct launches)
  void *inv_graph = nullptr;       // chol.cu: the instantiated graphs and their key
  int diag_blocked = 1;            // env GDCA_DIAG_BLOCKED=0: the rank-1 diagonal-block kernel of round 1 (256 CTA barriers per block)
  int chol_inner_lookahead = 1;    // env GDCA_CHOL_LOOKAHEAD=0: serial inner steps (round-1 first version)
  std::string err;
  int32_t shard_rank = 0, shard_world = 1;
  int64_t launches = 0;

  // ---- problem state ----
  int64_t L = 0, M = 0, Mpad = 0, n = 0, npad = 0;  // npad: n rounded up to the Cholesky block (ld of dC/dX/dT/dmJ)
  int32_t q = 0, s = 0, nplanes = 0;
  int64_t nwords = 0;  // ceil(L/32)

  // ---- device buffers (grown on demand, never shrunk) ----
  int8_t *dZ = nullptr;  size_t capZ = 0;      // [M][L] sequence-major (as uploaded)
  bool dZ_borrowed = false;                    // dZ points at caller memory (load_resident)
  int8_t *dZt = nullptr; size_t capZt = 0;     // [L][M] site-major copy (list building)
  uint8_t *dZq = nullptr; size_t capZq = 0;    // [M][roundup(L,256)] recoded + permuted copy (covariance)
  uint32_t *dPlanes = nullptr; size_t capPlanes = 0;  // [nwords][nplanes][Mpad]
  int32_t *dPerm = nullptr; size_t capPerm = 0;       // [L] packed position -> site (most variable sites first)
  int32_t *dCounts = nullptr; size_t capCounts = 0;   // [3][Mpad]
  unsigned long long *dHam = nullptr;          // [2] sum of hamming distances, pairs visited
  int *dQ = nullptr;                           // [1] max(Z)
  double *dW = nullptr; size_t capW = 0;       // [M]
  double *dMeff = nullptr;                     // [2] double-double
  int32_t *dList = nullptr; size_t capList = 0;    // [L][M] sequence ids grouped by state at site
  int32_t *dListOff = nullptr; size_t capListOff = 0;  // [L][32+1]
  double *dPi = nullptr; size_t capPi = 0;     // [n] (with pseudocount)
  double *dC = nullptr; size_t capC = 0;       // [npad][npad], padded with identity
  double *dX = nullptr; size_t capX = 0;       // [npad][npad] L^-1
  double *dmJ = nullptr; size_t capmJ = 0;     // [npad][npad]
  double *dCdiag = nullptr; size_t capCdiag = 0;  // [L][s][s] diagonal blocks of C (for DI)
  double *dT = nullptr; size_t capT = 0;       // [npad][npad] GEMM workspace (trtri)
  int *dInfo = nullptr;                        // [1] not-SPD info
  // ---- INT8-sliced FP64 GEMMs of the inversion (ozaki.cu) ----
  int di_engine = 1;                           // 1 (default): DI eigenvalues by tridiagonalisation + implicit QL, one lane per block; 0: one-sided Jacobi (env GDCA_DI_ENGINE)
  int ozaki_mode = 1;                          // 1 (default): big products of potrf / trtri / lauum on tcgen05 kind::i8; 0: DMMA only (env GDCA_OZAKI)
  int ozaki_tpc = 4;                           // tiles per CTA of the bulk trailing update (short CTAs: the look-ahead chain keeps getting SMs)
  int8_t *dDigA = nullptr; size_t capDigA = 0; // digit matrices [rows][k/64][8][64] int8
  int8_t *dDigB = nullptr; size_t capDigB = 0;
  double *dScaleA = nullptr; size_t capScaleA = 0;  // 2^e per operand row
  double *dScaleB = nullptr; size_t capScaleB = 0;
  unsigned long long *dOzMax = nullptr; size_t capOzMax = 0;  // column maxima of a transposed slice
  cudaEvent_t ev_sliced = nullptr;             // panel digits ready (potrf trailing update on two streams)
  cudaEvent_t ev_p1b = nullptr;                // next panel's columns updated (all but the diagonal tile, which the chain does itself)
  int8_t *dDigP = nullptr; size_t capDigP = 0; // digits of the 128 panel rows the chain needs first
  double *dScaleP = nullptr; size_t capScaleP = 0;
  double oz_int8_ops = 0.0;                    // INT8 operations of the last inversion
  double oz_fp64_flop = 0.0;                   // FP64 flop those products stand for
  bool last_inverse_ozaki = false;
  bool last_inverse_shared = false;            // the trailing update of the last factorisation was shared by the device group
  bool oz_attr_set = false;                    // dynamic shared memory size of ozaki_gemm_kernel raised on this device
  double *dS = nullptr; size_t capS = 0;       // [L][L] raw score
  double *dS2 = nullptr; size_t capS2 = 0;     // [L][L] APC-corrected
  double *dRed = nullptr; size_t capRed = 0;   // reductions for APC
  unsigned long long *dKeys = nullptr; size_t capKeys = 0;
  uint32_t *dVals = nullptr; size_t capVals = 0;
  gdca_rank_t *dR = nullptr; size_t capR = 0;

  // ---- tensor-core prefilter of the neighbour-count sweep (tcfilter.cu) ----
  uint8_t *dV = nullptr; size_t capV = 0;          // [rows][k-blocks*128 B] simplex code of the state classes, e4m3 or packed e2m1
  uint32_t *dFlags = nullptr; size_t capFlags = 0; // [T][T] 16-bit masks of the 32x32 cells the filter could not clear
  int2 *dItems = nullptr; size_t capItems = 0;     // compacted (bi, bj) list of blocks with a non-empty mask
  uint32_t *dItemMask = nullptr; size_t capItemMask = 0;  // their masks
  unsigned long long *dNItems = nullptr;           // [1] low word: its length; high word: flagged 32 x 32 cells in it
  uint32_t *dCellBase = nullptr; size_t capCellBase = 0;  // flagged cells in front of every listed block
  int have_V = 0;                                  // 0: dV stale; 8 / 4 / 80: dV holds the FP8 / FP4 / INT8 encoding of the loaded alignment
  int pair_list = 1;                               // the prefilter lists the candidate PAIRS (projected distance below thresh) and the exact stage checks those alone (env GDCA_PAIR_LIST=0: sweep the flagged cells)
  int2 *dPairs = nullptr; size_t capPairs = 0;     // candidate pairs (k < l)
  unsigned long long *dNPairs = nullptr;           // [1] candidates found (keeps counting past the capacity)
  unsigned long long pair_cap = 0;                 // capacity of dPairs in the last filter launch (0: no list)
  int cell_sweep = 1;                              // behind the prefilter: sweep flagged CELLS, one warp each (env GDCA_CELL_SWEEP=0: blocks)
  int tc_filter_mode = 1;                          // 0 off, 1 auto (large M), 2 always (tests)
  int tc_filter_bits = 4;                          // operands of the filter: 4 = e2m1 (kind::mxf4), 8 = e4m3 (kind::f8f6f4), 80 = int8 (kind::i8)
  int tc_filter_want_multicast = 2;                // 0: independent CTAs; 1: 2-CTA clusters + TMA multicast of the B tile; 2: 2-CTA pairs issuing ONE cta_group::2 MMA (FP4; env GDCA_TC_MULTICAST)
  int tc_filter_multicast = 0;                     // what the last filter launch used
  bool last_sweep_filtered = false;
  double tc_filter_tflop = 0.0;                    // flop of the last filter launch on this rank, in 1e12
  double tc_filter_l2_bytes = 0.0;                 // operand bytes its TMA loads moved
  long long tc_filter_tiles = 0;
  cudaEvent_t ev_sweep0 = nullptr, ev_filter = nullptr, ev_sweep1 = nullptr;  // filter / exact sweep split
  cudaEvent_t ev_cov0 = nullptr, ev_cov1 = nullptr;                            // around cov_rows_kernel / cov_tc_kernel alone

  // ---- covariance on the tensor cores: co-occurrence counts per weight class (covtc.cu) ----
  void *hostTab = nullptr; size_t capHostTab = 0;   // pinned host staging of the small planning tables
  int cov_max_clusters = 0;                        // cudaOccupancyMaxActiveClusters of cov_tc_kernel on this device
  unsigned int *dCovSync = nullptr;                // [1] round counter of the covariance producers
  int cov_round_sync = -1;                         // -1 auto (operand > 0.5 GB), 0 off, 1 on: the TMA producers of the covariance pairs start their tiles together (env GDCA_COV_ROUND_SYNC)
  int cov_engine = 0;                              // 0 auto (cost model), 1 scatter-add engine (cov.cu), 2 tensor cores whenever the weights are count classes (env GDCA_COV_ENGINE)
  int32_t *dClsHist = nullptr; size_t capClsHist = 0;  // [M] histogram of the counts + compacted (value, size) pairs
  int32_t *dClsPerm = nullptr; size_t capClsPerm = 0;  // [Mk] sequences in class order (-1: padding) + cursors + class values + segment ends
  double *dClsTab = nullptr; size_t capClsTab = 0;     // class bases (int64) and segment weights
  int32_t *dSegCnt = nullptr; size_t capSegCnt = 0;    // [L][segments][32] sequences of a class segment with a state at a site (Pi of the tensor-core engine)
  uint8_t *dXt = nullptr; size_t capXt = 0;        // [n][Mk/2] one-hot operand, packed e2m1, K-major
  int2 *dCovTiles = nullptr; size_t capCovTiles = 0;   // super-tiles of this rank
  bool weights_from_counts = false;                // W = 1/(count+1) of dCounts[counts_row] (or all 1): the classes are the distinct counts
  bool cov_full = false;                           // the last covariance wrote both triangles (nothing to mirror)
  int last_cov_engine = 0;                         // 1 scatter-add, 2 tensor cores
  int cov_tc_classes = 0, cov_tc_segments = 0, cov_tc_clusters = 0;
  long long cov_tc_kblocks = 0;
  double cov_tc_tflop = 0.0, cov_tc_l2_bytes = 0.0;
  int32_t *cov_cls_host = nullptr;                 // device groups: the leader's compacted classes, so the members plan without a device round trip

  // ---- peer memory (one process per GPU): IPC-mapped counts / C buffers of the other ranks ----
  bool peers_ready = false;
  int32_t *peer_counts[GDCA_MAX_PEERS] = {};   // [r] -> rank r's dCounts (own rank: dCounts)
  double *peer_C[GDCA_MAX_PEERS] = {};         // [r] -> rank r's dC
  void *peer_opened[2 * GDCA_MAX_PEERS] = {};  // mappings to close
  int32_t *exported_counts = nullptr;          // buffers the exported handles refer to (re-export if they moved)
  double *exported_C = nullptr;

  // ---- device group (ONE process, several GPUs): gdca_run() on the leader drives every member ----
  int group_size = 1, group_rank = 0;
  gdca_ctx *group[GDCA_MAX_PEERS] = {};        // leader only: [0] is the leader itself
  gdca_ctx *leader = nullptr;                  // members: their leader
  cudaEvent_t ev_group = nullptr;              // this member's arrival at a group barrier / "my copy of the data is complete"
  cudaStream_t stream_copy = nullptr;          // peer copies that run beside the compute stream (factor panels)
  cudaEvent_t ev_copy = nullptr;
  cudaEvent_t ev_sent = nullptr;               // this member's block columns have arrived at the leader (shared factorisation)
  cudaEvent_t ev_upd = nullptr;                // this member's compute stream has finished the columns it is about to send
  int share_min_nb = 64;                       // the trailing update of the factorisation is shared by the group from this many 128-blocks on (env GDCA_SHARE_MIN_NB)

  // ---- pageable host memory: pipelined H2D through a pinned ring (stage_copy.cpp) ----
  void *stage_buf[GDCA_STAGE_SLOTS] = {};
  cudaEvent_t stage_ev[GDCA_STAGE_SLOTS] = {};
  int staged_h2d = 1;                          // env GDCA_STAGED_H2D=0: plain cudaMemcpyAsync from pageable memory

  // ---- state flags ----
  bool have_hist = false;      // dZt + dListOff (per-site state histograms as bucket offsets) are valid
  bool have_alignment = false, have_lists = false, have_weights = false, have_cov = false, have_inv = false;
  double meff = 0.0, pseudocount = 0.0;
  int counts_row = -1;  // row of dCounts the weights came from (-1: theta == 0)
  gdca_stats_t stats{};
  cudaEvent_t ev[16] = {};
};

#define GDCA_CUDA(ctx, expr)                                                                      \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      char _b[512];                                                                               \
      snprintf(_b, sizeof _b, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      (ctx)->err = _b;                                                                            \
      return (_e == cudaErrorMemoryAllocation) ? GDCA_ERR_OOM : GDCA_ERR_CUDA;                    \
    }                                                                                             \
  } while (0)

#define GDCA_TRY(expr)                 \
  do {                                 \
    int32_t _s = (expr);               \
    if (_s != GDCA_OK) return _s;      \
  } while (0)

#define GDCA_LAUNCH_CHECK(ctx)                    \
  do {                                            \
    (ctx)->launches++;                            \
    GDCA_CUDA(ctx, cudaGetLastError());           \
  } while (0)

template <typename T>
static inline int32_t gdca_reserve(gdca_ctx *ctx, T *&ptr, size_t &cap, size_t count) {
  if (count <= cap && ptr) return GDCA_OK;
  if (ptr) {
    GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    GDCA_CUDA(ctx, cudaFree(ptr));
    ptr = nullptr;
    cap = 0;
  }
  GDCA_CUDA(ctx, cudaMalloc((void **)&ptr, count * sizeof(T)));
  cap = count;
  return GDCA_OK;
}

// Kernel launch that carries the PRIORITY of its stream as a launch attribute: a stream-captured graph node keeps it (plain <<<>>>
// launches lose the stream priority in a captured graph, and the chain of the factorisation then queues behind the bulk update).
template <typename... KArgs, typename... Args>
static inline cudaError_t gdca_launch_prio(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  int prio = 0;
  cudaStreamGetPriority(st, &prio);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributePriority;
  at[0].val.priority = prio;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

static inline int32_t gdca_fail(gdca_ctx *ctx, int32_t status, const char *msg) {
  if (ctx) ctx->err = msg;
  return status;
}

// ---- INT8-sliced FP64 GEMM (ozaki.cu) ----
enum : int { GDCA_OZ_LOWER_OUT = 1, GDCA_OZ_KBEG_N = 2, GDCA_OZ_KBEG_M = 4, GDCA_OZ_KEND_M = 8 };  // = the G_* flags of chol.cu
struct gdca_oz_operand {
  const int8_t *dig;      // [rows_total][pitch] bytes
  const double *scale;    // [rows_total] 2^e
  long long pitch;        // bytes per row = (k / 64) * 512
  long long rows_total;   // rows of all batch members
  long long rows_b;       // rows between consecutive batch members
};
int32_t gdca_oz_slice(gdca_ctx *ctx, cudaStream_t stream, const double *src, long long ld, long long stride_b, bool cols, int rows,
                      int k, int batch, long long rows_b, int8_t *dig, double *scale, gdca_oz_operand *out, bool lower_only = false);
struct gdca_oz_shard {   // one member's share of a product in a device group (nullptr: the whole product, one output buffer)
  int n_off;             // first column of this share within the full product
  int m_off;             // first row of this share within the full product (lower-triangular output test)
  int own_mod, own_rank; // row tiles im with im % own_mod == own_rank (own_mod <= 1: all)
  int col_mod, col_rank, col_unit0, col_per;  // 64-column tiles jn with ((col_unit0 + jn) / col_per) % col_mod == col_rank (col_mod <= 1: all)
  int npeer;             // output tile stored to npeer buffers at C + peer_off[p] BYTES (0 / 1: C only)
  long long peer_off[GDCA_MAX_PEERS];
};
int32_t gdca_oz_gemm(gdca_ctx *ctx, cudaStream_t stream, const gdca_oz_operand &A, const gdca_oz_operand &B, double *C, long long ldc,
                     long long strideC, int m, int n, int k, int batch, int flags, double alpha, int beta, int tiles_per_cta,
                     const gdca_oz_shard *sh = nullptr);

int32_t gdca_h2d(gdca_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t stream);  // stage_copy.cpp
void gdca_h2d_release(gdca_ctx *ctx);

// ---- stage entry points implemented across the .cu files ----
int32_t gdca_k_maxq(gdca_ctx *ctx);                       // pack.cu: dQ <- max(Z)
int32_t gdca_k_pack(gdca_ctx *ctx);                       // pack.cu: dZ -> dPlanes
int32_t gdca_k_pair_pass(gdca_ctx *ctx, int mode, int thresh, int sample_stride);   // pairs.cu
int32_t gdca_k_tc_filter(gdca_ctx *ctx, int thresh, float *dump, long long dump_ld);  // tcfilter.cu
static inline bool gdca_tc_filter_wanted(const gdca_ctx *ctx) {
  // auto: the filter pays once the sweep is more than a few tiles per SM; its T x T mask array and work list (T = Mpad/128)
  // stay below ~3 GB up to M = 2M sequences -- beyond that the plain sweep over all blocks runs
  const bool fits = ctx->Mpad / GDCA_TILE <= 16384;
  return fits && (ctx->tc_filter_mode == 2 || (ctx->tc_filter_mode == 1 && ctx->M >= 16384));
}
int32_t gdca_k_finish_weights(gdca_ctx *ctx, int which);  // weights.cu
int32_t gdca_k_site_hist(gdca_ctx *ctx);                  // cov.cu: site-major copy + per-site state histograms (bucket offsets)
int32_t gdca_k_build_lists(gdca_ctx *ctx);                // cov.cu: per-site lists of sequence ids grouped by state (on demand)
int32_t gdca_k_ident_sum(gdca_ctx *ctx, unsigned long long *ident_out);  // cov.cu: sum_{k<l} ident from site histograms
int32_t gdca_k_covariance(gdca_ctx *ctx, double pc, bool raw = false);  // cov.cu (raw: Pij_true instead of C)
int32_t gdca_k_covariance_tc(gdca_ctx *ctx, double pc, bool raw, bool *done);  // covtc.cu: tensor-core engine; *done = false: not applicable
int32_t gdca_k_cov_classes(gdca_ctx *ctx, int32_t *host_out /*[1 + 2 * GDCA_COV_MAXCLS]*/);  // covtc.cu: distinct counts and their sizes
struct CUtensorMap_st;
int32_t gdca_make_tensor_map_2d(gdca_ctx *ctx, CUtensorMap_st *map, void *base, long long rows, long long row_bytes, int box_rows);  // tcfilter.cu: 128-byte boxes, SWIZZLE_128B
int32_t gdca_k_add_pseudocount(gdca_ctx *ctx, const double *Pi_true, const double *Pij_true, long long n, int q, double pc,
                               double *Pi, double *Pij);  // cov.cu, contiguous device buffers
int32_t gdca_k_compute_C(gdca_ctx *ctx, const double *Pi, const double *Pij, long long n, double *C);  // cov.cu
int32_t gdca_k_symmetrize_C(gdca_ctx *ctx);               // cov.cu: mirror upper site blocks, save diag blocks
int32_t gdca_k_extract_diag(gdca_ctx *ctx);               // cov.cu: save the s x s diagonal blocks of dC
int32_t gdca_k_inverse(gdca_ctx *ctx);                    // chol.cu
void gdca_k_inverse_release(gdca_ctx *ctx);               // chol.cu: drops the captured graphs of the inversion
int32_t gdca_k_inverse_group(gdca_ctx *lead);             // chol.cu: potrf on the leader, trtri / lauum shared by the device group
int32_t gdca_k_score(gdca_ctx *ctx, int score);           // score.cu
int32_t gdca_k_apc(gdca_ctx *ctx);                        // rank.cu
int32_t gdca_k_rank(gdca_ctx *ctx, int64_t min_sep, int64_t R_len);  // rank.cu
int32_t gdca_k_synth(gdca_ctx *ctx, int8_t *Zdev, int64_t L, int64_t M, uint64_t seed);  // synth.cu
int32_t gdca_k_probe(gdca_ctx *ctx, double *lop3, double *popc, double *dmma, double *dfma);  // probe.cu
