/*
 * gdca_b200.h -- C ABI of libgdca_b200.so: the gDCA hot path on NVIDIA B200 (sm_100a).
 *
 * The reference (carlobaldassi/GaussDCA.jl) has no FFI: its only boundary is the Julia call
 * surface gDCA(filename; ...) -> Vector{Tuple{Int,Int,Float64}} (src/GaussDCA.jl:8-47).  This
 * library replaces everything between the encoded alignment Z (src/GaussDCA.jl:24) and the
 * ranking R (src/GaussDCA.jl:44).  The Julia wrapper (julia/GaussDCA.jl) keeps
 * lines :18-23 (argument check, FASTA parse, dedup) on the host and ccall's gdca_run().
 *
 * Conventions
 *  - Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *  - Layouts are Julia's: column-major.  Z is L x M Int8 with ONE SEQUENCE PER COLUMN, i.e. sequence
 *    k occupies bytes [k*L, (k+1)*L).  Symmetric matrices (C, mJ, S) are written full.
 *  - (site i, state a), 1-based, lives at matrix index (i-1)*s + a, s = q-1 (state q is dropped).
 *  - Host buffers are caller-owned; the library never retains or frees them.
 *  - Every entry point returns a gdca_status_t; gdca_last_error() gives the text.
 *  - There is no CPU fallback: without a usable sm_100 device gdca_create() fails.
 *  - One caller thread per context at a time (the reference is single-caller).
 */
#ifndef GDCA_B200_H
#define GDCA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDCA_ABI_VERSION 2

/* Size limits (every violation is reported as GDCA_ERR_INVALID_ARG with a message, never as a CUDA fault):
 *   L <= GDCA_MAX_L = 11616     the bit-plane packer keeps ceil(L/32) x 5 planes of 32 sequences in shared memory
 *   M <  2^31 - 128
 *   M * roundup(L,128) < 2^32   32-bit row offsets of the recoded alignment in the scatter-add covariance engine
 *                               (L=500: M < 8.3e6; L=1500: M < 2.7e6); the tensor-core engine (weights = 1/count) has no such limit
 *                               but needs n * (M/2 + 128 * classes) bytes for its one-hot operand (1.03 GB at L=500, M=200k)
 *   M <= 2 097 152 for the tensor-core prefilter (its T x T block-mask array, T = ceil(M/128) <= 16384); above that the
 *                               neighbour-count sweep still runs, unfiltered
 *   residue codes 1 <= Z <= 31  (q = max(Z) <= 31, src/GaussDCA.jl:25-26; a code < 1 is rejected)
 *   n = (q-1) L: four n x n FP64 buffers must fit in HBM (n = 30 000 -> 29 GB; n ~ 70 000 on one 180 GB B200) */
#define GDCA_MAX_L 11616
/* Numerical conventions that differ from DCAUtils in the last place only (both are documented deviations, not errors):
 *   theta = :auto   meanfracid = (ident_sum / L) / (M (M-1) / 2) from the EXACT integer identity sum; DCAUtils accumulates the
 *                   per-row fractions in floating point.  The two agree to a few ulp; thresh = floor(theta * L) can differ only when
 *                   theta * L lies within that rounding of an integer (none of the reference's golden cases does).
 *   Meff            the correctly rounded value of the exact rational sum of 1/count (double-double), independent of the order of
 *                   summation and of the number of GPUs; Julia's sum(W) is a pairwise float sum of the same W[k]. */

typedef enum {
  GDCA_OK = 0,
  GDCA_ERR_INVALID_ARG = 1, /* ArgumentError-class: bad shape / range (src/GaussDCA.jl:49-65)          */
  GDCA_ERR_Q_TOO_BIG = 2,   /* q >= 32: error("parameter q=$q is too big ...") (src/GaussDCA.jl:26)     */
  GDCA_ERR_NOT_SPD = 3,     /* PosDefException from cholesky(C) (src/GaussDCA.jl:34); info in stats    */
  GDCA_ERR_CUDA = 4,
  GDCA_ERR_OOM = 5,
  GDCA_ERR_NO_DEVICE = 6,
  GDCA_ERR_STATE = 7 /* staged call issued before the stage it depends on                       */
} gdca_status_t;

typedef enum { GDCA_SCORE_FROB = 0, GDCA_SCORE_DI = 1 } gdca_score_t;

/* One ranking row.  Bit-compatible with Julia's isbits Tuple{Int,Int,Float64} (24 bytes, offsets
 * 0/8/16), so a preallocated Vector{Tuple{Int,Int,Float64}} is passed straight through.
 * Replaces the element type built at src/GaussDCA.jl:90-97.  i < j, 1-based. */
typedef struct {
  int64_t i;
  int64_t j;
  double score;
} gdca_rank_t;

/* Per-call diagnostics (the reference prints theta/threshold/Meff from DCAUtils; it has no timers). */
typedef struct {
  int64_t L, M, n;      /* columns, sequences, n = (q-1)*L                                           */
  int32_t q;            /* alphabet size used, = max(Z) (src/GaussDCA.jl:25)                          */
  int32_t posdef_info;  /* 0, or the 1-based order of the leading minor that is not SPD              */
  double theta;         /* theta used (the :auto value when requested)                               */
  int64_t thresh;       /* floor(theta*L); neighbours have hamming < thresh                          */
  double meff;          /* sum_k 1/count[k], correctly rounded from the count histogram              */
  uint64_t ident_sum;   /* sum_{k<l} #identical positions (only when theta == :auto)                 */
  int32_t theta_passes; /* M x M pair sweeps executed (1, or 0 when theta == 0)                     */
  int32_t reserved;
  /* device milliseconds per stage, CUDA events on the library's stream (ms_chol: factorisation, ms_inv: inversion) */
  float ms_h2d, ms_pack, ms_theta, ms_weights, ms_cov, ms_chol, ms_inv, ms_score, ms_apc, ms_rank, ms_d2h, ms_total;
} gdca_stats_t;

typedef struct gdca_ctx gdca_ctx;

/* ---- context ------------------------------------------------------------------------------- */
int32_t gdca_abi_version(void);
/* device: CUDA ordinal.  Owns one stream, all device memory, reusable across calls. */
int32_t gdca_create(gdca_ctx **out, int32_t device);
/* Several GPUs of one node behind ONE context ("create(n_gpus or device list)"): the returned context leads a group
 * of n devices with peer access enabled between all pairs.  gdca_run() / gdca_run_resident() on it copy the alignment to devices[0]
 * once, broadcast it over NVLink, shard the pair sweep (peer atomics), the covariance (peer stores) and the inversion (trtri by
 * column slices with the result stored to every member, lauum by row tiles stored to devices[0]) and return the same bits as a
 * single-GPU run.  Every other entry point acts on devices[0] alone.  n == 1 is gdca_create().  The host wrapper picks the
 * devices (env GDCA_B200_DEVICES in julia/GaussDCA.jl and api.py).  gdca_destroy() releases the whole group. */
int32_t gdca_create_multi(gdca_ctx **out, const int32_t *devices, int32_t n);
int32_t gdca_group_size(const gdca_ctx *ctx);
void gdca_destroy(gdca_ctx *ctx);
const char *gdca_last_error(const gdca_ctx *ctx); /* ctx may be NULL: error of a failed gdca_create */
const char *gdca_status_string(int32_t status);
/* Work partition for one-process-per-GPU drivers: this context computes shard `rank` of `world`
 * in the sharded stages (pair sweep tiles, covariance row blocks).  Default (0,1). */
int32_t gdca_set_shard(gdca_ctx *ctx, int32_t rank, int32_t world);

/* ---- fused hot path: src/GaussDCA.jl:24-44 ---------------------------------------------------
 * theta < 0 means :auto.  score: gdca_score_t.  R_len must equal
 * (L-min_separation)*(L-min_separation+1)/2 (src/GaussDCA.jl:90).  stats may be NULL. */
int32_t gdca_run(gdca_ctx *ctx, const int8_t *Z, int64_t L, int64_t M, double theta, double pseudocount,
                 int32_t score, int64_t min_separation, gdca_rank_t *R, int64_t R_len, gdca_stats_t *stats);
int64_t gdca_ranking_length(int64_t L, int64_t min_separation);
/* Same pipeline with Z already resident in device memory (L*M int8).  The ranking stays on the device
 * (gdca_dev_R_ptr) and is also copied to R_host_or_null when that is not NULL. */
int32_t gdca_run_resident(gdca_ctx *ctx, const int8_t *Z_dev, int64_t L, int64_t M, double theta, double pseudocount,
                          int32_t score, int64_t min_separation, gdca_rank_t *R_host_or_null, int64_t R_len,
                          gdca_stats_t *stats);
void *gdca_dev_R_ptr(gdca_ctx *ctx); /* gdca_rank_t[R_len] on the device after a run / score_rank */

/* ---- staged entry points, HOST buffers (parity tests; DCAUtils-shaped pieces) ------------------ */
/* compute_theta + compute_weights (call site src/GaussDCA.jl:28).  theta < 0 => :auto.
 * counts int32[M], W f64[M] (= 1/count); either may be NULL. */
int32_t gdca_compute_weights(gdca_ctx *ctx, const int8_t *Z, int64_t L, int64_t M, double theta, int32_t *counts,
                             double *W, double *meff, double *theta_used, int64_t *thresh, uint64_t *ident_sum);
/* frequencies + add_pseudocount + compute_C fused (src/GaussDCA.jl:28-32).  q = max(Z) is computed
 * by the library.  C is n x n, Pi (with pseudocount) is n; Pi may be NULL. */
int32_t gdca_compute_covariance(gdca_ctx *ctx, const int8_t *Z, int64_t L, int64_t M, const double *W, double meff,
                                double pseudocount, double *C, double *Pi, int32_t *q_out);
/* The same stage in DCAUtils' own pieces, for hosts that use them one by one:
 * compute_weighted_frequencies(Z, q, theta) -> Pi_true [n], Pij_true [n x n, full symmetric], Meff, W (call site
 * src/GaussDCA.jl:28; W, theta_used, q_out may be NULL); add_pseudocount(Pi_true, Pij_true, pc, q) -> Pi, Pij (:30);
 * compute_C(Pi, Pij) = Pij - Pi Pi' (:32,:76).  The fused gdca_compute_covariance gives the same C in one pass. */
int32_t gdca_compute_weighted_frequencies(gdca_ctx *ctx, const int8_t *Z, int64_t L, int64_t M, double theta, double *Pi_true,
                                          double *Pij_true, double *meff, double *W, double *theta_used, int32_t *q_out);
int32_t gdca_add_pseudocount(gdca_ctx *ctx, const double *Pi_true, const double *Pij_true, int64_t n, int32_t q, double pseudocount,
                             double *Pi, double *Pij);
int32_t gdca_compute_C(gdca_ctx *ctx, const double *Pi, const double *Pij, int64_t n, double *C);
/* mJ = inv(cholesky(C)) (src/GaussDCA.jl:34).  info: 0 or failing leading minor (1-based). */
int32_t gdca_inverse(gdca_ctx *ctx, const double *C, int64_t n, double *mJ, int32_t *info);
/* compute_FN (src/GaussDCA.jl:39) / compute_DI_gauss (:37).  C is only read for DI (may be NULL for FROB).
 * S is L x L, L = n/(q-1), zero diagonal. */
int32_t gdca_score(gdca_ctx *ctx, const double *mJ, const double *C, int64_t n, int32_t q, int32_t score, double *S);
/* correct_APC (src/GaussDCA.jl:78-86) */
int32_t gdca_apc(gdca_ctx *ctx, const double *S, int64_t L, double *S_out);
/* compute_ranking (src/GaussDCA.jl:88-99): stable descending sort of (i, j, S[j,i]). */
int32_t gdca_ranking(gdca_ctx *ctx, const double *S, int64_t L, int64_t min_separation, gdca_rank_t *R,
                     int64_t R_len);

/* ---- staged entry points, DEVICE-resident state ------------------------------------------------
 * For one-process-per-GPU drivers (torch.distributed / NCCL does the exchange on the exposed device
 * buffers) and for timing with inputs already in HBM.  Order: load -> pair_pass(es) ->
 * finish_weights -> covariance -> inverse -> score -> rank.  Nothing here synchronises the host
 * except where a scalar result is returned. */
int32_t gdca_dev_load(gdca_ctx *ctx, const int8_t *Z_host, int64_t L, int64_t M); /* H2D + q + bit-plane pack */
int32_t gdca_dev_load_resident(gdca_ctx *ctx, const int8_t *Z_dev, int64_t L, int64_t M); /* Z already on device */
/* One sweep over this shard's tiles of the M x M pair matrix.
 * mode 0: accumulate sum of hamming distances (theta :auto);  mode 1: neighbour counts for `thresh`;
 * mode 2: both in one sweep, counts for the three thresholds thresh-1, thresh, thresh+1 (speculative). */
int32_t gdca_dev_pair_pass(gdca_ctx *ctx, int32_t mode, int64_t thresh);
/* Tensor-core prefilter of the mode-1 sweep (csrc/tcfilter.cu): 32 x 32 cells of sequence pairs are proved
 * neighbour-free by a tcgen05 contraction of a 4-class projection of the alignment (a lower bound of the hamming
 * distance); only the 128 x 128 blocks with a cell left go through the exact bit-plane sweep, and only the warps of
 * that sweep that own such a cell do any work.  Counts are identical in every mode.
 * mode 0: off; 1: auto (default; on for M >= 16384; env GDCA_TC_FILTER overrides the default); 2: always.
 * bits 4 (default): packed e2m1 operands, kind::mxf4 with unit block scales; 8: e4m3 operands, kind::f8f6f4;
 * 80: signed int8 operands, kind::i8 with S32 accumulators (env GDCA_TC_FILTER_BITS overrides the default). */
int32_t gdca_set_tc_filter(gdca_ctx *ctx, int32_t mode);
int32_t gdca_set_tc_filter_bits(gdca_ctx *ctx, int32_t bits);
/* on (default): the prefilter runs as 2-CTA clusters that share the column tile by TMA multicast; off: independent CTAs
 * (env GDCA_TC_MULTICAST overrides the default).  Same results either way. */
int32_t gdca_set_tc_filter_multicast(gdca_ctx *ctx, int32_t on);
/* Test hook: run only the prefilter for `thresh` on the loaded alignment.  flags_host: [T*T] uint32, T = ceil(M/128);
 * bit 4*(r/32) + (c/32) of entry (bi, bj), bi <= bj, is set iff the 32 x 32 cell at rows r.., columns c.. of that
 * block must be swept.  S_host (optional): the projected score 4*ident_proj - L of every visited tile,
 * [128*T][ld] floats, ld >= 128*T + 256 (parts of the lower triangle that no tile covers are left at 0). */
int32_t gdca_dev_tc_filter(gdca_ctx *ctx, int64_t thresh, uint32_t *flags_host, float *S_host, int64_t ld);
/* Host-only (no GPU): the tile order of CTA `cta` of a `grid`-CTA prefilter launch for T = ceil(M/128) row blocks, as rows
 * {bi, cj, valid, peer_valid} (4 x int32) -- the kernel's own iterator compiled for the host, for CPU tests. */
int32_t gdca_tc_filter_tile_order(int32_t T, int32_t bits, int32_t rank, int32_t world, int32_t grid, int32_t cta,
                                  int32_t *out, int64_t cap_rows, int64_t *n_rows);
/* What the last mode-1 sweep did: *filtered = 0, or the operand code (4 / 8 / 80) of the prefilter that ran; its tiles,
 * flop (1e12) and TMA operand bytes; 128x128 blocks that went through the exact sweep; device ms of the two parts. */
int32_t gdca_dev_sweep_info(gdca_ctx *ctx, int32_t *filtered, int64_t *filter_tiles, double *filter_tflop,
                            int64_t *swept_blocks, float *ms_filter, float *ms_exact, double *filter_l2_bytes);
/* Exact stage behind the prefilter: 1 (default; env GDCA_PAIR_LIST) = the prefilter lists every candidate PAIR (projected distance
 * below the threshold) and the exact stage checks those pairs alone, byte by byte on the alignment itself -- the cost no longer
 * depends on how the sequences are ordered; 0 = sweep the flagged 32 x 32 cells on the bit planes.  A list that overflows its
 * capacity (32 candidates per sequence) is dropped on the device and the cells are swept.  Counts are identical either way.
 * gdca_dev_pair_list_info: candidates found by the last sweep and the capacity of the list (0: no list). */
int32_t gdca_set_pair_list(gdca_ctx *ctx, int32_t on);
int32_t gdca_dev_pair_list_info(gdca_ctx *ctx, int64_t *candidates, int64_t *capacity);
/* how the last prefilter launch ran: 0 independent CTAs, 1 2-CTA clusters with TMA multicast of the column tile, 2 CTA pairs issuing
 * one tcgen05.mma.cta_group::2 of 256 x 224 (FP4 operands; the default) */
int32_t gdca_dev_tc_filter_launch_mode(gdca_ctx *ctx);
/* device ms of the last cov_rows_kernel / cov_tc_kernel launch alone (the covariance stage also runs small preparation kernels) */
int32_t gdca_dev_cov_kernel_ms(gdca_ctx *ctx, float *ms);
/* Engine of the frequency / covariance stage (DCAUtils compute_weighted_frequencies, call site src/GaussDCA.jl:28):
 *   1 = scatter-add (cov.cu; any weights), 2 = exact co-occurrence counts per weight class on the FP4 tensor cores (covtc.cu;
 *   needs W = 1/count, i.e. weights computed by this library), 0 = auto: the tensor cores when a cost model of the class sizes
 *   says they win (default; env GDCA_COV_ENGINE).  With mode 2 the scatter-add engine still runs when the weights were supplied
 *   by the caller (gdca_dev_set_weights / gdca_compute_covariance) or there are more than 512 distinct counts. */
int32_t gdca_set_cov_engine(gdca_ctx *ctx, int32_t mode);
/* what the last covariance ran on: engine (1 / 2), distinct counts found, class segments, 256-sequence k-blocks of the operand,
 * 4-CTA clusters launched, FP4 tensor work in 1e12 flop, TMA operand bytes requested from L2.  Any pointer may be NULL. */
int32_t gdca_dev_cov_info(gdca_ctx *ctx, int32_t *engine, int32_t *classes, int32_t *segments, int64_t *kblocks, int32_t *clusters,
                          double *tflop, double *l2_bytes);
/* mode-0 sweep over every stride-th tile of this shard (cheap estimate of the mean identity). */
int32_t gdca_dev_pair_sample(gdca_ctx *ctx, int32_t stride);
/* partial results of this shard, device pointers: u64[2] {hamming sum, pairs visited} (after a mode-1 sweep
 * slot [1] holds the number of (pair, 32-site word) units really executed, i.e. not skipped by the early exit); int32[3*Mpad] counts (row t = thresh-1+t
 * in mode 2, row 0 only in mode 1), NOT including the self count. */
void *gdca_dev_ham_sum_ptr(gdca_ctx *ctx);
void *gdca_dev_counts_ptr(gdca_ctx *ctx);
int64_t gdca_dev_counts_stride(gdca_ctx *ctx);
/* sum_{k<l} #identical positions of the loaded alignment from per-site state histograms: O(M L), no pair
 * sweep (replaces the O(M^2 L) loop of DCAUtils compute_theta, call site src/GaussDCA.jl:28); exact. */
int32_t gdca_dev_ident_sum(gdca_ctx *ctx, uint64_t *ident_sum);
int32_t gdca_theta_from_ident_sum(int64_t L, int64_t M, uint64_t ident_sum, double *theta, int64_t *thresh);
/* theta from the (all-reduced) hamming sum of a mode-0/2 sweep -- host arithmetic identical to the oracle's. */
int32_t gdca_theta_from_ham_sum(int64_t L, int64_t M, uint64_t ham_sum, double *theta, int64_t *thresh,
                                uint64_t *ident_sum);
/* counts (all-reduced, row `which` of the counts buffer) -> W = 1/(1+count), Meff.  theta == 0 path: which = -1. */
int32_t gdca_dev_finish_weights(gdca_ctx *ctx, int32_t which, double *meff);
int32_t gdca_dev_set_weights(gdca_ctx *ctx, const double *W_host, double meff);
/* this shard's rows of C (others zero); sum over shards == C.  Device pointer to n x n f64. */
int32_t gdca_dev_covariance(gdca_ctx *ctx, double pseudocount);
void *gdca_dev_C_ptr(gdca_ctx *ctx);  /* [npad][npad] f64, leading dimension gdca_dev_npad() */
int64_t gdca_dev_npad(gdca_ctx *ctx);
void *gdca_dev_W_ptr(gdca_ctx *ctx);  /* f64[M] */
int32_t gdca_dev_inverse(gdca_ctx *ctx, int32_t *info); /* C -> mJ on this device */
/* Engine of the big FP64 products of the inversion (trailing updates of the Cholesky, trtri levels, lauum) for n >= 2048:
 * 1 (default; env GDCA_OZAKI overrides): INT8-sliced on the tcgen05 tensor cores -- every operand row is split exactly into eight
 * signed 7-bit digits, the 36 digit products with t + u < 8 accumulate exactly in S32 (tcgen05.mma kind::i8, TMEM) and are recombined
 * in FP64 with one rounding (csrc/ozaki.cu; mJ stays within ~1e-12 normwise of the DMMA path);  0: FP64 tensor cores (DMMA) only. */
int32_t gdca_set_ozaki(gdca_ctx *ctx, int32_t mode);
/* Eigenvalue engine of the DI score (compute_DI_gauss, reference src/GaussDCA.jl:37).  1 (default): V = G'G per site pair, reduced to
 * tridiagonal form by Householder reflections (one warp, 32 pairs in turn) and solved by implicit QL (one lane per pair)
 * (csrc/score.cu di_eig_kernel; env GDCA_DI_ENGINE);  0: one-sided Jacobi on G, three pairs per warp (di_kernel, the round-1 engine;
 * the two agree to ~1e-13 normwise, the Jacobi engine is kept as the cross-check). */
int32_t gdca_set_di_engine(gdca_ctx *ctx, int32_t mode);
/* what the last inversion ran: *ozaki 0/1, the INT8 operations executed and the FP64 flop they stand for */
int32_t gdca_dev_inverse_info(gdca_ctx *ctx, int32_t *ozaki, double *int8_ops, double *fp64_flop_on_int8);
/* 1 if the last factorisation on a device group shared its trailing update by block columns: every member owns the 512-column
 * outer blocks ob = rank (mod n), applies each broadcast panel to them and ships the next panel's columns to the leader one step
 * ahead (automatic for n >= 16 384, where the bulk updates outweigh the serial chain; env GDCA_SHARE_MIN_NB, in 128-blocks). */
int32_t gdca_dev_inverse_shared(gdca_ctx *ctx);
/* Test hook (tests/test_gpu_ozaki.py): C[m x n] = beta C + alpha opA opB^T on one FP64 GEMM engine, host buffers, contiguous
 * matrices.  engine 0: dgemm_kernel (DMMA), 1: INT8-sliced tcgen05 kernel (beta 0 or 1).  A is [m][k] (a_cols = 0) or [k][m]
 * (a_cols = 1), B is [n][k] (b_cols = 0) or [k][n] (b_cols = 1).  flags: 1 skip output tiles above the diagonal, 2 k starts at
 * n0 (B lower triangular as [k][n]), 4 k starts at m0 (A lower triangular as [k][m]), 8 k ends at m0 + 128.  m, n, k % 128 == 0. */
int32_t gdca_test_fp64_gemm(gdca_ctx *ctx, int32_t engine, const double *A, int32_t a_cols, const double *B, int32_t b_cols,
                            double *C, int64_t m, int64_t n, int64_t k, int32_t flags, double alpha, double beta);
void *gdca_dev_mJ_ptr(gdca_ctx *ctx);
int32_t gdca_dev_score_rank(gdca_ctx *ctx, int32_t score, int64_t min_separation, gdca_rank_t *R_host, int64_t R_len);
void *gdca_dev_S_ptr(gdca_ctx *ctx); /* L x L APC-corrected scores after gdca_dev_score_rank */
/* Peer memory for one-process-per-GPU hosts: the exchange steps fused into the kernels, no collective library on
 * the data path.  Every rank exports CUDA-IPC handles of its counts and C buffers (128 bytes: counts | C); the host
 * all-gathers them (any transport) and imports the table.  From then on (a) the pair sweep adds its neighbour hits
 * into EVERY rank's counters with peer atomics over NVLink (fused all-reduce) and (b) the covariance kernel stores
 * its rows straight into rank 0's C (fused reduce; rows are dealt by site, writes are disjoint).  The host only
 * barriers: zero_counts -> barrier -> pair_pass -> sync -> barrier; rank 0 zero_C -> barrier -> covariance -> sync -> barrier. */
int32_t gdca_dev_peer_export(gdca_ctx *ctx, uint8_t *handles128);
int32_t gdca_dev_peer_import(gdca_ctx *ctx, int32_t world, const uint8_t *handles /* world x 128 bytes */);
int32_t gdca_dev_peer_valid(gdca_ctx *ctx); /* 1: imported mappings still match the live buffers */
int32_t gdca_dev_peer_close(gdca_ctx *ctx);
int32_t gdca_dev_zero_counts(gdca_ctx *ctx);
int32_t gdca_dev_zero_C(gdca_ctx *ctx);
int32_t gdca_dev_sync(gdca_ctx *ctx);
/* stream-ordered D2H copy of any exposed device buffer, then sync (hosts without a CUDA binding) */
int32_t gdca_dev_copy_to_host(gdca_ctx *ctx, void *dst_host, const void *src_dev, int64_t nbytes);
int32_t gdca_dev_get_stats(gdca_ctx *ctx, gdca_stats_t *stats);
void *gdca_dev_stream(gdca_ctx *ctx); /* cudaStream_t the library launches on (for CUDA-event timing) */
int64_t gdca_dev_kernel_launches(gdca_ctx *ctx); /* kernels launched by this context so far */

/* ---- synthetic alignment of SURVEY 8(d) (bench / tests): fills Z_dev (L*M int8, device) ---------- */
int32_t gdca_synth_alignment_dev(gdca_ctx *ctx, int8_t *Z_dev, int64_t L, int64_t M, uint64_t seed);
int32_t gdca_synth_alignment(gdca_ctx *ctx, int8_t *Z_host, int64_t L, int64_t M, uint64_t seed);

/* ---- host front-end (no GPU needed): the I/O of src/GaussDCA.jl:20-23 and printrank (:67-74) ----------------
 * Errors of these four are reported through gdca_host_last_error(). */
/* DCAUtils read_fasta_alignment (call site src/GaussDCA.jl:20): plain or gzipped FASTA -> Z (L x M Int8, one sequence
 * per column).  The buffer is malloc'ed by the library; release it with gdca_free_host. */
int32_t gdca_read_fasta_alignment(const char *path, double max_gap_fraction, int8_t **Z_out, int64_t *L_out,
                                  int64_t *M_out);
/* DCAUtils remove_duplicate_sequences (call site src/GaussDCA.jl:21-23): first occurrence of every distinct sequence,
 * order kept.  Z_out needs room for L*M bytes and may alias Z; kept (optional, int64[M]) receives the kept indices. */
int32_t gdca_remove_duplicate_sequences(const int8_t *Z, int64_t L, int64_t M, int8_t *Z_out, int64_t *M_out,
                                        int64_t *kept);
/* printrank(outfile, R) / printrank(io, R): "%i %i %e\n" per row (src/GaussDCA.jl:67-74).  gdca_format_rank writes
 * into buf (<= 64 bytes per row) and reports the bytes used. */
int32_t gdca_write_rank(const char *path, const gdca_rank_t *R, int64_t n);
int32_t gdca_format_rank(const gdca_rank_t *R, int64_t n, char *buf, int64_t cap, int64_t *used);
void gdca_free_host(void *p);
const char *gdca_host_last_error(void);

/* ---- measured-peak helpers (bench only): raw pipe throughput probes --------------------------- */
int32_t gdca_probe_peaks(gdca_ctx *ctx, double *lop3_tops, double *popc_tops, double *dmma_tflops, double *dfma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* GDCA_B200_H */
