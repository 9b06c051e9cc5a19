// chol.cu -- K5: mJ = inv(cholesky(C))  (reference src/GaussDCA.jl:34: LAPACK dpotrf + dpotri + symmetrise).
//
// Hand-written blocked factorisation and inversion, FP64 throughout, all matrix products on the
// FP64 tensor pipe (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4; tcgen05 has no FP64 kind):
//
//   potrf  right-looking, NB = 128:   A[k,k] -> L[k,k] and inv(L[k,k])        (one CTA, registers)
//                                     L[I,k] = A[I,k] * inv(L[k,k])'          (GEMM, replaces TRSM)
//                                     A[I,J] -= L[I,k] * L[J,k]'   (I>=J>k)   (GEMM, lower tiles only)
//   trtri  X = L^-1 by recursive doubling: X21 = -X22 * (L21 * X11), batched over the diagonal;
//          every level is two GEMMs with full-chip parallelism (no sequential column sweep).
//   lauum  mJ = X' X, lower tiles only, k >= i (X is lower triangular), then mirrored.
//
// The matrix is padded to a multiple of 128 with an identity block, so no kernel here has edge
// cases: inv([[C,0],[0,I]]) = [[inv(C),0],[0,I]].  All matrices are row-major with ld = npad.
// A non-positive pivot records the 1-based order of the failing leading minor (PosDefException.info).
#include <algorithm>
#include <string.h>

#include "gdca_internal.cuh"

namespace {

constexpr int NB = GDCA_NB;  // 128
constexpr int BK = 16;
constexpr int GSTAGES = 4;      // 4-slot ring: stage kt+1 is already visible while kt is computed (fragment prefetch)
constexpr int LDS_N = BK + 4;    // [row][k] tile stride (doubles): conflict-free 64-bit fragment loads
constexpr int LDS_T = NB + 8;    // [k][row] tile stride: the transposed orientations (AT / BT) keep 2-way conflicts on their fragment
                                 // loads (ncu r1: 24 % / 49 % of the wavefronts of <0,1> / <1,1>); since round 2 those instantiations
                                 // only serve n < 2048 and the small trtri levels -- the big transposed products run in ozaki.cu
constexpr int TILE_D = NB * LDS_N;  // 2560 doubles >= BK*LDS_T = 2176
constexpr int GTHREADS = 256;

enum : int {
  G_LOWER_OUT = 1,  // skip output tiles strictly above the block diagonal
  G_KBEG_N = 2,     // k starts at n0      (B operand lower triangular as [k][n])
  G_KBEG_M = 4,     // k starts at m0      (A operand lower triangular as [k][m])
  G_KEND_M = 8      // k ends at m0 + NB   (A operand lower triangular as [m][k])
};

struct GemmP {
  const double *A, *B;
  double *C;
  long long lda, ldb, ldc;
  long long strideA, strideB, strideC;  // per blockIdx.z
  int m, n, k;
  int flags;
  double alpha, beta;
  const int *info;  // not-SPD flag of the factorisation (or nullptr): once it is set every later launch of the chain returns at once
};

__device__ __forceinline__ void cp16(void *smem, const void *gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// C[m x n] = beta*C + alpha * opA * opB,  opA(m,k) = AT ? A[k][m] : A[m][k],  opB(k,n) = BT ? B[k][n] : B[n][k]
template <bool AT, bool BT>
__global__ void __launch_bounds__(GTHREADS, 1) dgemm_kernel(GemmP p) {
  extern __shared__ __align__(16) double sm[];
  // longest-K tiles first: with G_KEND_M the K range grows with m0, so walk the block rows from the bottom
  const int by = (p.flags & G_KEND_M) ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int m0 = by * NB, n0 = blockIdx.x * NB;
  if ((p.flags & G_LOWER_OUT) && n0 > m0) return;
  int kbeg = 0, kend = p.k;
  if (p.flags & G_KBEG_N) kbeg = max(kbeg, n0);
  if (p.flags & G_KBEG_M) kbeg = max(kbeg, m0);
  if (p.flags & G_KEND_M) kend = min(kend, m0 + NB);
  const int nk = (kend - kbeg) / BK;

  const double *A = p.A + blockIdx.z * p.strideA;
  const double *B = p.B + blockIdx.z * p.strideB;
  double *C = p.C + blockIdx.z * p.strideC;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, c = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;  // warp tile: rows wm*64.., cols wn*32..

  auto load_stage = [&](int kt) {
    if (kt < nk) {
      const int k0 = kbeg + kt * BK;
      double *As = sm + (size_t)(kt % GSTAGES) * 2 * TILE_D;
      double *Bs = As + TILE_D;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int ch = tid + it * GTHREADS;  // 0..1023 chunks of 2 doubles
        if (!AT) {
          const int row = ch >> 3, cc = ch & 7;
          cp16(As + row * LDS_N + cc * 2, A + (long long)(m0 + row) * p.lda + k0 + cc * 2);
        } else {
          const int row = ch >> 6, cc = ch & 63;
          cp16(As + row * LDS_T + cc * 2, A + (long long)(k0 + row) * p.lda + m0 + cc * 2);
        }
        if (!BT) {
          const int row = ch >> 3, cc = ch & 7;
          cp16(Bs + row * LDS_N + cc * 2, B + (long long)(n0 + row) * p.ldb + k0 + cc * 2);
        } else {
          const int row = ch >> 6, cc = ch & 63;
          cp16(Bs + row * LDS_T + cc * 2, B + (long long)(k0 + row) * p.ldb + n0 + cc * 2);
        }
      }
    }
    cp_commit();
  };

  double acc[8][4][2];
#pragma unroll
  for (int mi = 0; mi < 8; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

  auto load_frags = [&](const double *As, const double *Bs, int kk, double (&a)[8], double (&b)[4]) {
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
      a[mi] = AT ? As[(kk * 4 + c) * LDS_T + wm * 64 + mi * 8 + g] : As[(wm * 64 + mi * 8 + g) * LDS_N + kk * 4 + c];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
      b[ni] = BT ? Bs[(kk * 4 + c) * LDS_T + wn * 32 + ni * 8 + g] : Bs[(wn * 32 + ni * 8 + g) * LDS_N + kk * 4 + c];
  };

  // Ring invariant at the top of iteration kt: stages <= kt+1 have landed and are visible to every thread, so the
  // first fragments of stage kt+1 can be fetched while the last DMMAs of stage kt issue -- the barrier and the
  // shared-memory latency no longer sit in front of the tensor pipe.
  // The reference throws at the failing pivot (cholesky(C), src/GaussDCA.jl:34).  Here the ~250 launches behind a failed diagonal
  // block are already queued: each of them reads the flag (under the latency of its first operand loads) and leaves.
  // One thread reads, the CTA votes: a diagonal block on another stream may set the flag while this launch is starting, and
  // the threads of a CTA must agree on leaving.
  int failed = (p.info && tid == 0) ? *reinterpret_cast<const volatile int *>(p.info) : 0;
  load_stage(0);
  load_stage(1);
  load_stage(2);
  failed = __syncthreads_or(failed);
  if (failed) {
    cp_wait<0>();
    return;
  }
  double fa[2][8], fb[2][4];  // ping-pong fragment registers (BK/4 is even: every stage starts on buffer 0)
  cp_wait<2>();  // stage 0
  __syncthreads();
  if (nk > 0) load_frags(sm, sm + TILE_D, 0, fa[0], fb[0]);
  for (int kt = 0; kt < nk; ++kt) {
    cp_wait<1>();  // stages <= kt+1 (only kt+2 may still be in flight)
    __syncthreads();
    load_stage(kt + GSTAGES - 1);  // slot of stage kt-1: everyone finished reading it before this barrier
    const double *As = sm + (size_t)(kt % GSTAGES) * 2 * TILE_D;
    const double *Bs = As + TILE_D;
    const double *An = sm + (size_t)((kt + 1) % GSTAGES) * 2 * TILE_D;
    const double *Bn = An + TILE_D;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      if (kk + 1 < BK / 4)
        load_frags(As, Bs, kk + 1, fa[(kk + 1) & 1], fb[(kk + 1) & 1]);
      else if (kt + 1 < nk)
        load_frags(An, Bn, 0, fa[0], fb[0]);
#pragma unroll
      for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], fa[kk & 1][mi], fb[kk & 1][ni]);
    }
  }
  cp_wait<0>();

  // epilogue: thread owns (row g, cols 2c,2c+1) of every 8x8 atom
#pragma unroll
  for (int mi = 0; mi < 8; ++mi) {
    const long long row = m0 + wm * 64 + mi * 8 + g;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const long long col = n0 + wn * 32 + ni * 8 + 2 * c;
      double2 *dst = reinterpret_cast<double2 *>(C + row * p.ldc + col);
      double2 v;
      v.x = p.alpha * acc[mi][ni][0];
      v.y = p.alpha * acc[mi][ni][1];
      if (p.beta != 0.0) {
        const double2 old = *dst;
        v.x += p.beta * old.x;
        v.y += p.beta * old.y;
      }
      *dst = v;
    }
  }
}

// ---- the two 128 x 128 x 128 products on the critical chain of the factorisation (panel tile, diagonal update) ----
// C[128 x 128] = beta C + alpha A B', A and B as [row][k].  The big kernel above gives such a product ONE CTA whose 4-slot ring
// never fills (8 k-steps): ~26 us.  Here 8 CTAs of 4 warps take a strip of 16 output rows each: the strip's 16 operand rows of A
// and all 128 rows of B arrive in one cp.async burst, then 256 DMMAs per warp.  In-place use (C == A, the panel tile) is safe:
// a CTA is the only reader of its A rows and has them in shared memory before its first store.
constexpr int SG_M = 16;          // output rows per CTA
constexpr int SG_LD = NB + 4;     // operand row stride (doubles): conflict-free 64-bit fragment loads
__global__ void __launch_bounds__(128) dgemm_small_kernel(GemmP p) {
  extern __shared__ __align__(16) double sm[];
  double *As = sm, *Bs = sm + SG_M * SG_LD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, c = lane & 3;
  const int m0 = blockIdx.x * SG_M;
  int failed = (p.info && tid == 0) ? *reinterpret_cast<const volatile int *>(p.info) : 0;
  for (int ch = tid; ch < (SG_M + NB) * (NB / 2); ch += 128) {  // 16-byte chunks: row = ch / 64, two doubles at 2 (ch % 64)
    const int row = ch >> 6, cc = ch & 63;
    if (row < SG_M)
      cp16(As + row * SG_LD + cc * 2, p.A + (long long)(m0 + row) * p.lda + cc * 2);
    else
      cp16(Bs + (row - SG_M) * SG_LD + cc * 2, p.B + (long long)(row - SG_M) * p.ldb + cc * 2);
  }
  cp_commit();
  failed = __syncthreads_or(failed);
  cp_wait<0>();
  if (failed) return;
  __syncthreads();
  const int wn = warp * 32;  // warp tile 16 x 32 = 2 x 4 atoms of 8 x 8
  double acc[2][4][2] = {};
#pragma unroll 4
  for (int kk = 0; kk < NB / 4; ++kk) {
    double a[2], b[4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) a[mi] = As[(mi * 8 + g) * SG_LD + kk * 4 + c];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[(wn + ni * 8 + g) * SG_LD + kk * 4 + c];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
  }
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    const long long row = m0 + mi * 8 + g;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const long long col = wn + ni * 8 + 2 * c;
      double2 *dst = reinterpret_cast<double2 *>(p.C + row * p.ldc + col);
      double2 v;
      v.x = p.alpha * acc[mi][ni][0];
      v.y = p.alpha * acc[mi][ni][1];
      if (p.beta != 0.0) {
        const double2 old = *dst;
        v.x += p.beta * old.x;
        v.y += p.beta * old.y;
      }
      *dst = v;
    }
  }
}

// ---- diagonal block: A[k,k] (lower) -> inv(chol(A[k,k])) written to X[k,k] (lower, zeros above) ----
// One CTA of 16 warps, the 128 x 128 block lives in REGISTERS (32 doubles per thread), both phases are
// right-looking rank-1 sweeps with ONE barrier per step and a one-step look-ahead:
//   phase 1 (Cholesky): warp w owns the 8 columns c = w + 16b, lane l the 4 rows i = l + 32a.  In step j1
//     everybody applies column j1-1; the warp that owns column j1 updates that column FIRST, finalises it
//     (pivot by shuffle, one rsqrt) and publishes it through a double-buffered shared line while the
//     other warps are still busy with their updates -- the serial chain hides behind the rank-1 update.
//   phase 2 (X = L^-1, forward substitution on all 128 right-hand sides): transposed ownership (warp w
//     owns rows i = w + 16a, lane l columns c = l + 32b), same look-ahead on rows.
// The step loops are written as (slot, 16 steps) nests so that the owner's register slot is a
// compile-time index (no dynamic register indexing, no select chains).
constexpr int DT = 512;
constexpr int DLD = NB + 1;
#ifdef DIAG_DBG
__device__ long long g_diag_clk[8];
#define DIAG_STAMP(i) do { if (threadIdx.x == 0) g_diag_clk[i] = clock64(); } while (0)
#else
#define DIAG_STAMP(i)
#endif

__global__ void __launch_bounds__(DT, 1) diag_block_kernel(const double *__restrict__ Akk, long long lda,
                                                           double *__restrict__ Xkk, long long ldx, int col0,
                                                           int n_true, int *__restrict__ info) {
  extern __shared__ double Ls[];  // [NB][DLD]: staging, then the factor for phase 2
  __shared__ double line[2][NB];
  __shared__ double rdiag[NB];    // 1 / L[j][j]
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  DIAG_STAMP(0);
  if (*reinterpret_cast<const volatile int *>(info) != 0) return;  // an earlier block already failed: nothing left to factor

  // coalesced load through shared memory (rows contiguous), then pick own elements
  for (int e = tid; e < NB * NB; e += DT) {
    const int r = e >> 7, c = e & 127;
    Ls[r * DLD + c] = (c <= r) ? Akk[(long long)r * lda + c] : 0.0;
  }
  __syncthreads();
  double v[4][8];  // v[a][b] = element (i = l + 32a, c = w + 16b)
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) v[a][b] = Ls[(l + 32 * a) * DLD + w + 16 * b];
  DIAG_STAMP(1);

  // ---------------- phase 1: Cholesky, lower ----------------
#pragma unroll
  for (int b1 = 0; b1 < 8; ++b1) {        // register slot of the column being finalised
#pragma unroll 1
    for (int jj = 0; jj < 16; ++jj) {
      const int j1 = 16 * b1 + jj;        // column finalised in this step (owner: warp jj)
      const int j = j1 - 1;               // column applied in this step (-1: none)
      const bool owner = (w == jj);
      const double *col = line[j & 1];
      double li[4], lc[8];
      if (j >= 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a) li[a] = col[l + 32 * a];
#pragma unroll
        for (int b = 0; b < 8; ++b) lc[b] = col[w + 16 * b];
      }
      if (owner) {
        if (j >= 0) {
#pragma unroll
          for (int a = 0; a < 4; ++a) v[a][b1] = fma(-li[a], lc[b1], v[a][b1]);
        }
        // pivot: row j1 = lane (j1 & 31) of slot a = j1 >> 5 = b1 >> 1 (compile time)
        const double d = __shfl_sync(0xffffffffu, v[b1 >> 1][b1], j1 & 31);
        if (!(d > 0.0)) {  // also catches NaN
          if (l == 0 && col0 + j1 < n_true) atomicCAS(info, 0, col0 + j1 + 1);
        }
        const double rs = rsqrt(d);
        if (l == 0) rdiag[j1] = rs;
        double *out = line[j1 & 1];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int i = l + 32 * a;
          const double lij = (i == j1) ? d * rs : ((i > j1) ? v[a][b1] * rs : 0.0);  // final L[i][j1]
          v[a][b1] = lij;
          out[i] = lij;
        }
      }
      if (j >= 0) {
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const int c = w + 16 * b;
          // warp-uniform: columns right of j, except the one the owner has already brought up to date
          if (c > j && !(owner && b == b1)) {
            // rows above the diagonal (i < c) only hold don't-care values: no per-element predicate needed
#pragma unroll
            for (int a = 0; a < 4; ++a) v[a][b] = fma(-li[a], lc[b], v[a][b]);
          }
        }
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int i = l + 32 * a, c = w + 16 * b;
      Ls[i * DLD + c] = (c <= i) ? v[a][b] : 0.0;
    }
  __syncthreads();
  DIAG_STAMP(2);

  // ---------------- phase 2: X = L^-1 ----------------
  double t[8][4];  // t[a][b] = element (i = w + 16a, c = l + 32b)
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) t[a][b] = (w + 16 * a == l + 32 * b) ? 1.0 : 0.0;
#pragma unroll
  for (int a1 = 0; a1 < 8; ++a1) {        // register slot of the row being finalised
#pragma unroll 1
    for (int kk = 0; kk < 16; ++kk) {
      const int k1 = 16 * a1 + kk;        // row finalised in this step (owner: warp kk)
      const int k = k1 - 1;               // row applied in this step (-1: none)
      const bool owner = (w == kk);
      const double *row = line[k & 1];
      double li[8], xr[4];
      if (k >= 0) {
#pragma unroll
        for (int a = 0; a < 8; ++a) li[a] = Ls[(w + 16 * a) * DLD + k];  // warp-uniform; zero for i < k
#pragma unroll
        for (int b = 0; b < 4; ++b) xr[b] = row[l + 32 * b];             // zero for c > k
      }
      if (owner) {
        const double dk = rdiag[k1];
        double *out = line[k1 & 1];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          double x = t[a1][b];
          if (k >= 0) x = fma(-li[a1], xr[b], x);
          x *= dk;  // X[k1, c] final (zero for c > k1)
          t[a1][b] = x;
          out[l + 32 * b] = x;
        }
      }
      if (k >= 0) {
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          if (w + 16 * a > k && !(owner && a == a1)) {  // warp-uniform
            // xr is zero for columns > k, so the update is exact without a per-element predicate
#pragma unroll
            for (int b = 0; b < 4; ++b) t[a][b] = fma(-li[a], xr[b], t[a][b]);
          }
        }
      }
      __syncthreads();
    }
  }
  DIAG_STAMP(3);
  // stage through shared memory for coalesced row writes
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = w + 16 * a, c = l + 32 * b;
      Ls[i * DLD + c] = (c <= i) ? t[a][b] : 0.0;
    }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += DT) {
    const int r = e >> 7, c = e & 127;
    Xkk[(long long)r * ldx + c] = Ls[r * DLD + c];
  }
  DIAG_STAMP(4);
}

// ---- diagonal block, blocked (round 2; the chain of the factorisation waits for this kernel 79 times at n = 10 000) ----
// Same contract as diag_block_kernel: A[k,k] (lower) -> inv(chol(A[k,k])) written to X[k,k] (lower, zeros above); a non-positive
// pivot records the order of the failing leading minor.  Same arithmetic per element (rs = rsqrt(pivot), L_jj = pivot * rs,
// L_ij = a_ij * rs, right-looking updates), different schedule: instead of 128 + 128 barrier-separated rank-1 steps of the whole
// CTA (68 us), the block is factorised in four 32-column panels held in shared memory:
//   (a) the 32 x 32 diagonal block: warp 0 owns it, a row per lane in registers.  The serial chain per column is
//       pivot -> rsqrt -> multiplier -> next pivot (one shuffle); the finished column goes to shared memory, the rank-1 update
//       reads it back as broadcast words.  Warp 1 follows one column behind (a flag in shared memory, no barrier) and builds the
//       INVERSE of the block by forward substitution, a column per lane, off the chain;
//   (b) the rows below: L = A * inv(L_D)' as DMMA strips of 8 rows (a warp owns its strip: in place);
//   (c) the trailing block: A -= L L' on the 8 x 8 lower tiles, DMMA, three tiles in flight per warp;
// and X = L^-1 by recursive doubling on the 32-blocks (X21 = -X22 (L21 X11), two levels, four DMMA stages, independent
// accumulator chains interleaved) with the unused upper triangle of the block as workspace.  The first 32 rows are loaded
// first: the other warps fetch the rest of the block while warps 0 and 1 already factorise.  16 CTA barriers instead of 256.
constexpr int D2_LD = NB + 4;   // == 4 (mod 16): conflict-free 64-bit DMMA fragment loads in both orientations
constexpr int D2_ID = 36;       // leading dimension of the inverted 32 x 32 diagonal blocks (same residue)
constexpr size_t D2_SMEM = (size_t)(NB * D2_LD + 4 * 32 * D2_ID + 32 * 32 + 32) * sizeof(double);

__global__ void __launch_bounds__(DT, 1) diag_block_kernel2(const double *__restrict__ Akk, long long lda, double *__restrict__ Xkk,
                                                            long long ldx, int col0, int n_true, int *__restrict__ info) {
  extern __shared__ __align__(16) double sm2[];
  double *Ls = sm2;                       // [128][D2_LD]: the block, then its factor (lower); upper blocks: workspace of the inversion
  double *Dv = sm2 + NB * D2_LD;          // [4][32][D2_ID]: inverses of the diagonal 32 x 32 blocks of the factor
  double *Lc = Dv + 4 * 32 * D2_ID;       // [32 columns][32 rows]: the diagonal block's factor, column by column as it is finished
  double *rsd = Lc + 32 * 32;             // [32] 1 / L[j][j]
  __shared__ int s_ready;                 // columns of Lc published by warp 0
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  const int g = l >> 2, c = l & 3;
  constexpr unsigned FULL = 0xffffffffu;
  if (*reinterpret_cast<const volatile int *>(info) != 0) return;  // an earlier block already failed: nothing left to factor
  DIAG_STAMP(0);

  // rows r0 .. r1-1 of the block into shared memory: 16-byte chunks, all loads of a thread in flight before its first store;
  // the strict upper triangle is never read from memory
  auto load_rows = [&](int rbeg, int rend, int t0, int nthr) {
    for (int e0 = rbeg * 64 + t0; e0 < rend * 64; e0 += 4 * nthr) {
      double2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * nthr, r = e >> 6, cc = (e & 63) * 2;
        v[u] = make_double2(0.0, 0.0);
        if (e < rend * 64 && cc <= r) v[u] = *reinterpret_cast<const double2 *>(Akk + (long long)r * lda + cc);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * nthr, r = e >> 6, cc = (e & 63) * 2;
        if (e < rend * 64) {
          if (cc + 1 > r) v[u].y = 0.0;
          *reinterpret_cast<double2 *>(Ls + r * D2_LD + cc) = v[u];
        }
      }
    }
  };
  load_rows(0, 32, tid, DT);
  if (tid == 0) s_ready = 0;
  __syncthreads();
  DIAG_STAMP(1);

  // ---------------- phase 1: Cholesky, four panels of 32 columns ----------------
#pragma unroll 1
  for (int p = 0; p < 4; ++p) {
    const int r0 = 32 * p;
    double *Dp = Dv + p * 32 * D2_ID;
    if (w == 0) {
      // (a) factor: lane l owns row l of the diagonal block
      double v[32];
#pragma unroll
      for (int cc = 0; cc < 32; ++cc) v[cc] = Ls[(r0 + l) * D2_LD + r0 + cc];
      double d = __shfl_sync(FULL, v[0], 0);
      double rs = rsqrt(d);
      int first_bad = (d > 0.0) ? 0 : 1;  // 1-based column of the first non-positive pivot of this block (also catches NaN)
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        double lij = (l == j) ? d * rs : v[j] * rs;  // final L[l][j]
        if (l < j) lij = 0.0;
        const double rs_j = rs;
        if (j < 31) {
          // the next pivot: its own lane needs nobody else's multiplier for the update of its diagonal element.  Its rsqrt --
          // the long dependent chain of the column -- starts HERE, so that it runs under the rank-1 update below
          const double dn = fma(-lij, lij, v[j + 1]);
          d = __shfl_sync(FULL, dn, j + 1);
          rs = rsqrt(d);
          if (first_bad == 0 && !(d > 0.0)) first_bad = j + 2;
        }
        Lc[j * 32 + l] = lij;
        if (l == j) rsd[j] = rs_j;
        __syncwarp();
        if ((j & 3) == 3) {  // the inverse (warp 1) follows at most four columns behind
          if (l == 0) {
            __threadfence_block();
            *reinterpret_cast<volatile int *>(&s_ready) = j + 1;
          }
        }
        // rank-1 update of the columns to the right, multipliers read back as broadcast pairs
#pragma unroll
        for (int cc = (j + 1) & ~1; cc < 32; cc += 2) {
          const double2 lc = *reinterpret_cast<const double2 *>(Lc + j * 32 + cc);
          if (cc > j) v[cc] = fma(-lij, lc.x, v[cc]);
          v[cc + 1] = fma(-lij, lc.y, v[cc + 1]);
        }
      }
      if (first_bad && l == 0 && col0 + r0 + first_bad - 1 < n_true) atomicCAS(info, 0, col0 + r0 + first_bad);
    } else if (w == 1) {
      // (a') inverse of the diagonal block, one column behind the factor: lane l owns column l, forward substitution
      double x[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = (i == l) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        if ((k & 3) == 0) {  // columns are published four at a time
          while (*reinterpret_cast<volatile int *>(&s_ready) <= k + 3) {
          }
          __threadfence_block();
        }
        x[k] *= rsd[k];
#pragma unroll
        for (int i = (k + 1) & ~1; i < 32; i += 2) {
          const double2 li = *reinterpret_cast<const double2 *>(Lc + k * 32 + i);
          if (i > k) x[i] = fma(-li.x, x[k], x[i]);
          x[i + 1] = fma(-li.y, x[k], x[i + 1]);
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) Dp[i * D2_ID + l] = x[i];
    } else if (p == 0) {
      load_rows(32, NB, tid - 64, DT - 64);  // the rest of the block arrives while the first diagonal block is factorised
    }
    __syncthreads();
    if (tid == 0) s_ready = 0;  // (nobody reads it before the next barrier)
    if (p == 0) DIAG_STAMP(5);
    const int m = NB - r0 - 32;  // rows below the diagonal block
    if (m > 0) {
      // (b) panel: rows i0 .. i0+7 of  A[:, r0:r0+32] * inv(L_D)'   (inv(L_D)[n][k] = 0 for k > n)
      for (int s = w; s < m / 8; s += DT / 32) {
        const int i0 = r0 + 32 + 8 * s;
        double acc[4][2] = {};
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const double a = Ls[(i0 + g) * D2_LD + r0 + kk * 4 + c];
#pragma unroll
          for (int ni = 0; ni < 4; ++ni)
            if (kk <= 2 * ni + 1) dmma(acc[ni][0], acc[ni][1], a, Dp[(ni * 8 + g) * D2_ID + kk * 4 + c]);
        }
        __syncwarp();
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
          *reinterpret_cast<double2 *>(Ls + (i0 + g) * D2_LD + r0 + ni * 8 + 2 * c) = make_double2(acc[ni][0], acc[ni][1]);
      }
      __syncthreads();
      // (c) trailing update on the lower 8 x 8 tiles of the m x m block, three independent tiles per round and warp
      const int T = m / 8, ntile = T * (T + 1) / 2;
      for (int tb = 3 * w; tb < ntile; tb += 3 * (DT / 32)) {
        int i0[3], j0[3];
        double d0[3] = {}, d1[3] = {};
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int t = min(tb + u, ntile - 1);
          int ti = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
          while (ti * (ti + 1) / 2 > t) --ti;
          while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
          i0[u] = r0 + 32 + 8 * ti;
          j0[u] = r0 + 32 + 8 * (t - ti * (ti + 1) / 2);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
          for (int u = 0; u < 3; ++u)
            dmma(d0[u], d1[u], Ls[(i0[u] + g) * D2_LD + r0 + kk * 4 + c], Ls[(j0[u] + g) * D2_LD + r0 + kk * 4 + c]);
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          if (tb + u < ntile) {
            double2 *dst = reinterpret_cast<double2 *>(Ls + (i0[u] + g) * D2_LD + j0[u] + 2 * c);
            double2 old = *dst;
            old.x -= d0[u];
            old.y -= d1[u];
            *dst = old;
          }
        }
      }
      __syncthreads();
    }
  }
  DIAG_STAMP(2);

  // ---------------- phase 2: X = L^-1 by recursive doubling on the 32-blocks ----------------
  // level A: X[1][0] = -D1 (L[1][0] D0),  X[3][2] = -D3 (L[3][2] D2)   (D_b = inverse of diagonal block b)
  {
    const int q = w >> 3;                    // pair 0: blocks (0,1), pair 1: blocks (2,3)
    const int lo = 2 * q, hi = 2 * q + 1;
    const double *Dlo = Dv + lo * 32 * D2_ID, *Dhi = Dv + hi * 32 * D2_ID;
    double *Wq = Ls + (32 * q) * D2_LD + 64;           // T of this pair: rows 32q.., columns 64..95 (upper workspace)
    double *Xq = Ls + (64 * q) * D2_LD + 64 * q + 32;  // X[hi][lo]: block (0,1) resp. (2,3) of the upper workspace
    const int ti = (w & 7) >> 1, tj0 = 2 * (w & 1);    // this warp's two tiles: (ti, tj0) and (ti, tj0 + 1), one A fragment
    {
      // T = L[hi][lo] * Dlo, Dlo[k][j] = 0 for k < j
      double d[2][2] = {};
      for (int kk = 2 * tj0; kk < 8; ++kk) {
        const double a = Ls[(32 * hi + 8 * ti + g) * D2_LD + 32 * lo + kk * 4 + c];
        dmma(d[0][0], d[0][1], a, Dlo[(kk * 4 + c) * D2_ID + 8 * tj0 + g]);
        if (kk >= 2 * tj0 + 2) dmma(d[1][0], d[1][1], a, Dlo[(kk * 4 + c) * D2_ID + 8 * tj0 + 8 + g]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
        *reinterpret_cast<double2 *>(Wq + (8 * ti + g) * D2_LD + 8 * (tj0 + u) + 2 * c) = make_double2(d[u][0], d[u][1]);
    }
    __syncthreads();
    {
      // X[hi][lo] = -Dhi * T, Dhi[i][k] = 0 for k > i
      double d[2][2] = {};
      for (int kk = 0; kk <= 2 * ti + 1; ++kk) {
        const double a = Dhi[(8 * ti + g) * D2_ID + kk * 4 + c];
        dmma(d[0][0], d[0][1], a, Wq[(kk * 4 + c) * D2_LD + 8 * tj0 + g]);
        dmma(d[1][0], d[1][1], a, Wq[(kk * 4 + c) * D2_LD + 8 * tj0 + 8 + g]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
        *reinterpret_cast<double2 *>(Xq + (8 * ti + g) * D2_LD + 8 * (tj0 + u) + 2 * c) = make_double2(-d[u][0], -d[u][1]);
    }
    __syncthreads();
  }
  // level B: X21 = -X22 (L21 X11) on the 64-blocks; X11 = [[D0, 0], [X10, D1]], X22 = [[D2, 0], [X32, D3]]
  double *Wb = Ls + 64;                      // T: rows 0..63, columns 64..127 of the upper workspace
  {
    // T = L21 * X11: column tiles a4 and 7 - a4 (balanced k ranges), two row tiles each; four accumulator chains in flight
    const int a4 = w & 3, tib = (w >> 2) * 2;
    const int tjA = a4, tjB = 7 - a4;
    double d[4][2] = {};
    for (int kk = 2 * tjA; kk < 16; ++kk) {
      const double a0 = Ls[(64 + 8 * tib + g) * D2_LD + kk * 4 + c], a1 = Ls[(64 + 8 * tib + 8 + g) * D2_LD + kk * 4 + c];
      {
        const double *bp = (kk < 8) ? Dv + (kk * 4 + c) * D2_ID + 8 * tjA + g                 // D0
                                    : Ls + (kk * 4 - 32 + c) * D2_LD + 32 + 8 * tjA + g;      // X10
        const double b = *bp;
        dmma(d[0][0], d[0][1], a0, b);
        dmma(d[1][0], d[1][1], a1, b);
      }
      if (kk >= 2 * tjB) {  // tjB >= 4: k >= 32, the D1 block
        const double b = Dv[32 * D2_ID + (kk * 4 - 32 + c) * D2_ID + 8 * tjB - 32 + g];
        dmma(d[2][0], d[2][1], a0, b);
        dmma(d[3][0], d[3][1], a1, b);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      *reinterpret_cast<double2 *>(Wb + (8 * (tib + (u & 1)) + g) * D2_LD + 8 * ((u & 2) ? tjB : tjA) + 2 * c) = make_double2(d[u][0], d[u][1]);
  }
  __syncthreads();
  {
    // X21 = -X22 * T: row tiles a4 and 7 - a4, two column tiles each; straight to global memory
    const int a4 = w & 3, tjb = (w >> 2) * 2;
    const int tiA = a4, tiB = 7 - a4;
    double d[4][2] = {};
    for (int kk = 0; kk <= 2 * tiB + 1; ++kk) {
      const double b0 = Wb[(kk * 4 + c) * D2_LD + 8 * tjb + g], b1 = Wb[(kk * 4 + c) * D2_LD + 8 * tjb + 8 + g];
      if (kk <= 2 * tiA + 1) {  // tiA < 4: the D2 block
        const double a = Dv[2 * 32 * D2_ID + (8 * tiA + g) * D2_ID + kk * 4 + c];
        dmma(d[0][0], d[0][1], a, b0);
        dmma(d[1][0], d[1][1], a, b1);
      }
      {
        const double *ap = (kk < 8) ? Ls + (64 + 8 * tiB - 32 + g) * D2_LD + 96 + kk * 4 + c                    // X32
                                    : Dv + 3 * 32 * D2_ID + (8 * tiB - 32 + g) * D2_ID + kk * 4 - 32 + c;       // D3
        const double a = *ap;
        dmma(d[2][0], d[2][1], a, b0);
        dmma(d[3][0], d[3][1], a, b1);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      *reinterpret_cast<double2 *>(Xkk + (long long)(64 + 8 * ((u & 2) ? tiB : tiA) + g) * ldx + 8 * (tjb + (u & 1)) + 2 * c) =
          make_double2(-d[u][0], -d[u][1]);
  }
  DIAG_STAMP(3);
  // the rest of X: the four inverted diagonal blocks, X10, X32, zeros elsewhere (rows 64.., columns 0..63 were written above)
  for (int e = tid; e < NB * NB / 2; e += DT) {
    const int r = e >> 6, cc = (e & 63) * 2;
    if (r >= 64 && cc < 64) continue;
    const int rb = r >> 5, cb = cc >> 5;
    double2 v = make_double2(0.0, 0.0);
    if (rb == cb)
      v = *reinterpret_cast<const double2 *>(Dv + rb * 32 * D2_ID + (r & 31) * D2_ID + (cc & 31));
    else if (rb == cb + 1 && (rb & 1))
      v = *reinterpret_cast<const double2 *>(Ls + (r - 32) * D2_LD + 32 + cc);
    *reinterpret_cast<double2 *>(Xkk + (long long)r * ldx + cc) = v;
  }
  DIAG_STAMP(4);
}

__global__ void pad_identity_kernel(double *__restrict__ C, long long n, long long npad) {
  const long long r = n + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < npad) C[r * npad + r] = 1.0;
}

// upper <- lower (plain mirror), 32x32 tiles
__global__ void mirror_lower_kernel(double *__restrict__ A, long long n, long long ld) {
  __shared__ double tile[32][33];
  const int tr = blockIdx.y, tc = blockIdx.x;
  if (tc > tr) return;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  for (int yy = ly; yy < 32; yy += 8) {
    const long long rr = (long long)tr * 32 + yy, cc = (long long)tc * 32 + lx;
    tile[yy][lx] = (rr < n && cc < n) ? A[rr * ld + cc] : 0.0;
  }
  __syncthreads();
  for (int yy = ly; yy < 32; yy += 8) {
    const long long rr = (long long)tc * 32 + yy, cc = (long long)tr * 32 + lx;  // transposed position
    if (rr < n && cc < n && rr < cc) A[rr * ld + cc] = tile[lx][yy];
  }
}

template <bool AT, bool BT>
int32_t gemm(gdca_ctx *ctx, const GemmP &p_in, int batch, cudaStream_t stream = nullptr) {
  GemmP p = p_in;
  p.info = ctx->leader ? ctx->leader->dInfo : ctx->dInfo;  // a group member watches the leader's factorisation
  if (!stream) stream = ctx->stream;
  if (p.m <= 0 || p.n <= 0 || batch <= 0) return GDCA_OK;
  const size_t smem = (size_t)GSTAGES * 2 * TILE_D * sizeof(double);
  GDCA_CUDA(ctx, cudaFuncSetAttribute(dgemm_kernel<AT, BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)(p.n / NB), (unsigned)(p.m / NB), (unsigned)batch);
  GDCA_CUDA(ctx, gdca_launch_prio(dgemm_kernel<AT, BT>, grid, dim3(GTHREADS), smem, stream, p));
  GDCA_LAUNCH_CHECK(ctx);
  return GDCA_OK;
}

// 128 x 128 x 128, both operands [row][k], on the critical chain: 8 strip CTAs instead of one ring-pipelined CTA
int32_t gemm_small(gdca_ctx *ctx, const GemmP &p_in, cudaStream_t stream) {
  GemmP p = p_in;
  p.info = ctx->leader ? ctx->leader->dInfo : ctx->dInfo;
  const size_t smem = (size_t)(SG_M + NB) * SG_LD * sizeof(double);
  GDCA_CUDA(ctx, cudaFuncSetAttribute(dgemm_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GDCA_CUDA(ctx, gdca_launch_prio(dgemm_small_kernel, dim3(NB / SG_M), dim3(128), smem, stream, p));
  GDCA_LAUNCH_CHECK(ctx);
  return GDCA_OK;
}

}  // namespace

// Stream capture of the single-GPU inversion: the ~900 launches, memsets and cross-stream event hand-shakes of one call become
// two CUDA graphs (factorisation | inversion of the factor + X'X: the timing event between them stays a real event).
struct InvCapture {
  gdca_ctx *ctx;
  cudaGraph_t graph[2] = {nullptr, nullptr};
  int parts = 0;
  bool failed = false;
  // between the factorisation and the inversion: close the first graph, open the second
  int32_t split() {
    if (cudaStreamEndCapture(ctx->stream, &graph[0]) != cudaSuccess || !graph[0]) return fail();
    parts = 1;
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) return fail();
    return GDCA_OK;
  }
  int32_t fail() {
    failed = true;
    cudaGetLastError();
    return gdca_fail(ctx, GDCA_ERR_CUDA, "inverse: stream capture failed");
  }
};

static int32_t inverse_enqueue(gdca_ctx *ctx, InvCapture *cap) {
  const long long np = ctx->npad, n = ctx->n;
  const int nb = (int)(np / NB);
  GDCA_TRY(gdca_reserve(ctx, ctx->dX, ctx->capX, (size_t)np * np));
  GDCA_TRY(gdca_reserve(ctx, ctx->dT, ctx->capT, (size_t)np * np));
  GDCA_TRY(gdca_reserve(ctx, ctx->dmJ, ctx->capmJ, (size_t)np * np));
  // INT8-sliced tcgen05 GEMMs (ozaki.cu) for the big products; DMMA for diagonal blocks, K = 128 panels and small shapes
  const bool oz = ctx->ozaki_mode != 0 && nb >= 16;
  ctx->last_inverse_ozaki = oz;
  ctx->last_inverse_shared = false;
  ctx->oz_int8_ops = 0.0;
  ctx->oz_fp64_flop = 0.0;
  if (oz) {
    GDCA_TRY(gdca_reserve(ctx, ctx->dDigA, ctx->capDigA, (size_t)np * np * 8));   // lauum: all of X', 8 digit slots per element
    GDCA_TRY(gdca_reserve(ctx, ctx->dDigB, ctx->capDigB, (size_t)np * np * 8));   // trtri operands (the top block of the last level
                                                                                   // is up to (nb - 1) blocks wide), potrf panels
    GDCA_TRY(gdca_reserve(ctx, ctx->dScaleA, ctx->capScaleA, (size_t)np));
    GDCA_TRY(gdca_reserve(ctx, ctx->dScaleB, ctx->capScaleB, (size_t)np));
    GDCA_TRY(gdca_reserve(ctx, ctx->dDigP, ctx->capDigP, (size_t)NB * 8 * 1024));  // 128 rows x (K <= 1024) x 8 digit slots
    GDCA_TRY(gdca_reserve(ctx, ctx->dScaleP, ctx->capScaleP, (size_t)NB));
  }
  constexpr int OZ_MIN_REM = 8;   // trailing updates with at least this many block rows left
  constexpr int OZ_MIN_H = 4;     // trtri levels with K >= 512
  double *A = ctx->dC, *X = ctx->dX, *T = ctx->dT, *J = ctx->dmJ;
  // Device group: the factorisation runs on the leader (its serial chain of diagonal blocks does not shard) while every
  // finished panel is copied to the other members beside it; the inversion of the factor and the product X'X are then shared.
  const int N = (oz && ctx->group_size > 1) ? ctx->group_size : 1;
  gdca_ctx **grp = ctx->group;
  auto on_dev = [&](gdca_ctx *c) -> int32_t {
    GDCA_CUDA(ctx, cudaSetDevice(c->device));
    return GDCA_OK;
  };
  auto mtry = [&](gdca_ctx *c, int32_t st) -> int32_t {
    if (st != GDCA_OK && c != ctx) ctx->err = "device " + std::to_string(c->device) + ": " + c->err;
    return st;
  };
  for (int r = 1; r < N; ++r) {
    gdca_ctx *c = grp[r];
    GDCA_TRY(on_dev(c));
    c->npad = np;
    c->n = n;
    GDCA_TRY(mtry(c, gdca_reserve(c, c->dC, c->capC, (size_t)np * np)));
    GDCA_TRY(mtry(c, gdca_reserve(c, c->dX, c->capX, (size_t)np * np)));
    GDCA_TRY(mtry(c, gdca_reserve(c, c->dT, c->capT, (size_t)np * np)));
    GDCA_TRY(mtry(c, gdca_reserve(c, c->dDigA, c->capDigA, (size_t)np * np * 8)));
    GDCA_TRY(mtry(c, gdca_reserve(c, c->dDigB, c->capDigB, (size_t)np * np * 8)));
    GDCA_TRY(mtry(c, gdca_reserve(c, c->dScaleA, c->capScaleA, (size_t)np)));
    GDCA_TRY(mtry(c, gdca_reserve(c, c->dScaleB, c->capScaleB, (size_t)np)));
    c->oz_int8_ops = c->oz_fp64_flop = 0.0;
    // the copy stream starts behind whatever the member's compute stream still does with these buffers
    GDCA_CUDA(ctx, cudaEventRecord(c->ev_copy, c->stream));
    GDCA_CUDA(ctx, cudaStreamWaitEvent(c->stream_copy, c->ev_copy, 0));
    GDCA_CUDA(ctx, cudaMemsetAsync(c->dX, 0, (size_t)np * np * sizeof(double), c->stream_copy));
  }
  GDCA_TRY(on_dev(ctx));
  GDCA_CUDA(ctx, cudaMemsetAsync(ctx->dInfo, 0, sizeof(int), ctx->stream));
  GDCA_CUDA(ctx, cudaMemsetAsync(X, 0, (size_t)np * np * sizeof(double), ctx->stream));
  if (np > n) {
    pad_identity_kernel<<<(unsigned)((np - n + 127) / 128), 128, 0, ctx->stream>>>(A, n, np);
    GDCA_LAUNCH_CHECK(ctx);
  }
  const size_t dsmem = (size_t)NB * DLD * sizeof(double);
  GDCA_CUDA(ctx, cudaFuncSetAttribute(diag_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem));
  GDCA_CUDA(ctx, cudaFuncSetAttribute(diag_block_kernel2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D2_SMEM));
  const bool diag2 = ctx->diag_blocked != 0;
  auto blk = [&](double *base, int I, int Jb) { return base + ((long long)I * NB) * np + (long long)Jb * NB; };
  constexpr int OB = 4;  // 128-blocks per outer block of the factorisation
  // SHARED FACTORISATION (device group, n >= share_min_nb * 128: the bulk trailing updates outweigh the serial chain).
  // Outer block column ob (512 columns) of the trailing matrix belongs to member ob % N: the owner applies every broadcast
  // panel to it and ships it to the leader one step before the leader factorises it.  The chain itself stays on the leader.
  bool share = N > 1 && nb >= ctx->share_min_nb;
  const bool shared_factorisation = share;
  ctx->last_inverse_shared = shared_factorisation;
  auto owner = [&](int ob) { return ob % N; };
  auto copy_cols = [&](gdca_ctx *on, cudaStream_t st, double *dst, const double *src, int K, int wblocks) -> int32_t {
    // rows [K, nb) of the block columns [K, K + wblocks)
    GDCA_CUDA(ctx, cudaSetDevice(on->device));
    GDCA_CUDA(ctx, cudaMemcpy2DAsync(dst + ((long long)K * NB) * np + (long long)K * NB, (size_t)np * sizeof(double),
                                     src + ((long long)K * NB) * np + (long long)K * NB, (size_t)np * sizeof(double),
                                     (size_t)wblocks * NB * sizeof(double), (size_t)(nb - K) * NB, cudaMemcpyDefault, st));
    return GDCA_OK;
  };
  if (share) {
    GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_copy, ctx->stream));  // C (+ identity padding) is complete on the leader
    for (int r = 1; r < N; ++r) {
      gdca_ctx *c = grp[r];
      GDCA_TRY(on_dev(c));
      GDCA_CUDA(ctx, cudaStreamWaitEvent(c->stream_copy, ctx->ev_copy, 0));
      for (int ob = r; ob * OB < nb; ob += N) GDCA_TRY(copy_cols(c, c->stream_copy, c->dC, A, ob * OB, std::min(OB, nb - ob * OB)));
      GDCA_CUDA(ctx, cudaEventRecord(c->ev_copy, c->stream_copy));
      GDCA_CUDA(ctx, cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
    }
    GDCA_TRY(on_dev(ctx));
  }

  // ---------------- potrf: two-level right-looking ----------------
  // inner step (128 columns): diagonal block, panel, update of the remaining columns of the OUTER block only;
  // after OB inner steps one trailing update with K = OB*128 (C tiles read/written n/512 times, not n/128).
  cudaStream_t sA = ctx->stream, sB = ctx->stream2, sP = ctx->stream3;
  bool pending_trail = false;
  bool prev_split = false;   // the previous bulk update was launched in two parts: the chain only waits for the first (ev_trail_a)
  bool p1b_pending = false;  // the panel stream still updates this outer block's columns below its first diagonal tile
  for (int K0 = 0; K0 < nb; K0 += OB) {
    const int Kend = (K0 + OB < nb) ? K0 + OB : nb;
    bool sp_pending = false;  // work of this outer block is still queued on the panel stream
    for (int k = K0; k < Kend; ++k) {
      if (diag2)
        GDCA_CUDA(ctx, gdca_launch_prio(diag_block_kernel2, dim3(1), dim3(DT), D2_SMEM, sA, (const double *)blk(A, k, k), np, blk(X, k, k), np, k * NB, (int)n, ctx->dInfo));
      else
        GDCA_CUDA(ctx, gdca_launch_prio(diag_block_kernel, dim3(1), dim3(DT), dsmem, sA, (const double *)blk(A, k, k), np, blk(X, k, k), np, k * NB, (int)n, ctx->dInfo));
      GDCA_LAUNCH_CHECK(ctx);
      const int rem = nb - k - 1;
      if (rem == 0) break;
      if (p1b_pending) {  // the first panel product of the block reads columns the panel stream has just updated
        GDCA_CUDA(ctx, cudaStreamWaitEvent(sA, ctx->ev_p1b, 0));
        p1b_pending = false;
      }
      const int inner_cols = Kend - k - 1;
      if (!ctx->chol_inner_lookahead || inner_cols == 0 || rem < 2) {
        // last step of the outer block (or look-ahead off): the whole panel on the main stream, then the inner update
        if (sp_pending) {
          GDCA_CUDA(ctx, cudaStreamWaitEvent(sA, ctx->ev_u2b, 0));
          sp_pending = false;
        }
        GemmP p{};
        // panel: L[I,k] = A[I,k] * X[k,k]'   (in place)
        p.A = blk(A, k + 1, k); p.lda = np;
        p.B = blk(X, k, k);     p.ldb = np;
        p.C = blk(A, k + 1, k); p.ldc = np;
        p.m = rem * NB; p.n = NB; p.k = NB; p.flags = 0; p.alpha = 1.0; p.beta = 0.0;
        GDCA_TRY((gemm<false, false>(ctx, p, 1, sA)));
        if (inner_cols > 0) {
          // A[I,J] -= L[I,k] L[J,k]'  for k < J < Kend, I >= J
          GemmP t{};
          t.A = blk(A, k + 1, k); t.lda = np;
          t.B = blk(A, k + 1, k); t.ldb = np;
          t.C = blk(A, k + 1, k + 1); t.ldc = np;
          t.m = rem * NB; t.n = inner_cols * NB; t.k = NB; t.flags = G_LOWER_OUT; t.alpha = -1.0; t.beta = 1.0;
          GDCA_TRY((gemm<false, false>(ctx, t, 1, sA)));
        }
        continue;
      }
      // ---- inner look-ahead: the main stream only does what the NEXT diagonal block needs (its panel tile and its own
      // update); the rest of the panel and of the inner update run beside it on the panel stream.
      //   sA: diag(k) | P1: L[k+1,k] | U1: A[k+1,k+1] -= L[k+1,k] L[k+1,k]'            -> diag(k+1) ...
      //   sP:           P2: L[I,k], I >= k+2 | U2a: column k+1, rows >= k+2 | U2b: columns >= k+2 (lower tiles)
      // Hand-shakes: P2 needs X[k,k] (ev_diag); U2a needs L[k+1,k] (ev_p1); P1 of step k needs column k final = U2a of
      // step k-1 (ev_u2a); U1 of step k and U2b of step k-1 both write tile (k+1,k+1) (ev_u2b).
      GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_diag, sA));
      if (sp_pending) GDCA_CUDA(ctx, cudaStreamWaitEvent(sA, ctx->ev_u2a, 0));
      {
        GemmP p{};  // P1
        p.A = blk(A, k + 1, k); p.lda = np;
        p.B = blk(X, k, k);     p.ldb = np;
        p.C = blk(A, k + 1, k); p.ldc = np;
        p.m = NB; p.n = NB; p.k = NB; p.flags = 0; p.alpha = 1.0; p.beta = 0.0;
        GDCA_TRY(gemm_small(ctx, p, sA));
      }
      GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_p1, sA));
      if (sp_pending) GDCA_CUDA(ctx, cudaStreamWaitEvent(sA, ctx->ev_u2b, 0));
      {
        GemmP u{};  // U1
        u.A = blk(A, k + 1, k); u.lda = np;
        u.B = blk(A, k + 1, k); u.ldb = np;
        u.C = blk(A, k + 1, k + 1); u.ldc = np;
        u.m = NB; u.n = NB; u.k = NB; u.flags = 0; u.alpha = -1.0; u.beta = 1.0;
        GDCA_TRY(gemm_small(ctx, u, sA));
      }
      // panel stream
      GDCA_CUDA(ctx, cudaStreamWaitEvent(sP, ctx->ev_diag, 0));
      {
        GemmP p{};  // P2
        p.A = blk(A, k + 2, k); p.lda = np;
        p.B = blk(X, k, k);     p.ldb = np;
        p.C = blk(A, k + 2, k); p.ldc = np;
        p.m = (rem - 1) * NB; p.n = NB; p.k = NB; p.flags = 0; p.alpha = 1.0; p.beta = 0.0;
        GDCA_TRY((gemm<false, false>(ctx, p, 1, sP)));
      }
      GDCA_CUDA(ctx, cudaStreamWaitEvent(sP, ctx->ev_p1, 0));
      {
        GemmP u{};  // U2a: A[I,k+1] -= L[I,k] L[k+1,k]'  for I >= k+2
        u.A = blk(A, k + 2, k); u.lda = np;
        u.B = blk(A, k + 1, k); u.ldb = np;
        u.C = blk(A, k + 2, k + 1); u.ldc = np;
        u.m = (rem - 1) * NB; u.n = NB; u.k = NB; u.flags = 0; u.alpha = -1.0; u.beta = 1.0;
        GDCA_TRY((gemm<false, false>(ctx, u, 1, sP)));
      }
      GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_u2a, sP));
      if (inner_cols > 1) {
        GemmP u{};  // U2b: A[I,J] -= L[I,k] L[J,k]'  for k+2 <= J < Kend, I >= J
        u.A = blk(A, k + 2, k); u.lda = np;
        u.B = blk(A, k + 2, k); u.ldb = np;
        u.C = blk(A, k + 2, k + 2); u.ldc = np;
        u.m = (rem - 1) * NB; u.n = (inner_cols - 1) * NB; u.k = NB; u.flags = G_LOWER_OUT; u.alpha = -1.0; u.beta = 1.0;
        GDCA_TRY((gemm<false, false>(ctx, u, 1, sP)));
      }
      GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_u2b, sP));
      sp_pending = true;
    }
    if (sp_pending) GDCA_CUDA(ctx, cudaStreamWaitEvent(sA, ctx->ev_u2b, 0));
    if (N > 1) {
      // columns [K0, Kend) of the factor and the inverted diagonal blocks are final: every member pulls them over NVLink
      // with its copy engine while the leader goes on factorising
      GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_copy, sA));
      const size_t wbytes = (size_t)(Kend - K0) * NB * sizeof(double);
      for (int r = 1; r < N; ++r) {
        gdca_ctx *c = grp[r];
        GDCA_TRY(on_dev(c));
        GDCA_CUDA(ctx, cudaStreamWaitEvent(c->stream_copy, ctx->ev_copy, 0));
        GDCA_CUDA(ctx, cudaMemcpy2DAsync(blk(c->dC, K0, K0), (size_t)np * sizeof(double), blk(A, K0, K0), (size_t)np * sizeof(double), wbytes,
                                         (size_t)(nb - K0) * NB, cudaMemcpyDefault, c->stream_copy));
        GDCA_CUDA(ctx, cudaMemcpy2DAsync(blk(c->dX, K0, K0), (size_t)np * sizeof(double), blk(X, K0, K0), (size_t)np * sizeof(double), wbytes,
                                         (size_t)(Kend - K0) * NB, cudaMemcpyDefault, c->stream_copy));
        if (share) {  // the member's compute stream applies this panel to its own block columns
          GDCA_CUDA(ctx, cudaEventRecord(c->ev_copy, c->stream_copy));
          GDCA_CUDA(ctx, cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
        }
      }
      GDCA_TRY(on_dev(ctx));
    }
    const int rem = nb - Kend;
    if (rem > 0) {
      // Trailing update A[I,J] -= L[I,K0:Kend] L[J,K0:Kend]' (I >= J >= Kend), split for a one-panel look-ahead:
      //   part 1 (main stream): only the columns of the NEXT outer panel [Kend, Kn) -- the factorisation goes on;
      //   part 2 (helper stream, low priority): all columns >= Kn, overlapping the next panel's serial chain.
      const int Kn = (Kend + OB < nb) ? Kend + OB : nb;
      GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_fact, sA));          // panel [K0,Kend) is final
      // columns [Kend, Kn) carry the bulk update of panel K0-OB: all of it, or -- when it was launched in two parts -- its first part
      cudaEvent_t ev_prev = prev_split ? ctx->ev_trail_a : ctx->ev_trail;
      if (pending_trail) GDCA_CUDA(ctx, cudaStreamWaitEvent(sA, ev_prev, 0));
      if (share && rem < OZ_MIN_REM) {
        // The last few block columns: every owner hands its columns back and the leader finishes alone, on the same engines
        // as a single-GPU run (bit-identical results for every group size).
        for (int r = 1; r < N; ++r) {
          gdca_ctx *c = grp[r];
          GDCA_TRY(on_dev(c));
          GDCA_CUDA(ctx, cudaEventRecord(c->ev_upd, c->stream));
          GDCA_CUDA(ctx, cudaStreamWaitEvent(c->stream_copy, c->ev_upd, 0));
          for (int ob = r; ob * OB < nb; ob += N)
            if (ob * OB >= Kend) GDCA_TRY(copy_cols(c, c->stream_copy, A, c->dC, ob * OB, std::min(OB, nb - ob * OB)));
          GDCA_CUDA(ctx, cudaEventRecord(c->ev_sent, c->stream_copy));
        }
        GDCA_TRY(on_dev(ctx));
        for (int r = 1; r < N; ++r)
          for (cudaStream_t st : {sA, sB, sP}) GDCA_CUDA(ctx, cudaStreamWaitEvent(st, grp[r]->ev_sent, 0));
        share = false;
      }
      if (oz && rem >= OZ_MIN_REM) {
        const int sidx = K0 / OB;  // this step's outer block; the chain goes on with block sidx + 1
        if (share && sidx >= 1 && owner(sidx + 1) != 0) {
          // block column sidx + 1 carries the updates of all earlier panels on its owner: it has been shipped one step ahead
          GDCA_CUDA(ctx, cudaStreamWaitEvent(sA, grp[owner(sidx + 1)]->ev_sent, 0));
          GDCA_CUDA(ctx, cudaStreamWaitEvent(sP, grp[owner(sidx + 1)]->ev_sent, 0));
        }
        // The chain only needs the NEXT DIAGONAL TILE before it can go on: the 128 panel rows that produce it are sliced and
        // multiplied on the main stream (2 tiles), the next diagonal block starts factorising right behind them.  Everything
        // else -- slicing the whole panel once, the other columns of the next panel, the bulk -- runs beside it: the panel
        // stream updates the next panel's columns below that tile (the first panel product of the next block waits for it),
        // the low-priority stream the bulk.  (The waits on the previous bulk update come first: it wrote all of this.)
        const int kk = (Kend - K0) * NB;
        gdca_oz_operand Pd{};
        GDCA_TRY(gdca_oz_slice(ctx, sA, blk(A, Kend, K0), np, 0, false, NB, kk, 1, NB, ctx->dDigP, ctx->dScaleP, &Pd));
        GDCA_TRY(gdca_oz_gemm(ctx, sA, Pd, Pd, blk(A, Kend, Kend), np, 0, NB, NB, kk, 1, GDCA_OZ_LOWER_OUT, -1.0, 1, 0));
        GDCA_CUDA(ctx, cudaStreamWaitEvent(sP, ctx->ev_fact, 0));
        if (pending_trail) GDCA_CUDA(ctx, cudaStreamWaitEvent(sP, ev_prev, 0));
        // One GPU: the bulk update goes out in two parts -- first the block columns of the outer block after next, which is all the
        // chain waits for at the next boundary, then the rest -- so the chain never stalls behind a whole bulk update.  The panel
        // digits alternate between two buffers: the second part of the previous bulk may still be reading the other one (it is
        // complete once the first part of THIS bulk's predecessor has been waited for: the helper stream runs them in order).
        const bool split = !share;  // also on the leader of a device group while the factorisation itself is not shared
        int8_t *pdig = (split && (sidx & 1)) ? ctx->dDigA : ctx->dDigB;
        double *pscale = (split && (sidx & 1)) ? ctx->dScaleA : ctx->dScaleB;
        gdca_oz_operand P{};
        GDCA_TRY(gdca_oz_slice(ctx, sP, blk(A, Kend, K0), np, 0, false, rem * NB, kk, 1, rem * NB, pdig, pscale, &P));
        GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_sliced, sP));
        if (rem > 1) {
          gdca_oz_operand P1 = P;   // rows from block Kend + 1 on
          P1.dig += (long long)NB * P.pitch;
          P1.scale += NB;
          P1.rows_total = P1.rows_b = (long long)(rem - 1) * NB;
          gdca_oz_shard sh{};
          sh.own_mod = 1;
          sh.m_off = NB;  // lower-triangular test against the panel's own origin
          GDCA_TRY(gdca_oz_gemm(ctx, sP, P1, P, blk(A, Kend + 1, Kend), np, 0, (rem - 1) * NB, (Kn - Kend) * NB, kk, 1, GDCA_OZ_LOWER_OUT,
                                -1.0, 1, 0, &sh));
        }
        GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_p1b, sP));
        p1b_pending = true;
        ctx->oz_fp64_flop += 2.0 * (double)kk * NB * NB * ((double)(Kn - Kend) * rem - 0.5 * (Kn - Kend) * (Kn - Kend - 1));
        const int rem2 = nb - Kn;
        if (rem2 > 0) {
          GDCA_CUDA(ctx, cudaStreamWaitEvent(sB, ctx->ev_sliced, 0));
          gdca_oz_operand P2 = P;
          P2.dig += (long long)(Kn - Kend) * NB * P.pitch;
          P2.scale += (long long)(Kn - Kend) * NB;
          P2.rows_total = P2.rows_b = (long long)rem2 * NB;
          gdca_oz_shard cf{};  // shared factorisation: only the block columns this member owns
          cf.own_mod = 1;
          cf.col_mod = share ? N : 1;
          cf.col_rank = 0;
          cf.col_unit0 = Kn * (NB / 64);
          cf.col_per = OB * (NB / 64);
          const int Kn2 = (Kn + OB < nb) ? Kn + OB : nb;
          const int bulk_tpc = (cap && ctx->ozaki_tpc > 0) ? -(ctx->num_sms - 40) : ctx->ozaki_tpc;  // persistent grid, 40 SMs left to the chain and the panel products (measured: 9.24 ms against 9.85 / 9.99 ms with 16 / 64)
          if (split && Kn2 < nb) {
            // part a: block columns [Kn, Kn2) (rows >= Kn); part b: everything from Kn2 on
            GDCA_TRY(gdca_oz_gemm(ctx, sB, P2, P2, blk(A, Kn, Kn), np, 0, rem2 * NB, (Kn2 - Kn) * NB, kk, 1, GDCA_OZ_LOWER_OUT, -1.0, 1, bulk_tpc));
            GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_trail_a, sB));
            gdca_oz_operand P3 = P2;
            P3.dig += (long long)(Kn2 - Kn) * NB * P.pitch;
            P3.scale += (long long)(Kn2 - Kn) * NB;
            P3.rows_total = P3.rows_b = (long long)(nb - Kn2) * NB;
            GDCA_TRY(gdca_oz_gemm(ctx, sB, P3, P3, blk(A, Kn2, Kn2), np, 0, (nb - Kn2) * NB, (nb - Kn2) * NB, kk, 1, GDCA_OZ_LOWER_OUT, -1.0, 1, bulk_tpc));
            prev_split = true;
          } else {
            GDCA_TRY(gdca_oz_gemm(ctx, sB, P2, P2, blk(A, Kn, Kn), np, 0, rem2 * NB, rem2 * NB, kk, 1, GDCA_OZ_LOWER_OUT, -1.0, 1,
                                  bulk_tpc, share ? &cf : nullptr));
            prev_split = false;
          }
          ctx->oz_fp64_flop += 2.0 * (double)kk * NB * NB * (0.5 * (double)rem2 * (rem2 + 1));
          GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_trail, sB));
          pending_trail = true;
          if (share) {
            // every other member: slice the panel rows it received, update the block column the leader needs next FIRST and
            // ship it (copy engine, NVLink), then the rest of its columns
            const int wu = std::min(OB, nb - Kn);   // blocks of outer block sidx + 2
            for (int r = 1; r < N; ++r) {
              gdca_ctx *c = grp[r];
              GDCA_TRY(on_dev(c));
              gdca_oz_operand Pm{};
              GDCA_TRY(mtry(c, gdca_oz_slice(c, c->stream, blk(c->dC, Kn, K0), np, 0, false, rem2 * NB, kk, 1, rem2 * NB, c->dDigB, c->dScaleB, &Pm)));
              if (owner(sidx + 2) == r) {
                GDCA_TRY(mtry(c, gdca_oz_gemm(c, c->stream, Pm, Pm, blk(c->dC, Kn, Kn), np, 0, rem2 * NB, wu * NB, kk, 1, GDCA_OZ_LOWER_OUT, -1.0, 1, 0)));
                GDCA_CUDA(ctx, cudaEventRecord(c->ev_upd, c->stream));
                GDCA_CUDA(ctx, cudaStreamWaitEvent(c->stream_copy, c->ev_upd, 0));
                GDCA_TRY(copy_cols(c, c->stream_copy, A, c->dC, Kn, wu));
                GDCA_CUDA(ctx, cudaEventRecord(c->ev_sent, c->stream_copy));
              }
              const int Kr = Kn + OB;  // the columns behind that block
              if (Kr < nb) {
                gdca_oz_operand Pr = Pm;
                Pr.dig += (long long)(Kr - Kn) * NB * Pm.pitch;
                Pr.scale += (long long)(Kr - Kn) * NB;
                Pr.rows_total = Pr.rows_b = (long long)(nb - Kr) * NB;
                gdca_oz_shard cm = cf;
                cm.col_rank = r;
                cm.col_unit0 = Kr * (NB / 64);
                GDCA_TRY(mtry(c, gdca_oz_gemm(c, c->stream, Pr, Pr, blk(c->dC, Kr, Kr), np, 0, (nb - Kr) * NB, (nb - Kr) * NB, kk, 1,
                                              GDCA_OZ_LOWER_OUT, -1.0, 1, 0, &cm)));
              }
            }
            GDCA_TRY(on_dev(ctx));
          }
        }
        continue;
      }
      GemmP t{};
      t.A = blk(A, Kend, K0); t.lda = np;
      t.B = blk(A, Kend, K0); t.ldb = np;
      t.C = blk(A, Kend, Kend); t.ldc = np;
      t.m = rem * NB; t.n = (Kn - Kend) * NB; t.k = (Kend - K0) * NB; t.flags = G_LOWER_OUT; t.alpha = -1.0; t.beta = 1.0;
      GDCA_TRY((gemm<false, false>(ctx, t, 1, sA)));
      const int rem2 = nb - Kn;
      if (rem2 > 0) {
        GDCA_CUDA(ctx, cudaStreamWaitEvent(sB, ctx->ev_fact, 0));
        GemmP u{};
        u.A = blk(A, Kn, K0); u.lda = np;
        u.B = blk(A, Kn, K0); u.ldb = np;
        u.C = blk(A, Kn, Kn); u.ldc = np;
        u.m = rem2 * NB; u.n = rem2 * NB; u.k = (Kend - K0) * NB; u.flags = G_LOWER_OUT; u.alpha = -1.0; u.beta = 1.0;
        GDCA_TRY((gemm<false, false>(ctx, u, 1, sB)));
        GDCA_CUDA(ctx, cudaEventRecord(ctx->ev_trail, sB));
        pending_trail = true;
        prev_split = false;
      }
    }
  }
  if (pending_trail) GDCA_CUDA(ctx, cudaStreamWaitEvent(sA, ctx->ev_trail, 0));
  if (p1b_pending) GDCA_CUDA(ctx, cudaStreamWaitEvent(sA, ctx->ev_p1b, 0));
  if (cap)
    GDCA_TRY(cap->split());
  else if (ctx->ev[GDCA_EV_POTRF])
    GDCA_CUDA(ctx, cudaEventRecord(ctx->ev[GDCA_EV_POTRF], sA));
  for (int r = 1; r < N; ++r) {  // the members' compute streams continue once their copy of the factor is complete
    gdca_ctx *c = grp[r];
    GDCA_TRY(on_dev(c));
    GDCA_CUDA(ctx, cudaEventRecord(c->ev_copy, c->stream_copy));
    GDCA_CUDA(ctx, cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
  }
  GDCA_TRY(on_dev(ctx));
  // every member's stream waits for every member's stream (event records + stream waits; the host never blocks)
  auto barrier = [&]() -> int32_t {
    for (int r = 0; r < N; ++r) {
      GDCA_TRY(on_dev(grp[r]));
      GDCA_CUDA(ctx, cudaEventRecord(grp[r]->ev_group, grp[r]->stream));
    }
    for (int r = 0; r < N; ++r) {
      GDCA_TRY(on_dev(grp[r]));
      for (int q = 0; q < N; ++q)
        if (q != r) GDCA_CUDA(ctx, cudaStreamWaitEvent(grp[r]->stream, grp[q]->ev_group, 0));
    }
    return on_dev(ctx);
  };

  // ---------------- trtri by recursive doubling ----------------
  for (int h = 1; h < nb; h *= 2) {
    // groups start at g = 0, 2h, 4h, ...; top = [g, g+h), bottom = [g+h, min(g+2h, nb))
    const int ngroups_full = nb / (2 * h);            // groups whose bottom part is complete
    const int rest = nb - ngroups_full * 2 * h;       // blocks left after the full groups
    const long long gstride = (long long)2 * h * NB * (np + 1);  // along the diagonal
    // A level is SHARED in a device group when every member gets a whole number of 64-column tiles of each X21 block: member r
    // computes the columns [r w, (r+1) w) of T and of X21 and stores its X21 columns into EVERY member's X (the all-gather is
    // fused into the GEMM epilogue; a barrier closes the level).  Otherwise every member computes the level on its replica --
    // same kernels, same inputs, same bits.
    const bool shared = N > 1 && oz && h >= OZ_MIN_H && (h * NB) % (64 * N) == 0;
    auto level = [&](gdca_ctx *c, int rnk, int g0, int mb, int batch) -> int32_t {  // mb = bottom blocks
      double *A = c->dC, *X = c->dX, *T = c->dT;  // this member's replicas
      if (oz && h >= OZ_MIN_H) {
        cudaStream_t st = c->stream;
        const int w = shared ? h * NB / N : h * NB, col0 = shared ? rnk * w : 0;
        gdca_oz_shard sh{};
        sh.n_off = col0;
        sh.own_mod = 1;
        gdca_oz_operand oa{}, ob{};
        // T[bottom, top] = L[bottom, top] * X[top, top]: rows of L21 against the COLUMNS of the lower-triangular X11 (k >= n0)
        GDCA_TRY(mtry(c, gdca_oz_slice(c, st, blk(A, g0 + h, g0), np, gstride, false, mb * NB, h * NB, batch, mb * NB, c->dDigA, c->dScaleA, &oa)));
        // (X11 is lower triangular: column col0 + r is zero above row col0 + r, so the zero-skipping of the slice, which goes by the
        // local column r, stays on the safe side for any col0; the product below starts at k = 128 floor((col0 + n0) / 128))
        GDCA_TRY(mtry(c, gdca_oz_slice(c, st, blk(X, g0, g0) + col0, np, gstride, true, w, h * NB, batch, w, c->dDigB, c->dScaleB, &ob, /*lower_only=*/true)));
        GDCA_TRY(mtry(c, gdca_oz_gemm(c, st, oa, ob, blk(T, g0 + h, g0) + col0, np, gstride, mb * NB, w, h * NB, batch, GDCA_OZ_KBEG_N, 1.0, 0, 0, &sh)));
        // X[bottom, top] = - X[bottom, bottom] * T[bottom, top]: rows of the lower-triangular X22 (k < m0 + 128) against columns of T
        GDCA_TRY(mtry(c, gdca_oz_slice(c, st, blk(X, g0 + h, g0 + h), np, gstride, false, mb * NB, mb * NB, batch, mb * NB, c->dDigA, c->dScaleA, &oa, /*lower_only=*/true)));
        GDCA_TRY(mtry(c, gdca_oz_slice(c, st, blk(T, g0 + h, g0) + col0, np, gstride, true, w, mb * NB, batch, w, c->dDigB, c->dScaleB, &ob)));
        sh.n_off = 0;
        if (shared) {
          sh.npeer = N;
          for (int q = 0; q < N; ++q) sh.peer_off[q] = (long long)(reinterpret_cast<char *>(grp[q]->dX) - reinterpret_cast<char *>(X));
        }
        GDCA_TRY(mtry(c, gdca_oz_gemm(c, st, oa, ob, blk(X, g0 + h, g0) + col0, np, gstride, mb * NB, w, mb * NB, batch, GDCA_OZ_KEND_M, -1.0, 0, 0, &sh)));
        if (c == ctx)
          ctx->oz_fp64_flop += (double)batch * 2.0 * NB * NB * NB * ((double)mb * h * (h + 1) / 2 + (double)h * mb * (mb + 1) / 2);
        return GDCA_OK;
      }
      GemmP a{};
      // T[bottom, top] = L[bottom, top] * X[top, top]         (B as [k][n], lower triangular: k >= n0)
      a.A = blk(A, g0 + h, g0); a.lda = np; a.strideA = gstride;
      a.B = blk(X, g0, g0);     a.ldb = np; a.strideB = gstride;
      a.C = blk(T, g0 + h, g0); a.ldc = np; a.strideC = gstride;
      a.m = mb * NB; a.n = h * NB; a.k = h * NB; a.flags = G_KBEG_N; a.alpha = 1.0; a.beta = 0.0;
      GDCA_TRY(mtry(c, (gemm<false, true>(c, a, batch))));
      GemmP b{};
      // X[bottom, top] = - X[bottom, bottom] * T[bottom, top] (A lower triangular: k < m0 + NB)
      b.A = blk(X, g0 + h, g0 + h); b.lda = np; b.strideA = gstride;
      b.B = blk(T, g0 + h, g0);     b.ldb = np; b.strideB = gstride;
      b.C = blk(X, g0 + h, g0);     b.ldc = np; b.strideC = gstride;
      b.m = mb * NB; b.n = h * NB; b.k = mb * NB; b.flags = G_KEND_M; b.alpha = -1.0; b.beta = 0.0;
      GDCA_TRY(mtry(c, (gemm<false, true>(c, b, batch))));
      return GDCA_OK;
    };
    for (int r = 0; r < N; ++r) {
      gdca_ctx *c = N > 1 ? grp[r] : ctx;
      GDCA_TRY(on_dev(c));
      if (ngroups_full > 0) GDCA_TRY(level(c, r, 0, h, ngroups_full));
      if (rest > h) GDCA_TRY(level(c, r, ngroups_full * 2 * h, rest - h, 1));
    }
    if (shared) GDCA_TRY(barrier());
  }
  GDCA_TRY(on_dev(ctx));

  // ---------------- lauum: mJ = X' X (lower tiles), then mirror ----------------
  if (oz) {
    // operand rows = columns of X: ONE transposed slice serves both sides; k >= m0 (X is lower triangular).  In a device
    // group member r computes the row tiles im = r (mod N) from its replica of X and stores them straight into the leader's mJ.
    for (int r = 0; r < N; ++r) {
      gdca_ctx *c = N > 1 ? grp[r] : ctx;
      GDCA_TRY(on_dev(c));
      gdca_oz_shard sh{};
      sh.own_mod = N;
      sh.own_rank = r;
      gdca_oz_operand ox{};
      GDCA_TRY(mtry(c, gdca_oz_slice(c, c->stream, c->dX, np, 0, true, (int)np, (int)np, 1, np, c->dDigA, c->dScaleA, &ox, /*lower_only=*/true)));
      GDCA_TRY(mtry(c, gdca_oz_gemm(c, c->stream, ox, ox, J, np, 0, (int)np, (int)np, (int)np, 1, GDCA_OZ_LOWER_OUT | GDCA_OZ_KBEG_M, 1.0, 0, 0, &sh)));
    }
    if (N > 1) GDCA_TRY(barrier());
    GDCA_TRY(on_dev(ctx));
    for (int r = 1; r < N; ++r) ctx->oz_int8_ops += grp[r]->oz_int8_ops;
    ctx->oz_fp64_flop += (double)np * np * np / 3.0;
    const unsigned nt = (unsigned)((np + 31) / 32);
    mirror_lower_kernel<<<dim3(nt, nt), 256, 0, ctx->stream>>>(J, np, np);
    GDCA_LAUNCH_CHECK(ctx);
  } else {
    GemmP l{};
    l.A = X; l.lda = np;
    l.B = X; l.ldb = np;
    l.C = J; l.ldc = np;
    l.m = (int)np; l.n = (int)np; l.k = (int)np; l.flags = G_LOWER_OUT | G_KBEG_M; l.alpha = 1.0; l.beta = 0.0;
    GDCA_TRY((gemm<true, true>(ctx, l, 1)));
    const unsigned nt = (unsigned)((np + 31) / 32);
    mirror_lower_kernel<<<dim3(nt, nt), 256, 0, ctx->stream>>>(J, np, np);
    GDCA_LAUNCH_CHECK(ctx);
  }
  return GDCA_OK;
}

namespace {
struct InvGraphKey {
  long long np, n;
  int oz, diag_blocked, lookahead, tpc;
  const void *ptr[12];
  bool operator==(const InvGraphKey &o) const { return memcmp(this, &o, sizeof *this) == 0; }
};
struct InvGraph {
  InvGraphKey key;
  cudaGraphExec_t exec[2] = {nullptr, nullptr};
  double int8_ops = 0.0, fp64_flop = 0.0;
  long long launches = 0;
  bool ozaki = false;
};
void inv_graph_destroy(InvGraph *g) {
  if (!g) return;
  for (cudaGraphExec_t e : g->exec)
    if (e) cudaGraphExecDestroy(e);
  delete g;
}
}  // namespace

void gdca_k_inverse_release(gdca_ctx *ctx) {
  inv_graph_destroy(static_cast<InvGraph *>(ctx->inv_graph));
  ctx->inv_graph = nullptr;
}

int32_t gdca_k_inverse(gdca_ctx *ctx) {
  if (!ctx->have_cov) return gdca_fail(ctx, GDCA_ERR_STATE, "inverse: covariance not computed");
  const long long np = ctx->npad;
  // One GPU: the launch sequence depends only on the shape and on the buffer addresses, so it is captured once and replayed as
  // two CUDA graphs -- the serial chain of the factorisation (79 x diag -> panel tile -> diagonal update at n = 10 000) is bound by
  // launch and cross-stream hand-shake latency, which a graph launch takes off the host.  (A device group keeps direct launches.)
  const bool want_graph = ctx->inv_graph_mode != 0 && ctx->group_size <= 1 && !ctx->leader;
  bool done = false;
  if (want_graph) {
    const bool oz = ctx->ozaki_mode != 0 && np / NB >= 16;
    // every buffer the captured launches refer to exists (and keeps its address) before the capture starts
    GDCA_TRY(gdca_reserve(ctx, ctx->dX, ctx->capX, (size_t)np * np));
    GDCA_TRY(gdca_reserve(ctx, ctx->dT, ctx->capT, (size_t)np * np));
    GDCA_TRY(gdca_reserve(ctx, ctx->dmJ, ctx->capmJ, (size_t)np * np));
    if (oz) {
      GDCA_TRY(gdca_reserve(ctx, ctx->dDigA, ctx->capDigA, (size_t)np * np * 8));
      GDCA_TRY(gdca_reserve(ctx, ctx->dDigB, ctx->capDigB, (size_t)np * np * 8));
      GDCA_TRY(gdca_reserve(ctx, ctx->dScaleA, ctx->capScaleA, (size_t)np));
      GDCA_TRY(gdca_reserve(ctx, ctx->dScaleB, ctx->capScaleB, (size_t)np));
      GDCA_TRY(gdca_reserve(ctx, ctx->dDigP, ctx->capDigP, (size_t)NB * 8 * 1024));
      GDCA_TRY(gdca_reserve(ctx, ctx->dScaleP, ctx->capScaleP, (size_t)NB));
      GDCA_TRY(gdca_reserve(ctx, ctx->dOzMax, ctx->capOzMax, (size_t)np));
    }
    InvGraphKey key;
    memset(&key, 0, sizeof key);
    key.np = np;
    key.n = ctx->n;
    key.oz = oz ? 1 : 0;
    key.diag_blocked = ctx->diag_blocked;
    key.lookahead = ctx->chol_inner_lookahead;
    key.tpc = ctx->ozaki_tpc;
    const void *ptrs[12] = {ctx->dC, ctx->dX, ctx->dT, ctx->dmJ, ctx->dDigA, ctx->dDigB, ctx->dScaleA, ctx->dScaleB, ctx->dDigP, ctx->dScaleP,
                            ctx->dOzMax, ctx->dInfo};
    memcpy(key.ptr, ptrs, sizeof ptrs);
    InvGraph *g = static_cast<InvGraph *>(ctx->inv_graph);
    if (g && !(g->key == key)) {
      inv_graph_destroy(g);
      g = nullptr;
      ctx->inv_graph = nullptr;
    }
    if (!g) {
      InvCapture cap;
      cap.ctx = ctx;
      const long long launches0 = ctx->launches;
      int32_t st = GDCA_ERR_CUDA;
      if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
        st = inverse_enqueue(ctx, &cap);
        cudaGraph_t last = nullptr;
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &last);  // also ends a capture the body left behind on an error
        if (st == GDCA_OK && e == cudaSuccess && last && cap.parts == 1 && !cap.failed) {
          cap.graph[1] = last;
          g = new InvGraph();
          g->key = key;
          g->int8_ops = ctx->oz_int8_ops;
          g->fp64_flop = ctx->oz_fp64_flop;
          g->ozaki = ctx->last_inverse_ozaki;
          g->launches = ctx->launches - launches0;
          for (int i = 0; i < 2 && g; ++i)
            if (cudaGraphInstantiate(&g->exec[i], cap.graph[i], 0) != cudaSuccess) {
              inv_graph_destroy(g);
              g = nullptr;
            }
        } else if (last) {
          cudaGraphDestroy(last);
          last = nullptr;
        }
        for (cudaGraph_t &gr : cap.graph)
          if (gr) cudaGraphDestroy(gr);
      }
      cudaGetLastError();
      ctx->launches = launches0;  // nothing has run yet
      if (!g) {
        // capture is not possible here (or the body failed): run the launches directly from now on
        ctx->inv_graph_mode = 0;
        ctx->err.clear();
      }
      ctx->inv_graph = g;
    }
    if (g) {
      GDCA_CUDA(ctx, cudaGraphLaunch(g->exec[0], ctx->stream));
      if (ctx->ev[GDCA_EV_POTRF]) GDCA_CUDA(ctx, cudaEventRecord(ctx->ev[GDCA_EV_POTRF], ctx->stream));
      GDCA_CUDA(ctx, cudaGraphLaunch(g->exec[1], ctx->stream));
      ctx->oz_int8_ops = g->int8_ops;
      ctx->oz_fp64_flop = g->fp64_flop;
      ctx->last_inverse_ozaki = g->ozaki;
      ctx->last_inverse_shared = false;
      ctx->launches += g->launches;
      done = true;
    }
  }
  if (!done) GDCA_TRY(inverse_enqueue(ctx, nullptr));
  int info = 0;
  GDCA_CUDA(ctx, cudaMemcpyAsync(&info, ctx->dInfo, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stats.posdef_info = info;
  ctx->have_cov = false;  // dC now holds the factor, not C
  if (info != 0) {
    char b[128];
    snprintf(b, sizeof b, "matrix is not positive definite; Cholesky factorization failed (info=%d)", info);
    ctx->err = b;
    return GDCA_ERR_NOT_SPD;
  }
  ctx->have_inv = true;
  return GDCA_OK;
}

int32_t gdca_k_inverse_group(gdca_ctx *lead) { return gdca_k_inverse(lead); }

// ---------------------------------------------------------------------------------------------- test hook
// C[m x n] = beta C + alpha opA opB^T on one of the two FP64 GEMM engines, host buffers (tests/test_gpu_ozaki.py).
extern "C" int32_t gdca_test_fp64_gemm(gdca_ctx *ctx, int32_t engine, const double *A, int32_t a_cols, const double *B, int32_t b_cols,
                                       double *C, int64_t m, int64_t n, int64_t k, int32_t flags, double alpha, double beta) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  if (!A || !B || !C || m < 128 || n < 128 || k < 128 || m % 128 || n % 128 || k % 128)
    return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "test_fp64_gemm: m, n, k must be positive multiples of 128");
  if (engine == 1 && beta != 0.0 && beta != 1.0) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "test_fp64_gemm: engine 1 takes beta 0 or 1");
  GDCA_CUDA(ctx, cudaSetDevice(ctx->device));
  double *dA = nullptr, *dB = nullptr, *dCm = nullptr;
  int8_t *dgA = nullptr, *dgB = nullptr;
  double *scA = nullptr, *scB = nullptr;
  int32_t st = GDCA_OK;
  auto body = [&]() -> int32_t {
    GDCA_CUDA(ctx, cudaMalloc((void **)&dA, (size_t)m * k * 8));
    GDCA_CUDA(ctx, cudaMalloc((void **)&dB, (size_t)n * k * 8));
    GDCA_CUDA(ctx, cudaMalloc((void **)&dCm, (size_t)m * n * 8));
    GDCA_CUDA(ctx, cudaMemcpyAsync(dA, A, (size_t)m * k * 8, cudaMemcpyHostToDevice, ctx->stream));
    GDCA_CUDA(ctx, cudaMemcpyAsync(dB, B, (size_t)n * k * 8, cudaMemcpyHostToDevice, ctx->stream));
    GDCA_CUDA(ctx, cudaMemcpyAsync(dCm, C, (size_t)m * n * 8, cudaMemcpyHostToDevice, ctx->stream));
    GDCA_CUDA(ctx, cudaMemsetAsync(ctx->dInfo, 0, sizeof(int), ctx->stream));
    if (engine == 0) {
      GemmP p{};
      p.A = dA; p.lda = a_cols ? m : k;
      p.B = dB; p.ldb = b_cols ? n : k;
      p.C = dCm; p.ldc = n;
      p.m = (int)m; p.n = (int)n; p.k = (int)k; p.flags = flags; p.alpha = alpha; p.beta = beta;
      if (!a_cols && !b_cols) GDCA_TRY((gemm<false, false>(ctx, p, 1)));
      if (!a_cols && b_cols) GDCA_TRY((gemm<false, true>(ctx, p, 1)));
      if (a_cols && b_cols) GDCA_TRY((gemm<true, true>(ctx, p, 1)));
      if (a_cols && !b_cols) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "test_fp64_gemm: engine 0 has no (A as [k][m], B as [n][k]) kernel");
    } else {
      GDCA_CUDA(ctx, cudaMalloc((void **)&dgA, (size_t)m * k * 8));
      GDCA_CUDA(ctx, cudaMalloc((void **)&dgB, (size_t)n * k * 8));
      GDCA_CUDA(ctx, cudaMalloc((void **)&scA, (size_t)m * 8));
      GDCA_CUDA(ctx, cudaMalloc((void **)&scB, (size_t)n * 8));
      GDCA_CUDA(ctx, cudaMemsetAsync(dgA, 0, (size_t)m * k * 8, ctx->stream));
      GDCA_CUDA(ctx, cudaMemsetAsync(dgB, 0, (size_t)n * k * 8, ctx->stream));
      gdca_oz_operand oa{}, ob{};
      GDCA_TRY(gdca_oz_slice(ctx, ctx->stream, dA, a_cols ? m : k, 0, a_cols != 0, (int)m, (int)k, 1, m, dgA, scA, &oa));
      GDCA_TRY(gdca_oz_slice(ctx, ctx->stream, dB, b_cols ? n : k, 0, b_cols != 0, (int)n, (int)k, 1, n, dgB, scB, &ob));
      GDCA_TRY(gdca_oz_gemm(ctx, ctx->stream, oa, ob, dCm, n, 0, (int)m, (int)n, (int)k, 1, flags, alpha, beta != 0.0, 0));
    }
    GDCA_CUDA(ctx, cudaMemcpyAsync(C, dCm, (size_t)m * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GDCA_OK;
  };
  st = body();
  cudaStreamSynchronize(ctx->stream);
  for (void *q : {(void *)dA, (void *)dB, (void *)dCm, (void *)dgA, (void *)dgB, (void *)scA, (void *)scB})
    if (q) cudaFree(q);
  return st;
}
