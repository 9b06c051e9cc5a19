// score.cu -- K6/K7: per site-pair block scores from mJ = inv(C).
//
//   K6 compute_FN       (DCAUtils, un-vendored; reference call site src/GaussDCA.jl:39)
//        B = mJ[block i, block j] (s x s);  K = B - rowmean - colmean + mean  (means over the s x s block)
//        FN[i,j] = FN[j,i] = ||K||_F ;  zero diagonal
//   K7 compute_DI_gauss (DCAUtils, un-vendored; reference call site src/GaussDCA.jl:37)
//        V = (sqrt(C_ii) mJ_ij sqrt(C_jj)) (.)';  DI = s/2 log(1/2) + 1/2 sum_k log(1 + sqrt(1 + 4 lambda_k(V)))
//      evaluated through the equivalent form lambda_k = sigma_k(Lc_i' mJ_ij Lc_j)^2 with C_ii = Lc_i Lc_i'
//      (similarity transform; removes the per-site matrix square root).  Singular values come from a
//      one-sided Jacobi iteration held entirely in shared memory, three (i,j) blocks per warp.
//
// One warp per (i<j) block; a CTA is 4 warps sharing site i.  FN is HBM-bound (3200 B read per
// block at s = 20); DI adds ~3e5 FP64 flop per block on the plain FP64 pipe.
#include "gdca_internal.cuh"

namespace {

constexpr int SW = 4;  // warps (blocks j) per CTA

// One warp per (i<j) block.  The block is read ONCE from HBM with 16-byte loads issued back to back (all of a
// lane's loads are in flight before the first use), parked in shared memory, and the gauge is applied in a
// second pass over shared memory (two-pass for accuracy: ||K|| can be far below ||B||).
__global__ void __launch_bounds__(SW * 32) fn_kernel(const double *__restrict__ mJ, long long ld, int L, int s,
                                                     double *__restrict__ S) {
  extern __shared__ double sm[];  // [SW][s*s + 2*s]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.y, j = blockIdx.x * SW + warp;
  if (j <= i || j >= L) return;
  double *Bk = sm + (size_t)warp * (s * s + 2 * s + 2);
  double *rs = Bk + s * s, *cs = rs + s;
  const int ss = s * s;
  double tot = 0.0;
  const double *base = mJ + ((long long)i * s) * ld + (long long)j * s;
  if ((s & 1) == 0) {
    // rows of s doubles are 16-byte aligned (ld and j*s are even): s/2 double2 per row
    const int hs = s >> 1, nv = s * hs;
    constexpr int MAXV = 8;  // 8 x 32 double2 in flight per round (one round at s = 20, two at s = 30)
    for (int e0 = 0; e0 < nv; e0 += MAXV * 32) {
      double2 v[MAXV];
#pragma unroll
      for (int u = 0; u < MAXV; ++u) {
        const int e = e0 + lane + 32 * u;
        if (e < nv) {
          const int a = e / hs, b2 = e - a * hs;
          v[u] = *reinterpret_cast<const double2 *>(base + (long long)a * ld + 2 * b2);
        }
      }
#pragma unroll
      for (int u = 0; u < MAXV; ++u) {
        const int e = e0 + lane + 32 * u;
        if (e < nv) {
          const int a = e / hs, b2 = e - a * hs;
          Bk[a * s + 2 * b2] = v[u].x;
          Bk[a * s + 2 * b2 + 1] = v[u].y;
          tot += v[u].x + v[u].y;
        }
      }
    }
  } else {
    for (int e = lane; e < ss; e += 32) {
      const int a = e / s, b = e - a * s;
      const double x = base[(long long)a * ld + b];
      Bk[e] = x;
      tot += x;
    }
  }
  for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  __syncwarp();
  if (lane < s) {
    double r = 0.0, c = 0.0;
    for (int k = 0; k < s; ++k) {
      r += Bk[lane * s + k];
      c += Bk[k * s + lane];
    }
    rs[lane] = r / s;
    cs[lane] = c / s;
  }
  __syncwarp();
  const double mean = tot / ss;
  double acc = 0.0;
  for (int e = lane; e < ss; e += 32) {
    const int a = e / s, b = e - a * s;
    const double k = Bk[e] - rs[a] - cs[b] + mean;
    acc += k * k;
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const double f = sqrt(acc);
    S[(long long)i * L + j] = f;
    S[(long long)j * L + i] = f;
  }
}

__global__ void zero_diag_kernel(double *__restrict__ S, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < L) S[(long long)i * L + i] = 0.0;
}

// ---- per-site Cholesky of the diagonal blocks of C: Lc[i] lower, zeros above ----
__global__ void site_chol_kernel(const double *__restrict__ Cdiag, int s, double *__restrict__ Lc) {
  extern __shared__ double sm[];  // [s][s]
  const int i = blockIdx.x, lane = threadIdx.x;
  const int ss = s * s;
  for (int e = lane; e < ss; e += 32) sm[e] = Cdiag[(long long)i * ss + e];
  __syncwarp();
  for (int j = 0; j < s; ++j) {
    const double d = sqrt(sm[j * s + j]);
    __syncwarp();
    for (int r = j + lane; r < s; r += 32) sm[r * s + j] = (r == j) ? d : sm[r * s + j] / d;
    __syncwarp();
    for (int e = lane; e < (s - 1 - j) * (s - 1 - j); e += 32) {
      const int rr = j + 1 + e / (s - 1 - j), cc = j + 1 + e % (s - 1 - j);
      if (cc <= rr) sm[rr * s + cc] -= sm[rr * s + j] * sm[cc * s + j];
    }
    __syncwarp();
  }
  for (int e = lane; e < ss; e += 32) {
    const int a = e / s, b = e - a * s;
    Lc[(long long)i * ss + e] = (b <= a) ? sm[e] : 0.0;
  }
}

// One warp handles NSUB = min(3, 32 / ceil(s/2)) consecutive blocks (i, j0..j0+NSUB-1): the two 20x20x20
// products use all 32 lanes per block, and the Jacobi rounds -- s/2 disjoint column pairs each -- run for
// all NSUB blocks at once (30 of 32 lanes busy at s = 20 instead of 10).
__global__ void __launch_bounds__(SW * 32) di_kernel(const double *__restrict__ mJ, long long ld,
                                                     const double *__restrict__ Lc, int L, int s, int nsub,
                                                     double *__restrict__ S) {
  extern __shared__ double sm[];  // Li[s*s] + SW * (nsub * G[s*(s+1)] + Lj[s*s] + T1[s*s])
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.y;
  const int j0 = (blockIdx.x * SW + warp) * nsub;
  const int ss = s * s, gs = s + 1;
  double *Li = sm;
  for (int e = threadIdx.x; e < ss; e += SW * 32) Li[e] = Lc[(long long)i * ss + e];
  __syncthreads();
  if (j0 + nsub - 1 <= i || j0 >= L) return;  // no block of this warp lies in the upper triangle
  double *Gall = sm + ss + (size_t)warp * (nsub * s * gs + 2 * ss);
  double *Lj = Gall + nsub * s * gs, *T1 = Lj + ss;

  for (int sub = 0; sub < nsub; ++sub) {
    const int j = j0 + sub;
    double *G = Gall + sub * s * gs;
    if (j <= i || j >= L) {  // inactive slot: zero columns never rotate
      for (int e = lane; e < s * gs; e += 32) G[e] = 0.0;
      __syncwarp();
      continue;
    }
    // B (row-major, temporarily in G) and Lj
    for (int e = lane; e < ss; e += 32) {
      const int a = e / s, b = e - a * s;
      G[e] = mJ[((long long)i * s + a) * ld + (long long)j * s + b];
      Lj[e] = Lc[(long long)j * ss + e];
    }
    __syncwarp();
    // T1 = B * Lj      (Lj lower: k >= b)
    for (int e = lane; e < ss; e += 32) {
      const int a = e / s, b = e - a * s;
      double t = 0.0;
      for (int k = b; k < s; ++k) t += G[a * s + k] * Lj[k * s + b];
      T1[e] = t;
    }
    __syncwarp();
    // A = Li' * T1     (Li lower: k >= a);  stored column-major in G with stride gs
    for (int e = lane; e < ss; e += 32) {
      const int a = e / s, b = e - a * s;
      double t = 0.0;
      for (int k = a; k < s; ++k) t += Li[k * s + a] * T1[k * s + b];
      G[b * gs + a] = t;  // column b, row a
    }
    __syncwarp();
  }

  // ---- one-sided Jacobi: orthogonalise the columns of every G; sigma_k^2 = ||g_k||^2 ----
  const int nc = (s + 1) & ~1;  // even number of players (last one is a bye when s is odd)
  const int half = nc >> 1;
  const int mysub = lane / half, mypair = lane - mysub * half;
  const bool jlane = mysub < nsub;
  double *G = Gall + (jlane ? mysub : 0) * s * gs;
  for (int sweep = 0; sweep < 40; ++sweep) {
    bool rotated = false;
    for (int r = 0; r < nc - 1; ++r) {
      int p = -1, q2 = -1;
      if (jlane) {
        if (mypair == 0) {
          p = nc - 1;
          q2 = r;
        } else {
          p = (r + mypair) % (nc - 1);
          q2 = (r - mypair + (nc - 1)) % (nc - 1);
        }
      }
      if (p >= 0 && p < s && q2 < s) {
        double *gp = G + p * gs, *gq = G + q2 * gs;
        double al = 0.0, be = 0.0, ga = 0.0;
        for (int k = 0; k < s; ++k) {
          const double x = gp[k], y = gq[k];
          al += x * x;
          be += y * y;
          ga += x * y;
        }
        if (fabs(ga) > 1e-15 * sqrt(al * be) && ga != 0.0) {
          const double zeta = (be - al) / (2.0 * ga);
          const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cth = 1.0 / sqrt(1.0 + t * t), sth = cth * t;
          for (int k = 0; k < s; ++k) {
            const double x = gp[k], y = gq[k];
            gp[k] = cth * x - sth * y;
            gq[k] = sth * x + cth * y;
          }
          rotated = true;
        }
      }
      __syncwarp();
    }
    if (!__any_sync(0xffffffffu, rotated)) break;
  }
  for (int sub = 0; sub < nsub; ++sub) {
    const int j = j0 + sub;
    if (j <= i || j >= L) continue;  // warp-uniform
    const double *Gs = Gall + sub * s * gs;
    double part = 0.0;
    if (lane < s) {
      const double *gk = Gs + lane * gs;
      double lam = 0.0;
      for (int k = 0; k < s; ++k) lam += gk[k] * gk[k];
      part = log(1.0 + sqrt(1.0 + 4.0 * lam));
    }
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) {
      const double di = 0.5 * s * log(0.5) + 0.5 * part;
      S[(long long)i * L + j] = di;
      S[(long long)j * L + i] = di;
    }
  }
}


// ---- K7, second engine: eigenvalues of V = G'G by Householder tridiagonalisation + implicit QL ------------------------------------
// The one-sided Jacobi iteration above spends ~2e5 flop per 20 x 20 block with every operand in shared memory (ncu: l1tex at 91 %,
// half of the wavefronts bank conflicts; 11.8 ms at L = 500).  DI only needs the eigenvalues of the symmetric V (DCAUtils takes
// lambda_k(V) of exactly this product), so a warp takes 32 consecutive blocks (i, j0..j0+31) and, one block after the other with all
// lanes: forms G = Lc_i' mJ_ij Lc_j and V = G'G, reduces V to tridiagonal form by Householder reflections (lane k owns element k of
// the reflector, row k of the matrix-vector product and column k of the rank-2 update: three warp sums per step, no divergence) and
// parks the 2 s numbers (d, e) in an interleaved store (entry k of block m at [k * 33 + m]).  Then EVERY LANE runs the implicit QL
// iteration on ITS OWN tridiagonal matrix -- O(s^2) scalar recurrences, 32 matrices per warp in lock step instead of one lane
// working while 31 wait.  The classical EISPACK tred1 / tql1 pair, restated: ~2.5e4 flop per block, ~21 KB of shared memory per warp
// (eight warps per SM), absolute eigenvalue error ~ eps ||V||, which is what the log(1 + sqrt(1 + 4 lambda)) sum needs.
constexpr int DI_W = 5;    // warps per CTA (two CTAs per SM at s = 20)
constexpr int DI_IL = 33;  // interleave stride of the (d, e) store (doubles)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;  // the butterfly leaves the same bits on every lane
}

__host__ __device__ inline size_t di_eig_warp_doubles(int s) {
  // (x, q) pairs first (16-byte aligned), then B/G, Lj, T1/V (odd row stride), (d, e); rounded to an even number of doubles
  return (64 + (size_t)2 * s * s + (size_t)s * (s | 1) + (size_t)2 * s * DI_IL + 1) & ~(size_t)1;
}
__host__ __device__ inline size_t di_eig_li_doubles(int s) { return ((size_t)s * s + 1) & ~(size_t)1; }

// S > 0: the number of states is a compile-time constant (S = 20, proteins: q = 21) -- the inner loops of the three products unroll
// into loads with immediate offsets and run over the full k range (the zero triangles of Lc_i, Lc_j are stored); S = 0: any s.
template <int S>
__global__ void __launch_bounds__(DI_W * 32) di_eig_kernel(const double *__restrict__ mJ, long long ld, const double *__restrict__ Lc,
                                                           int L, int s_rt, double *__restrict__ S_out) {
  extern __shared__ __align__(16) double sm_di[];  // Li + DI_W * di_eig_warp_doubles(s)
  double *sm = sm_di;
  const int s = S > 0 ? S : s_rt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.y;
  const int j0 = (blockIdx.x * DI_W + warp) * 32;
  const int ss = s * s, ldv = s | 1;  // odd row stride: lanes that walk down a column of V hit distinct banks
  double *Li = sm;
  for (int e = threadIdx.x; e < ss; e += blockDim.x) Li[e] = Lc[(long long)i * ss + e];
  __syncthreads();
  if (j0 + 31 <= i || j0 >= L) return;  // no block of this warp lies in the upper triangle
  double *wbase = sm + di_eig_li_doubles(s) + (size_t)warp * di_eig_warp_doubles(s);
  double2 *xq = reinterpret_cast<double2 *>(wbase);  // (x_k, q_k) of the current reflector
  double *Bm = wbase + 64, *Lj = Bm + ss, *Vm = Lj + ss, *DE = Vm + (size_t)s * ldv;
  const int m_lo = max(0, i + 1 - j0), m_hi = min(32, L - j0);  // active slots [m_lo, m_hi)
  const int da = 32 / s, db = 32 - da * s;                      // (row, column) advance of an element index that grows by 32
  const int a0 = lane / s, b0 = lane - a0 * s;

  for (int m = m_lo; m < m_hi; ++m) {
    const int j = j0 + m;
    {
      int a = a0, b = b0;
      for (int e = lane; e < ss; e += 32) {
        Bm[e] = mJ[((long long)i * s + a) * ld + (long long)j * s + b];
        Lj[e] = Lc[(long long)j * ss + e];
        a += da; b += db;
        if (b >= s) { b -= s; ++a; }
      }
    }
    __syncwarp();
    {  // T1 = B * Lj      (Lj lower: only k >= b contributes), into the V buffer
      int a = a0, b = b0;
      for (int e = lane; e < ss; e += 32) {
        const double *bp = Bm + a * s, *lp = Lj + b;
        double t0 = 0.0, t1 = 0.0;
        if (S > 0) {
#pragma unroll
          for (int k = 0; k + 1 < S; k += 2) {
            t0 = fma(bp[k], lp[k * S], t0);
            t1 = fma(bp[k + 1], lp[(k + 1) * S], t1);
          }
          if (S & 1) t0 = fma(bp[S - 1], lp[(S - 1) * S], t0);
        } else {
          for (int k = b; k < s; ++k) t0 = fma(bp[k], lp[k * s], t0);
        }
        Vm[e] = t0 + t1;
        a += da; b += db;
        if (b >= s) { b -= s; ++a; }
      }
    }
    __syncwarp();
    {  // G = Li' * T1     (Li lower: only k >= a contributes), row-major over B
      int a = a0, b = b0;
      for (int e = lane; e < ss; e += 32) {
        const double *lp = Li + a, *tp = Vm + b;
        double t0 = 0.0, t1 = 0.0;
        if (S > 0) {
#pragma unroll
          for (int k = 0; k + 1 < S; k += 2) {
            t0 = fma(lp[k * S], tp[k * S], t0);
            t1 = fma(lp[(k + 1) * S], tp[(k + 1) * S], t1);
          }
          if (S & 1) t0 = fma(lp[(S - 1) * S], tp[(S - 1) * S], t0);
        } else {
          for (int k = a; k < s; ++k) t0 = fma(lp[k * s], tp[k * s], t0);
        }
        Bm[e] = t0 + t1;
        a += da; b += db;
        if (b >= s) { b -= s; ++a; }
      }
    }
    __syncwarp();
    {  // V = G' G, full symmetric storage with row stride ldv
      int a = a0, b = b0;
      for (int e = lane; e < ss; e += 32) {
        const double *ga = Bm + a, *gb = Bm + b;
        double t0 = 0.0, t1 = 0.0;
        if (S > 0) {
#pragma unroll
          for (int k = 0; k + 1 < S; k += 2) {
            t0 = fma(ga[k * S], gb[k * S], t0);
            t1 = fma(ga[(k + 1) * S], gb[(k + 1) * S], t1);
          }
          if (S & 1) t0 = fma(ga[(S - 1) * S], gb[(S - 1) * S], t0);
        } else {
          int k = 0;
          for (; k + 1 < s; k += 2) {
            t0 = fma(ga[k * s], gb[k * s], t0);
            t1 = fma(ga[(k + 1) * s], gb[(k + 1) * s], t1);
          }
          if (k < s) t0 = fma(ga[k * s], gb[k * s], t0);
        }
        Vm[a * ldv + b] = t0 + t1;
        a += da; b += db;
        if (b >= s) { b -= s; ++a; }
      }
    }
    __syncwarp();
    // Householder reduction, rows s-1 .. 1: u = row r (columns 0..l, l = r-1) with u_l -= g, g = -sign(u_l) ||u||, H = u'u / 2,
    // p = V u / H, K = u'p / (2H), q = p - K u, V -= u q' + q u'.  The sub-diagonal entry of row r is g.  (No rescaling of the
    // row: the entries of V are squares of couplings, far from the over/underflow thresholds; a row whose squares underflow to
    // zero is treated as zero -- its eigenvalue contribution is below 1e-300 either way.)
#pragma unroll 1
    for (int r = s - 1; r >= 1; --r) {
      const int l = r - 1;
      double *esub = DE + (size_t)(s + r - 1) * DI_IL + m;  // e[r-1] (already shifted for the QL iteration)
      if (l == 0) {
        if (lane == 0) *esub = Vm[r * ldv];
        continue;
      }
      double x = (lane <= l) ? Vm[r * ldv + lane] : 0.0;
      double h = warp_sum(x * x);
      if (!(h >= 1e-290)) {  // warp-uniform (also a NaN row: the NaN reaches the result through the diagonal)
        if (lane == 0) *esub = 0.0;
        continue;
      }
      const double f = __shfl_sync(0xffffffffu, x, l);
      const double g = (f >= 0.0) ? -sqrt(h) : sqrt(h);
      if (lane == 0) *esub = g;
      h -= f * g;
      const double rh = 1.0 / h;
      if (lane == l) x = f - g;
      xq[lane].x = x;
      __syncwarp();
      double pv = 0.0;
      if (lane <= l) {
        const double *row = Vm + lane * ldv;
        double p0 = 0.0, p1 = 0.0;
        int k = 0;
#pragma unroll 4
        for (; k + 1 <= l; k += 2) {
          p0 = fma(row[k], xq[k].x, p0);
          p1 = fma(row[k + 1], xq[k + 1].x, p1);
        }
        if (k <= l) p0 = fma(row[k], xq[k].x, p0);
        pv = (p0 + p1) * rh;
      }
      const double K = warp_sum(pv * x) * (0.5 * rh);
      const double q = fma(-K, x, pv);  // 0 on the lanes beyond l
      xq[lane].y = q;
      __syncwarp();
      if (lane <= l) {
        double *vp = Vm + lane;
#pragma unroll 4
        for (int jj = 0; jj <= l; ++jj) {
          const double2 w = xq[jj];
          *vp -= fma(w.x, q, w.y * x);
          vp += (S > 0 ? (S | 1) : ldv);
        }
      }
      __syncwarp();
    }
    if (lane < s) DE[(size_t)lane * DI_IL + m] = Vm[lane * ldv + lane];
    if (lane == 0) DE[(size_t)(2 * s - 1) * DI_IL + m] = 0.0;
    __syncwarp();
  }

  // ---- implicit QL on (d, e), eigenvalues only: lane m owns the tridiagonal matrix in slot m ----
  const int j = j0 + lane;
  if (j <= i || j >= L) return;
#define D_(k) DE[(size_t)(k) * DI_IL + lane]
#define E_(k) DE[(size_t)(s + (k)) * DI_IL + lane]
  for (int l = 0; l < s; ++l) {
    for (int iter = 0; iter < 60; ++iter) {
      int m = l;
      for (; m < s - 1; ++m) {
        const double dd = fabs(D_(m)) + fabs(D_(m + 1));
        if (fabs(E_(m)) <= 2.220446049250313e-16 * dd) break;
      }
      if (m == l) break;
      const double el = E_(l), dl = D_(l);
      double g = (D_(l + 1) - dl) / (2.0 * el);
      double r = sqrt(fma(g, g, 1.0));
      g = D_(m) - dl + el / (g + copysign(r, g));
      double sn = 1.0, cs = 1.0, p = 0.0;
      int k = m - 1;
      for (; k >= l; --k) {
        const double ek = E_(k);
        const double f = sn * ek, b = cs * ek;
        r = sqrt(fma(f, f, g * g));
        E_(k + 1) = r;
        if (r == 0.0) {
          D_(k + 1) -= p;
          E_(m) = 0.0;
          break;
        }
        const double rinv = 1.0 / r;
        sn = f * rinv;
        cs = g * rinv;
        g = D_(k + 1) - p;
        r = fma(D_(k) - g, sn, 2.0 * cs * b);
        p = sn * r;
        D_(k + 1) = g + p;
        g = fma(cs, r, -b);
      }
      if (r == 0.0 && k >= l) continue;
      D_(l) -= p;
      E_(l) = g;
      E_(m) = 0.0;
    }
  }
  double part = 0.0;
  for (int k = 0; k < s; ++k) part += log(1.0 + sqrt(fma(4.0, fmax(D_(k), 0.0), 1.0)));
#undef E_
#undef D_
  const double di = 0.5 * s * log(0.5) + 0.5 * part;
  S_out[(long long)i * L + j] = di;
  S_out[(long long)j * L + i] = di;
}

}  // namespace

int32_t gdca_k_score(gdca_ctx *ctx, int score) {
  if (!ctx->have_inv) return gdca_fail(ctx, GDCA_ERR_STATE, "score: inverse not computed");
  const int L = (int)ctx->L, s = ctx->s;
  GDCA_TRY(gdca_reserve(ctx, ctx->dS, ctx->capS, (size_t)L * L));
  dim3 grid((unsigned)((L + SW - 1) / SW), (unsigned)L);
  if (score == GDCA_SCORE_FROB) {
    const size_t smem = (size_t)SW * (s * s + 2 * s + 2) * sizeof(double);
    fn_kernel<<<grid, SW * 32, smem, ctx->stream>>>(ctx->dmJ, ctx->npad, L, s, ctx->dS);
    GDCA_LAUNCH_CHECK(ctx);
  } else if (score == GDCA_SCORE_DI) {
    // Lc lives behind the saved diagonal blocks
    GDCA_TRY(gdca_reserve(ctx, ctx->dRed, ctx->capRed, (size_t)L * s * s + 4096));
    double *Lc = ctx->dRed;
    site_chol_kernel<<<(unsigned)L, 32, (size_t)s * s * sizeof(double), ctx->stream>>>(ctx->dCdiag, s, Lc);
    GDCA_LAUNCH_CHECK(ctx);
    if (ctx->di_engine != 0) {
      // tridiagonalisation by the warp, implicit QL by the lane: 32 site pairs per warp, DI_W warps per CTA
      const size_t smem = (di_eig_li_doubles(s) + (size_t)DI_W * di_eig_warp_doubles(s)) * sizeof(double);
      dim3 egrid((unsigned)((L + DI_W * 32 - 1) / (DI_W * 32)), (unsigned)L);
      if (s == 20) {
        GDCA_CUDA(ctx, cudaFuncSetAttribute(di_eig_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        di_eig_kernel<20><<<egrid, DI_W * 32, smem, ctx->stream>>>(ctx->dmJ, ctx->npad, Lc, L, s, ctx->dS);
      } else {
        GDCA_CUDA(ctx, cudaFuncSetAttribute(di_eig_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        di_eig_kernel<0><<<egrid, DI_W * 32, smem, ctx->stream>>>(ctx->dmJ, ctx->npad, Lc, L, s, ctx->dS);
      }
      GDCA_LAUNCH_CHECK(ctx);
      zero_diag_kernel<<<(unsigned)((L + 255) / 256), 256, 0, ctx->stream>>>(ctx->dS, L);
      GDCA_LAUNCH_CHECK(ctx);
      return GDCA_OK;
    }
    const int half = (s + 1) / 2;
    const int nsub = (32 / half) < 3 ? (32 / half) : 3;
    const size_t smem = ((size_t)s * s + (size_t)SW * ((size_t)nsub * s * (s + 1) + 2 * s * s)) * sizeof(double);
    GDCA_CUDA(ctx, cudaFuncSetAttribute(di_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 dgrid((unsigned)((L + SW * nsub - 1) / (SW * nsub)), (unsigned)L);
    di_kernel<<<dgrid, SW * 32, smem, ctx->stream>>>(ctx->dmJ, ctx->npad, Lc, L, s, nsub, ctx->dS);
    GDCA_LAUNCH_CHECK(ctx);
  } else {
    return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "score: must be 0 (frob) or 1 (DI)");
  }
  zero_diag_kernel<<<(unsigned)((L + 255) / 256), 256, 0, ctx->stream>>>(ctx->dS, L);
  GDCA_LAUNCH_CHECK(ctx);
  return GDCA_OK;
}
