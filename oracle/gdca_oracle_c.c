/*
 * gdca_oracle_c.c -- CPU restatement of the gDCA hot loops.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The shipped path is the CUDA
 * library under gaussdca.jl_b200/csrc and has no CPU fallback.
 *
 * What it restates.  GaussDCA.jl is a thin wrapper (reference src/GaussDCA.jl:8-47); the heavy
 * loops live in the third-party package DCAUtils 1.x, which is NOT vendored under /root/reference
 * (Project.toml:6,12 -- compat "1", no Manifest, so no exact pin).  The functions below restate the
 * published behaviour of that package at the reference's own call sites:
 *
 *   oracle_compress_Z            DCAUtils compress_Z, reached from compute_weighted_frequencies,
 *                                call site src/GaussDCA.jl:28 (12 residues x 5 bits per UInt64)
 *   oracle_ident_sum_*           DCAUtils compute_theta (theta == :auto branch), src/GaussDCA.jl:28
 *   oracle_neighbour_counts_*    DCAUtils compute_weights,                       src/GaussDCA.jl:28
 *   oracle_weighted_freqs        DCAUtils compute_freqs (Pi_true, Pij_true),     src/GaussDCA.jl:28
 *   oracle_synth_alignment       no reference counterpart: the synthetic generator of SURVEY 8(d)
 *
 * Two pair-sweep flavours exist on purpose, mirroring the reference's own duality
 * (test/runtests.jl:78-86, DCAUTILS_FORCE_FALLBACK): "_packed" works on the 5-bit words,
 * "_bytes" compares the Int8 matrix directly.  They must agree bit for bit.
 *
 * Parity pin: the end-to-end oracle built on these loops reproduces all four golden files of
 * test/data at their 7-digit print precision (tests/test_oracle_golden.py).  Intermediates
 * (counts, W, Meff, C, mJ) are not pinned by any reference test.
 *
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC (see oracle/Makefile).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define RES_PER_WORD 12 /* 12 x 5 bits = 60 bits used of each UInt64 */

int64_t oracle_words_per_seq(int64_t L) { return (L + RES_PER_WORD - 1) / RES_PER_WORD; }

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* `julia -t N` (README.md:92-94): the thread count is the caller's choice.  Launchers such as torchrun export
 * OMP_NUM_THREADS=1 into the environment; bench.py sets the count explicitly through this instead. */
void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n >= 1) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* Z is L x M column-major Int8 (one sequence per column, src/GaussDCA.jl:24). */
void oracle_compress_Z(const int8_t *Z, int64_t L, int64_t M, uint64_t *cZ) {
  const int64_t nw = oracle_words_per_seq(L);
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < M; ++k) {
    const int8_t *z = Z + k * L;
    uint64_t *c = cZ + k * nw;
    for (int64_t w = 0; w < nw; ++w) {
      uint64_t acc = 0;
      for (int r = 0; r < RES_PER_WORD; ++r) {
        int64_t i = w * RES_PER_WORD + r;
        uint64_t v = (i < L) ? (uint64_t)(uint8_t)z[i] & 31u : 0u; /* pad with 0: equal in all seqs */
        acc |= v << (5 * r);
      }
      c[w] = acc;
    }
  }
}

/* number of differing 5-bit fields between two packed words */
static inline int diff_fields(uint64_t a, uint64_t b) {
  uint64_t z = a ^ b;
  uint64_t t = z | (z >> 1) | (z >> 2) | (z >> 3) | (z >> 4);
  t &= 0x0084210842108421ULL; /* bit 0 of each of the 12 fields */
  return __builtin_popcountll(t);
}

static inline int64_t ham_packed(const uint64_t *a, const uint64_t *b, int64_t nw) {
  int64_t d = 0;
  for (int64_t w = 0; w < nw; ++w) d += diff_fields(a[w], b[w]);
  return d;
}

static inline int64_t ham_bytes(const int8_t *a, const int8_t *b, int64_t L) {
  int64_t d = 0;
  for (int64_t i = 0; i < L; ++i) d += (a[i] != b[i]);
  return d;
}

/* sum over k<l of ident(k,l), ident = L - hamming (gap == gap counts as identical). */
uint64_t oracle_ident_sum_packed(const uint64_t *cZ, int64_t L, int64_t M) {
  const int64_t nw = oracle_words_per_seq(L);
  uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total)
  for (int64_t k = 0; k < M - 1; ++k) {
    uint64_t s = 0;
    for (int64_t l = k + 1; l < M; ++l) s += (uint64_t)(L - ham_packed(cZ + k * nw, cZ + l * nw, nw));
    total += s;
  }
  return total;
}

uint64_t oracle_ident_sum_bytes(const int8_t *Z, int64_t L, int64_t M) {
  uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total)
  for (int64_t k = 0; k < M - 1; ++k) {
    uint64_t s = 0;
    for (int64_t l = k + 1; l < M; ++l) s += (uint64_t)(L - ham_bytes(Z + k * L, Z + l * L, L));
    total += s;
  }
  return total;
}

/* counts[k] = 1 + #{l != k : hamming(k,l) < thresh}.  Restricted to rows [k0,k1) x all l>k so a
 * bounded sample of the sweep can be timed (bench cpu_baseline); full sweep = [0,M). */
void oracle_neighbour_counts_packed_range(const uint64_t *cZ, int64_t L, int64_t M, int64_t thresh,
                                          int64_t k0, int64_t k1, int32_t *counts) {
  const int64_t nw = oracle_words_per_seq(L);
  (void)L;
  for (int64_t k = 0; k < M; ++k) counts[k] = 1;
#pragma omp parallel
  {
    int32_t *priv = (int32_t *)calloc((size_t)M, sizeof(int32_t));
#pragma omp for schedule(dynamic, 16)
    for (int64_t k = k0; k < k1; ++k) {
      int32_t ck = 0;
      for (int64_t l = k + 1; l < M; ++l) {
        if (ham_packed(cZ + k * nw, cZ + l * nw, nw) < thresh) {
          ++ck;
          ++priv[l];
        }
      }
      priv[k] += ck;
    }
#pragma omp critical
    for (int64_t k = 0; k < M; ++k) counts[k] += priv[k];
    free(priv);
  }
}

void oracle_neighbour_counts_packed(const uint64_t *cZ, int64_t L, int64_t M, int64_t thresh, int32_t *counts) {
  oracle_neighbour_counts_packed_range(cZ, L, M, thresh, 0, M, counts);
}

void oracle_neighbour_counts_bytes(const int8_t *Z, int64_t L, int64_t M, int64_t thresh, int32_t *counts) {
  for (int64_t k = 0; k < M; ++k) counts[k] = 1;
#pragma omp parallel
  {
    int32_t *priv = (int32_t *)calloc((size_t)M, sizeof(int32_t));
#pragma omp for schedule(dynamic, 16)
    for (int64_t k = 0; k < M - 1; ++k) {
      for (int64_t l = k + 1; l < M; ++l) {
        if (ham_bytes(Z + k * L, Z + l * L, L) < thresh) {
          ++priv[k];
          ++priv[l];
        }
      }
    }
#pragma omp critical
    for (int64_t k = 0; k < M; ++k) counts[k] += priv[k];
    free(priv);
  }
}

/* ident sum restricted to rows [k0,k1) (bounded sample for the CPU baseline). */
uint64_t oracle_ident_sum_packed_range(const uint64_t *cZ, int64_t L, int64_t M, int64_t k0, int64_t k1) {
  const int64_t nw = oracle_words_per_seq(L);
  uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total)
  for (int64_t k = k0; k < k1; ++k) {
    uint64_t s = 0;
    for (int64_t l = k + 1; l < M; ++l) s += (uint64_t)(L - ham_packed(cZ + k * nw, cZ + l * nw, nw));
    total += s;
  }
  return total;
}

/*
 * Weighted one- and two-point frequencies over states 1..s (s = q-1; state q is dropped).
 * Index of (site i, state a), 0-based: i*s + (a-1).  Pij is n x n, n = s*L, written full/symmetric.
 *   Pi[(i,a)]        = sum_k W[k] [Z[i,k]=a] / Meff
 *   Pij[(i,a),(j,b)] = sum_k W[k] [Z[i,k]=a][Z[j,k]=b] / Meff
 * Sequence range [k0,k1) lets the baseline time a bounded sample; full = [0,M).
 * Threads split the site index i, so every output element is summed in sequence order k.
 */
void oracle_weighted_freqs_range(const int8_t *Z, int64_t L, int64_t M, int32_t q, const double *W, double Meff,
                                 int64_t k0, int64_t k1, double *Pi, double *Pij) {
  const int64_t s = q - 1, n = s * L;
  memset(Pi, 0, (size_t)n * sizeof(double));
  memset(Pij, 0, (size_t)n * (size_t)n * sizeof(double));
  (void)M;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t i = 0; i < L; ++i) {
    for (int64_t k = k0; k < k1; ++k) {
      const int8_t *z = Z + k * L;
      const int a = z[i];
      if (a == q) continue;
      const double w = W[k];
      const int64_t r = i * s + (a - 1);
      Pi[r] += w;
      double *row = Pij + r * n;
      for (int64_t j = i; j < L; ++j) {
        const int b = z[j];
        if (b != q) row[j * s + (b - 1)] += w;
      }
    }
  }
  /* normalise the computed (upper, by site) part and mirror it */
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; ++r) {
    Pi[r] /= Meff;
    const int64_t i = r / s;
    for (int64_t c = i * s; c < n; ++c) {
      double v = Pij[r * n + c] / Meff;
      Pij[r * n + c] = v;
    }
  }
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; ++r) {
    const int64_t i = r / s;
    for (int64_t c = (i + 1) * s; c < n; ++c) Pij[c * n + r] = Pij[r * n + c];
  }
}

void oracle_weighted_freqs(const int8_t *Z, int64_t L, int64_t M, int32_t q, const double *W, double Meff, double *Pi,
                           double *Pij) {
  oracle_weighted_freqs_range(Z, L, M, q, W, Meff, 0, M, Pi, Pij);
}

/* ---- synthetic alignment generator; identical bytes from C, CUDA and numpy ---- */
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
/* counter-based draw: stream tag t, indices (a,b) */
static inline uint64_t draw(uint64_t seed, uint64_t t, uint64_t a, uint64_t b) {
  return splitmix64(splitmix64(splitmix64(seed ^ (t * 0xD1B54A32D192ED03ULL)) + a) + b);
}
static inline double u01(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

void oracle_synth_alignment(int8_t *Z, int64_t L, int64_t M, uint64_t seed) {
  const int64_t K = (M / 50) > 0 ? (M / 50) : 1; /* ancestor families */
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < M; ++k) {
    const uint64_t anc = (uint64_t)(k % K);
    const double mu = 0.05 + 0.60 * u01(draw(seed, 3, (uint64_t)k, 0));
    for (int64_t i = 0; i < L; ++i) {
      /* ancestor residue: gap (21) w.p. 0.10, else uniform over 1..20 */
      uint64_t ra = draw(seed, 1, anc, (uint64_t)i);
      int8_t v = (u01(ra) < 0.10) ? 21 : (int8_t)(1 + (splitmix64(ra) % 20));
      /* mutation: w.p. mu resample uniformly over 1..21 */
      uint64_t rm = draw(seed, 2, (uint64_t)k, (uint64_t)i);
      if (u01(rm) < mu) v = (int8_t)(1 + (splitmix64(rm) % 21));
      Z[k * L + i] = v;
    }
  }
  Z[0] = 21; /* guarantee q = 21 */
}
