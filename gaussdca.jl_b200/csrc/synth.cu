// synth.cu -- synthetic clustered alignment of SURVEY 8(d), generated directly in HBM.
// No reference counterpart (the reference ships two Pfam fixtures only).  Counter-based SplitMix64
// draws, so this kernel, oracle/gdca_oracle_c.c:oracle_synth_alignment and any other port produce
// identical bytes: K = M/50 ancestors (gap w.p. 0.10, else uniform 1..20); sequence k copies
// ancestor k mod K and resamples each site uniformly over 1..21 w.p. mu_k ~ U(0.05, 0.65); Z[0,0]=21.
#include "gdca_internal.cuh"

namespace {
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
__device__ __forceinline__ unsigned long long draw(unsigned long long seed, unsigned long long t, unsigned long long a,
                                                   unsigned long long b) {
  return splitmix64(splitmix64(splitmix64(seed ^ (t * 0xD1B54A32D192ED03ULL)) + a) + b);
}
__device__ __forceinline__ double u01(unsigned long long r) {
  return __dmul_rn((double)(r >> 11), 1.0 / 9007199254740992.0);
}

__global__ void synth_kernel(int8_t *__restrict__ Z, long long L, long long M, unsigned long long seed) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= L * M) return;
  const long long k = e / L, i = e - k * L;
  const long long K = (M / 50) > 0 ? (M / 50) : 1;
  const unsigned long long anc = (unsigned long long)(k % K);
  const double mu = __dadd_rn(0.05, __dmul_rn(0.60, u01(draw(seed, 3, (unsigned long long)k, 0))));
  const unsigned long long ra = draw(seed, 1, anc, (unsigned long long)i);
  int v = (u01(ra) < 0.10) ? 21 : (int)(1 + (splitmix64(ra) % 20));
  const unsigned long long rm = draw(seed, 2, (unsigned long long)k, (unsigned long long)i);
  if (u01(rm) < mu) v = (int)(1 + (splitmix64(rm) % 21));
  if (e == 0) v = 21;
  Z[e] = (int8_t)v;
}
}  // namespace

int32_t gdca_k_synth(gdca_ctx *ctx, int8_t *Zdev, int64_t L, int64_t M, uint64_t seed) {
  const long long ne = (long long)L * M;
  synth_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, ctx->stream>>>(Zdev, L, M, seed);
  GDCA_LAUNCH_CHECK(ctx);
  return GDCA_OK;
}
