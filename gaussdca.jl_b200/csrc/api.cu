// api.cu -- the C ABI of libgdca_b200.so (include/gdca_b200.h).  Host-side orchestration only:
// argument checks, buffer management, stage ordering, CUDA-event timers.  No arithmetic of the
// hot path runs on the CPU; without an sm_100 device gdca_create() fails (no fallback).
#include <math.h>
#include <string.h>

#include <new>
#include <vector>

#include <stdlib.h>

#include "gdca_internal.cuh"

static thread_local std::string g_create_error;

namespace {

int32_t peer_close_all(gdca_ctx *ctx) {
  for (void *&m : ctx->peer_opened) {
    if (m) cudaIpcCloseMemHandle(m);
    m = nullptr;
  }
  for (int r = 0; r < GDCA_MAX_PEERS; ++r) ctx->peer_counts[r] = nullptr, ctx->peer_C[r] = nullptr;
  ctx->peers_ready = false;
  return GDCA_OK;
}

enum { EV_BEGIN = 0, EV_H2D, EV_PACK, EV_THETA, EV_WEIGHTS, EV_COV, EV_CHOL, EV_SCORE, EV_APC, EV_RANK, EV_D2H, EV_COUNT };

int32_t set_device(gdca_ctx *ctx) {
  GDCA_CUDA(ctx, cudaSetDevice(ctx->device));
  return GDCA_OK;
}

int32_t rec(gdca_ctx *ctx, int which) {
  GDCA_CUDA(ctx, cudaEventRecord(ctx->ev[which], ctx->stream));
  return GDCA_OK;
}

int planes_for_q(int q) {
  int p = 1;
  while ((1 << p) <= q) ++p;  // q needs p bits
  return p;
}

// shape bookkeeping once q is known
void set_shape(gdca_ctx *ctx, int64_t L, int64_t M, int q) {
  ctx->L = L;
  ctx->M = M;
  ctx->q = q;
  ctx->s = q - 1;
  ctx->nplanes = planes_for_q(q);
  ctx->Mpad = (M + GDCA_TILE - 1) / GDCA_TILE * GDCA_TILE;
  ctx->nwords = (L + 31) / 32;
  ctx->n = (int64_t)(q - 1) * L;
  ctx->npad = (ctx->n + GDCA_NB - 1) / GDCA_NB * GDCA_NB;
  ctx->stats.L = L;
  ctx->stats.M = M;
  ctx->stats.n = ctx->n;
  ctx->stats.q = q;
}

int32_t load_common(gdca_ctx *ctx, int64_t L, int64_t M) {
  // q = max(Z)  (src/GaussDCA.jl:25-26)
  ctx->L = L;
  ctx->M = M;
  GDCA_TRY(gdca_k_maxq(ctx));
  int qq[2] = {0, 0};
  GDCA_CUDA(ctx, cudaMemcpyAsync(qq, ctx->dQ, sizeof qq, cudaMemcpyDeviceToHost, ctx->stream));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int q = qq[0];
  if (qq[1] < 1) {
    char b[128];
    snprintf(b, sizeof b, "alignment holds the residue code %d: codes must be >= 1 (A=1 ... Y=20, everything else 21)", qq[1]);
    ctx->err = b;
    return GDCA_ERR_INVALID_ARG;
  }
  if (q >= 32) {
    char b[96];
    snprintf(b, sizeof b, "parameter q=%d is too big (max 31 is allowed)", q);
    ctx->err = b;
    return GDCA_ERR_Q_TOO_BIG;
  }
  if (q < 2) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "alignment has q = max(Z) < 2: nothing to couple");
  set_shape(ctx, L, M, q);
  // buffers that peers have mapped are about to move: drop the mappings (the host re-exchanges handles)
  if (ctx->peers_ready && ((size_t)3 * ctx->Mpad > ctx->capCounts || (size_t)ctx->npad * ctx->npad > ctx->capC)) peer_close_all(ctx);
  ctx->have_alignment = true;
  ctx->have_lists = ctx->have_hist = ctx->have_weights = ctx->have_cov = ctx->have_inv = false;
  ctx->weights_from_counts = false;
  ctx->have_V = 0;
  GDCA_TRY(gdca_k_pack(ctx));  // builds the per-site lists first (site order), then the bit planes
  return GDCA_OK;
}

int32_t check_LM(gdca_ctx *ctx, const void *Z, int64_t L, int64_t M) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  if (!Z) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "Z is NULL");
  if (L < 1 || M < 1) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "alignment must have L >= 1 and M >= 1");
  // pack_planes_kernel parks ceil(L/32) x 5 planes x 128 B of one 32-sequence group in shared memory (<= 227 KB)
  if (L > GDCA_MAX_L) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "L must be <= 11616 (GDCA_MAX_L, see include/gdca_b200.h)");
  if (M >= (1ll << 31) - GDCA_TILE) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "M must be < 2^31");
  return GDCA_OK;
}

// theta / weights on the loaded alignment.  theta < 0 => :auto.
int32_t weights_stage(gdca_ctx *ctx, double theta) {
  gdca_stats_t &st = ctx->stats;
  st.theta_passes = 0;
  st.ident_sum = 0;
  if (theta < 0) {
    // theta = :auto without a pair sweep: sum_{k<l} ident(k,l) = sum_i sum_v n_iv (n_iv - 1)/2 from the per-site
    // state histograms (exact integers, O(M L)); then ONE neighbour-count sweep with the exact threshold.
    unsigned long long ident = 0;
    GDCA_TRY(gdca_k_ident_sum(ctx, &ident));
    double th;
    int64_t thresh;
    GDCA_TRY(gdca_theta_from_ident_sum(ctx->L, ctx->M, ident, &th, &thresh));
    st.theta = th;
    st.thresh = thresh;
    st.ident_sum = ident;
    GDCA_TRY(rec(ctx, EV_THETA));
    GDCA_TRY(gdca_k_pair_pass(ctx, 1, (int)thresh, 1));
    st.theta_passes = 1;
    GDCA_TRY(gdca_k_finish_weights(ctx, 0));
  } else {
    st.theta = theta;
    GDCA_TRY(rec(ctx, EV_THETA));
    if (theta == 0.0) {
      st.thresh = 0;
      GDCA_TRY(gdca_k_finish_weights(ctx, -1));
    } else {
      st.thresh = (int64_t)floor(theta * (double)ctx->L);
      GDCA_TRY(gdca_k_pair_pass(ctx, 1, (int)st.thresh, 1));
      st.theta_passes = 1;
      GDCA_TRY(gdca_k_finish_weights(ctx, 0));
    }
  }
  st.meff = ctx->meff;
  GDCA_TRY(rec(ctx, EV_WEIGHTS));
  return GDCA_OK;
}

// A staged host-buffer entry point is about to change L / n / q under the loaded alignment: every piece of device state
// derived from that alignment (lists, planes, weights, filter operands, covariance, inverse) stops being usable, so a later
// gdca_dev_* call reports GDCA_ERR_STATE instead of mixing the new shape with the old buffers.
void drop_alignment_state(gdca_ctx *ctx) {
  ctx->have_alignment = ctx->have_lists = ctx->have_hist = ctx->have_weights = ctx->have_cov = ctx->have_inv = false;
  ctx->weights_from_counts = false;
  ctx->have_V = 0;
  if (ctx->dZ_borrowed) {
    ctx->dZ = nullptr;
    ctx->capZ = 0;
    ctx->dZ_borrowed = false;
  }
}

float ev_ms(gdca_ctx *ctx, int a, int b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]) != cudaSuccess) {
    cudaGetLastError();
    return 0.f;
  }
  return ms;
}

int32_t upload_padded(gdca_ctx *ctx, double *dst, int64_t npad, const double *src_host, int64_t n) {
  GDCA_CUDA(ctx, cudaMemsetAsync(dst, 0, (size_t)npad * npad * sizeof(double), ctx->stream));
  GDCA_CUDA(ctx, cudaMemcpy2DAsync(dst, (size_t)npad * sizeof(double), src_host, (size_t)n * sizeof(double),
                                   (size_t)n * sizeof(double), (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  return GDCA_OK;
}

int32_t download_padded(gdca_ctx *ctx, double *dst_host, const double *src, int64_t npad, int64_t n) {
  GDCA_CUDA(ctx, cudaMemcpy2DAsync(dst_host, (size_t)n * sizeof(double), src, (size_t)npad * sizeof(double),
                                   (size_t)n * sizeof(double), (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GDCA_OK;
}

}  // namespace

extern "C" {

int32_t gdca_abi_version(void) { return GDCA_ABI_VERSION; }

const char *gdca_status_string(int32_t status) {
  switch (status) {
    case GDCA_OK: return "ok";
    case GDCA_ERR_INVALID_ARG: return "invalid argument";
    case GDCA_ERR_Q_TOO_BIG: return "q too big";
    case GDCA_ERR_NOT_SPD: return "matrix is not positive definite";
    case GDCA_ERR_CUDA: return "CUDA error";
    case GDCA_ERR_OOM: return "out of device memory";
    case GDCA_ERR_NO_DEVICE: return "no usable sm_100 device";
    case GDCA_ERR_STATE: return "stage called out of order";
    default: return "unknown status";
  }
}

const char *gdca_last_error(const gdca_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int32_t gdca_create(gdca_ctx **out, int32_t device) {
  if (!out) return GDCA_ERR_INVALID_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) +
                     "); libgdca_b200 has no CPU fallback";
    return GDCA_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= ndev) {
    g_create_error = "device ordinal out of range";
    return GDCA_ERR_INVALID_ARG;
  }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    g_create_error = cudaGetErrorString(e);
    return GDCA_ERR_CUDA;
  }
  if (prop.major != 10) {
    char b[160];
    snprintf(b, sizeof b, "device %d is sm_%d%d; libgdca_b200 is built for sm_100a only (no fallback)", device, prop.major,
             prop.minor);
    g_create_error = b;
    return GDCA_ERR_NO_DEVICE;
  }
  gdca_ctx *ctx = new (std::nothrow) gdca_ctx();
  if (!ctx) return GDCA_ERR_OOM;
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  auto fail = [&](cudaError_t err) {
    g_create_error = cudaGetErrorString(err);
    delete ctx;
    return GDCA_ERR_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(e);
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // numerically lower = higher priority
  if ((e = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) return fail(e);
  if ((e = cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, prio_lo)) != cudaSuccess) return fail(e);
  if ((e = cudaStreamCreateWithPriority(&ctx->stream3, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) return fail(e);
  for (cudaEvent_t *ev : {&ctx->ev_diag, &ctx->ev_p1, &ctx->ev_u2a, &ctx->ev_u2b})
    if ((e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
  if (const char *env = getenv("GDCA_CHOL_LOOKAHEAD")) ctx->chol_inner_lookahead = atoi(env) != 0;
  if (const char *env = getenv("GDCA_DIAG_BLOCKED")) ctx->diag_blocked = atoi(env) != 0;
  if (const char *env = getenv("GDCA_INV_GRAPH")) ctx->inv_graph_mode = atoi(env) != 0;
  if (const char *env = getenv("GDCA_DI_ENGINE")) ctx->di_engine = atoi(env) != 0;
  if ((e = cudaEventCreateWithFlags(&ctx->ev_fact, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
  if ((e = cudaEventCreateWithFlags(&ctx->ev_trail, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
  if ((e = cudaEventCreateWithFlags(&ctx->ev_trail_a, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
  if ((e = cudaEventCreateWithFlags(&ctx->ev_sliced, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
  if ((e = cudaEventCreateWithFlags(&ctx->ev_group, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
  if ((e = cudaEventCreateWithFlags(&ctx->ev_p1b, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
  if ((e = cudaEventCreateWithFlags(&ctx->ev_sent, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
  if ((e = cudaEventCreateWithFlags(&ctx->ev_upd, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
  if (const char *env = getenv("GDCA_SHARE_MIN_NB")) {
    const int v = atoi(env);
    if (v >= 16) ctx->share_min_nb = v;
  }
  if ((e = cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming)) != cudaSuccess) return fail(e);
  if ((e = cudaStreamCreateWithFlags(&ctx->stream_copy, cudaStreamNonBlocking)) != cudaSuccess) return fail(e);
  if (const char *env = getenv("GDCA_OZAKI")) ctx->ozaki_mode = atoi(env) != 0;
  if (const char *env = getenv("GDCA_OZ_TPC")) {
    const int v = atoi(env);
    if (v >= -1024 && v <= 1024) ctx->ozaki_tpc = v;
  }
  if ((e = cudaMalloc((void **)&ctx->dHam, 2 * sizeof(unsigned long long))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void **)&ctx->dQ, 2 * sizeof(int))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void **)&ctx->dMeff, 2 * sizeof(double))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void **)&ctx->dInfo, sizeof(int))) != cudaSuccess) return fail(e);
  for (int i = 0; i < EV_COUNT; ++i)
    if ((e = cudaEventCreate(&ctx->ev[i])) != cudaSuccess) return fail(e);
  if ((e = cudaEventCreate(&ctx->ev[GDCA_EV_POTRF])) != cudaSuccess) return fail(e);
  if ((e = cudaEventCreate(&ctx->ev_sweep0)) != cudaSuccess) return fail(e);
  if ((e = cudaEventCreate(&ctx->ev_filter)) != cudaSuccess) return fail(e);
  if ((e = cudaEventCreate(&ctx->ev_sweep1)) != cudaSuccess) return fail(e);
  if ((e = cudaEventCreate(&ctx->ev_cov0)) != cudaSuccess) return fail(e);
  if ((e = cudaEventCreate(&ctx->ev_cov1)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void **)&ctx->dNItems, sizeof(unsigned long long))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void **)&ctx->dCovSync, sizeof(unsigned int))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void **)&ctx->dNPairs, sizeof(unsigned long long))) != cudaSuccess) return fail(e);
  if ((e = cudaMemset(ctx->dNPairs, 0, sizeof(unsigned long long))) != cudaSuccess) return fail(e);
  if (const char *env = getenv("GDCA_TC_FILTER")) {
    const int m = atoi(env);
    if (m >= 0 && m <= 2) ctx->tc_filter_mode = m;
  }
  if (const char *env = getenv("GDCA_TC_FILTER_BITS")) {
    const int b = atoi(env);
    if (b == 4 || b == 8 || b == 80) ctx->tc_filter_bits = b;
  }
  if (const char *env = getenv("GDCA_STAGED_H2D")) ctx->staged_h2d = atoi(env) != 0;
  if (const char *env = getenv("GDCA_CELL_SWEEP")) ctx->cell_sweep = atoi(env) != 0;
  if (const char *env = getenv("GDCA_PAIR_LIST")) ctx->pair_list = atoi(env) != 0;
  if (const char *env = getenv("GDCA_TC_MULTICAST")) {
    const int v = atoi(env);
    if (v >= 0 && v <= 2) ctx->tc_filter_want_multicast = v;
  }
  if (const char *env = getenv("GDCA_COV_ROUND_SYNC")) ctx->cov_round_sync = atoi(env);
  if (const char *env = getenv("GDCA_COV_ENGINE")) {
    const int m = atoi(env);
    if (m >= 0 && m <= 2) ctx->cov_engine = m;
  }
  *out = ctx;
  return GDCA_OK;
}

void gdca_destroy(gdca_ctx *ctx) {
  if (!ctx) return;
  if (ctx->group_size > 1 && ctx->group_rank == 0) {  // a leader takes its members with it
    for (int r = 1; r < ctx->group_size; ++r) {
      gdca_ctx *m = ctx->group[r];
      ctx->group[r] = nullptr;
      if (m) {
        m->group_size = 1;
        gdca_destroy(m);
      }
    }
    ctx->group_size = 1;
  }
  cudaSetDevice(ctx->device);
  for (void *m : ctx->peer_opened)
    if (m) cudaIpcCloseMemHandle(m);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  gdca_k_inverse_release(ctx);
  if (ctx->hostTab) cudaFreeHost(ctx->hostTab);
  gdca_h2d_release(ctx);
  if (!ctx->dZ_borrowed) cudaFree(ctx->dZ);
  void *bufs[] = {ctx->dZt,  ctx->dZq, ctx->dPerm, ctx->dPlanes, ctx->dCounts, ctx->dHam,  ctx->dQ,   ctx->dW,    ctx->dMeff, ctx->dList,
                  ctx->dListOff, ctx->dPi, ctx->dC,   ctx->dX,    ctx->dmJ,  ctx->dCdiag, ctx->dT,   ctx->dInfo,
                  ctx->dS,   ctx->dS2,     ctx->dRed, ctx->dKeys, ctx->dVals, ctx->dR,
                  ctx->dV,   ctx->dFlags,  ctx->dItems, ctx->dNItems, ctx->dItemMask,
                  ctx->dDigA, ctx->dDigB,  ctx->dScaleA, ctx->dScaleB, ctx->dOzMax, ctx->dCellBase, ctx->dDigP, ctx->dScaleP,
                  ctx->dClsHist, ctx->dClsPerm, ctx->dClsTab, ctx->dXt, ctx->dCovTiles, ctx->dSegCnt, ctx->dCovSync, ctx->dPairs, ctx->dNPairs};
  for (void *b : bufs)
    if (b) cudaFree(b);
  for (int i = 0; i < 16; ++i)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (cudaEvent_t e : {ctx->ev_sweep0, ctx->ev_filter, ctx->ev_sweep1, ctx->ev_cov0, ctx->ev_cov1})
    if (e) cudaEventDestroy(e);
  if (ctx->ev_fact) cudaEventDestroy(ctx->ev_fact);
  if (ctx->ev_trail) cudaEventDestroy(ctx->ev_trail);
  if (ctx->ev_trail_a) cudaEventDestroy(ctx->ev_trail_a);
  if (ctx->ev_sliced) cudaEventDestroy(ctx->ev_sliced);
  if (ctx->ev_group) cudaEventDestroy(ctx->ev_group);
  if (ctx->ev_p1b) cudaEventDestroy(ctx->ev_p1b);
  if (ctx->ev_sent) cudaEventDestroy(ctx->ev_sent);
  if (ctx->ev_upd) cudaEventDestroy(ctx->ev_upd);
  if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
  if (ctx->stream_copy) cudaStreamDestroy(ctx->stream_copy);
  for (cudaEvent_t e : {ctx->ev_diag, ctx->ev_p1, ctx->ev_u2a, ctx->ev_u2b})
    if (e) cudaEventDestroy(e);
  if (ctx->stream3) cudaStreamDestroy(ctx->stream3);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int32_t gdca_create_multi(gdca_ctx **out, const int32_t *devices, int32_t n) {
  if (!out) return GDCA_ERR_INVALID_ARG;
  *out = nullptr;
  if (!devices || n < 1 || n > GDCA_MAX_PEERS) {
    g_create_error = "gdca_create_multi: need 1 <= n <= 16 device ordinals";
    return GDCA_ERR_INVALID_ARG;
  }
  // Test mode (env GDCA_GROUP_ALLOW_SAME_DEVICE=1): the same ordinal may be listed several times -- the members then are
  // independent contexts that share one GPU, so every piece of the group path (broadcast, barriers, peer tables, shared levels,
  // stores into the other members' buffers) runs on a one-GPU box.  Without it a repeated ordinal is an error.
  const char *same_env = getenv("GDCA_GROUP_ALLOW_SAME_DEVICE");
  const bool allow_same = same_env && atoi(same_env) != 0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j)
      if (devices[i] == devices[j] && !allow_same) {
        g_create_error = "gdca_create_multi: a device ordinal is listed twice";
        return GDCA_ERR_INVALID_ARG;
      }
  gdca_ctx *g[GDCA_MAX_PEERS] = {};
  auto undo = [&]() {
    for (int i = 0; i < n; ++i)
      if (g[i]) gdca_destroy(g[i]);
  };
  for (int i = 0; i < n; ++i) {
    const int32_t st = gdca_create(&g[i], devices[i]);
    if (st != GDCA_OK) {
      undo();
      return st;
    }
  }
  // every member reads and writes every other member's buffers inside its kernels: peer access both ways for every pair
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      if (i == j || devices[i] == devices[j]) continue;
      int can = 0;
      cudaSetDevice(devices[i]);
      cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
      if (!can) {
        char b[128];
        snprintf(b, sizeof b, "gdca_create_multi: device %d cannot access device %d (no NVLink / P2P path)", devices[i], devices[j]);
        g_create_error = b;
        undo();
        return GDCA_ERR_NO_DEVICE;
      }
      const cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        g_create_error = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
        cudaGetLastError();
        undo();
        return GDCA_ERR_CUDA;
      }
      cudaGetLastError();
    }
  for (int i = 0; i < n; ++i) {
    g[0]->group[i] = g[i];
    g[i]->leader = g[0];
    g[i]->group_rank = i;
  }
  g[0]->group_size = n;
  *out = g[0];
  return GDCA_OK;
}

int32_t gdca_group_size(const gdca_ctx *ctx) { return ctx ? ctx->group_size : 0; }

int32_t gdca_set_shard(gdca_ctx *ctx, int32_t rank, int32_t world) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  if (world < 1 || rank < 0 || rank >= world) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "set_shard: need 0 <= rank < world");
  ctx->shard_rank = rank;
  ctx->shard_world = world;
  return GDCA_OK;
}

int64_t gdca_ranking_length(int64_t L, int64_t min_separation) {
  const int64_t d = L - min_separation;
  return d > 0 ? d * (d + 1) / 2 : 0;
}

int32_t gdca_theta_from_ident_sum(int64_t L, int64_t M, uint64_t ident, double *theta, int64_t *thresh) {
  if (L < 1 || M < 2) return GDCA_ERR_INVALID_ARG;
  // same IEEE operations, same order, as oracle theta_from_ident_sum
  const double meanfracid = ((double)ident / (double)L) / (0.5 * (double)M * (double)(M - 1));
  const double th = fmin(0.5, 0.38 * 0.32 / meanfracid);
  if (theta) *theta = th;
  if (thresh) *thresh = (int64_t)floor(th * (double)L);
  return GDCA_OK;
}

int32_t gdca_theta_from_ham_sum(int64_t L, int64_t M, uint64_t ham_sum, double *theta, int64_t *thresh,
                                uint64_t *ident_sum) {
  if (L < 1 || M < 2) return GDCA_ERR_INVALID_ARG;
  const uint64_t npairs = (uint64_t)M * (uint64_t)(M - 1) / 2;
  const uint64_t ident = npairs * (uint64_t)L - ham_sum;
  if (ident_sum) *ident_sum = ident;
  return gdca_theta_from_ident_sum(L, M, ident, theta, thresh);
}

int32_t gdca_dev_ident_sum(gdca_ctx *ctx, uint64_t *ident_sum) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  GDCA_TRY(set_device(ctx));
  unsigned long long v = 0;
  GDCA_TRY(gdca_k_ident_sum(ctx, &v));
  if (ident_sum) *ident_sum = v;
  return GDCA_OK;
}

// ------------------------------------------------------------------ device-resident stages
int32_t gdca_dev_load(gdca_ctx *ctx, const int8_t *Z_host, int64_t L, int64_t M) {
  GDCA_TRY(check_LM(ctx, Z_host, L, M));
  GDCA_TRY(set_device(ctx));
  if (ctx->dZ_borrowed) {
    ctx->dZ = nullptr;
    ctx->capZ = 0;
    ctx->dZ_borrowed = false;
  }
  GDCA_TRY(rec(ctx, EV_BEGIN));
  GDCA_TRY(gdca_reserve(ctx, ctx->dZ, ctx->capZ, (size_t)L * M + 16));
  GDCA_TRY(gdca_h2d(ctx, ctx->dZ, Z_host, (size_t)L * M, ctx->stream));  // pinned: one async copy; pageable: pipelined through a pinned ring
  GDCA_TRY(rec(ctx, EV_H2D));
  GDCA_TRY(load_common(ctx, L, M));
  GDCA_TRY(rec(ctx, EV_PACK));
  return GDCA_OK;
}

int32_t gdca_dev_load_resident(gdca_ctx *ctx, const int8_t *Z_dev, int64_t L, int64_t M) {
  GDCA_TRY(check_LM(ctx, Z_dev, L, M));
  GDCA_TRY(set_device(ctx));
  if (!ctx->dZ_borrowed && ctx->dZ) {
    GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    GDCA_CUDA(ctx, cudaFree(ctx->dZ));
  }
  ctx->dZ = const_cast<int8_t *>(Z_dev);
  ctx->capZ = 0;
  ctx->dZ_borrowed = true;
  GDCA_TRY(rec(ctx, EV_BEGIN));
  GDCA_TRY(rec(ctx, EV_H2D));
  GDCA_TRY(load_common(ctx, L, M));
  GDCA_TRY(rec(ctx, EV_PACK));
  return GDCA_OK;
}

int32_t gdca_dev_pair_pass(gdca_ctx *ctx, int32_t mode, int64_t thresh) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  GDCA_TRY(set_device(ctx));
  return gdca_k_pair_pass(ctx, mode, (int)thresh, 1);
}

int32_t gdca_set_tc_filter(gdca_ctx *ctx, int32_t mode) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  if (mode < 0 || mode > 2) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "set_tc_filter: mode must be 0, 1 or 2");
  ctx->tc_filter_mode = mode;
  return GDCA_OK;
}

int32_t gdca_set_pair_list(gdca_ctx *ctx, int32_t on) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  ctx->pair_list = on != 0;
  return GDCA_OK;
}

int32_t gdca_dev_pair_list_info(gdca_ctx *ctx, int64_t *candidates, int64_t *capacity) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  GDCA_TRY(set_device(ctx));
  unsigned long long n = 0;
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  GDCA_CUDA(ctx, cudaMemcpy(&n, ctx->dNPairs, sizeof n, cudaMemcpyDeviceToHost));
  if (candidates) *candidates = (int64_t)n;
  if (capacity) *capacity = (int64_t)ctx->pair_cap;
  return GDCA_OK;
}

int32_t gdca_set_tc_filter_bits(gdca_ctx *ctx, int32_t bits) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  if (bits != 4 && bits != 8 && bits != 80)
    return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "set_tc_filter_bits: bits must be 4 (e2m1), 8 (e4m3) or 80 (int8)");
  ctx->tc_filter_bits = bits;
  return GDCA_OK;
}

int32_t gdca_set_tc_filter_multicast(gdca_ctx *ctx, int32_t on) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  ctx->tc_filter_want_multicast = on < 0 ? 0 : (on > 2 ? 2 : on);
  return GDCA_OK;
}

int32_t gdca_dev_tc_filter(gdca_ctx *ctx, int64_t thresh, uint32_t *flags_host, float *S_host, int64_t ld) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  if (!flags_host) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "tc_filter: flags_host is NULL");
  GDCA_TRY(set_device(ctx));
  if (!ctx->have_alignment) return gdca_fail(ctx, GDCA_ERR_STATE, "tc_filter: no alignment loaded");
  const int64_t T = ctx->Mpad / GDCA_TILE, rows = T * GDCA_TILE;
  float *dS = nullptr;
  if (S_host) {
    if (ld < rows + 256) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "tc_filter: ld must be >= 128*T + 256");
    GDCA_CUDA(ctx, cudaMalloc((void **)&dS, (size_t)rows * ld * sizeof(float)));
    GDCA_CUDA(ctx, cudaMemsetAsync(dS, 0, (size_t)rows * ld * sizeof(float), ctx->stream));
  }
  int32_t st = gdca_k_tc_filter(ctx, (int)thresh, dS, ld);
  if (st == GDCA_OK) {
    cudaError_t e = cudaMemcpyAsync(flags_host, ctx->dFlags, (size_t)(T * T) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && dS)
      e = cudaMemcpyAsync(S_host, dS, (size_t)rows * ld * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      ctx->err = std::string("tc_filter: ") + cudaGetErrorString(e);
      st = GDCA_ERR_CUDA;
    }
  }
  if (dS) cudaFree(dS);
  return st;
}

int32_t gdca_dev_sweep_info(gdca_ctx *ctx, int32_t *filtered, int64_t *filter_tiles, double *filter_tflop,
                            int64_t *swept_blocks, float *ms_filter, float *ms_exact, double *filter_l2_bytes) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  GDCA_TRY(set_device(ctx));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const bool f = ctx->last_sweep_filtered;
  if (filtered) *filtered = f ? 1 : 0;
  if (filter_tiles) *filter_tiles = f ? ctx->tc_filter_tiles : 0;
  if (filter_tflop) *filter_tflop = f ? ctx->tc_filter_tflop : 0.0;
  if (filter_l2_bytes) *filter_l2_bytes = f ? ctx->tc_filter_l2_bytes : 0.0;
  if (filtered && f) *filtered = ctx->tc_filter_bits;
  const int64_t T = ctx->Mpad / GDCA_TILE;
  int64_t blocks = T * (T + 1) / 2 / ctx->shard_world;
  if (f) {
    int nit = 0;
    GDCA_CUDA(ctx, cudaMemcpy(&nit, ctx->dNItems, sizeof(int), cudaMemcpyDeviceToHost));  // low word of the packed counter
    blocks = nit;
  }
  if (swept_blocks) *swept_blocks = blocks;
  float a = 0.f, b = 0.f;
  if (f) {
    if (cudaEventElapsedTime(&a, ctx->ev_sweep0, ctx->ev_filter) != cudaSuccess) cudaGetLastError(), a = 0.f;
    if (cudaEventElapsedTime(&b, ctx->ev_filter, ctx->ev_sweep1) != cudaSuccess) cudaGetLastError(), b = 0.f;
  } else {
    if (cudaEventElapsedTime(&b, ctx->ev_sweep0, ctx->ev_sweep1) != cudaSuccess) cudaGetLastError(), b = 0.f;
  }
  if (ms_filter) *ms_filter = a;
  if (ms_exact) *ms_exact = b;
  return GDCA_OK;
}

int32_t gdca_dev_tc_filter_launch_mode(gdca_ctx *ctx) { return ctx ? ctx->tc_filter_multicast : -1; }

int32_t gdca_dev_cov_kernel_ms(gdca_ctx *ctx, float *ms) {
  if (!ctx || !ms) return GDCA_ERR_INVALID_ARG;
  GDCA_TRY(set_device(ctx));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (cudaEventElapsedTime(ms, ctx->ev_cov0, ctx->ev_cov1) != cudaSuccess) {
    cudaGetLastError();
    *ms = 0.f;
  }
  return GDCA_OK;
}

int32_t gdca_dev_pair_sample(gdca_ctx *ctx, int32_t stride) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  GDCA_TRY(set_device(ctx));
  return gdca_k_pair_pass(ctx, 0, 0, stride);
}

void *gdca_dev_ham_sum_ptr(gdca_ctx *ctx) { return ctx ? ctx->dHam : nullptr; }
void *gdca_dev_counts_ptr(gdca_ctx *ctx) { return ctx ? ctx->dCounts : nullptr; }
int64_t gdca_dev_counts_stride(gdca_ctx *ctx) { return ctx ? ctx->Mpad : 0; }
void *gdca_dev_C_ptr(gdca_ctx *ctx) { return ctx ? ctx->dC : nullptr; }
void *gdca_dev_mJ_ptr(gdca_ctx *ctx) { return ctx ? ctx->dmJ : nullptr; }
void *gdca_dev_S_ptr(gdca_ctx *ctx) { return ctx ? ctx->dS2 : nullptr; }
void *gdca_dev_W_ptr(gdca_ctx *ctx) { return ctx ? ctx->dW : nullptr; }
void *gdca_dev_stream(gdca_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int64_t gdca_dev_kernel_launches(gdca_ctx *ctx) { return ctx ? ctx->launches : 0; }
int64_t gdca_dev_npad(gdca_ctx *ctx) { return ctx ? ctx->npad : 0; }

int32_t gdca_dev_finish_weights(gdca_ctx *ctx, int32_t which, double *meff) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  GDCA_TRY(set_device(ctx));
  GDCA_TRY(gdca_k_finish_weights(ctx, which));
  ctx->stats.meff = ctx->meff;
  if (meff) *meff = ctx->meff;
  return GDCA_OK;
}

int32_t gdca_dev_set_weights(gdca_ctx *ctx, const double *W_host, double meff) {
  if (!ctx || !W_host) return GDCA_ERR_INVALID_ARG;
  if (!ctx->have_alignment) return gdca_fail(ctx, GDCA_ERR_STATE, "set_weights: no alignment loaded");
  GDCA_TRY(set_device(ctx));
  GDCA_TRY(gdca_reserve(ctx, ctx->dW, ctx->capW, (size_t)ctx->Mpad));
  GDCA_CUDA(ctx, cudaMemcpyAsync(ctx->dW, W_host, (size_t)ctx->M * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  const double mm[2] = {meff, 0.0};
  GDCA_CUDA(ctx, cudaMemcpyAsync(ctx->dMeff, mm, sizeof mm, cudaMemcpyHostToDevice, ctx->stream));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->meff = meff;
  ctx->stats.meff = meff;
  ctx->have_weights = true;
  ctx->weights_from_counts = false;  // arbitrary weights: the scatter-add covariance engine
  return GDCA_OK;
}

int32_t gdca_dev_covariance(gdca_ctx *ctx, double pseudocount) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  if (!(pseudocount >= 0.0 && pseudocount <= 1.0))
    return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "invalid pseudocount value (must be between 0 and 1)");
  GDCA_TRY(set_device(ctx));
  return gdca_k_covariance(ctx, pseudocount);
}

int32_t gdca_set_cov_engine(gdca_ctx *ctx, int32_t mode) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  if (mode < 0 || mode > 2) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "set_cov_engine: mode must be 0 (auto), 1 (scatter-add) or 2 (tensor cores)");
  ctx->cov_engine = mode;
  for (int r = 1; r < ctx->group_size; ++r)
    if (ctx->group[r]) ctx->group[r]->cov_engine = mode;
  return GDCA_OK;
}

int32_t gdca_dev_cov_info(gdca_ctx *ctx, int32_t *engine, int32_t *classes, int32_t *segments, int64_t *kblocks, int32_t *clusters,
                          double *tflop, double *l2_bytes) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  const bool tc = ctx->last_cov_engine == 2;
  if (engine) *engine = ctx->last_cov_engine;
  if (classes) *classes = ctx->cov_tc_classes;
  if (segments) *segments = tc ? ctx->cov_tc_segments : 0;
  if (kblocks) *kblocks = tc ? ctx->cov_tc_kblocks : 0;
  if (clusters) *clusters = tc ? ctx->cov_tc_clusters : 0;
  if (tflop) *tflop = tc ? ctx->cov_tc_tflop : 0.0;
  if (l2_bytes) *l2_bytes = tc ? ctx->cov_tc_l2_bytes : 0.0;
  return GDCA_OK;
}

int32_t gdca_dev_inverse(gdca_ctx *ctx, int32_t *info) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  GDCA_TRY(set_device(ctx));
  if (!ctx->have_cov) return gdca_fail(ctx, GDCA_ERR_STATE, "inverse: covariance not computed");
  GDCA_TRY(gdca_k_symmetrize_C(ctx));
  const int32_t st = gdca_k_inverse(ctx);
  if (info) *info = ctx->stats.posdef_info;
  return st;
}

int32_t gdca_set_ozaki(gdca_ctx *ctx, int32_t mode) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  if (mode != 0 && mode != 1) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "set_ozaki: mode must be 0 (DMMA only) or 1 (INT8-sliced tcgen05 GEMMs)");
  ctx->ozaki_mode = mode;
  return GDCA_OK;
}

int32_t gdca_set_di_engine(gdca_ctx *ctx, int32_t mode) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  if (mode != 0 && mode != 1)
    return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "set_di_engine: mode must be 0 (one-sided Jacobi) or 1 (tridiagonalisation + implicit QL)");
  ctx->di_engine = mode;
  return GDCA_OK;
}

int32_t gdca_dev_inverse_info(gdca_ctx *ctx, int32_t *ozaki, double *int8_ops, double *fp64_flop_on_int8) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  if (ozaki) *ozaki = ctx->last_inverse_ozaki ? 1 : 0;
  if (int8_ops) *int8_ops = ctx->oz_int8_ops;
  if (fp64_flop_on_int8) *fp64_flop_on_int8 = ctx->oz_fp64_flop;
  return GDCA_OK;
}

int32_t gdca_dev_inverse_shared(gdca_ctx *ctx) { return (ctx && ctx->last_inverse_shared) ? 1 : 0; }

int32_t gdca_dev_score_rank(gdca_ctx *ctx, int32_t score, int64_t min_separation, gdca_rank_t *R_host, int64_t R_len) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  GDCA_TRY(set_device(ctx));
  GDCA_TRY(gdca_k_score(ctx, score));
  GDCA_TRY(gdca_k_apc(ctx));
  GDCA_TRY(gdca_k_rank(ctx, min_separation, R_len));
  if (R_host && R_len > 0) {
    GDCA_CUDA(ctx, cudaMemcpyAsync(R_host, ctx->dR, (size_t)R_len * sizeof(gdca_rank_t), cudaMemcpyDeviceToHost, ctx->stream));
    GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return GDCA_OK;
}

// ------------------------------------------------------------------ peer memory (one process per GPU)
int32_t gdca_dev_peer_export(gdca_ctx *ctx, uint8_t *handles128) {
  if (!ctx || !handles128) return GDCA_ERR_INVALID_ARG;
  if (!ctx->have_alignment) return gdca_fail(ctx, GDCA_ERR_STATE, "peer_export: no alignment loaded");
  GDCA_TRY(set_device(ctx));
  // the buffers the peers will write into must exist (and keep their address) before the handles are taken
  GDCA_TRY(gdca_reserve(ctx, ctx->dCounts, ctx->capCounts, (size_t)3 * ctx->Mpad));
  GDCA_TRY(gdca_reserve(ctx, ctx->dC, ctx->capC, (size_t)ctx->npad * ctx->npad));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  GDCA_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->dCounts));
  memcpy(handles128, &h, 64);
  GDCA_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->dC));
  memcpy(handles128 + 64, &h, 64);
  ctx->exported_counts = ctx->dCounts;
  ctx->exported_C = ctx->dC;
  return GDCA_OK;
}

int32_t gdca_dev_peer_import(gdca_ctx *ctx, int32_t world, const uint8_t *handles /* world x 128 bytes */) {
  if (!ctx || !handles) return GDCA_ERR_INVALID_ARG;
  if (world < 2 || world > GDCA_MAX_PEERS || world != ctx->shard_world)
    return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "peer_import: world must equal the shard world (2..16)");
  GDCA_TRY(set_device(ctx));
  peer_close_all(ctx);
  for (int r = 0; r < world; ++r) {
    if (r == ctx->shard_rank) {
      ctx->peer_counts[r] = ctx->dCounts;
      ctx->peer_C[r] = ctx->dC;
      continue;
    }
    cudaIpcMemHandle_t h;
    void *p = nullptr;
    memcpy(&h, handles + (size_t)r * 128, 64);
    GDCA_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peer_opened[2 * r] = p;
    ctx->peer_counts[r] = (int32_t *)p;
    memcpy(&h, handles + (size_t)r * 128 + 64, 64);
    GDCA_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peer_opened[2 * r + 1] = p;
    ctx->peer_C[r] = (double *)p;
  }
  ctx->peers_ready = true;
  return GDCA_OK;
}

int32_t gdca_dev_peer_close(gdca_ctx *ctx) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  GDCA_TRY(set_device(ctx));
  return peer_close_all(ctx);
}

// 1 when the exported buffers are still the live ones (no re-export / re-import needed for this alignment)
int32_t gdca_dev_peer_valid(gdca_ctx *ctx) {
  if (!ctx) return 0;
  return (ctx->peers_ready && ctx->exported_counts == ctx->dCounts && ctx->exported_C == ctx->dC &&
          ctx->capCounts >= (size_t)3 * ctx->Mpad && ctx->capC >= (size_t)ctx->npad * ctx->npad) ? 1 : 0;
}

int32_t gdca_dev_zero_counts(gdca_ctx *ctx) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  GDCA_TRY(set_device(ctx));
  GDCA_TRY(gdca_reserve(ctx, ctx->dCounts, ctx->capCounts, (size_t)3 * ctx->Mpad));
  GDCA_CUDA(ctx, cudaMemsetAsync(ctx->dCounts, 0, (size_t)3 * ctx->Mpad * sizeof(int32_t), ctx->stream));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GDCA_OK;
}

int32_t gdca_dev_zero_C(gdca_ctx *ctx) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  GDCA_TRY(set_device(ctx));
  GDCA_TRY(gdca_reserve(ctx, ctx->dC, ctx->capC, (size_t)ctx->npad * ctx->npad));
  GDCA_CUDA(ctx, cudaMemsetAsync(ctx->dC, 0, (size_t)ctx->npad * ctx->npad * sizeof(double), ctx->stream));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GDCA_OK;
}

int32_t gdca_dev_sync(gdca_ctx *ctx) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  GDCA_TRY(set_device(ctx));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GDCA_OK;
}

int32_t gdca_dev_copy_to_host(gdca_ctx *ctx, void *dst_host, const void *src_dev, int64_t nbytes) {
  if (!ctx || !dst_host || !src_dev || nbytes < 0) return GDCA_ERR_INVALID_ARG;
  GDCA_TRY(set_device(ctx));
  GDCA_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, (size_t)nbytes, cudaMemcpyDeviceToHost, ctx->stream));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GDCA_OK;
}

int32_t gdca_dev_get_stats(gdca_ctx *ctx, gdca_stats_t *stats) {
  if (!ctx || !stats) return GDCA_ERR_INVALID_ARG;
  *stats = ctx->stats;
  return GDCA_OK;
}

// ------------------------------------------------------------------ fused run on a device group (one process, N GPUs)
namespace {

// Every member's stream waits until every member's stream has reached this point: N event records + N (N-1) stream waits,
// nothing blocks the host.  (A wait takes the record that precedes it in program order, so one event per member is reused.)
int32_t group_barrier(gdca_ctx *lead) {
  const int N = lead->group_size;
  for (int r = 0; r < N; ++r) {
    gdca_ctx *c = lead->group[r];
    GDCA_TRY(set_device(c));
    GDCA_CUDA(lead, cudaEventRecord(c->ev_group, c->stream));
  }
  for (int r = 0; r < N; ++r) {
    gdca_ctx *c = lead->group[r];
    GDCA_TRY(set_device(c));
    for (int p = 0; p < N; ++p)
      if (p != r) GDCA_CUDA(lead, cudaStreamWaitEvent(c->stream, lead->group[p]->ev_group, 0));
  }
  return GDCA_OK;
}

// a member's failure is reported through the leader
int32_t member_try(gdca_ctx *lead, gdca_ctx *c, int32_t st) {
  if (st != GDCA_OK && c != lead) lead->err = "device " + std::to_string(c->device) + ": " + c->err;
  return st;
}

void group_peer_tables(gdca_ctx *lead, bool on) {
  const int N = lead->group_size;
  for (int r = 0; r < N; ++r) {
    gdca_ctx *c = lead->group[r];
    for (int p = 0; p < GDCA_MAX_PEERS; ++p) {
      c->peer_counts[p] = (on && p < N) ? lead->group[p]->dCounts : nullptr;
      c->peer_C[p] = (on && p < N) ? lead->group[p]->dC : nullptr;
    }
    c->peers_ready = on;
    c->shard_rank = on ? r : 0;
    c->shard_world = on ? N : 1;
  }
}

}  // namespace


// The alignment reaches the leader once (one PCIe copy, or it is already resident there) and travels to the other members over
// NVLink along a binary tree; the pair sweep and the covariance run sharded with their exchange fused into the kernels (peer
// atomics / peer stores into the mapped buffers of the other devices); the inversion is sharded by gdca_k_inverse_group; scores,
// APC and ranking run on the leader.  Results are bit-identical to the single-GPU run (integer exchanges, disjoint writes).
static int32_t run_group(gdca_ctx *lead, const int8_t *Z, bool resident, int64_t L, int64_t M, double theta, double pseudocount,
                         int32_t score, int64_t min_separation, gdca_rank_t *R, int64_t R_len, gdca_stats_t *stats) {
  const int N = lead->group_size;
  gdca_ctx **g = lead->group;
  lead->stats = gdca_stats_t{};
  auto body = [&]() -> int32_t {
    // ---- 1. the alignment: H2D to the leader only, then a tree broadcast over NVLink
    GDCA_TRY(set_device(lead));
    if (resident) {
      if (!lead->dZ_borrowed && lead->dZ) {
        GDCA_CUDA(lead, cudaStreamSynchronize(lead->stream));
        GDCA_CUDA(lead, cudaFree(lead->dZ));
      }
      lead->dZ = const_cast<int8_t *>(Z);
      lead->capZ = 0;
      lead->dZ_borrowed = true;
      GDCA_TRY(rec(lead, EV_BEGIN));
    } else {
      if (lead->dZ_borrowed) {
        lead->dZ = nullptr;
        lead->capZ = 0;
        lead->dZ_borrowed = false;
      }
      GDCA_TRY(rec(lead, EV_BEGIN));
      GDCA_TRY(gdca_reserve(lead, lead->dZ, lead->capZ, (size_t)L * M + 16));
      GDCA_TRY(gdca_h2d(lead, lead->dZ, Z, (size_t)L * M, lead->stream));
    }
    GDCA_CUDA(lead, cudaEventRecord(lead->ev_group, lead->stream));  // "my copy is complete"
    for (int r = 1; r < N; ++r) {
      GDCA_TRY(set_device(g[r]));
      if (g[r]->dZ_borrowed) {
        g[r]->dZ = nullptr;
        g[r]->capZ = 0;
        g[r]->dZ_borrowed = false;
      }
      GDCA_TRY(member_try(lead, g[r], gdca_reserve(g[r], g[r]->dZ, g[r]->capZ, (size_t)L * M + 16)));
    }
    for (int s = 1; s < N; s <<= 1)          // round: members [0, s) send to [s, 2s)
      for (int src = 0; src < s && src + s < N; ++src) {
        gdca_ctx *d = g[src + s];
        GDCA_TRY(set_device(d));
        GDCA_CUDA(lead, cudaStreamWaitEvent(d->stream, g[src]->ev_group, 0));
        GDCA_CUDA(lead, cudaMemcpyPeerAsync(d->dZ, d->device, g[src]->dZ, g[src]->device, (size_t)L * M, d->stream));
        GDCA_CUDA(lead, cudaEventRecord(d->ev_group, d->stream));
      }
    GDCA_TRY(set_device(lead));
    GDCA_TRY(rec(lead, EV_H2D));
    // ---- 2. every member: q, per-site lists, bit planes (replicated: each needs all of it for its share of the pair matrix)
    for (int r = 0; r < N; ++r) {
      GDCA_TRY(set_device(g[r]));
      GDCA_TRY(member_try(lead, g[r], load_common(g[r], L, M)));
    }
    GDCA_TRY(set_device(lead));
    GDCA_TRY(rec(lead, EV_PACK));
    // ---- 3. theta (leader; O(M L) histogram sum), then the sharded neighbour-count sweep
    gdca_stats_t &st = lead->stats;
    int64_t thresh = 0;
    if (theta < 0) {
      unsigned long long ident = 0;
      GDCA_TRY(gdca_k_ident_sum(lead, &ident));
      double th;
      GDCA_TRY(gdca_theta_from_ident_sum(L, M, ident, &th, &thresh));
      st.theta = th;
      st.ident_sum = ident;
    } else {
      st.theta = theta;
      thresh = (int64_t)floor(theta * (double)L);
    }
    st.thresh = (st.theta == 0.0) ? 0 : thresh;
    GDCA_TRY(rec(lead, EV_THETA));
    for (int r = 0; r < N; ++r) {  // the buffers the peers write into must exist before their addresses go into the tables
      GDCA_TRY(set_device(g[r]));
      GDCA_TRY(member_try(lead, g[r], gdca_reserve(g[r], g[r]->dCounts, g[r]->capCounts, (size_t)3 * g[r]->Mpad)));
      GDCA_TRY(member_try(lead, g[r], gdca_reserve(g[r], g[r]->dC, g[r]->capC, (size_t)g[r]->npad * g[r]->npad)));
    }
    group_peer_tables(lead, true);
    if (st.theta == 0.0) {
      for (int r = 0; r < N; ++r) {
        GDCA_TRY(set_device(g[r]));
        GDCA_TRY(member_try(lead, g[r], gdca_k_finish_weights(g[r], -1)));
      }
    } else {
      for (int r = 0; r < N; ++r) {
        GDCA_TRY(set_device(g[r]));
        GDCA_CUDA(lead, cudaMemsetAsync(g[r]->dCounts, 0, (size_t)3 * g[r]->Mpad * sizeof(int32_t), g[r]->stream));
      }
      GDCA_TRY(group_barrier(lead));  // every member's counters are zero before anyone adds
      for (int r = 0; r < N; ++r) {
        GDCA_TRY(set_device(g[r]));
        GDCA_TRY(member_try(lead, g[r], gdca_k_pair_pass(g[r], 1, (int)thresh, 1)));
      }
      GDCA_TRY(group_barrier(lead));  // all peer atomics have landed
      st.theta_passes = 1;
      for (int r = 0; r < N; ++r) {
        GDCA_TRY(set_device(g[r]));
        GDCA_TRY(member_try(lead, g[r], gdca_k_finish_weights(g[r], 0)));
      }
    }
    st.meff = lead->meff;
    GDCA_TRY(set_device(lead));
    GDCA_TRY(rec(lead, EV_WEIGHTS));
    // ---- 4. covariance: rows dealt by site, every member stores its rows straight into the leader's C
    GDCA_CUDA(lead, cudaMemsetAsync(lead->dC, 0, (size_t)lead->npad * lead->npad * sizeof(double), lead->stream));
    GDCA_TRY(group_barrier(lead));
    // the weight classes of the tensor-core engine are the same on every member (every member holds all counts): one device
    // round trip on the leader, the members plan from its copy
    std::vector<int32_t> cls(1 + 2 * GDCA_COV_MAXCLS);
    const bool share_cls = lead->cov_engine != 1 && lead->weights_from_counts;
    if (share_cls) {
      GDCA_TRY(set_device(lead));
      GDCA_TRY(gdca_k_cov_classes(lead, cls.data()));
    }
    for (int r = 0; r < N; ++r) {
      GDCA_TRY(set_device(g[r]));
      g[r]->cov_cls_host = share_cls ? cls.data() : nullptr;
      const int32_t cst = member_try(lead, g[r], gdca_k_covariance(g[r], pseudocount));
      g[r]->cov_cls_host = nullptr;
      GDCA_TRY(cst);
    }
    GDCA_TRY(group_barrier(lead));
    GDCA_TRY(set_device(lead));
    GDCA_TRY(gdca_k_symmetrize_C(lead));
    GDCA_TRY(rec(lead, EV_COV));
    // ---- 5. inversion, sharded; scores, APC, ranking on the leader
    GDCA_TRY(gdca_k_inverse_group(lead));
    GDCA_TRY(set_device(lead));
    GDCA_TRY(rec(lead, EV_CHOL));
    GDCA_TRY(gdca_k_score(lead, score));
    GDCA_TRY(rec(lead, EV_SCORE));
    GDCA_TRY(gdca_k_apc(lead));
    GDCA_TRY(rec(lead, EV_APC));
    GDCA_TRY(gdca_k_rank(lead, min_separation, R_len));
    GDCA_TRY(rec(lead, EV_RANK));
    if (R_len > 0 && R)
      GDCA_CUDA(lead, cudaMemcpyAsync(R, lead->dR, (size_t)R_len * sizeof(gdca_rank_t), cudaMemcpyDeviceToHost, lead->stream));
    GDCA_TRY(rec(lead, EV_D2H));
    GDCA_CUDA(lead, cudaStreamSynchronize(lead->stream));
    return GDCA_OK;
  };
  const int32_t status = body();
  // leave every member as a plain single-device context
  for (int r = 0; r < N; ++r) {
    cudaSetDevice(g[r]->device);
    cudaStreamSynchronize(g[r]->stream);
    cudaStreamSynchronize(g[r]->stream_copy);
  }
  group_peer_tables(lead, false);
  for (int r = 0; r < N; ++r) drop_alignment_state(g[r]);
  cudaSetDevice(lead->device);
  gdca_stats_t &st = lead->stats;
  if (status == GDCA_OK) {
    st.ms_h2d = ev_ms(lead, EV_BEGIN, EV_H2D);
    st.ms_pack = ev_ms(lead, EV_H2D, EV_PACK);
    st.ms_theta = ev_ms(lead, EV_PACK, EV_THETA);
    st.ms_weights = ev_ms(lead, EV_THETA, EV_WEIGHTS);
    st.ms_cov = ev_ms(lead, EV_WEIGHTS, EV_COV);
    st.ms_chol = ev_ms(lead, EV_COV, GDCA_EV_POTRF);
    st.ms_inv = ev_ms(lead, GDCA_EV_POTRF, EV_CHOL);
    st.ms_score = ev_ms(lead, EV_CHOL, EV_SCORE);
    st.ms_apc = ev_ms(lead, EV_SCORE, EV_APC);
    st.ms_rank = ev_ms(lead, EV_APC, EV_RANK);
    st.ms_d2h = ev_ms(lead, EV_RANK, EV_D2H);
    st.ms_total = ev_ms(lead, EV_BEGIN, EV_D2H);
  }
  if (stats) *stats = st;
  return status;
}

// ------------------------------------------------------------------ fused run
static int32_t run_impl(gdca_ctx *ctx, const int8_t *Z, bool resident, int64_t L, int64_t M, double theta,
                        double pseudocount, int32_t score, int64_t min_separation, gdca_rank_t *R, int64_t R_len,
                        gdca_stats_t *stats) {
  GDCA_TRY(check_LM(ctx, Z, L, M));
  // the reference's check_arguments ranges (src/GaussDCA.jl:49-65); theta < 0 encodes :auto
  if (!(pseudocount >= 0.0 && pseudocount <= 1.0))
    return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "invalid pseudocount value (must be between 0 and 1)");
  if (!(theta < 0.0) && !(theta >= 0.0 && theta <= 1.0))
    return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "invalid theta value (must be :auto, or a number between 0 and 1)");
  if (score != GDCA_SCORE_FROB && score != GDCA_SCORE_DI)
    return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "invalid score value (must be either :DI or :frob)");
  if (min_separation < 1) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "invalid min_separation value (must be >= 1)");
  if (R_len != gdca_ranking_length(L, min_separation))
    return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "R_len != (L-min_separation)*(L-min_separation+1)/2");
  if (R_len > 0 && !R && !resident) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "R is NULL");
  if (M < 2 && theta < 0) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "theta = :auto needs at least 2 sequences");
  if (ctx->group_size > 1) return run_group(ctx, Z, resident, L, M, theta, pseudocount, score, min_separation, R, R_len, stats);
  ctx->stats = gdca_stats_t{};
  const int32_t saved_rank = ctx->shard_rank, saved_world = ctx->shard_world;
  ctx->shard_rank = 0;
  ctx->shard_world = 1;
  auto body = [&]() -> int32_t {
    if (resident)
      GDCA_TRY(gdca_dev_load_resident(ctx, Z, L, M));
    else
      GDCA_TRY(gdca_dev_load(ctx, Z, L, M));
    GDCA_TRY(weights_stage(ctx, theta));
    GDCA_TRY(gdca_k_covariance(ctx, pseudocount));
    GDCA_TRY(gdca_k_symmetrize_C(ctx));
    GDCA_TRY(rec(ctx, EV_COV));
    GDCA_TRY(gdca_k_inverse(ctx));
    GDCA_TRY(rec(ctx, EV_CHOL));
    GDCA_TRY(gdca_k_score(ctx, score));
    GDCA_TRY(rec(ctx, EV_SCORE));
    GDCA_TRY(gdca_k_apc(ctx));
    GDCA_TRY(rec(ctx, EV_APC));
    GDCA_TRY(gdca_k_rank(ctx, min_separation, R_len));
    GDCA_TRY(rec(ctx, EV_RANK));
    if (R_len > 0 && R)
      GDCA_CUDA(ctx, cudaMemcpyAsync(R, ctx->dR, (size_t)R_len * sizeof(gdca_rank_t), cudaMemcpyDeviceToHost, ctx->stream));
    GDCA_TRY(rec(ctx, EV_D2H));
    GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GDCA_OK;
  };
  const int32_t status = body();
  ctx->shard_rank = saved_rank;
  ctx->shard_world = saved_world;
  // a resident run borrows the caller's device buffer only for the duration of the call (the caller may free it afterwards)
  if (resident) drop_alignment_state(ctx);
  gdca_stats_t &st = ctx->stats;
  if (status == GDCA_OK) {
    st.ms_h2d = ev_ms(ctx, EV_BEGIN, EV_H2D);
    st.ms_pack = ev_ms(ctx, EV_H2D, EV_PACK);
    st.ms_theta = ev_ms(ctx, EV_PACK, EV_THETA);
    st.ms_weights = ev_ms(ctx, EV_THETA, EV_WEIGHTS);
    st.ms_cov = ev_ms(ctx, EV_WEIGHTS, EV_COV);
    st.ms_chol = ev_ms(ctx, EV_COV, GDCA_EV_POTRF);   // blocked Cholesky factorisation
    st.ms_inv = ev_ms(ctx, GDCA_EV_POTRF, EV_CHOL);   // trtri + lauum + mirror
    st.ms_score = ev_ms(ctx, EV_CHOL, EV_SCORE);
    st.ms_apc = ev_ms(ctx, EV_SCORE, EV_APC);
    st.ms_rank = ev_ms(ctx, EV_APC, EV_RANK);
    st.ms_d2h = ev_ms(ctx, EV_RANK, EV_D2H);
    st.ms_total = ev_ms(ctx, EV_BEGIN, EV_D2H);
  }
  if (stats) *stats = st;
  return status;
}

int32_t gdca_run(gdca_ctx *ctx, const int8_t *Z, int64_t L, int64_t M, double theta, double pseudocount, int32_t score,
                 int64_t min_separation, gdca_rank_t *R, int64_t R_len, gdca_stats_t *stats) {
  return run_impl(ctx, Z, false, L, M, theta, pseudocount, score, min_separation, R, R_len, stats);
}

int32_t gdca_run_resident(gdca_ctx *ctx, const int8_t *Z_dev, int64_t L, int64_t M, double theta, double pseudocount,
                          int32_t score, int64_t min_separation, gdca_rank_t *R_host_or_null, int64_t R_len,
                          gdca_stats_t *stats) {
  return run_impl(ctx, Z_dev, true, L, M, theta, pseudocount, score, min_separation, R_host_or_null, R_len, stats);
}

void *gdca_dev_R_ptr(gdca_ctx *ctx) { return ctx ? ctx->dR : nullptr; }

// ------------------------------------------------------------------ staged, host buffers
int32_t gdca_compute_weights(gdca_ctx *ctx, const int8_t *Z, int64_t L, int64_t M, double theta, int32_t *counts,
                             double *W, double *meff, double *theta_used, int64_t *thresh, uint64_t *ident_sum) {
  GDCA_TRY(check_LM(ctx, Z, L, M));
  if (!(theta < 0.0) && !(theta >= 0.0 && theta <= 1.0))
    return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "invalid theta value (must be :auto, or a number between 0 and 1)");
  if (M < 2 && theta < 0) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "theta = :auto needs at least 2 sequences");
  ctx->stats = gdca_stats_t{};
  GDCA_TRY(gdca_dev_load(ctx, Z, L, M));
  GDCA_TRY(weights_stage(ctx, theta));
  if (W) GDCA_CUDA(ctx, cudaMemcpyAsync(W, ctx->dW, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (counts) {
    if (ctx->counts_row < 0) {
      for (int64_t k = 0; k < M; ++k) counts[k] = 1;
    } else {
      GDCA_CUDA(ctx, cudaMemcpy(counts, ctx->dCounts + (size_t)ctx->counts_row * ctx->Mpad, (size_t)M * sizeof(int32_t),
                                cudaMemcpyDeviceToHost));
      for (int64_t k = 0; k < M; ++k) counts[k] += 1;  // the sequence itself
    }
  }
  if (meff) *meff = ctx->stats.meff;
  if (theta_used) *theta_used = ctx->stats.theta;
  if (thresh) *thresh = ctx->stats.thresh;
  if (ident_sum) *ident_sum = ctx->stats.ident_sum;
  return GDCA_OK;
}

int32_t gdca_compute_covariance(gdca_ctx *ctx, const int8_t *Z, int64_t L, int64_t M, const double *W, double meff,
                                double pseudocount, double *C, double *Pi, int32_t *q_out) {
  GDCA_TRY(check_LM(ctx, Z, L, M));
  if (!W || !C) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "W and C must not be NULL");
  if (!(pseudocount >= 0.0 && pseudocount <= 1.0))
    return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "invalid pseudocount value (must be between 0 and 1)");
  GDCA_TRY(gdca_dev_load(ctx, Z, L, M));
  GDCA_TRY(gdca_dev_set_weights(ctx, W, meff));
  GDCA_TRY(gdca_k_covariance(ctx, pseudocount));
  GDCA_TRY(gdca_k_symmetrize_C(ctx));
  if (Pi) GDCA_CUDA(ctx, cudaMemcpyAsync(Pi, ctx->dPi, (size_t)ctx->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GDCA_TRY(download_padded(ctx, C, ctx->dC, ctx->npad, ctx->n));
  if (q_out) *q_out = ctx->q;
  return GDCA_OK;
}

int32_t gdca_compute_weighted_frequencies(gdca_ctx *ctx, const int8_t *Z, int64_t L, int64_t M, double theta, double *Pi_true,
                                          double *Pij_true, double *meff, double *W, double *theta_used, int32_t *q_out) {
  GDCA_TRY(check_LM(ctx, Z, L, M));
  if (!Pi_true || !Pij_true) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "Pi_true and Pij_true must not be NULL");
  if (!(theta < 0.0) && !(theta >= 0.0 && theta <= 1.0))
    return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "invalid theta value (must be :auto, or a number between 0 and 1)");
  if (M < 2 && theta < 0) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "theta = :auto needs at least 2 sequences");
  ctx->stats = gdca_stats_t{};
  GDCA_TRY(gdca_dev_load(ctx, Z, L, M));
  GDCA_TRY(weights_stage(ctx, theta));
  GDCA_TRY(gdca_k_covariance(ctx, 0.0, /*raw=*/true));  // dPi = Pi_true, dC = Pij_true (upper site blocks)
  GDCA_TRY(gdca_k_symmetrize_C(ctx));
  ctx->have_cov = false;  // dC holds frequencies, not a covariance
  if (W) GDCA_CUDA(ctx, cudaMemcpyAsync(W, ctx->dW, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GDCA_CUDA(ctx, cudaMemcpyAsync(Pi_true, ctx->dPi, (size_t)ctx->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GDCA_TRY(download_padded(ctx, Pij_true, ctx->dC, ctx->npad, ctx->n));
  if (meff) *meff = ctx->stats.meff;
  if (theta_used) *theta_used = ctx->stats.theta;
  if (q_out) *q_out = ctx->q;
  return GDCA_OK;
}

int32_t gdca_add_pseudocount(gdca_ctx *ctx, const double *Pi_true, const double *Pij_true, int64_t n, int32_t q, double pseudocount,
                             double *Pi, double *Pij) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  if (!Pi_true || !Pij_true || !Pi || !Pij || n < 1 || q < 2 || n % (q - 1) != 0)
    return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "add_pseudocount: NULL buffer, or n is not a multiple of q-1");
  if (!(pseudocount >= 0.0 && pseudocount <= 1.0))
    return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "invalid pseudocount value (must be between 0 and 1)");
  GDCA_TRY(set_device(ctx));
  const size_t nn = (size_t)n * n;
  GDCA_TRY(gdca_reserve(ctx, ctx->dX, ctx->capX, nn));
  GDCA_TRY(gdca_reserve(ctx, ctx->dT, ctx->capT, nn));
  GDCA_TRY(gdca_reserve(ctx, ctx->dPi, ctx->capPi, (size_t)2 * n));
  GDCA_CUDA(ctx, cudaMemcpyAsync(ctx->dX, Pij_true, nn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GDCA_CUDA(ctx, cudaMemcpyAsync(ctx->dPi, Pi_true, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GDCA_TRY(gdca_k_add_pseudocount(ctx, ctx->dPi, ctx->dX, n, q, pseudocount, ctx->dPi + n, ctx->dT));
  GDCA_CUDA(ctx, cudaMemcpyAsync(Pi, ctx->dPi + n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GDCA_CUDA(ctx, cudaMemcpyAsync(Pij, ctx->dT, nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->have_cov = ctx->have_inv = false;  // scratch buffers reused
  return GDCA_OK;
}

int32_t gdca_compute_C(gdca_ctx *ctx, const double *Pi, const double *Pij, int64_t n, double *C) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  if (!Pi || !Pij || !C || n < 1) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "compute_C: Pi, Pij, C must not be NULL and n >= 1");
  GDCA_TRY(set_device(ctx));
  const size_t nn = (size_t)n * n;
  GDCA_TRY(gdca_reserve(ctx, ctx->dX, ctx->capX, nn));
  GDCA_TRY(gdca_reserve(ctx, ctx->dT, ctx->capT, nn));
  GDCA_TRY(gdca_reserve(ctx, ctx->dPi, ctx->capPi, (size_t)n));
  GDCA_CUDA(ctx, cudaMemcpyAsync(ctx->dX, Pij, nn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GDCA_CUDA(ctx, cudaMemcpyAsync(ctx->dPi, Pi, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GDCA_TRY(gdca_k_compute_C(ctx, ctx->dPi, ctx->dX, n, ctx->dT));
  GDCA_CUDA(ctx, cudaMemcpyAsync(C, ctx->dT, nn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->have_cov = ctx->have_inv = false;
  return GDCA_OK;
}

int32_t gdca_inverse(gdca_ctx *ctx, const double *C, int64_t n, double *mJ, int32_t *info) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  if (!C || !mJ || n < 1) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "inverse: C, mJ must not be NULL and n >= 1");
  GDCA_TRY(set_device(ctx));
  drop_alignment_state(ctx);
  ctx->n = n;
  ctx->npad = (n + GDCA_NB - 1) / GDCA_NB * GDCA_NB;
  GDCA_TRY(gdca_reserve(ctx, ctx->dC, ctx->capC, (size_t)ctx->npad * ctx->npad));
  GDCA_TRY(upload_padded(ctx, ctx->dC, ctx->npad, C, n));
  ctx->have_cov = true;
  const int32_t st = gdca_k_inverse(ctx);
  if (info) *info = ctx->stats.posdef_info;
  if (st != GDCA_OK) return st;
  return download_padded(ctx, mJ, ctx->dmJ, ctx->npad, n);
}

int32_t gdca_score(gdca_ctx *ctx, const double *mJ, const double *C, int64_t n, int32_t q, int32_t score, double *S) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  if (!mJ || !S) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "score: mJ and S must not be NULL");
  if (q < 2 || q >= 32 || n < 1 || n % (q - 1) != 0)
    return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "score: need 2 <= q <= 31 and n divisible by q-1");
  if (score == GDCA_SCORE_DI && !C) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "score: DI needs C");
  GDCA_TRY(set_device(ctx));
  drop_alignment_state(ctx);
  ctx->n = n;
  ctx->npad = (n + GDCA_NB - 1) / GDCA_NB * GDCA_NB;
  ctx->q = q;
  ctx->s = q - 1;
  ctx->L = n / (q - 1);
  GDCA_TRY(gdca_reserve(ctx, ctx->dmJ, ctx->capmJ, (size_t)ctx->npad * ctx->npad));
  GDCA_TRY(upload_padded(ctx, ctx->dmJ, ctx->npad, mJ, n));
  if (score == GDCA_SCORE_DI) {
    GDCA_TRY(gdca_reserve(ctx, ctx->dC, ctx->capC, (size_t)ctx->npad * ctx->npad));
    GDCA_TRY(upload_padded(ctx, ctx->dC, ctx->npad, C, n));
    GDCA_TRY(gdca_k_extract_diag(ctx));
  }
  ctx->have_inv = true;
  ctx->have_cov = false;
  GDCA_TRY(gdca_k_score(ctx, score));
  GDCA_CUDA(ctx, cudaMemcpyAsync(S, ctx->dS, (size_t)ctx->L * ctx->L * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GDCA_OK;
}

int32_t gdca_apc(gdca_ctx *ctx, const double *S, int64_t L, double *S_out) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  if (!S || !S_out || L < 2) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "apc: S, S_out must not be NULL and L >= 2");
  GDCA_TRY(set_device(ctx));
  drop_alignment_state(ctx);
  ctx->L = L;
  GDCA_TRY(gdca_reserve(ctx, ctx->dS, ctx->capS, (size_t)L * L));
  GDCA_CUDA(ctx, cudaMemcpyAsync(ctx->dS, S, (size_t)L * L * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GDCA_TRY(gdca_k_apc(ctx));
  GDCA_CUDA(ctx, cudaMemcpyAsync(S_out, ctx->dS2, (size_t)L * L * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GDCA_OK;
}

int32_t gdca_ranking(gdca_ctx *ctx, const double *S, int64_t L, int64_t min_separation, gdca_rank_t *R, int64_t R_len) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  if (!S || L < 1) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "ranking: S must not be NULL and L >= 1");
  if (R_len > 0 && !R) return gdca_fail(ctx, GDCA_ERR_INVALID_ARG, "R is NULL");
  GDCA_TRY(set_device(ctx));
  drop_alignment_state(ctx);
  ctx->L = L;
  GDCA_TRY(gdca_reserve(ctx, ctx->dS2, ctx->capS2, (size_t)L * L));
  GDCA_CUDA(ctx, cudaMemcpyAsync(ctx->dS2, S, (size_t)L * L * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GDCA_TRY(gdca_k_rank(ctx, min_separation, R_len));
  if (R_len > 0)
    GDCA_CUDA(ctx, cudaMemcpyAsync(R, ctx->dR, (size_t)R_len * sizeof(gdca_rank_t), cudaMemcpyDeviceToHost, ctx->stream));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GDCA_OK;
}

// ------------------------------------------------------------------ synthetic data, probes
int32_t gdca_synth_alignment_dev(gdca_ctx *ctx, int8_t *Z_dev, int64_t L, int64_t M, uint64_t seed) {
  GDCA_TRY(check_LM(ctx, Z_dev, L, M));
  GDCA_TRY(set_device(ctx));
  GDCA_TRY(gdca_k_synth(ctx, Z_dev, L, M, seed));
  GDCA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return GDCA_OK;
}

int32_t gdca_synth_alignment(gdca_ctx *ctx, int8_t *Z_host, int64_t L, int64_t M, uint64_t seed) {
  GDCA_TRY(check_LM(ctx, Z_host, L, M));
  GDCA_TRY(set_device(ctx));
  int8_t *tmp = nullptr;
  GDCA_CUDA(ctx, cudaMalloc((void **)&tmp, (size_t)L * M));
  int32_t st = gdca_k_synth(ctx, tmp, L, M, seed);
  if (st == GDCA_OK) {
    cudaError_t e = cudaMemcpyAsync(Z_host, tmp, (size_t)L * M, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      ctx->err = cudaGetErrorString(e);
      st = GDCA_ERR_CUDA;
    }
  }
  cudaFree(tmp);
  return st;
}

int32_t gdca_probe_peaks(gdca_ctx *ctx, double *lop3_tops, double *popc_tops, double *dmma_tflops, double *dfma_tflops) {
  if (!ctx) return GDCA_ERR_INVALID_ARG;
  GDCA_TRY(set_device(ctx));
  return gdca_k_probe(ctx, lop3_tops, popc_tops, dmma_tflops, dfma_tflops);
}

}  // extern "C"
