// stage_copy.cpp -- host-to-device copy of the alignment from PAGEABLE host memory.
//
// The hosts of this library hand it ordinary heap memory: a Julia Matrix{Int8} (GC-managed, src/GaussDCA.jl:20-24) or a numpy
// array.  cudaMemcpyAsync from such memory is staged by the driver through one internal buffer on one thread (~11 GB/s: 9 ms for
// the 100 MB of config C, 15 % of the whole step).  Here the copy is pipelined by the library instead: the source is cut into
// chunks, each chunk is copied by several host threads into a ring of pinned buffers and sent on the stream while the threads
// already fill the next buffer.  Pinned or registered sources (cudaPointerGetAttributes) take the direct path.
#include <cuda_runtime.h>
#include <omp.h>
#include <string.h>

#include "gdca_internal.cuh"

namespace {
constexpr size_t CHUNK = (size_t)8 << 20;   // 8 MiB per ring slot
}

int32_t gdca_h2d(gdca_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t stream) {
  cudaPointerAttributes attr{};
  const cudaError_t pe = cudaPointerGetAttributes(&attr, src);
  if (pe != cudaSuccess) cudaGetLastError();
  const bool pageable = (pe != cudaSuccess) || attr.type == cudaMemoryTypeUnregistered;
  if (!pageable || bytes < 4 * CHUNK || !ctx->staged_h2d) {
    GDCA_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    return GDCA_OK;
  }
  for (int i = 0; i < GDCA_STAGE_SLOTS; ++i) {
    if (!ctx->stage_buf[i]) GDCA_CUDA(ctx, cudaHostAlloc(&ctx->stage_buf[i], CHUNK, cudaHostAllocDefault));
    if (!ctx->stage_ev[i]) GDCA_CUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming));
  }
  int threads = omp_get_num_procs();
  if (threads > 8) threads = 8;
  if (threads < 1) threads = 1;
  const size_t nchunks = (bytes + CHUNK - 1) / CHUNK;
  for (size_t c = 0; c < nchunks; ++c) {
    const int slot = (int)(c % GDCA_STAGE_SLOTS);
    const size_t off = c * CHUNK, len = (off + CHUNK <= bytes) ? CHUNK : bytes - off;
    if (c >= GDCA_STAGE_SLOTS) GDCA_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[slot]));  // its previous transfer has left the buffer
    char *b = static_cast<char *>(ctx->stage_buf[slot]);
    const char *s = static_cast<const char *>(src) + off;
    const size_t piece = (len + threads - 1) / threads;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int t = 0; t < threads; ++t) {
      const size_t o = (size_t)t * piece;
      if (o < len) memcpy(b + o, s + o, (o + piece <= len) ? piece : len - o);
    }
    GDCA_CUDA(ctx, cudaMemcpyAsync(static_cast<char *>(dst) + off, b, len, cudaMemcpyHostToDevice, stream));
    GDCA_CUDA(ctx, cudaEventRecord(ctx->stage_ev[slot], stream));
  }
  return GDCA_OK;
}

void gdca_h2d_release(gdca_ctx *ctx) {
  for (int i = 0; i < GDCA_STAGE_SLOTS; ++i) {
    if (ctx->stage_buf[i]) cudaFreeHost(ctx->stage_buf[i]);
    if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
    ctx->stage_buf[i] = nullptr;
    ctx->stage_ev[i] = nullptr;
  }
}
