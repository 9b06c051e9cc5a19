"""One-process-per-GPU driver for the stages of gDCA that shard.

The reference has no multi-device path (one Julia process, shared-memory threads, README.md:92-94);
this is new.  Work partition, chosen so that every exchanged quantity is an exact integer or a
sum with zeros (results are bit-identical for any number of ranks):

  theta :auto  no exchange: every rank holds the alignment and computes the exact identity sum from per-site
               state histograms (O(M L)).
  pair sweep   rank r owns the 128-sequence row blocks bi = r (mod world) of the upper-triangular M x M pair matrix:
               its tiles of the tensor-core prefilter and the exact sweep of the blocks they flag (without the
               prefilter: tiles bi <= bj dealt round-robin); every rank holds the whole packed alignment.
               Exchange FUSED into the kernel: the (rare) neighbour hits are added into
               every rank's int32 counters with peer atomics over NVLink (CUDA-IPC mapped buffers).
  covariance   output rows dealt to ranks by site (i mod world).  Exchange FUSED into the kernel: every rank
               stores its rows straight into rank 0's C over NVLink (disjoint rows, no reduction needed).
  inverse, scores, APC, ranking: rank 0 (n^3 flop on one GPU; the block scores are HBM-bound
               microseconds).  Other ranks wait at the closing barrier.

`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is plumbing only: it carries the 128-byte IPC
handles once per context and the barriers; no collective touches the data path.

The control flow lives in `run_sharded`, written against a small backend interface so that the
host logic (speculative threshold, row selection, who does what) is testable on CPU with gloo.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np


def theta_from_ident(L: int, M: int, ident: int):
    """Host arithmetic of gdca_theta_from_ident_sum (same IEEE operations in the same order)."""
    meanfracid = (ident / L) / (0.5 * M * (M - 1))
    theta = 0.5 if meanfracid == 0.0 else min(0.5, 0.38 * 0.32 / meanfracid)
    return theta, int(math.floor(theta * L))


def theta_from_ham(L: int, M: int, ham_sum: int):
    """Same, from the hamming sum of a mode-0/2 sweep."""
    ident = M * (M - 1) // 2 * L - ham_sum
    return theta_from_ident(L, M, ident) + (ident,)


def run_sharded(be, dist, L: int, M: int, theta, pseudocount: float, score: str, min_separation: int):
    """gDCA from the loaded alignment on `dist.get_world_size()` ranks.  `be` is a backend (below).
    Returns (R or None, info) -- R on rank 0 only."""
    rank, world = dist.get_rank(), dist.get_world_size()
    be.set_shard(rank, world)
    info = {"passes": 0}
    if isinstance(theta, str):
        # theta = :auto needs no exchange: every rank holds the alignment and gets the exact identity sum
        # from the per-site state histograms (O(M L)).
        ident = be.ident_sum()
        th, thresh = theta_from_ident(L, M, ident)
        info["ident_sum"] = ident
    else:
        th = float(theta)
        thresh = int(math.floor(th * L))
    info.update(theta=th, thresh=thresh)
    if th == 0.0:
        which = -1                                # W == 1, Meff == M: no sweep at all
    else:
        # this rank's tiles of the M x M pair matrix; afterwards EVERY rank holds the complete integer counts
        # (GPU: peer atomics inside the sweep kernel; exact, independent of the rank count)
        be.sweep_counts(thresh, dist)
        info["passes"] = 1
        which = 0
    info["meff"] = be.finish_weights(which)      # every rank: W = 1/count, Meff (identical everywhere)
    be.covariance_to_root(pseudocount, dist)     # this rank's rows of C end up in rank 0's buffer
    R = None
    if rank == 0:
        be.inverse()
        R = be.score_rank(score, min_separation)
    dist.barrier()
    return R, info


class GpuBackend:
    """The real thing: libgdca_b200.so on this rank's GPU; tensors are views of the library's buffers."""

    def __init__(self, ctx):
        import torch
        self.torch = torch
        self.ctx = ctx
        self.lib = ctx.lib
        self.stream = torch.cuda.ExternalStream(int(self.lib.gdca_dev_stream(ctx.h)), device=ctx.device)
        self.L = self.M = 0

    # -- views of library-owned device memory (no copies)
    def _view(self, ptr, shape, typestr):
        class _Holder:
            pass
        h = _Holder()
        h.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                      "strides": None}
        return self.torch.as_tensor(h, device=f"cuda:{self.ctx.device}")

    def n_units(self):
        return self.torch.cuda.get_device_properties(self.ctx.device).multi_processor_count

    def load(self, Z, resident=False):
        M, L = Z.shape
        self.L, self.M = L, M
        if resident:   # Z: torch int8 tensor on this device
            self.ctx.check(self.lib.gdca_dev_load_resident(self.ctx.h, ctypes.c_void_p(Z.data_ptr()), L, M))
        else:
            self.ctx.check(self.lib.gdca_dev_load(self.ctx.h, Z.ctypes.data_as(ctypes.c_void_p), L, M))

    def set_shard(self, rank, world):
        self.ctx.check(self.lib.gdca_set_shard(self.ctx.h, rank, world))

    def ident_sum(self):
        v = ctypes.c_uint64()
        self.ctx.check(self.lib.gdca_dev_ident_sum(self.ctx.h, ctypes.byref(v)))
        return int(v.value)

    def pair_pass(self, mode, thresh):
        self.ctx.check(self.lib.gdca_dev_pair_pass(self.ctx.h, mode, thresh))

    # -- exchange steps fused into the kernels over peer memory (CUDA IPC + NVLink); the host only barriers
    def _ensure_peers(self, dist):
        torch = self.torch
        world = dist.get_world_size()
        if world == 1:
            return
        dev = f"cuda:{self.ctx.device}"
        ok = torch.tensor([self.lib.gdca_dev_peer_valid(self.ctx.h)], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op="min")
        if int(ok.item()) == 1:
            return
        self.ctx.check(self.lib.gdca_dev_peer_close(self.ctx.h))
        dist.barrier()                                   # nobody still maps a buffer that is about to move
        mine = np.zeros(128, dtype=np.uint8)
        self.ctx.check(self.lib.gdca_dev_peer_export(self.ctx.h, mine.ctypes.data_as(ctypes.c_void_p)))
        table = dist.all_gather_bytes(mine)              # (world, 128) uint8
        table = np.ascontiguousarray(table)
        self.ctx.check(self.lib.gdca_dev_peer_import(self.ctx.h, world, table.ctypes.data_as(ctypes.c_void_p)))
        dist.barrier()

    def sweep_counts(self, thresh, dist):
        self._ensure_peers(dist)
        if dist.get_world_size() > 1:
            self.ctx.check(self.lib.gdca_dev_zero_counts(self.ctx.h))
            dist.barrier()                               # every rank's counters are zero before anyone adds
        self.pair_pass(1, thresh)
        if dist.get_world_size() > 1:
            self.ctx.check(self.lib.gdca_dev_sync(self.ctx.h))
            dist.barrier()                               # all peer atomics have landed

    def covariance_to_root(self, pc, dist):
        self._ensure_peers(dist)
        if dist.get_world_size() > 1:
            if dist.get_rank() == 0:
                self.ctx.check(self.lib.gdca_dev_zero_C(self.ctx.h))
            dist.barrier()
        self.covariance(pc)
        if dist.get_world_size() > 1:
            self.ctx.check(self.lib.gdca_dev_sync(self.ctx.h))
            dist.barrier()                               # all rows have landed in rank 0's C

    def ham_tensor(self):
        return self._view(self.lib.gdca_dev_ham_sum_ptr(self.ctx.h), (2,), "<i8")

    def counts_tensor(self):
        return self._view(self.lib.gdca_dev_counts_ptr(self.ctx.h), (3 * self.lib.gdca_dev_counts_stride(self.ctx.h),), "<i4")

    def C_tensor(self):
        npad = self.lib.gdca_dev_npad(self.ctx.h)
        return self._view(self.lib.gdca_dev_C_ptr(self.ctx.h), (npad * npad,), "<f8")

    def to_host(self, t):
        with self.torch.cuda.stream(self.stream):   # ordered after the collective on the library stream
            return t.cpu().tolist()

    def finish_weights(self, which):
        meff = ctypes.c_double()
        self.ctx.check(self.lib.gdca_dev_finish_weights(self.ctx.h, which, ctypes.byref(meff)))
        return meff.value

    def covariance(self, pc):
        self.ctx.check(self.lib.gdca_dev_covariance(self.ctx.h, float(pc)))

    def inverse(self):
        info = ctypes.c_int32()
        self.ctx.check(self.lib.gdca_dev_inverse(self.ctx.h, ctypes.byref(info)))

    def score_rank(self, score, min_separation, to_host=True):
        from ._lib import RANK_DTYPE, SCORE_CODES, ptr
        n_out = int(self.lib.gdca_ranking_length(self.L, int(min_separation)))
        R = np.empty(n_out, dtype=RANK_DTYPE) if to_host else None
        self.ctx.check(self.lib.gdca_dev_score_rank(self.ctx.h, SCORE_CODES[score], int(min_separation), ptr(R), n_out))
        return R


class _StreamDist:
    """torch.distributed with every collective enqueued on the library's CUDA stream."""

    def __init__(self, dist, torch, stream, device):
        self.d, self.torch, self.stream, self.device = dist, torch, stream, device

    def get_rank(self):
        return self.d.get_rank()

    def get_world_size(self):
        return self.d.get_world_size()

    def all_reduce(self, t, op="sum"):
        with self.torch.cuda.stream(self.stream):
            self.d.all_reduce(t, op=self.d.ReduceOp.MIN if op == "min" else self.d.ReduceOp.SUM)

    def all_gather_bytes(self, arr):
        """numpy uint8[k] on every rank -> numpy uint8[world, k] (control-plane data: IPC handles)."""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            mine = torch.from_numpy(arr).to(f"cuda:{self.device}")   # the context's device, not torch's current one
            out = torch.empty((self.d.get_world_size(), arr.size), dtype=torch.uint8, device=mine.device)
            self.d.all_gather_into_tensor(out, mine)
            return out.cpu().numpy()

    def barrier(self):
        self.stream.synchronize()
        self.d.barrier()


def gdca_sharded(Z, pseudocount=0.8, theta="auto", score="frob", min_separation=5, *, ctx, resident=False):
    """Run gDCA on all ranks of the default process group (NCCL).  Every rank passes the same Z.
    -> (R structured array on rank 0 / None elsewhere, info)."""
    import torch
    import torch.distributed as dist
    # the ranges of check_arguments (src/GaussDCA.jl:49-65), as gdca_run() re-checks them
    if not (0 <= pseudocount <= 1):
        raise ValueError(f"invalid pseudocount value: {pseudocount} (must be between 0 and 1)")
    if not ((isinstance(theta, str) and theta == "auto") or (not isinstance(theta, str) and 0 <= theta <= 1)):
        raise ValueError(f"invalid θ value: {theta} (must be either :auto, or a number between 0 and 1)")
    if score not in ("DI", "frob"):
        raise ValueError(f"invalid score value: {score} (must be either :DI or :frob)")
    if not (min_separation >= 1):
        raise ValueError(f"invalid min_separation value: {min_separation} (must be >= 1)")
    if Z.shape[0] < 2 and isinstance(theta, str):
        raise ValueError("theta = :auto needs at least 2 sequences")
    be = GpuBackend(ctx)
    be.load(Z, resident=resident)
    sd = _StreamDist(dist, torch, be.stream, ctx.device)
    return run_sharded(be, sd, be.L, be.M, theta, pseudocount, score, min_separation)
