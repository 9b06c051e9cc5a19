"""CPU, world_size 2, gloo: the host-side logic of the sharded driver (gaussdca.jl_b200/dist.py:run_sharded)
with a stand-in backend that computes each rank's shard on the CPU from brute-force numpy + the oracle."""
import math
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

TILE = 128


def tile_items(M):
    T = (M + TILE - 1) // TILE
    return [(bi, bj) for bi in range(T) for bj in range(bi, T)]


class FakeBackend:
    """Same partition as pairs.cu / cov.cu: item t -> rank t % world; covariance rows by site i % world."""

    def __init__(self, Z, orc, torch):
        self.Z, self.orc, self.torch = Z, orc, torch
        self.M, self.L = Z.shape
        self.H = (Z[:, None, :] != Z[None, :, :]).sum(-1).astype(np.int64)
        self.Mpad = (self.M + TILE - 1) // TILE * TILE
        self.counts = torch.zeros(3 * self.Mpad, dtype=torch.int32)
        self.ham = torch.zeros(2, dtype=torch.int64)
        self.calls = []

    def n_units(self):
        return 1

    def set_shard(self, rank, world):
        self.rank, self.world = rank, world

    def _sweep(self, mode, thresh, stride):
        items = tile_items(self.M)
        w = self.world * stride
        ham = npairs = 0
        cnt = np.zeros((3, self.Mpad), dtype=np.int32)
        for t, (bi, bj) in enumerate(items):
            if t % w != self.rank:
                continue
            r = np.arange(bi * TILE, min((bi + 1) * TILE, self.M))
            c = np.arange(bj * TILE, min((bj + 1) * TILE, self.M))
            Hs = self.H[np.ix_(r, c)]
            valid = (r[:, None] < c[None, :]) if bi == bj else np.ones_like(Hs, dtype=bool)
            ham += int(Hs[valid].sum())
            npairs += int(valid.sum())
            ths = [thresh - 1, thresh, thresh + 1] if mode == 2 else [thresh]
            for k, th in enumerate(ths):
                hit = valid & (Hs < th)
                np.add.at(cnt[k], r, hit.sum(1))
                np.add.at(cnt[k], c, hit.sum(0))
        if mode != 1:
            self.ham = self.torch.tensor([ham, npairs], dtype=self.torch.int64)
        if mode != 0:
            self.counts = self.torch.from_numpy(cnt.reshape(-1).copy())

    def ident_sum(self):
        self.calls.append(("ident",))
        iu = np.triu_indices(self.M, 1)
        return int((self.L - self.H[iu]).sum())

    def pair_pass(self, mode, thresh):
        self.calls.append(("pass", mode, thresh))
        self._sweep(mode, thresh, 1)

    def ham_tensor(self):
        return self.ham

    def counts_tensor(self):
        return self.counts

    def to_host(self, t):
        return t.tolist()

    def sweep_counts(self, thresh, dist):
        self.pair_pass(1, thresh)
        dist.all_reduce(self.counts)          # CPU stand-in for the peer atomics of the GPU backend

    def covariance_to_root(self, pc, dist):
        self.covariance(pc)
        dist.reduce(self.C, dst=0)            # CPU stand-in for the peer stores into rank 0's buffer

    def finish_weights(self, which):
        if which < 0:
            self.cnt = np.ones(self.M, dtype=np.int32)
        else:
            self.cnt = self.counts.numpy().reshape(3, self.Mpad)[which, :self.M] + 1
        self.W = 1.0 / self.cnt
        self.Meff = self.orc.meff_from_counts(self.cnt)
        return self.Meff

    def covariance(self, pc):
        q = int(self.Z.max())
        s = q - 1
        Pi, Pij = self.orc.compute_freqs(self.Z, q, self.W, self.Meff)
        Pi, Pij = self.orc.add_pseudocount(Pi, Pij, pc, q)
        C = self.orc.compute_C(Pi, Pij)
        mine = (np.arange(C.shape[0]) // s) % self.world == self.rank
        C[~mine] = 0.0
        self.q = q
        self.C = self.torch.from_numpy(C)

    def C_tensor(self):
        return self.C

    def inverse(self):
        self.mJ = self.orc.inv_cholesky(self.C.numpy())

    def score_rank(self, score, ms):
        C = self.C.numpy()
        S = self.orc.compute_DI_gauss(self.mJ, C, self.q) if score == "DI" else self.orc.compute_FN(self.mJ, self.q)
        return self.orc.compute_ranking(self.orc.correct_APC(S), ms)


def _worker(rank, world, port, case, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.load_package()
    from gaussdca_jl_b200.dist import run_sharded
    orc = g.load_oracle()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L, M, theta, score = case
    Z = orc.synth_alignment(L, M, seed=5)
    be = FakeBackend(Z, orc, torch)
    R, info = run_sharded(be, dist, L, M, theta, 0.8, score, 3)
    info["calls"] = be.calls
    out.put((rank, R, info))
    dist.destroy_process_group()


@pytest.mark.parametrize("case", [(24, 1400, "auto", "frob"), (24, 300, "auto", "frob"), (16, 300, 0.3, "DI"),
                                  (16, 200, 0.0, "frob")], ids=["auto", "auto-small", "fixed", "theta0"])
def test_sharded_driver_world2_gloo(orc, case):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = []
    import queue
    import time
    deadline = time.time() + 240
    while len(res) < len(procs) and time.time() < deadline:
        try:
            res.append(out.get(timeout=2))
        except queue.Empty:
            assert all(p.is_alive() or p.exitcode == 0 for p in procs), "a rank died"
    assert len(res) == len(procs)
    res.sort(key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    L, M, theta, score = case
    Z = orc.synth_alignment(L, M, seed=5)
    st = {}
    Ro = orc.gdca_from_Z(Z, 0.8, theta, score, 3, stages=st)
    (r0, R0, i0), (r1, R1, i1) = res
    assert R1 is None and R0 is not None                      # ranking lives on rank 0 only
    assert i0["thresh"] == i1["thresh"] == st["thresh"] and i0["theta"] == st["theta"]
    assert i0["meff"] == i1["meff"] == st["Meff"]
    assert [(a, b) for a, b, _ in R0] == [(a, b) for a, b, _ in Ro]
    assert max(abs(x - y) for (_, _, x), (_, _, y) in zip(R0, Ro)) < 1e-10
    if case[2] == "auto":
        assert i0["ident_sum"] == orc.ident_sum(Z)
        assert i0["passes"] == 1 and i0["calls"] == [("ident",), ("pass", 1, st["thresh"])]   # one sweep, exact threshold
    elif theta == 0.0:
        assert i0["passes"] == 0 and i0["calls"] == []
