#!/usr/bin/env python
"""bench.py -- gDCA hot path on B200: the BASELINE.json metric on the BASELINE.json config.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C|B|D|Cs|E] [--configs all|none]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the whole hot path (encoded alignment -> ranking, reference src/GaussDCA.jl:24-44) over one
synthetic alignment.  Headline workload = BASELINE.json configs[2]: synthetic L=500, M=200k, theta=:auto, :frob,
pseudocount 0.8 (fits one GPU).  One JSON line on stdout (rank 0).

  value         seconds per gDCA with Z already resident in HBM (CUDA events on the library's stream)
  e2e           seconds per gDCA through the C ABI call gdca_run() with pinned HOST buffers: H2D of Z and D2H of the
                ranking are inside the timed region;  e2e_pageable: the same call on pageable numpy memory (what a Julia
                Matrix{Int8} or a numpy array is)
  parity        this run's GPU results against the CPU oracle run FOR REAL on the same full-size alignment
  configs       the other BASELINE.json configs (B, D, C in shuffled sequence order, E), each with value / e2e / parity
  roofline      the kernel family with the largest share of the step;  roofline_kernels: one entry per hot kernel
  cpu_baseline  the oracle port, all host threads, the FULL workload (not a sample) timed stage by stage on this box
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

WORKLOADS = {
    # name: (L, M, score, pseudocount, BASELINE.json configs index, shuffled)
    "B": (200, 50_000, "frob", 0.8, 1, False),
    "C": (500, 200_000, "frob", 0.8, 2, False),     # <- headline: the metric is quoted on this
    "D": (500, 200_000, "DI", 0.2, 3, False),
    "Cs": (500, 200_000, "frob", 0.8, 2, True),     # configs[2] with the sequences in random order
    "E": (1500, 1_000_000, "frob", 0.8, 4, False),  # the 8-GPU config; fits one GPU
    "S": (100, 20_000, "frob", 0.8, None, False),   # small smoke shape
}
SEED = 20140321
MIN_SEP = 5
METRIC = "gDCA end-to-end s @L=500,M=200k"
TOL = 1e-9


def workload_string(name):
    L, M, score, pc, idx, shuf = WORKLOADS[name]
    s = f"synthetic L={L} M={M} theta=auto score={score} pseudocount={pc} min_separation={MIN_SEP}"
    if shuf:
        s += " sequences-shuffled"
    if idx is not None:
        s += f" (BASELINE.json configs[{idx}])"
    return s


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def committed_probe():
    """Pipe peaks recorded by a committed probe run (profiles/r2_probe_peaks.json: gdca_probe_peaks on a B200 of this pool)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_probe_peaks.json")))
    except Exception:
        return {}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full captures."""
    for f in ("r2_traffic.json", "r1_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", f))).get(kernel)
            return (t["dram_bytes_read"] or 0) + (t["dram_bytes_write"] or 0)
        except Exception:
            continue
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self, wait_s=5.0):
        """Launch nvidia-smi and wait for its first row: the sampler must already be running when the timed region begins
        (a step is ~45 ms: a sampler started at the region's first step would deliver nothing before the last)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < wait_s:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def mark(self):
        """Row index now: rows [mark() at region start, mark() at region end) were sampled during the region."""
        return len(self.rows)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self, lo=0, hi=None):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows[lo:hi]:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU side (oracle; checker + baseline)
def julia_probe():
    """BASELINE.md 4.1 / SURVEY H1: the preferred CPU baseline is `julia -t N` running the real gDCA -- look for it at run time."""
    exe = shutil.which("julia")
    if not exe:
        return {"julia": None, "outcome": "julia not found on PATH (shutil.which): the oracle port is timed instead"}
    try:
        r = subprocess.run([exe, "-e", "using DCAUtils, GaussDCA; print(1)"], capture_output=True, text=True, timeout=120)
        ok = r.returncode == 0
        return {"julia": exe, "outcome": ("julia and GaussDCA/DCAUtils load" if ok else
                                         "julia present but GaussDCA/DCAUtils are not installed (no network): oracle port timed")}
    except Exception as e:  # noqa: BLE001
        return {"julia": exe, "outcome": f"julia probe failed: {e}"}


def oracle_full(name, use_cache=True):
    """The oracle pipeline run for real on the whole workload (oracle/fullsize.py), all host threads."""
    graft.load_oracle().build()
    from oracle import fullsize
    L, M, score, pc, _, _ = WORKLOADS[name]
    return fullsize, fullsize.pipeline_full(L, M, score, pc, SEED, MIN_SEP, use_cache=use_cache)


def cpu_baseline_full(name, use_cache=True):
    fullsize, o = oracle_full(name, use_cache)
    L, M, score, pc, _, _ = WORKLOADS[name]
    cb = {
        "value": o["t_total"], "unit": "s", "cores": int(o["threads"]), "kind": "port",
        "pairs_per_s": 2 * (M * (M - 1) // 2) / (o["t_theta"] + o["t_counts"]),
        "stages_s": {"pack": o["t_pack"], "theta_pair_sweep": o["t_theta"], "count_pair_sweep": o["t_counts"],
                     "frequencies": o["t_freqs"], "pseudocount_C": o["t_pc_C"], "chol_inverse": o["t_inv"],
                     "score": o["t_score"], "apc_ranking": o["t_apc_rank"]},
        "sample": (f"the FULL workload, not a sample and not extrapolated: oracle port (C/OpenMP packed 5-bit pair sweeps x 2, "
                   f"scatter-add frequencies, SciPy LAPACK dpotrf+dpotri, {score}, APC, stable ranking) on L={L}, M={M} with "
                   f"{int(o['threads'])} threads"
                   + (" -- stage seconds read from this box's cache of an earlier run of the same oracle" if o["cached"] or
                      o["weights_cached"] else "")),
    }
    return cb, o


def cpu_baseline_sampled(L, M, score, pc):
    """Config E only (5e11 pairs: ~20 min of CPU per sweep): the oracle port on a BOUNDED sample, extrapolated by the exact
    work ratios (pairs, sequences, n^3).  The real threshold is used for the count pass."""
    import numpy as np
    orc = graft.load_oracle()
    orc.build()
    from oracle import fullsize
    cores = fullsize.set_all_threads()
    lib = orc.lib()
    Ms = min(M, 50_000)
    Z = orc.synth_alignment(L, Ms, SEED)
    t0 = time.perf_counter()
    cZ = orc.compress_Z(Z)
    t_pack = time.perf_counter() - t0
    k1 = min(Ms, 8_000)
    t0 = time.perf_counter()
    ident = lib.oracle_ident_sum_packed_range(orc._ptr(cZ), L, Ms, 0, k1)
    t_theta = time.perf_counter() - t0
    pairs_sample = k1 * Ms - k1 * (k1 + 1) // 2
    theta = min(0.5, 0.38 * 0.32 / ((ident / L) / pairs_sample))   # mean identity of the sampled pairs
    thresh = int(theta * L)
    counts = np.empty(Ms, dtype=np.int32)
    t0 = time.perf_counter()
    lib.oracle_neighbour_counts_packed_range(orc._ptr(cZ), L, Ms, thresh, 0, k1, orc._ptr(counts))
    t_cnt = time.perf_counter() - t0
    t_pairs_full = (t_theta + t_cnt) * (M * (M - 1) // 2) / pairs_sample
    q = 21
    n = (q - 1) * L
    ks = min(Ms, max(200, int(6e9 / (L * L))))
    W = np.ones(Ms)
    Pi = np.empty(n); Pij = np.empty((n, n))
    t0 = time.perf_counter()
    lib.oracle_weighted_freqs_range(orc._ptr(Z), L, Ms, q, orc._ptr(W), float(Ms), 0, ks, orc._ptr(Pi), orc._ptr(Pij))
    t_freq = time.perf_counter() - t0
    del Pij
    ns = min(n, 6000)
    A = np.random.default_rng(0).standard_normal((ns, ns + 8))
    Cs = A @ A.T / ns + np.eye(ns)
    t0 = time.perf_counter()
    with fullsize._blas_threads(cores):
        orc.inv_cholesky(Cs)
    t_inv = time.perf_counter() - t0
    total = t_pack * M / Ms + t_pairs_full + t_freq * M / ks + t_inv * (n / ns) ** 3
    return {"value": total, "unit": "s", "cores": cores, "kind": "port",
            "sample": (f"EXTRAPOLATED (config E only): pair sweeps on rows [0,{k1}) of {Ms} sequences ({pairs_sample:.3g} pairs x 2, "
                       f"real thresh {thresh}, scaled by pair count), frequencies on {ks} sequences (scaled by M), dpotrf+dpotri at "
                       f"n={ns} (scaled by n^3); {t_theta + t_cnt + t_freq + t_inv:.1f} s of CPU work measured")}


def run_reference(args, name):
    """--impl reference: the reference's own CPU path on the box's host cores, all threads (set explicitly: torchrun exports
    OMP_NUM_THREADS=1).  Julia + DCAUtils are probed at run time; without them the oracle port runs the FULL workload once
    (config C: ~100 s on 16 threads) -- a measured number, not an extrapolation; K is ignored because one step is that long."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    L, M, score, pc, _, _ = WORKLOADS[name if name != "Cs" else "C"]
    probe = julia_probe()
    if name == "E":
        cb = cpu_baseline_sampled(L, M, score, pc)
    else:
        cb, _ = cpu_baseline_full(name if name != "Cs" else "C", use_cache=False)   # always timed afresh in this arm
    v = cb["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "s", "n_gpus": args.gpus, "steps": 1,
        "warmup": 0, "ms_per_step": v * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(name), "seed": SEED, "generator": "SURVEY 8(d) clustered SplitMix64",
                   "steps_note": "one full un-sampled pass of the workload (K and W ignored: a CPU step is ~minutes)",
                   "julia_probe": probe},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ parity of a GPU run vs the oracle
def parity_vs_oracle(np, glib, ctx, o, M, L, R, st, perm=None):
    """counts / theta / thresh / Meff bit-exact, APC scores normwise, ranking tie-aware identical + top-L identical.
    R: host ranking of the run (structured), st: its stats; perm: the run's sequence k is the oracle's sequence perm[k]."""
    lib = ctx.lib
    out = {"oracle": "oracle/fullsize.py: the CPU restatement run on the same full-size alignment, all host threads"}
    counts = np.empty(M, dtype=np.int32)
    ctx.check(lib.gdca_dev_copy_to_host(ctx.h, glib.ptr(counts), lib.gdca_dev_counts_ptr(ctx.h), M * 4))
    counts += 1
    oc = o["counts"] if perm is None else o["counts"][perm]
    out["max_count_diff"] = int(np.max(np.abs(counts.astype(np.int64) - oc)))
    out["theta_equal"] = bool(st["theta"] == o["theta"])
    out["thresh_equal"] = bool(st["thresh"] == o["thresh"])
    out["meff_equal"] = bool(st["meff"] == o["Meff"])
    S = np.empty((L, L), dtype=np.float64)
    ctx.check(lib.gdca_dev_copy_to_host(ctx.h, glib.ptr(S), lib.gdca_dev_S_ptr(ctx.h), L * L * 8))
    out["apc_score_normwise_err"] = float(np.max(np.abs(S - o["S"])) / np.max(np.abs(o["S"])))
    Ro = o["R"]
    smax = float(np.max(np.abs(Ro["score"])))
    kg, ko = R["i"] * (L + 1) + R["j"], Ro["i"] * (L + 1) + Ro["j"]
    og, oo = np.argsort(kg, kind="stable"), np.argsort(ko, kind="stable")
    keys_equal = bool(np.array_equal(kg[og], ko[oo]))
    out["ranking_keys_equal"] = keys_equal
    tie_ok = False
    if keys_equal:
        out["ranking_score_normwise_err"] = float(np.max(np.abs(R["score"][og] - Ro["score"][oo])) / smax)
        pos_o = np.empty(len(Ro), dtype=np.int64)      # oracle position of the pair at GPU position t
        pos_o[og] = oo
        t = np.arange(len(R))
        moved = pos_o != t
        out["ranking_positions_differ"] = int(moved.sum())
        # a pair may sit elsewhere only inside a run of (near-)tied scores
        tie_ok = bool(np.all(np.abs(Ro["score"][t[moved]] - Ro["score"][pos_o[moved]]) <= 4 * TOL * smax))
    out["ranking_tie_aware_identical"] = tie_ok
    out["top_L_identical"] = bool(np.array_equal(R["i"][:L], Ro["i"][:L]) and np.array_equal(R["j"][:L], Ro["j"][:L]))
    out["tolerance"] = TOL
    out["ok"] = bool(out["max_count_diff"] == 0 and out["theta_equal"] and out["thresh_equal"] and out["meff_equal"] and
                     out["apc_score_normwise_err"] <= TOL and tie_ok and out["top_L_identical"])
    return out


# ------------------------------------------------------------------------------------ our arm
class Runner:
    """One workload on this rank's GPU(s): resident-timed steps, e2e steps, the results needed for the parity leg."""

    def __init__(self, args, name, torch, np, pkg, glib, gdist, ctx, world, rank, local, dist):
        self.a, self.name = args, name
        self.torch, self.np, self.pkg, self.glib, self.gdist = torch, np, pkg, glib, gdist
        self.ctx, self.lib = ctx, ctx.lib
        self.world, self.rank, self.local, self.dist = world, rank, local, dist
        self.L, self.M, self.score, self.pc, _, self.shuffled = WORKLOADS[name]
        self.dev = f"cuda:{local}"
        self.stream = torch.cuda.ExternalStream(int(self.lib.gdca_dev_stream(ctx.h)), device=local)
        self.n_out = int(self.lib.gdca_ranking_length(self.L, MIN_SEP))
        self.perm = None
        L, M = self.L, self.M
        self.Zd = torch.empty((M, L), dtype=torch.int8, device=self.dev)
        ctx.check(self.lib.gdca_synth_alignment_dev(ctx.h, ctypes.c_void_p(self.Zd.data_ptr()), L, M, SEED))
        if self.shuffled:
            self.perm = np.random.default_rng(SEED).permutation(M)
            self.Zd = self.Zd[torch.from_numpy(self.perm).to(self.dev)].contiguous()
        self.st = glib.Stats()

    def barrier(self):
        for d in range(self.world):
            self.torch.cuda.synchronize(d)

    def step_resident(self):
        # one device or a device group: the same C ABI call (gdca_create_multi contexts shard the step over their GPUs)
        self.ctx.check(self.lib.gdca_run_resident(self.ctx.h, ctypes.c_void_p(self.Zd.data_ptr()), self.L, self.M, -1.0, self.pc,
                                                  self.glib.SCORE_CODES[self.score], MIN_SEP, None, self.n_out,
                                                  ctypes.byref(self.st)))

    def timed_resident(self, W, K, flush):
        torch = self.torch
        for _ in range(W):
            self.step_resident()
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        acc = {}
        self.barrier()
        for a, b in ev:
            flush()
            a.record(self.stream)
            self.step_resident()
            b.record(self.stream)
            for k, v in self.st.asdict().items():
                if k.startswith("ms_"):
                    acc[k] = acc.get(k, 0.0) + v / K
        self.barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev) / K   # events on the leader's stream: it waits for every member at each exchange
        return ms, acc

    def e2e(self, K, flush, pinned=True):
        """gdca_run() on HOST buffers: H2D of Z and D2H of R inside the timed region (CUDA events of the library: before the
        H2D copy .. after the D2H copy).  Returns (seconds, R, stats)."""
        torch, np = self.torch, self.np
        L, M = self.L, self.M
        if pinned:
            Zh = torch.empty((M, L), dtype=torch.int8).pin_memory()
            Zh.copy_(self.Zd)
            Rh = torch.empty(self.n_out * 24, dtype=torch.uint8).pin_memory()
            zp, rp = Zh.data_ptr(), Rh.data_ptr()
        else:
            Zn = self.Zd.cpu().numpy().copy()                       # plain pageable numpy memory
            Rn = np.empty(self.n_out, dtype=self.glib.RANK_DTYPE)
            zp, rp = Zn.ctypes.data, Rn.ctypes.data
        e_ms = []
        for it in range(1 + K):
            flush()
            torch.cuda.synchronize()
            self.ctx.check(self.lib.gdca_run(self.ctx.h, ctypes.c_void_p(zp), L, M, -1.0, self.pc,
                                             self.glib.SCORE_CODES[self.score], MIN_SEP, ctypes.c_void_p(rp), self.n_out,
                                             ctypes.byref(self.st)))
            if it >= 1:
                e_ms.append(self.st.ms_total)
        R = (np.frombuffer(Rh.numpy(), dtype=self.glib.RANK_DTYPE) if pinned else Rn).copy()
        # a device group copies the alignment to its first GPU ONCE (the other members get it over NVLink)
        return sum(e_ms) / len(e_ms) / 1e3, R, self.st.asdict(), L * M


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C", choices=sorted(WORKLOADS))
    ap.add_argument("--configs", default="all", choices=["all", "none"],
                    help="also measure the other BASELINE.json configs (B, D, C shuffled, E) into the `configs` array")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the oracle run (no cpu_baseline, no parity)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, args.workload)

    # stdout carries exactly ONE line (the JSON): everything libraries print to fd 1 (the NCCL version banner, ...) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    pkg = graft.load_package()
    from gaussdca_jl_b200 import _lib as glib
    from gaussdca_jl_b200 import dist as gdist

    nproc = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = max(1, args.gpus)          # GPUs the step runs on
    if nproc > 1 and nproc != world:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={nproc}")
    dist = None
    if nproc > 1:
        # Launched as one rank per GPU (torch.distributed.run).  The library drives a device group from ONE process
        # (gdca_create_multi: peer access, cross-device stream waits, exchanges fused into the kernels), so rank 0 leads all
        # `world` GPUs and the other ranks only keep the launch contract: they wait at the closing barrier.  The control plane is
        # gloo on purpose -- an NCCL barrier would park a spinning kernel on every GPU for the length of the run.
        import torch.distributed as dist
        dist.init_process_group("gloo")
        if rank != 0:
            dist.barrier()
            dist.destroy_process_group()
            return
    local = 0
    torch.cuda.set_device(0)
    devices = list(range(world))
    ctx = pkg.Context(devices=devices)   # raises loudly without the CUDA library / B200s with peer access
    lib = ctx.lib
    W = max(3, args.warmup)
    K = max(1, args.steps)
    name = args.workload
    flushbuf = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")  # > 126 MB L2
    run = Runner(args, name, torch, np, pkg, glib, gdist, ctx, world, rank, local, dist)
    L, M, score, pc = run.L, run.M, run.score, run.pc
    stream = run.stream

    def l2_flush():
        with torch.cuda.stream(stream):
            flushbuf.zero_()

    # ---- value: device-resident, per-step CUDA events on the library stream, L2 flushed between steps
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.gdca_dev_kernel_launches(ctx.h)
    for _ in range(W):
        run.step_resident()
    launches1 = lib.gdca_dev_kernel_launches(ctx.h)
    m0 = sampler.mark()
    ms_step, stage_acc = run.timed_resident(0, K, l2_flush)
    launches2 = lib.gdca_dev_kernel_launches(ctx.h)
    time.sleep(0.03)             # the sample taken during the last step reaches the pipe
    m1 = sampler.mark()
    cov_ms = ctypes.c_float()
    ctx.check(lib.gdca_dev_cov_kernel_ms(ctx.h, ctypes.byref(cov_ms)))   # cov_rows_kernel of the last timed step
    stats = run.st.asdict()

    # ---- e2e: gdca_run() with pinned host buffers; the same on pageable memory
    e_s, R, est, h2d = run.e2e(K, l2_flush, pinned=True)
    e2e = {"value": e_s, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": run.n_out * 24,
           "host_memory": "pinned (cudaHostAlloc)"}
    e2e_pageable = None
    if True:
        p_s, _, _, _ = run.e2e(min(K, 3), l2_flush, pinned=False)
        e2e_pageable = {"value": p_s, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": run.n_out * 24,
                        "host_memory": "pageable (numpy / a Julia Matrix{Int8}): cudaMemcpyAsync stages it through the driver"}
    top = [int(R["i"][0]), int(R["j"][0]), float(R["score"][0])] if R is not None else None
    m2 = sampler.mark()
    # clocks DURING the timed region of `value`; if nvidia-smi delivered fewer than 3 rows in it, the e2e timed regions that
    # follow it back to back are included (and the window says so)
    few = (m1 - m0) < 3
    clocks = sampler.stop(m0, m2 if few else m1)
    clocks["window"] = "timed resident steps + timed e2e steps" if few else "timed resident steps"

    # ---- parity of THIS run (the e2e call's results) against the oracle run for real on the same alignment; cpu_baseline
    cb, parity, o_main = None, None, None
    if rank == 0 and not args.no_cpu_baseline:
        if name == "E":
            cb = cpu_baseline_sampled(L, M, score, pc)
        else:
            cb, o_main = cpu_baseline_full("C" if name == "Cs" else name)
            # the last e2e call left counts / S on the (leading) device; R and stats are the host copies
            parity = parity_vs_oracle(np, glib, ctx, o_main, M, L, R, est, perm=run.perm)

    # ---- per-kernel rooflines, measured live
    pk = peaks()
    probe_file = committed_probe()
    lop3, popc, dmma, dfma = (ctypes.c_double() for _ in range(4))
    ctx.check(lib.gdca_probe_peaks(ctx.h, ctypes.byref(lop3), ctypes.byref(popc), ctypes.byref(dmma), ctypes.byref(dfma)))
    # time the sweep alone: the production launch (mode 1: neighbour counts, exact early exit)
    ctx.check(lib.gdca_dev_load_resident(ctx.h, ctypes.c_void_p(run.Zd.data_ptr()), L, M))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    thr = int(est["thresh"])
    ctx.check(lib.gdca_set_shard(ctx.h, rank, world))
    tk = []
    for it in range(4):
        l2_flush()
        a.record(stream)
        ctx.check(lib.gdca_dev_pair_pass(ctx.h, 1, thr))
        b.record(stream)
        stream.synchronize()
        if it:
            tk.append(a.elapsed_time(b))
    t_pair = sum(tk) / len(tk) / 1e3
    hs = np.zeros(2, dtype=np.uint64)
    ctx.check(lib.gdca_dev_copy_to_host(ctx.h, glib.ptr(hs), lib.gdca_dev_ham_sum_ptr(ctx.h), 16))
    filt, f_tiles, s_blocks = ctypes.c_int32(), ctypes.c_int64(), ctypes.c_int64()
    f_tflop, ms_f, ms_x, f_l2 = ctypes.c_double(), ctypes.c_float(), ctypes.c_float(), ctypes.c_double()
    ctx.check(lib.gdca_dev_sweep_info(ctx.h, ctypes.byref(filt), ctypes.byref(f_tiles), ctypes.byref(f_tflop),
                                      ctypes.byref(s_blocks), ctypes.byref(ms_f), ctypes.byref(ms_x), ctypes.byref(f_l2)))
    pair_words = int(hs[1])                       # (pair, 32-site word) units really executed by the exact sweep
    npairs = M * (M - 1) // 2 // world            # this rank's shard of the sweep
    nwords = (L + 31) // 32
    alu_ops = pair_words * 5                      # executed ALU-pipe work: 5 LOP3 per pair-word
    full_ops = npairs * nwords * 5                # a sweep without filter and early exit
    Mpad = (M + 127) // 128 * 128
    hbm_bytes = 4 * nwords * 5 * Mpad + 4 * M     # packed planes once + counts
    t_exact = (ms_x.value / 1e3) if filt.value else t_pair
    cand, pcap = ctypes.c_int64(), ctypes.c_int64()
    ctx.check(lib.gdca_dev_pair_list_info(ctx.h, ctypes.byref(cand), ctypes.byref(pcap)))
    pair_path = bool(filt.value) and pcap.value > 0 and cand.value <= pcap.value
    exact = {
        "kernel": "pair_list_kernel (exact Hamming distance of the candidate pairs the prefilter lists: one warp per pair, byte compares on the alignment itself) + compact_flags_kernel",
        "bound": "hbm", "achieved": cand.value * 2.0 * L / t_exact / 1e9, "peak": pk.get("hbm_gbs"), "unit": "GB/s",
        "frac": (cand.value * 2.0 * L / t_exact / 1e9 / pk["hbm_gbs"]) if pk.get("hbm_gbs") else None,
        "ms_per_launch": t_exact * 1e3, "candidate_pairs": int(cand.value), "list_capacity": int(pcap.value),
        "note": "2 L bytes per candidate, mostly L2 hits (the alignment is 100 MB); the stage is latency / launch bound at this size "
                "(< 1 ms): its cost follows the number of candidates -- true neighbour pairs plus what the 4-class projection cannot "
                "tell apart -- not their spread over the pair matrix",
    } if pair_path else {
        "kernel": "cell_sweep_kernel<5> / pair_sweep_kernel<5,1> (exact neighbour counts on the bit planes, per-warp early exit)",
        "bound": "int32_alu", "achieved": alu_ops / t_exact / 1e12, "peak": lop3.value, "unit": "Tlop3/s",
        "frac": (alu_ops / t_exact / 1e12) / lop3.value, "ms_per_launch": t_exact * 1e3,
        "blocks_swept": int(s_blocks.value), "blocks_total": (Mpad // 128) * (Mpad // 128 + 1) // 2 // world,
        "peak_source": "measured live: gdca_probe_peaks LOP3 issue rate (MEASURED_PEAKS.json has no INT32 figure); "
                       f"committed probe run: {probe_file.get('lop3_tops')}",
        "executed_fraction_of_full_sweep": alu_ops / full_ops,
    }
    common = {
        "sweep_ms": t_pair * 1e3, "pairs_per_s": npairs / t_pair,
        "effective_tlop3_per_s_full_sweep_equivalent": full_ops / t_pair / 1e12,
        "hbm": {"achieved": hbm_bytes / t_pair / 1e9, "peak": pk.get("hbm_gbs"), "unit": "GB/s",
                "frac": hbm_bytes / t_pair / 1e9 / pk["hbm_gbs"] if pk.get("hbm_gbs") else None,
                "note": "compulsory bytes only; the sweep is compute-bound, operands live in L2"},
        "survey_floor_ops_per_pair": 5 * ((L + 5) // 6),
    }
    if filt.value:
        fp4 = filt.value == 4
        bf16 = pk.get("bf16_tflops")
        tc_peak = 9000.0 if fp4 else 4500.0     # nominal dense FP4 / FP8 = INT8 (B200_PROFILING.md table)
        scaled = (4.0 if fp4 else 2.0) * bf16 if bf16 else None
        t_f = ms_f.value / 1e3
        fmode = int(lib.gdca_dev_tc_filter_launch_mode(ctx.h))
        sweep_entry = {
            "kernel": ("tc_filter_kernel<fp4> (one tcgen05.mma.cta_group::2 kind::mxf4.block_scale 256x224x64 per CTA pair" if fp4 and fmode == 2 else
                       "tc_filter_kernel<fp4> (tcgen05 kind::mxf4.block_scale 128x224x64" if fp4 else
                       "tc_filter_kernel<int8> (tcgen05 kind::i8 128x256x32" if filt.value == 80 else
                       "tc_filter_kernel<fp8> (tcgen05 kind::f8f6f4 128x256x32") + ", TMA ring, TMEM epilogue)"
                      + ("" if world == 1 else f", shard {rank} of {world}"),
            "bound": "tensor", "achieved": f_tflop.value / t_f, "peak": tc_peak, "unit": "TFLOP/s",
            "frac": f_tflop.value / t_f / tc_peak,
            "peak_source": ("nominal dense " + ("FP4" if fp4 else "FP8") + " tensor rate of the B200_PROFILING.md table: "
                            "MEASURED_PEAKS.json measures bf16 only"),
            "frac_of_measured_bf16_scaled": (f_tflop.value / t_f / scaled) if scaled else None,
            "measured_bf16_scaled_note": (f"{4 if fp4 else 2} x MEASURED_PEAKS.json bf16_tflops (burst) = {scaled:.0f} TFLOP/s"
                                          if scaled else None),
            "ms_per_launch": t_f * 1e3, "tiles": int(f_tiles.value), "flop_per_launch": f_tflop.value * 1e12,
            "l2_operand_bytes_per_launch": f_l2.value, "l2_operand_tb_per_s": f_l2.value / t_f / 1e12,
            "tensor_pipe_note": ("ncu (profiles/r2_top_kernels.md): sm__pipe_tensor_cycles_active 99.9 % -- the pipe is saturated; "
                                 "the nominal 9 PFLOP/s is not reachable with 128x224x64 MMAs at this clock") if fp4 and fmode == 2 else None,
            "traffic": (ncu_traffic(f"tc_filter_kernel<{int(fp4)}, {fmode}>") or ncu_traffic(f"tc_filter_kernel<{int(fp4)}, 1>") or
                        ncu_traffic(f"tc_filter_kernel<{int(fp4)}, 0>"))
                       if world == 1 and name == "C" else None,
            "exact_sweep": exact, **common,
        }
    else:
        sweep_entry = {**exact, "traffic": ncu_traffic("pair_sweep_kernel<5, 1>") if world == 1 and name == "C" else None, **common}
    ctx.check(lib.gdca_set_shard(ctx.h, 0, 1))

    n = 20 * L
    roof, roof_kernels, stages = sweep_entry, None, None
    if world > 1:
        t_chol = (stage_acc.get("ms_chol", 0) + stage_acc.get("ms_inv", 0)) / 1e3
        stages = {"ms": {k: round(v, 4) for k, v in stage_acc.items()},
                  "note": "CUDA events on the leading device's stream; it waits for every member at each exchange"}
        inv_info = _inverse_info(ctx, lib)
        roof = {
            "kernel": inv_info.get("kernel", "Cholesky + inverse") + f"; factorisation on the leading GPU, trtri / lauum shared by {world} GPUs",
            "bound": "tensor", "bound_detail": "fp64_tensor (SURVEY 8(d): n^3 FP64 flop against the DMMA rate)",
            "achieved": n ** 3 / t_chol / 1e12, "peak": dmma.value * world, "unit": "TFLOP/s",
            "frac": (n ** 3 / t_chol / 1e12) / (dmma.value * world), "traffic": None,
            "peak_source": f"{world} x the FP64 tensor (DMMA) rate measured live by gdca_probe_peaks on the leading GPU",
            "ms_per_step": t_chol * 1e3, "share_of_step": t_chol * 1e3 / ms_step,
            "sweep_shard_on_leader": sweep_entry,
        }
    if world == 1:
        t_cov = stage_acc.get("ms_cov", 0) / 1e3
        t_chol = (stage_acc.get("ms_chol", 0) + stage_acc.get("ms_inv", 0)) / 1e3
        stages = {
            "ms": {k: round(v, 4) for k, v in stage_acc.items()},
            "theta_passes": stats["theta_passes"],
            "weights_pairs_per_s": npairs / ((stage_acc.get("ms_theta", 0) + stage_acc.get("ms_weights", 0)) / 1e3 + 1e-30),
            "cov_fp64_equiv_tflops": M * n * (n + 1) / t_cov / 1e12 if t_cov else None,
            "cov_fp64_equiv_note": ("SURVEY 8(d) counts the DENSE contraction M n (n+1) against the FP64 tensor peak.  The tensor-core "
                                    "engine executes that contraction as exact 0/1 products on the FP4 tensor cores (one class of equal "
                                    "weights at a time, FP64 combination), the scatter-add engine only its M L(L+1)/2 non-zero terms: "
                                    "either way the figure exceeds the DMMA peak by construction -- see roofline_kernels for the pipe "
                                    "each engine is really bound by"),
            "chol_inv_tflops": n ** 3 / t_chol / 1e12 if t_chol else None,
            "cov_plus_inv_fp64_equiv_tflops": (M * n * (n + 1) + n ** 3) / (t_cov + t_chol) / 1e12 if t_cov else None,
            "dmma_peak_tflops_measured": dmma.value, "dfma_peak_tflops_measured": dfma.value,
            "chol_inv_frac_of_dmma_peak": (n ** 3 / t_chol / 1e12) / dmma.value if t_chol else None,
        }
        sm_clk = (clocks.get("sm_mhz") or pk.get("sm_max_mhz") or 1965.0) * 1e6
        smem_peak = 148 * 128 * sm_clk / 1e12                       # TB/s: 128 B/clk/SM shared-memory crossbar
        t_cv = cov_ms.value / 1e3
        rmw = M * L * (L + 1) // 2                                   # FP64 additions = 8-byte smem read + 8-byte write each
        cinfo = ctx.cov_info()
        if cinfo["engine"] == 2:
            # exact co-occurrence counts per weight class on the FP4 tensor cores (csrc/covtc.cu): the SURVEY 8(d) convention --
            # the dense contraction M n (n+1) -- now IS what the kernel executes (lower triangle + mirror image)
            bf16 = pk.get("bf16_tflops")
            alg = float(M) * n * (n + 1)
            cov_entry = {
                "kernel": ("cov_tc_kernel (X'WX as integer co-occurrence counts per weight class: tcgen05 kind::mxf4 128x128x64, 2x2 "
                           "clusters with TMA multicast of the operand halves, FP64 combination sum_c w_c N_c in the epilogue registers)"),
                "bound": "tensor", "achieved": alg / t_cv / 1e12, "peak": 9000.0, "unit": "TFLOP/s",
                "frac": alg / t_cv / 1e12 / 9000.0,
                "peak_source": "nominal dense FP4 tensor rate of the B200_PROFILING.md table: MEASURED_PEAKS.json measures bf16 only",
                "frac_of_measured_bf16_scaled": (alg / t_cv / 1e12 / (4.0 * bf16)) if bf16 else None,
                "algorithmic_flop_per_launch": alg,
                "algorithmic_note": "SURVEY 8(d): M n (n+1) flop of the dense one-hot contraction (FP64-equivalent)",
                "executed_tflop_per_launch": cinfo["tflop"], "executed_tflops": cinfo["tflop"] / t_cv,
                "executed_note": "every 128x128x256 MMA block issued, incl. the padding of the classes to whole k-blocks and the "
                                 "unused quarter of the diagonal super-tiles",
                "weight_classes": cinfo["classes"], "class_segments": cinfo["segments"], "kblocks_of_256_sequences": cinfo["kblocks"],
                "clusters_of_4_ctas": cinfo["clusters"],
                "operand_bytes_into_sms_per_launch": 2.0 * cinfo["l2_bytes"],
                "sm_ingest_bytes_per_clk_per_sm": 2.0 * cinfo["l2_bytes"] / t_cv / (4 * cinfo["clusters"]) / sm_clk,
                "sm_ingest_note": "32 KB of operands per 128x128x256 block arrive in every SM (half fetched, half multicast by the "
                                  "neighbour): the L2->SM path delivers 64 B/clk/SM (ncu: l1tex__m_xbar2l1tex_read_bytes), which "
                                  "caps a 128x128 FP4 tile at 0.53 of the tensor rate",
                "l2_read_bytes_per_launch": cinfo["l2_bytes"],
                "ms_per_launch": cov_ms.value,
                "traffic": ncu_traffic("cov_tc_kernel") if name == "C" else None,
                "hbm": {"achieved": (float(n) * cinfo["kblocks"] * 128 + 8.0 * n * n) / t_cv / 1e9, "peak": pk.get("hbm_gbs"), "unit": "GB/s",
                        "note": "compulsory bytes (one-hot operand once + C once); not the limiter"},
            }
        else:
            cov_entry = {
                "kernel": "cov_rows_kernel<2> (weighted one-hot covariance as M*L(L+1)/2 private shared-memory FP64 adds)",
                "bound": "shared_memory", "achieved": 16 * rmw / t_cv / 1e12, "peak": smem_peak, "unit": "TB/s",
                "frac": 16 * rmw / t_cv / 1e12 / smem_peak,
                "peak_source": "148 SMs x 128 B/clk (B300_MICROARCH.md shared-memory crossbar) x SM clock under load (nominal; "
                               "MEASURED_PEAKS.json has no shared-memory figure)",
                "ms_per_launch": cov_ms.value, "adds_per_launch": rmw,
                "traffic": ncu_traffic("cov_rows_kernel<2, 0, 1>") or ncu_traffic("cov_rows_kernel<2, 0>") if name == "C" else None,
                "hbm": {"achieved": (L * M + 8 * n * n / 2) / t_cv / 1e9, "peak": pk.get("hbm_gbs"), "unit": "GB/s",
                        "note": "compulsory bytes (recoded alignment once + upper half of C once); not the limiter"},
            }
        inv_info = _inverse_info(ctx, lib)
        inv_entry = {
            "kernel": inv_info.get("kernel", "dgemm_kernel<*> + diag_block_kernel (blocked Cholesky, trtri by recursive doubling, lauum)"),
            "bound": "tensor", "bound_detail": "fp64_tensor (SURVEY 8(d): n^3 FP64 flop against the DMMA rate; the big products "
                                               "themselves run as INT8 tcgen05 digit products, see int8_ops_per_step)",
            "achieved": n ** 3 / t_chol / 1e12, "peak": dmma.value, "unit": "TFLOP/s",
            "frac": (n ** 3 / t_chol / 1e12) / dmma.value,
            "traffic": ncu_traffic("ozaki_gemm_kernel") if name == "C" else None,
            "traffic_note": "DRAM bytes of the longest ozaki_gemm_kernel launch (the lauum product) in the committed ncu --set full capture",
            "peak_source": "FP64 tensor (DMMA.8x8x4) rate measured live by gdca_probe_peaks (MEASURED_PEAKS.json has HBM and bf16 "
                           f"only); committed probe run profiles/r2_probe_peaks.json: {probe_file.get('dmma_tflops')} TFLOP/s",
            "ms_per_step": t_chol * 1e3, "flop_per_step": float(n) ** 3,
            "share_of_step": t_chol * 1e3 / ms_step,
            **{k: v for k, v in inv_info.items() if k != "kernel"},
        }
        cov_entry["share_of_step"] = cov_ms.value / ms_step
        sweep_entry["share_of_step"] = sweep_entry.get("sweep_ms", 0) / ms_step
        roof_kernels = [inv_entry, cov_entry, sweep_entry]
        roof = max(roof_kernels, key=lambda e: e.get("share_of_step", 0))   # the kernel FAMILY with the largest share of the step

    # ---- the other BASELINE.json configs, each with value / e2e / parity (driver-visible records of B, D, C shuffled, E)
    configs = None
    if args.configs == "all" and name == "C":
        configs = []
        del run
        torch.cuda.empty_cache()
        for other in (["B", "D", "Cs", "E"] if world == 1 else ["E"]):   # E is the 8-GPU config of BASELINE.json
            try:
                configs.append(_measure_config(args, other, torch, np, pkg, glib, gdist, ctx, world, rank, local, dist, l2_flush))
            except Exception as e:  # noqa: BLE001
                configs.append({"name": other, "workload": workload_string(other), "error": repr(e)})

    line = {
        "metric": METRIC, "value": ms_step / 1e3, "unit": "s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_string(name), "seed": SEED, "generator": "SURVEY 8(d) clustered SplitMix64",
                   "l2": "256 MiB flush write between steps",
                   "sharding": ("single GPU" if world == 1 else
                                f"device group of {world} GPUs behind one gdca_create_multi context (one process): alignment copied to "
                                "GPU 0 once and broadcast over NVLink; pair-matrix row blocks (peer atomics), covariance rows (peer "
                                "stores), trtri column slices (stored to every member) and lauum row tiles (stored to GPU 0) shared; "
                                "Cholesky chain, scores, APC, ranking on GPU 0"),
                   "data_plane": (None if world == 1 else
                                  f"peer memory over NVLink inside the kernels (cudaDeviceEnablePeerAccess between all {world} devices, "
                                  f"group size reported by the library: {int(lib.gdca_group_size(ctx.h))}); no collective library on the "
                                  "data path; cross-device ordering by CUDA events"),
                   "launch": (None if nproc == 1 else f"{nproc} ranks launched; rank 0 drives the group, the others wait (gloo barrier)")},
        "theta": est["theta"], "thresh": est["thresh"], "meff": est["meff"], "top_pair": top,
        "e2e": e2e, "e2e_pageable": e2e_pageable, "gpu_launches": int(launches2 - launches1),
        "clocks": clocks, "parity": parity, "roofline": roof, "roofline_kernels": roof_kernels, "stages": stages,
        "cpu_baseline": cb, "configs": configs,
    }
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.barrier()           # releases the other launched ranks
        dist.destroy_process_group()


def _inverse_info(ctx, lib):
    """What the last inversion ran on (DMMA only, or the INT8-sliced tcgen05 GEMMs): gdca_dev_inverse_info."""
    try:
        mode, ops, flop = ctypes.c_int32(), ctypes.c_double(), ctypes.c_double()
        ctx.check(lib.gdca_dev_inverse_info(ctx.h, ctypes.byref(mode), ctypes.byref(ops), ctypes.byref(flop)))
        if not mode.value:
            return {}
        return {"kernel": "ozaki_gemm_kernel (FP64 GEMMs of potrf / trtri / lauum as 36 INT8 tcgen05 kind::i8 digit products, S32 in "
                          "TMEM, one-rounding FP64 recombination) + dgemm_kernel<*> / diag_block_kernel (DMMA) for panels and diagonal blocks",
                "int8_ops_per_step": ops.value, "fp64_flop_carried_by_int8": flop.value,
                "int8_peak_nominal_tops": 4500.0,
                "note": ("frac is against the FP64 tensor (DMMA) peak, the roofline of a plain FP64 implementation: > 1 means the "
                         "sliced engine is past that wall; the INT8 kernels themselves run at ~0.57 of the nominal dense INT8 rate "
                         "(profiles/r2_top_kernels.md: tensor pipe 61 % active, bound by the shared-memory operand reads of N = 64 MMAs)")}
    except Exception:  # noqa: BLE001
        return {}


def _measure_config(args, name, torch, np, pkg, glib, gdist, ctx, world, rank, local, dist, l2_flush):
    """One entry of the `configs` array: value (resident), e2e (pinned host buffers), stage ms, parity vs the oracle."""
    run = Runner(args, name, torch, np, pkg, glib, gdist, ctx, world, rank, local, dist)
    big = name == "E"
    K = 1 if big else max(1, min(args.steps, 5))
    ms_step, stage_acc = run.timed_resident(1 if big else 3, K, l2_flush)
    e_s, R, est, h2d = run.e2e(1 if big else min(K, 3), l2_flush, pinned=True)
    entry = {
        "name": name, "workload": workload_string(name), "n_gpus": world, "steps": K, "value": ms_step / 1e3, "unit": "s",
        "e2e": {"value": e_s, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": run.n_out * 24},
        "stages_ms": {k: round(v, 4) for k, v in stage_acc.items()} or None,
        "theta": est["theta"], "thresh": est["thresh"], "meff": est["meff"],
        "top_pair": [int(R["i"][0]), int(R["j"][0]), float(R["score"][0])] if R is not None else None,
    }
    L, M, n = run.L, run.M, 20 * run.L
    if stage_acc:
        t_chol = (stage_acc.get("ms_chol", 0) + stage_acc.get("ms_inv", 0)) / 1e3
        entry["chol_inv_tflops"] = n ** 3 / t_chol / 1e12 if t_chol else None
        entry["weights_pairs_per_s"] = (M * (M - 1) // 2) / ((stage_acc.get("ms_theta", 0) + stage_acc.get("ms_weights", 0)) / 1e3 + 1e-30)
    if rank == 0 and not args.no_cpu_baseline:
        if big:
            entry["parity"] = {"note": "no full oracle run at M=1e6 (5e11 pairs per CPU sweep): properties only",
                               "ranking_sorted": bool(np.all(np.diff(R["score"]) <= 0)),
                               "ranking_rows": int(len(R)), "separation_ok": bool(np.all(R["j"] - R["i"] >= MIN_SEP))}
        else:
            cb, o = cpu_baseline_full("C" if name == "Cs" else name)
            entry["cpu_baseline"] = cb
            entry["parity"] = parity_vs_oracle(np, glib, ctx, o, M, L, R, est, perm=run.perm)
    del run
    torch.cuda.empty_cache()
    return entry


if __name__ == "__main__":
    main()
