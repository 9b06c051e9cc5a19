#!/usr/bin/env python
"""bench.py -- gDCA hot path on B200: the BASELINE.json metric on the BASELINE.json config.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C|B|...]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the whole hot path (encoded alignment -> ranking, reference
src/GaussDCA.jl:24-44) over one synthetic alignment.  Workload = BASELINE.json configs[2]: synthetic
L=500, M=200k, theta=:auto, :frob, pseudocount 0.8 (fits one GPU).  One JSON line on stdout (rank 0).

  value   seconds per gDCA with Z already resident in HBM (CUDA events on the library's stream)
  e2e     seconds per gDCA through the C ABI call gdca_run() with pinned HOST buffers: H2D of Z and
          D2H of the ranking are inside the timed region
  roofline / cpu_baseline / stages: see DESIGN.md section 6
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

WORKLOADS = {
    # name: (L, M, score, pseudocount)      BASELINE.json configs[...]
    "B": (200, 50_000, "frob", 0.8),      # configs[1]
    "C": (500, 200_000, "frob", 0.8),     # configs[2]  <- headline, metric is quoted on this
    "D": (500, 200_000, "DI", 0.2),       # configs[3]
    "S": (100, 20_000, "frob", 0.8),      # small smoke shape
}
SEED = 20140321
METRIC = "gDCA end-to-end s @L=500,M=200k"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu --set full capture
    of the same workload (profiles/r1_traffic.json, made by tools/summarise_profiles.py)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get(kernel)
        return (t["dram_bytes_read"] or 0) + (t["dram_bytes_write"] or 0)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
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


# ------------------------------------------------------------------------------------ CPU baseline
def cpu_baseline(L, M, score, pc, budget_s=20.0):
    """The oracle port (oracle/, C + OpenMP + LAPACK via SciPy) timed on this box's host cores on a BOUNDED
    sample of the same workload, extrapolated to the full workload by the exact work ratios."""
    import numpy as np
    orc = graft.load_oracle()
    orc.build()
    lib = orc.lib()
    cores = int(lib.oracle_max_threads())
    Ms = min(M, 50_000)                      # sample alignment: first Ms sequences of the same generator
    Z = orc.synth_alignment(L, M, SEED)[:Ms].copy() if M <= 200_000 else orc.synth_alignment(L, Ms, SEED)
    t0 = time.perf_counter()
    cZ = orc.compress_Z(Z)
    t_pack = time.perf_counter() - t0
    # pair sweep: theta pass + threshold pass over rows [0, k1) x all later sequences of the sample
    k1 = min(Ms, 20_000)
    t0 = time.perf_counter()
    lib.oracle_ident_sum_packed_range(orc._ptr(cZ), L, Ms, 0, k1)
    t_theta = time.perf_counter() - t0
    counts = np.empty(Ms, dtype=np.int32)
    t0 = time.perf_counter()
    lib.oracle_neighbour_counts_packed_range(orc._ptr(cZ), L, Ms, int(0.5 * L), 0, k1, orc._ptr(counts))
    t_cnt = time.perf_counter() - t0
    pairs_sample = k1 * Ms - k1 * (k1 + 1) // 2
    pairs_full = M * (M - 1) // 2
    t_pairs_full = (t_theta + t_cnt) * pairs_full / pairs_sample
    # frequencies: all sites, a slice of the sequences
    q = 21
    n = (q - 1) * L
    ks = min(Ms, max(200, int(1.2e10 / (L * L))))
    W = np.ones(Ms)
    Pi = np.empty(n); Pij = np.empty((n, n))
    t0 = time.perf_counter()
    lib.oracle_weighted_freqs_range(orc._ptr(Z), L, Ms, q, orc._ptr(W), float(Ms), 0, ks, orc._ptr(Pi), orc._ptr(Pij))
    t_freq = time.perf_counter() - t0
    t_freq_full = t_freq * M / ks
    # inversion: LAPACK dpotrf + dpotri at a reduced n, scaled by n^3
    ns = min(n, 8000)
    A = np.random.default_rng(0).standard_normal((ns, ns + 8))
    Cs = A @ A.T / ns + np.eye(ns)
    t0 = time.perf_counter()
    orc.inv_cholesky(Cs)
    t_inv = time.perf_counter() - t0
    t_inv_full = t_inv * (n / ns) ** 3
    total = t_pack * M / Ms + t_pairs_full + t_freq_full + t_inv_full
    return {
        "value": total, "unit": "s", "cores": cores, "kind": "port",
        "pairs_per_s": 2 * pairs_sample / (t_theta + t_cnt),
        "stages_s": {"pair_sweeps_x2": t_pairs_full, "frequencies": t_freq_full, "chol_inverse": t_inv_full},
        "sample": (f"oracle port (C/OpenMP + SciPy LAPACK), {cores} threads: pair sweeps on rows [0,{k1}) of the first {Ms} "
                   f"sequences ({pairs_sample:.3g} pairs x 2 passes, scaled by pair count); frequencies on {ks} sequences "
                   f"(scaled by M); dpotrf+dpotri at n={ns} (scaled by n^3); measured {t_theta + t_cnt + t_freq + t_inv:.1f} s "
                   f"of CPU work, extrapolated to L={L}, M={M}"),
    }


def run_reference(args, L, M, score, pc):
    """--impl reference: the reference's own CPU path.  Julia + DCAUtils are not installed in this image
    (probed; no network), so this times the oracle port with all host threads on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_baseline(L, M, score, pc, 5.0)
    cb = None
    for _ in range(max(1, min(args.steps, 3))):
        cb = cpu_baseline(L, M, score, pc)
        vals.append(cb["value"])
    v = sum(vals) / len(vals)
    cb["value"] = v
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "s", "n_gpus": args.gpus, "steps": len(vals),
        "warmup": 1, "ms_per_step": v * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic L={L} M={M} theta=auto score={score} pseudocount={pc}", "seed": SEED,
                   "note": "julia/DCAUtils absent: oracle port on host cores, bounded sample extrapolated"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    L, M, score, pc = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, L, M, score, pc)

    # stdout carries exactly ONE line (the JSON): everything libraries print to fd 1 (the NCCL version banner, ...) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    pkg = graft.load_package()
    from gaussdca_jl_b200 import _lib as glib
    from gaussdca_jl_b200 import dist as gdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(1, args.gpus) and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ["NCCL_DEBUG"] = os.environ.get("GDCA_NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = pkg.Context(local)   # raises loudly without the CUDA library / a B200
    lib = ctx.lib
    stream = torch.cuda.ExternalStream(int(lib.gdca_dev_stream(ctx.h)), device=local)
    W = max(3, args.warmup)
    K = max(1, args.steps)

    # synthetic alignment, generated on the device (identical bytes to oracle_synth_alignment)
    Zd = torch.empty((M, L), dtype=torch.int8, device=f"cuda:{local}")
    ctx.check(lib.gdca_synth_alignment_dev(ctx.h, ctypes.c_void_p(Zd.data_ptr()), L, M, SEED))
    n_out = int(lib.gdca_ranking_length(L, 5))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")  # > 126 MB L2

    def l2_flush():
        with torch.cuda.stream(stream):
            flush.zero_()

    st = glib.Stats()
    theta_code = -1.0
    launches0 = lib.gdca_dev_kernel_launches(ctx.h)

    def step_resident():
        if world == 1:
            ctx.check(lib.gdca_run_resident(ctx.h, ctypes.c_void_p(Zd.data_ptr()), L, M, theta_code, pc,
                                            glib.SCORE_CODES[score], 5, None, n_out, ctypes.byref(st)))
            return None
        return gdist.gdca_sharded(Zd, pc, "auto", score, 5, ctx=ctx, resident=True)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident, per-step CUDA events on the library stream, L2 flushed between steps
    for _ in range(W):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches1 = lib.gdca_dev_kernel_launches(ctx.h)
    stage_acc = {}
    barrier()
    for a, b in ev:
        l2_flush()
        a.record(stream)
        step_resident()
        b.record(stream)
        if world == 1:
            for k, v in st.asdict().items():
                if k.startswith("ms_"):
                    stage_acc[k] = stage_acc.get(k, 0.0) + v / K
    barrier()
    launches2 = lib.gdca_dev_kernel_launches(ctx.h)
    clocks = sampler.stop()
    cov_ms = ctypes.c_float()
    ctx.check(lib.gdca_dev_cov_kernel_ms(ctx.h, ctypes.byref(cov_ms)))   # cov_rows_kernel of the last timed step
    ms = sum(a.elapsed_time(b) for a, b in ev) / K
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item())
    stats = st.asdict()

    # ---- e2e: gdca_run() with pinned host buffers (H2D of Z, D2H of R inside the timed region)
    e2e = None
    if world == 1:
        Zh = torch.empty((M, L), dtype=torch.int8).pin_memory()
        Zh.copy_(Zd)
        Rh = torch.empty(n_out * 24, dtype=torch.uint8).pin_memory()
        e_ms = []
        for it in range(2 + K):
            l2_flush()
            torch.cuda.synchronize()
            ctx.check(lib.gdca_run(ctx.h, ctypes.c_void_p(Zh.data_ptr()), L, M, theta_code, pc, glib.SCORE_CODES[score], 5,
                                   ctypes.c_void_p(Rh.data_ptr()), n_out, ctypes.byref(st)))
            if it >= 2:
                e_ms.append(st.ms_total)   # CUDA events: before the H2D copy .. after the D2H copy
        e2e = {"value": sum(e_ms) / len(e_ms) / 1e3, "unit": "s", "h2d_bytes_per_step": L * M,
               "d2h_bytes_per_step": n_out * 24}
        R = np.frombuffer(Rh.numpy(), dtype=glib.RANK_DTYPE)
        top = [int(R["i"][0]), int(R["j"][0]), float(R["score"][0])]
    else:
        # N > 1: the public sharded API with HOST input on every rank (H2D inside), ranking copied to the host on rank 0
        Zh = torch.empty((M, L), dtype=torch.int8).pin_memory()
        Zh.copy_(Zd)
        Zn = Zh.numpy()
        e_ms = []
        for it in range(2 + K):
            l2_flush()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            R, _ = gdist.gdca_sharded(Zn, pc, "auto", score, 5, ctx=ctx, resident=False)
            b.record(stream)
            barrier()
            tt = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            if it >= 2:
                e_ms.append(float(tt.item()))
        e2e = {"value": sum(e_ms) / len(e_ms) / 1e3, "unit": "s", "h2d_bytes_per_step": L * M * world,
               "d2h_bytes_per_step": n_out * 24}
        top = [int(R["i"][0]), int(R["j"][0]), float(R["score"][0])] if rank == 0 else None

    if rank != 0:
        if dist is not None:
            # rank 0 still times its shard of the sweep below, and that kernel adds its hits into EVERY rank's counters over
            # peer memory: keep this rank's buffers alive until rank 0 is through
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (pair sweep), measured live
    roof, stages = None, None
    pk = peaks()
    if True:
        lop3, popc, dmma, dfma = (ctypes.c_double() for _ in range(4))
        ctx.check(lib.gdca_probe_peaks(ctx.h, ctypes.byref(lop3), ctypes.byref(popc), ctypes.byref(dmma), ctypes.byref(dfma)))
        # time the sweep kernel alone: the production launch (mode 1: neighbour counts, exact early exit)
        ctx.check(lib.gdca_dev_load_resident(ctx.h, ctypes.c_void_p(Zd.data_ptr()), L, M))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        thr = int(stats["thresh"]) if world == 1 else L // 2
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
        exact = {
            "kernel": "pair_sweep_kernel<5,1> (exact neighbour counts, per-warp early exit)",
            "bound": "int32_alu", "achieved": alu_ops / t_exact / 1e12, "peak": lop3.value, "unit": "Tlop3/s",
            "frac": (alu_ops / t_exact / 1e12) / lop3.value, "ms_per_launch": t_exact * 1e3,
            "blocks_swept": int(s_blocks.value), "blocks_total": (Mpad // 128) * (Mpad // 128 + 1) // 2 // world,
            "peak_source": "measured live: gdca_probe_peaks LOP3 issue rate (MEASURED_PEAKS.json has no INT32 figure)",
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
            # dominant kernel of the sweep: the tcgen05 prefilter.  algorithmic flop = 2 * 128 * BN * Kpad per tile x tiles
            fp4 = filt.value == 4
            bf16 = pk.get("bf16_tflops")
            tc_peak = 9000.0 if fp4 else 4500.0     # nominal dense FP4 / FP8 = INT8 (B200_PROFILING.md table)
            scaled = (4.0 if fp4 else 2.0) * bf16 if bf16 else None
            t_f = ms_f.value / 1e3
            roof = {
                "kernel": ("tc_filter_kernel<fp4> (tcgen05 kind::mxf4.block_scale 128x224x64" if fp4 else
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
                "l2_operand_bytes_per_launch": f_l2.value,
                "l2_operand_tb_per_s": f_l2.value / t_f / 1e12,
                "traffic": (ncu_traffic(f"tc_filter_kernel<{int(fp4)}, 1>") or ncu_traffic(f"tc_filter_kernel<{int(fp4)}, 0>"))
                           if world == 1 and args.workload == "C" else None,
                "exact_sweep": exact, **common,
            }
        else:
            roof = {**exact, "traffic": ncu_traffic("pair_sweep_kernel<5, 1>") if world == 1 and args.workload == "C" else None,
                    **common}
        ctx.check(lib.gdca_set_shard(ctx.h, 0, 1))
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()          # releases the other ranks (see above)
        n = 20 * L
        t_cov, t_chol = stage_acc.get("ms_cov", 0) / 1e3, (stage_acc.get("ms_chol", 0) + stage_acc.get("ms_inv", 0)) / 1e3
        stages = None if world > 1 else {
            "ms": {k: round(v, 4) for k, v in stage_acc.items()},
            "theta_passes": stats["theta_passes"],
            "weights_pairs_per_s": npairs / ((stage_acc.get("ms_theta", 0) + stage_acc.get("ms_weights", 0)) / 1e3 + 1e-30),
            "cov_fp64_equiv_tflops": M * n * (n + 1) / t_cov / 1e12 if t_cov else None,
            "chol_inv_tflops": n ** 3 / t_chol / 1e12 if t_chol else None,
            "cov_plus_inv_fp64_equiv_tflops": (M * n * (n + 1) + n ** 3) / (t_cov + t_chol) / 1e12 if t_cov else None,
            "dmma_peak_tflops_measured": dmma.value, "dfma_peak_tflops_measured": dfma.value,
            "chol_inv_frac_of_dmma_peak": (n ** 3 / t_chol / 1e12) / dmma.value if t_chol else None,
        }

    # ---- one roofline entry per hot kernel; "roofline" = the kernel with the largest share of the step
    roof_kernels = None
    if world == 1 and stages is not None:
        sm_clk = (clocks.get("sm_mhz") or pk.get("sm_max_mhz") or 1965.0) * 1e6
        smem_peak = 148 * 128 * sm_clk / 1e12                       # TB/s: 128 B/clk/SM shared-memory crossbar
        t_cv = cov_ms.value / 1e3
        rmw = M * L * (L + 1) // 2                                   # FP64 additions = 8-byte smem read + 8-byte write each
        cov_entry = {
            "kernel": "cov_rows_kernel<2> (weighted one-hot covariance as M*L(L+1)/2 private shared-memory FP64 adds)",
            "bound": "shared_memory", "achieved": 16 * rmw / t_cv / 1e12, "peak": smem_peak, "unit": "TB/s",
            "frac": 16 * rmw / t_cv / 1e12 / smem_peak,
            "peak_source": "148 SMs x 128 B/clk (B300_MICROARCH.md shared-memory crossbar) x SM clock under load",
            "ms_per_launch": cov_ms.value, "adds_per_launch": rmw,
            "fp64_equivalent_dense_tflops": M * n * (n + 1) / t_cv / 1e12,
            "traffic": ncu_traffic("cov_rows_kernel<2>") if args.workload == "C" else None,
            "hbm": {"achieved": (L * M + 8 * n * n / 2) / t_cv / 1e9, "peak": pk.get("hbm_gbs"), "unit": "GB/s",
                    "note": "compulsory bytes (recoded alignment once + upper half of C once); not the limiter"},
            "note": "ncu (profiles/r1_top_kernels.md): 0.79 of the 1 wavefront/clk/SM shared-memory pipe incl. staging",
        }
        inv_entry = {
            "kernel": "dgemm_kernel<*> + diag_block_kernel (blocked Cholesky, trtri by recursive doubling, lauum: ~275 launches)",
            "bound": "fp64_tensor", "achieved": n ** 3 / t_chol / 1e12, "peak": dmma.value, "unit": "TFLOP/s",
            "frac": (n ** 3 / t_chol / 1e12) / dmma.value, "peak_source": "measured live: gdca_probe_peaks DMMA.8x8x4 rate",
            "ms_per_step": t_chol * 1e3,
        }
        roof_kernels = [cov_entry, roof, inv_entry]
        roof = cov_entry if cov_ms.value >= roof.get("ms_per_launch", 0) else roof

    cb = None
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(L, M, score, pc)

    line = {
        "metric": METRIC, "value": ms_step / 1e3, "unit": "s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"synthetic L={L} M={M} theta=auto score={score} pseudocount={pc} min_separation=5 "
                               f"(BASELINE.json configs[{'BCD'.find(args.workload) + 1}])",
                   "seed": SEED, "generator": "SURVEY 8(d) clustered SplitMix64", "l2": "256 MiB flush write between steps",
                   "sharding": ("single GPU" if world == 1 else f"pair-matrix row blocks + covariance rows over {world} ranks, "
                                "exchange fused into the kernels over CUDA-IPC peer memory; inverse+scores on rank 0")},
        "theta": stats["theta"] if world == 1 else None, "thresh": stats["thresh"] if world == 1 else None,
        "meff": stats["meff"] if world == 1 else None, "top_pair": top,
        "e2e": e2e, "gpu_launches": int(launches2 - launches1),
        "clocks": clocks, "roofline": roof, "roofline_kernels": roof_kernels, "stages": stages, "cpu_baseline": cb,
    }
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
