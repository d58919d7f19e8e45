#!/usr/bin/env python
"""bench.py — MUSE sims/sec (MAP+score) on B200, with roofline and a CPU baseline beside it.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the config the metric is quoted on at 1/2/4/8 B200; it fits one
GPU): Neal's funnel, 2^16 latent dims, nsims = 2048, full solve θ̂ / J / H
(``muse(prob, 1.0; nsims, get_covariance=true)``, prior N(0,3), ∇z_logLike_atol = 1e-2, data at
θ_true = 0).  One *step* = one such full solve.  The metric counts the MAP+score units actually
executed (outer iterations × (nsims+1) + 1 fiducial + 2·nθ·(nsims÷10) finite-difference solves; the
reference's redundant centre evaluations and duplicate fiducial solves are neither executed nor
counted) divided by the device time of the step.  With N GPUs the default is the north star's split
(``--scaling strong``: the 2048 sims are sharded, N/8 per rank at 8); the weak-scaling figure (2048 sims
per GPU) rides along as ``other_scaling``.  A block of K steps is repeated until ≈ 0.6 s have been timed
(so that the clock record holds > 20 samples); the reported time is the median block.  The default
single-GPU run also measures the other BASELINE configs (C1, C2, C4, C5) into ``configs``.

value   device-resident: base normals and data already in HBM, K steps bracketed by CUDA events on
        the launch stream (barrier + synchronize on both sides), max over ranks.
e2e     the same solve through the public API with host buffers: every step uploads the observed
        data from pinned host memory, regenerates the base normals from the seed on the device
        (the reference draws them inside sample_x_z), solves, and reads all results back.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

D = 65536
NSIMS = 2048
THETA0 = 1.0
ATOL = 1e-2
DATA_SEED = 20261017
SIM_SEED = 314159


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default, the north-star split): --nsims sims in total, N/world per rank; weak: --nsims per rank")
    ap.add_argument("--d", "--dim", dest="d", type=int, default=D, help="latent dimension (use --dim under torchrun: its parser finds --d ambiguous)")
    ap.add_argument("--nsims", type=int, default=NSIMS, help="sims in total (strong) or per GPU (weak)")
    ap.add_argument("--family", default="funnel", choices=["funnel", "hiergauss", "corrgauss", "twolayer"])
    ap.add_argument("--group", type=int, default=0)
    ap.add_argument("--cluster", type=int, default=0)
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 generic solver only, 2 streaming kernel first")
    ap.add_argument("--cpu-sims", type=int, default=0, help="sims in the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the C1/C2/C4/C5 records of the default single-GPU run")
    ap.add_argument("--no-other-scaling", action="store_true", help="N > 1: skip the sub-record of the other scaling mode")
    return ap.parse_args()


def corr_consts(d, seed=5):
    """F3 constants: Σ₀ = A·Aᵀ/d + 0.1·I, A ~ N(0,1) from a stated seed; P = Σ₀⁻¹, L = chol Σ₀."""
    rng = np.random.Generator(np.random.Philox(seed))
    A = rng.standard_normal((d, d))
    S0 = A @ A.T / d + 0.1 * np.eye(d)
    return np.linalg.inv(S0), np.linalg.cholesky(S0)


def observed_data(family, d, L=None):
    """Synthetic observation at θ_true (0 for the funnel and corrgauss, (0,0) for hiergauss); host NumPy, seeded."""
    rng = np.random.Generator(np.random.Philox(DATA_SEED))
    xi, nu = rng.standard_normal(d), rng.standard_normal(d)
    if family == "twolayer":                           # the docstring hierarchy at σ_true = 0.3: data = (x, y) stacked, d = 2n
        n = d // 2
        w = np.exp(0.3 / 4) * xi[:n] + xi[n:]
        x = w + nu[:n]
        return np.concatenate([x, x + nu[n:]])
    return (L @ xi if L is not None else xi) + nu      # sig = 1, mu = 0 at θ_true


def theta_start(family):
    if family == "twolayer":
        return np.array([0.5])                         # muse(prob, (σ=0.5, θ=0)), src/turing.jl:78
    return np.array([0.5, 0.3]) if family == "hiergauss" else np.array([THETA0])


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """SM clock, power and throttle reasons of one GPU DURING the timed region.  In-process NVML (pynvml) from a
    sampling thread: an external `nvidia-smi -lms 100` loop cost ≈ 5 % of an 8-rank step on the bench box (it takes
    driver locks that every rank's launches contend on); `nvidia-smi` remains the fallback when pynvml is missing."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0, period=0.02):
        self.index, self.period = index, period
        self.sm, self.mx, self.reasons = [], [], set()
        self.rows, self.proc, self.thread, self.stop_flag, self.nvml = [], None, None, False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [t.strip() for t in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ----------------------------------------------------------------------------- CPU arm
_CPU_PROBLEMS = {}


def cpu_problem(family, d, nsims):
    """Oracle problem of the bench shape on seeded host draws (built once per shape: 2·nsims·d normals)."""
    import oracle as O

    key = (family, d, nsims)
    if key not in _CPU_PROBLEMS:
        if family == "corrgauss":
            P, L = corr_consts(d)
            fam = O.make_family(family, d, P=P, L=L)
        else:
            fam = O.make_family(family, d)
        rng = np.random.Generator(np.random.Philox(SIM_SEED))
        draws = O.Draws(rng.standard_normal((nsims, d)), rng.standard_normal((nsims, d)),
                        rng.standard_normal(d), rng.standard_normal(d))
        prior = O.FlatPrior() if family == "hiergauss" else O.NormalPrior(0, 3)
        _CPU_PROBLEMS[key] = O.OracleProblem(fam, observed_data(family, d, fam.L if family == "corrgauss" else None), draws, prior)
    return _CPU_PROBLEMS[key]


def cpu_solve_rate(family, d, nsims, nthreads, reps=1):
    """Full solve with the oracle's C port on the host cores; returns (units/s, units, seconds, threads)."""
    from oracle import cmuse, cport

    cport.build()
    prob = cpu_problem(family, d, nsims)
    threads = nthreads or cport.max_threads()
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        _, units = cmuse.muse_cpu(prob, theta_start(family), nsims=nsims, gradz_logLike_atol=ATOL, nthreads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return units / best, units, best, threads


def cpu_sample_nsims(args):
    """Bounded sample of the workload for the CPU arm: the full per-GPU batch when its two base-normal arrays fit
    ~4.5 GB of host memory (C3: 2048 × 65536 → 2.1 GB), otherwise as many sims as do."""
    if args.cpu_sims:
        return args.cpu_sims
    if args.family == "corrgauss":      # 2·d² flop per evaluation, ≈ 20 evaluations per unit: keep the sample to ≈ 10–30 s of CPU work
        return int(max(8, min(args.nsims, 2.0e11 / (40.0 * args.d * args.d * 2.5))))
    return int(max(16, min(args.nsims, 4.5e9 / (16.0 * args.d))))


def workload_name(family, d, nsims, scaling, world):
    cfg_name = {"funnel": "BASELINE configs[2]" if d == 65536 else ("BASELINE configs[1]" if (d == 512 and nsims == 10000) else
                          ("BASELINE configs[0]" if (d == 512 and nsims == 100) else "funnel, custom shape")),
                "hiergauss": "BASELINE configs[3]", "corrgauss": "BASELINE configs[4]",
                "twolayer": "the toy hierarchy of the reference's Turing docstring, src/turing.jl:63-79; not a BASELINE config"}[family]
    per = "" if world == 1 else (" per GPU" if scaling == "weak" else " in total, sharded over the GPUs")
    return (f"{family} d={d} nsims={nsims}{per}: full solve θ̂/J/H = muse(prob, θ₀={theta_start(family).tolist()}; nsims, "
            f"get_covariance=true), ∇z_logLike_atol={ATOL} ({cfg_name})")


def bench_config(args, world):
    """The `config` object — the SAME in the GPU arm and in the reference arm (the driver compares them): only what
    names the workload, nothing measured."""
    nsims_total = args.nsims * world if args.scaling == "weak" else args.nsims
    per_gpu = nsims_total / world
    return {"workload": workload_name(args.family, args.d, args.nsims, args.scaling, world),
            "family": args.family, "d": args.d, "nsims_total": nsims_total, "theta0": theta_start(args.family).tolist(),
            "atol": ATOL, "scaling": args.scaling,
            "l2": "inputs_larger_than_l2" if 2 * per_gpu * args.d * 8 > 126e6 else
                  "inputs fit the 126 MB L2 (%.1f MB of base normals per GPU): a latency-bound configuration, L2 is not flushed" % (2 * per_gpu * args.d * 8 / 1e6)}


def run_reference(args):
    """--impl reference: the CPU implementation of the path (oracle C port; the Julia reference cannot
    run in this image) on all host threads, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    nsims = cpu_sample_nsims(args)
    for _ in range(min(args.warmup, 1)):
        cpu_solve_rate(args.family, args.d, nsims, 0)
    rates, secs, units, threads = [], 0.0, 0, 0
    for _ in range(args.steps):
        r, u, dt, threads = cpu_solve_rate(args.family, args.d, nsims, 0)
        rates.append(r)
        secs += dt
        units += u
    value = units / secs
    sample = (f"full solve (θ̂,J,H) of {args.family} d={args.d} on nsims={nsims} of {args.nsims} per step; oracle C port (pthreads), "
              "analytic gradients — the Julia reference cannot run here and would be slower (serial pool, AD gradients)")
    line = {
        "impl": "reference", "metric": "MUSE sims/sec (MAP+score)", "value": value, "unit": "sims/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, world),
        "cpu_baseline": {"value": value, "unit": "sims/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ----------------------------------------------------------------------------- GPU arm
MIN_TIMED_SECONDS = 0.6       # blocks of K steps are repeated until the timed regions add up to this (clock record > 20 samples)


def src_sha256():
    """sha256 of the library's CUDA sources (what scripts/ncu_traffic.py stamps a capture with)."""
    import hashlib
    h = hashlib.sha256()
    cs = os.path.join(ROOT, "museinference.jl_b200", "csrc")
    for f in sorted(os.listdir(cs)) + [os.path.join("..", "..", "include", "muse_b200.h")]:
        with open(os.path.join(cs, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def measured_traffic(family, d, nsims):
    """DRAM bytes per solver launch from an `ncu --set full` capture of THIS build of the library (scripts/ncu_traffic.py
    writes profiles/r02_solver_traffic.json with the sha256 of the library's sources); null when the capture is of another build or shape."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_solver_traffic.json")) as fh:
            t = json.load(fh)
        if t.get("src_sha256") == src_sha256() and (t.get("family"), t.get("d"), t.get("nsims")) == (family, d, nsims):
            return t.get("dram_bytes_per_launch_avg")
    except Exception:
        pass
    return None


class Workload:
    """One problem of the bench on this rank's GPU: the solve closure and its timing helpers."""

    def __init__(self, m, torch, dist, family, d, nsims_total, pool, stream, args, world):
        self.m, self.torch, self.dist, self.world, self.stream = m, torch, dist, world, stream
        self.family, self.d, self.nsims_total, self.pool = family, d, nsims_total, pool
        P = Lc = None
        if family == "corrgauss":
            P, Lc = corr_consts(d)
        self.x_host = torch.from_numpy(observed_data(family, d, Lc)).pin_memory()
        prior = m.FlatPrior() if family == "hiergauss" else m.NormalPrior(0, 3)
        self.prob = m.SimpleMuseProblem(self.x_host.numpy(), family, prior, P=P, L=Lc, group=args.group, cluster=args.cluster,
                                        kernel=args.kernel, stream=stream.cuda_stream)
        self.th0 = theta_start(family)
        self.res = None

    def solve(self, seed):
        self.res = self.m.muse(self.prob, self.th0, rng=seed, nsims=self.nsims_total, gradz_logLike_atol=ATOL, get_covariance=True,
                               pool=self.pool)
        return self.res

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def warm(self, n):
        for w in range(n):
            self.solve(SIM_SEED)
            if w == 0:
                # per-launch event timing on from the second warm-up solve on: the library's graph of a solve bakes the
                # event-record nodes in, so the timed steps must see the configuration they were warmed with
                self.prob._backend.profile_reset(True)

    def timed_blocks(self, steps, body, min_seconds=MIN_TIMED_SECONDS, max_blocks=200):
        """Blocks of exactly `steps` steps, each bracketed by barrier + synchronize and CUDA events on the launch stream;
        returns the per-block times (ms, this rank)."""
        torch = self.torch
        out, total = [], 0.0
        while (total < min_seconds * 1e3 and len(out) < max_blocks) or not out:
            self.sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(self.stream)
            for s in range(steps):
                body(s)
            e1.record(self.stream)
            self.sync_all()
            wall = 1e3 * (time.perf_counter() - t0)
            out.append((e0.elapsed_time(e1), wall))
            total += wall
            if self.world > 1:      # every rank must take the same number of blocks
                t = torch.tensor([total], dtype=torch.float64, device="cuda")
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
                total = float(t.item())
        return out

    def reduce_max(self, vals):
        if self.world == 1:
            return list(vals)
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def reduce_sum(self, vals):
        if self.world == 1:
            return list(vals)
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.tolist()

    def close(self):
        self.prob.close()


def measure_value(wl, steps, warmup, min_seconds=MIN_TIMED_SECONDS):
    """Device-resident arm: base normals and data in HBM, blocks of `steps` full solves.  Returns a dict with the median
    block (max over ranks per block), the profile of the timed blocks and the units executed."""
    wl.warm(warmup)
    be = wl.prob._backend
    wl.sync_all()
    be.profile_reset(True)
    blocks = wl.timed_blocks(steps, lambda s: wl.solve(SIM_SEED), min_seconds)
    prof = be.profile()
    passes = be.profile_passes()
    be.profile_reset(False)
    ev = wl.reduce_max([b[0] for b in blocks])            # per block: max over ranks of the device time
    nb = len(ev)
    ms_block = float(np.median(ev))
    units_all, solve_ms_sum, solve_bytes_sum, launches_sum, flops_sum = wl.reduce_sum(
        [prof["solve_units"], prof["solve_ms"], prof["solve_bytes"], float(prof["launches"]), prof["solve_flops"]])
    return dict(ms_block=ms_block, blocks=nb, block_ms_all=ev, units_per_block=units_all / nb, prof=prof, passes=passes,
                solve_ms_sum=solve_ms_sum, solve_bytes_sum=solve_bytes_sum, launches_per_block=launches_sum / nb, flops_sum=flops_sum)


def roofline_of(wl, mv, hbm_peak, peak_src, steps, kernel_flag):
    """`roofline` object of one measured workload (HBM-bound isotropic families; FP64 tensor pipe for corrgauss)."""
    torch = wl.torch
    prof, world = mv["prof"], wl.world
    n_launch = max(1, prof["solve_launches"] * world)
    by_pass = {}
    for kind, v in mv["passes"].items():
        if v["launches"]:
            e = {"launches_per_step": v["launches"] / (steps * mv["blocks"]), "units_per_launch": v["units"] / v["launches"],
                 "avg_launch_ms": v["ms"] / v["launches"], "sims_per_s": v["units"] / (v["ms"] * 1e-3) if v["ms"] > 0 else None}
            if v["bytes"] > 0 and v["ms"] > 0:
                e["algorithmic_gbs"] = v["bytes"] / 1e9 / (v["ms"] * 1e-3)
                e["frac_of_hbm_peak"] = e["algorithmic_gbs"] / hbm_peak
            by_pass[kind] = e
    share = mv["solve_ms_sum"] / world / (mv["ms_block"] * mv["blocks"])
    if wl.family == "twolayer":
        # F4 runs on the lock-step solver (a launch per round, host poll between rounds): at the docstring's size a pass is a chain
        # of ≈ 10 short launches — latency, not bandwidth; no roofline figure is claimed for it
        return {"bound": "hbm", "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src,
                "kernel": "pair_apply_kernel + corr_iter_kernel (lock-step L-BFGS, elementwise 2x2-block product in place of the DGEMM)",
                "kernel_share_of_step": share, "passes": by_pass,
                "note": "launch-latency-bound at this size (2 L-BFGS iterations per unit, one launch per round)"}
    if wl.family == "corrgauss":
        # FP64 tensor roofline: denominator = cuBLAS DGEMM of the same shape measured here (MEASURED_PEAKS.json carries only
        # bf16), numerator = 2·rows·d² per P-product ÷ CUDA-event time of the whole solver chains
        d = wl.d
        rows = (wl.nsims_total // world + 1 + 127) // 128 * 128
        a = torch.full((rows, d), 4.7e-4, dtype=torch.float64, device="cuda")
        b = torch.full((d, d), 4.7e-4, dtype=torch.float64, device="cuda")
        for _ in range(2):
            torch.matmul(a, b)
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        for _ in range(5):
            torch.matmul(a, b)
        t1e.record()
        torch.cuda.synchronize()
        cublas_tf = 2.0 * rows * d * d / (t0e.elapsed_time(t1e) / 5 * 1e-3) / 1e12
        del a, b
        ach = prof["solve_flops"] / (prof["solve_ms"] * 1e-3) / 1e12 if prof["solve_ms"] > 0 else 0.0
        return {"bound": "tensor", "achieved": ach, "peak": cublas_tf, "unit": "TFLOP/s", "frac": ach / cublas_tf, "traffic": None,
                "peak_source": "cuBLAS DGEMM (torch.float64 matmul) of the same shape, measured in this run",
                "kernel": "dgemm_dmma_kernel (mma.sync m8n8k4 f64 → DMMA.8x8x4) inside the lock-step solver chains",
                "flops_per_solver_pass": prof["solve_flops"] / max(1, prof["solve_launches"]),
                "avg_pass_ms": prof["solve_ms"] / max(1, prof["solve_launches"]), "kernel_share_of_step": share, "passes": by_pass}
    achieved = (mv["solve_bytes_sum"] / 1e9) / (mv["solve_ms_sum"] / 1e3) if mv["solve_ms_sum"] > 0 else 0.0
    redo = prof.get("redo_units", 0)
    if kernel_flag == 1:
        kname = "iso_solver_kernel (generic two-sweep MAP+score solver)"
    elif wl.d >= 4096 or kernel_flag == 2:
        kname = "iso_stream_kernel (single-pass MAP+score, TMA ring)"
    else:
        kname = "iso_warp_stream_kernel (single-pass MAP+score, one warp per unit)"
    one_launch = prof["solve_launches"] > 0 and prof["launches"] == prof["solve_launches"] and "fd" in by_pass
    if one_launch:
        # muse_b200_muse_solve ran every solve as ONE cooperative launch (solve_persist_kernel: all passes, the θ updates and the
        # covariance stage's launches as phases of one kernel); `achieved` = the algorithmic bytes of all its phases ÷ the CUDA-event
        # time of the launch; the per-pass split below comes from the kernel's own %globaltimer stamps
        kname = ("solve_persist_kernel (one cooperative launch per solve; phases = the passes of " +
                 ("iso_stream_kernel's TMA ring" if (wl.d >= 4096 or kernel_flag == 2) else "iso_warp_stream_kernel's warp-per-unit sweep") +
                 ", θ updates and get_H! stage in between; lazy ẑ + lean α=1 trial unless MUSE_LAZY=0 / MUSE_LEAN=0)")
    if redo:
        kname += f" + iso_solver_kernel re-solve of {redo} handed-back units"
    return {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": measured_traffic(wl.family, wl.d, wl.nsims_total) if world == 1 else None,
            "peak_source": peak_src, "kernel": kname, "redo_units": redo,
            "algorithmic_bytes_per_launch": mv["solve_bytes_sum"] / n_launch, "avg_launch_ms": mv["solve_ms_sum"] / n_launch,
            "kernel_share_of_step": share, "passes": by_pass,
            "passes_timed_by": "globaltimer stamps of the launch's CTA 0" if one_launch else "CUDA events around each launch chain"}


def run_b200(args):
    import gc

    import torch
    import torch.distributed as dist

    import museinference_jl_b200 as m

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1 and os.environ.get("MUSE_BENCH_PIN", "1") == "1":
        # one slice of the host cores per rank: the rank's thread spins in stream synchronisations between passes and
        # must not be migrated or share a core with another rank's (16 cores for 8 ranks on the bench box)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local_rank * per:(local_rank + 1) * per] or cores))
        except (AttributeError, OSError):
            pass
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        pool = m.ShardPool(device=local_rank)
    else:
        pool = m.LocalPool()
        pool.device = local_rank
    m.build_library()

    family, d = args.family, args.d
    nsims_total = args.nsims * world if args.scaling == "weak" else args.nsims
    stream = torch.cuda.Stream()          # the library launches on this stream; events are recorded on it
    torch.cuda.set_stream(stream)
    # multi-rank runs get at least 5 warm-up solves (the first exchanges of a fresh communicator / peer mapping set up
    # lazily); single rank at least 3 (the measurement rules ask for W ≥ 3; the library also needs one solve to allocate, one
    # with the profiling configuration of the timed steps and one to capture its CUDA graph of a solve)
    warmup = max(args.warmup, 5) if world > 1 else max(args.warmup, 3)

    wl = Workload(m, torch, dist, family, d, nsims_total, pool, stream, args, world)
    sampler = ClockSampler(local_rank)
    # like timeit: no cyclic-GC pauses inside the timed regions (a 1–2 ms pause in one rank stalls all ranks at the next exchange)
    gc.collect()
    gc.disable()
    wl.warm(warmup)
    if rank == 0:
        sampler.start()
    mv = measure_value(wl, args.steps, 0)
    be = wl.prob._backend
    res = wl.res
    geo = be.geometry()
    p2p_ready = be.p2p_info()[1] if world > 1 else False

    # ---- end-to-end arm ("e2e"): host data in, results out, draws regenerated from the seed ----
    th0 = wl.th0
    h2d = d * 8 + th0.size * 8
    d2h_box = [0]
    be.profile_reset(True)

    e2e_count = [0]

    def e2e_step(s):
        e2e_count[0] += 1                           # counted over all blocks: consecutive steps never share a seed, whatever --steps is
        wl.prob.set_data(wl.x_host.numpy())         # H2D of the step's input from pinned host memory
        r = wl.solve(SIM_SEED + 1 + (e2e_count[0] % 2))   # new seed ⇒ base normals regenerated on the device
        d2h_box[0] = (len(r.gs) * th0.size + len(r.Hs) * th0.size ** 2) * 8 + (len(r.history) * (nsims_total // world + 1) * 20)

    blocks_e = wl.timed_blocks(args.steps, e2e_step)
    prof_e = be.profile()
    be.profile_reset(False)
    clocks = sampler.stop() if rank == 0 else None
    ev_e = wl.reduce_max([max(b) for b in blocks_e])                 # per block: device events or wall clock, whichever is longer
    ms_e2e = float(np.median(ev_e))
    (units_e2e_all,) = wl.reduce_sum([prof_e["solve_units"]])
    units_e2e_block = units_e2e_all / len(ev_e)

    # ---- the other scaling mode as a sub-record (N > 1) ------------------------------------------------------------
    other = None
    if world > 1 and not args.no_other_scaling:
        o_mode = "weak" if args.scaling == "strong" else "strong"
        o_total = args.nsims * world if o_mode == "weak" else args.nsims
        wl.close()
        wl2 = Workload(m, torch, dist, family, d, o_total, pool, stream, args, world)
        mv2 = measure_value(wl2, args.steps, warmup, min_seconds=0.3)
        other = {"scaling": o_mode, "nsims_total": o_total, "value": mv2["units_per_block"] / (mv2["ms_block"] / 1e3), "unit": "sims/s",
                 "ms_per_step": mv2["ms_block"] / args.steps, "blocks": mv2["blocks"],
                 "kernel_share_of_step": mv2["solve_ms_sum"] / world / (mv2["ms_block"] * mv2["blocks"])}
        wl2.close()
    gc.enable()

    if rank == 0:
        peaks, peak_src = {}, "fallback"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
            peak_src = "measured"
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        value = mv["units_per_block"] / (mv["ms_block"] / 1e3)
        line = {
            "metric": "MUSE sims/sec (MAP+score)", "value": value, "unit": "sims/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": mv["ms_block"] / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(args, world),
            "timing": {"blocks": mv["blocks"], "steps_per_block": args.steps, "block_ms": [round(b, 4) for b in mv["block_ms_all"]],
                       "statistic": "median block, max over ranks per block; every block is bracketed by barrier + synchronize and "
                                    "timed with CUDA events on the launch stream"},
            "detail": {"outer_iterations": len(res.history), "units_per_step": mv["units_per_block"] / args.steps,
                       "solver_geometry": geo,
                       "theta_hat": [float(t) for t in res.theta], "sigma": [float(s) for s in np.sqrt(np.diag(res.Sigma))],
                       "exchange": None if world == 1 else ("nccl all-gather between launches" if os.environ.get("MUSE_EXCHANGE") == "nccl" or not p2p_ready
                                                            else "peer-mapped stores + flags inside the one-launch solve (NVLink, no collective launch)")},
            "clocks": clocks,
            "e2e": {"value": units_e2e_block / (ms_e2e / 1e3), "unit": "sims/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h_box[0], "ms_per_step": ms_e2e / args.steps, "blocks": len(ev_e),
                    "note": "draws regenerated on device from the seed each step (reference draws inside sample_x_z)"},
            "gpu_launches": int(round(mv["launches_per_block"])),
            "roofline": roofline_of(wl, mv, hbm_peak, peak_src, args.steps, args.kernel) if other is None else None,
        }
        if other is not None:
            line["other_scaling"] = other
            # the primary workload's handle was closed for the sub-record: its roofline comes from the saved profile
            line["roofline"] = roofline_of(wl, mv, hbm_peak, peak_src, args.steps, args.kernel)
        if world == 1 and not args.no_cpu_baseline:
            nsims_cpu = cpu_sample_nsims(args)
            cpu_solve_rate(family, d, nsims_cpu, 0)                       # warm-up (page-in of the host draws)
            rate, units, secs, threads = cpu_solve_rate(family, d, nsims_cpu, 0, reps=3)
            n1 = max(16, min(nsims_cpu, 128))
            rate1, units1, secs1, _ = cpu_solve_rate(family, d, n1, 1)
            line["cpu_baseline"] = {
                "value": rate, "unit": "sims/s", "cores": threads, "kind": "port",
                "sample": f"full solve (θ̂,J,H) on nsims={nsims_cpu} of the same shape, {units} units in {secs:.2f}s; "
                          "oracle C port with analytic gradients (faster than the Julia reference's AD path)",
                "value_1thread": rate1, "sample_1thread": f"the same on 1 thread (the reference's default pool is a serial map), nsims={n1}: {units1} units in {secs1:.2f}s"}
        if world == 1 and not args.no_extra_configs and (family, d, args.nsims) == ("funnel", D, NSIMS):
            if other is None:
                wl.close()
            line["configs"] = extra_configs(m, torch, dist, pool, stream, args, hbm_peak, peak_src)
        _emit(line)
    try:
        wl.close()
    except Exception:
        pass
    if world > 1:
        dist.destroy_process_group()


def extra_configs(m, torch, dist, pool, stream, args, hbm_peak, peak_src):
    """The other BASELINE configs through the same code, one record each (single GPU): C1, C2 (latency path), C4, C5 — and the fourth
    registered family (F4, the docstring hierarchy of the reference's Turing adapter) at its docstring size."""
    import copy
    out = []
    for name, family, d, nsims, steps, secs in (("C1", "funnel", 512, 100, 100, 0.3), ("C2", "funnel", 512, 10000, 50, 0.3),
                                                ("C4", "hiergauss", 100000, 4096, 5, 0.2), ("C5", "corrgauss", 4096, 8192, 2, 0.0),
                                                # the reference's Turing docstring model at its own size (src/turing.jl:63-79: n = 512, nsims = 100)
                                                ("F4", "twolayer", 1024, 100, 20, 0.2)):
        rec = {"name": name, "workload": workload_name(family, d, nsims, "weak", 1), "steps": steps}
        try:
            a = copy.copy(args)
            a.family, a.d, a.nsims, a.kernel, a.group, a.cluster = family, d, nsims, 0, 0, 0
            wl = Workload(m, torch, dist, family, d, nsims, pool, stream, a, 1)
            mv = measure_value(wl, steps, 3, min_seconds=secs)
            rec.update(value=mv["units_per_block"] / (mv["ms_block"] / 1e3), unit="sims/s", ms_per_step=mv["ms_block"] / steps,
                       blocks=mv["blocks"], outer_iterations=len(wl.res.history), units_per_step=mv["units_per_block"] / steps,
                       gpu_launches=int(round(mv["launches_per_block"])),
                       roofline=roofline_of(wl, mv, hbm_peak, peak_src, steps, 0))
            wl.close()
        except Exception as e:      # an extra config must never take the headline line down with it
            rec["error"] = f"{type(e).__name__}: {e}"
        out.append(rec)
    return out


def main():
    args = parse()
    # Libraries (NCCL's version banner, torchrun notices) may print to stdout; the contract is ONE JSON line there.
    # Route fd 1 to stderr for the duration of the run and emit the JSON line on the saved descriptor.
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    global _emit

    def _emit(line):
        sys.stdout.flush()
        os.write(saved, (json.dumps(line) + "\n").encode())

    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
