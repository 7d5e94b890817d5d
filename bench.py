#!/usr/bin/env python
"""Benchmark of the constrained-TMVN sampling hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU via torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU chain on the host cores

Workload (config.workload): BASELINE.json configs[3] — 1-D bounded + monotone lineqGP, m = 1000 hat
knots (d = 2000 eta coordinates, 2999 finite bound rows), HMC, 10^6 samples sharded over 8 GPUs,
i.e. 125 000 samples per GPU (weak scaling: per-GPU work fixed).  tmvrnorm.HMC defaults:
burn.in = 100 per chain; one chain per resident CTA slot of the device (296 on a B200 at d = 2000)
x 423 samples.  One "step" = one full pass: every chain's burn-in + its samples.  Synthetic data, seeded (lineqgpr_b200.workloads.cfg4).

value   = samples/s with all inputs resident in HBM (lqg_hmc_box, LQG_MEM_DEVICE), CUDA events.
e2e     = samples/s through the C ABI with pinned HOST buffers (H2D of mu,U,Sigma,lb,ub,init and
          D2H of the d x N sample matrix inside the timed region).
The one-off O(d^3) host whitening (solve/chol, what tmg does in R) is outside both, as it is
outside the reference arm (which times the chain at the same .Call boundary); it is reported
as setup_s.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "constrained TMVN samples/sec (HMC, m=1000 knots)"
UNIT = "samples/s"
WORKLOAD = ("cfg4: 1-D bounded+monotone lineqGP, m=1000 knots (d=2000, 2999 bound rows), HMC T=pi/2, "
            "125000 samples per GPU (10^6 over 8), burn.in=100 per chain, one chain per resident CTA slot")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=["cfg4", "cfg5"])
    ap.add_argument("--nsim-per-gpu", type=int, default=125000)
    ap.add_argument("--chains-per-gpu", type=int, default=0, help="0 = lqg_hmc_box_slots(d)")
    ap.add_argument("--burn-in", type=int, default=100)
    ap.add_argument("--prepass", type=int, default=-1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-transitions", type=int, default=0, help="transitions per core for the CPU sample (0 = auto)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def build_problem(name):
    from lineqgpr_b200 import workloads as W
    from lineqgpr_b200.samplers import whiten_box, EPSILON
    par, aux = (W.cfg4() if name == "cfg4" else W.cfg5(n_test=10))
    mu = np.asarray(par["mu"], float); Sigma = np.asarray(par["Sigma"], float)
    lb = np.asarray(par["lb"], float); ub = np.asarray(par["ub"], float)
    map_ = np.asarray(par.get("map", mu), float)
    init = np.minimum(np.maximum(map_, lb + EPSILON), ub - EPSILON)          # R/lineqGPsamplers.R:133-135
    t0 = time.perf_counter()
    U = whiten_box(mu, Sigma)                                                # one Cholesky (UL factor = tmg's Ri)
    setup_s = time.perf_counter() - t0
    return dict(mu=mu, Sigma=Sigma, lb=lb, ub=ub, init=init, U=U, setup_s=setup_s, d=mu.size,
                c=int(np.isfinite(lb).sum() + np.isfinite(ub).sum()))


def bind_to_gpu_numa(index):
    """Restrict this rank to the CPUs NVML reports as local to its GPU, so that first-touch places
    the pinned host buffers of the e2e leg on that NUMA node.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(n_cpu, 64) + 63) // 64 + 8)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            return len(use)
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.t_lo, self.t_hi = 0.0, float("inf")

    def window(self, lo, hi):
        self.t_lo, self.t_hi = lo, hi

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
        rows = [r for t, r in self.rows if self.t_lo <= t <= self.t_hi] or [r for _, r in self.rows]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(P, n_tr, cores, seed0=100, tuned=False):
    """The reference chain on the host cores: one serial chain per core (tmg is one serial Markov
    chain by construction), canonical-frame inputs of the SAME workload, run by oracle/_ref (the
    UNMODIFIED tmg HmcSampler.cpp) when that library exists, else by the C++ restatement.
    tuned=True times the restatement WITHOUT tmg's per-constraint struct copies
    (tmg:src/HmcSampler.cpp:156,256), so that the speed-up is not inflated by an avoidable
    CPU inefficiency.
    Returns (samples_per_s, kind, wall_s, info)."""
    from oracle import lineqgpr_oracle as O
    O.build(ref=True)
    H, r, init, f, g = O.hmc_setup(P["mu"], P["Sigma"], P["lb"], P["ub"], P["init"])
    w = O.whiten(H, r, init, f, g)
    use_ref = O.ref_lib() is not None
    results = [None] * cores
    burn = 3                                   # untimed: leave the cold clamped-MAP start

    def run(n, seed, z0):
        if use_ref and not tuned:
            return O.ref_rtmg_canonical(n, seed, z0, w["f2"], w["g2"])
        return O.rtmg_canonical(n, seed, z0, w["f2"], w["g2"], faithful_copy=not tuned)[0]

    starts = [None] * cores

    def warm(i):
        starts[i] = run(burn, seed0 + i, w["initial2"])[-1].copy()

    def work(i):
        t = time.perf_counter()
        z = run(n_tr, 7919 + seed0 + i, starts[i])
        results[i] = (time.perf_counter() - t, z.shape[0], {})

    th = [threading.Thread(target=warm, args=(i,)) for i in range(cores)]      # ctypes drops the GIL
    [t.start() for t in th]; [t.join() for t in th]
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    [t.start() for t in th]; [t.join() for t in th]
    wall = time.perf_counter() - t0
    n = sum(r[1] for r in results)
    return n / wall, ("reference" if (use_ref and not tuned) else "port"), wall, results


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    P = build_problem(a.workload)
    cores = os.cpu_count() or 1
    n_tr = a.cpu_transitions or 8
    times = []
    for s in range(a.warmup + a.steps):
        v, kind, wall, _ = cpu_reference_sample(P, n_tr, cores, seed0=1000 * s)
        if s >= a.warmup:
            times.append((v, wall))
    tot_samples = len(times) * cores * n_tr
    tot_wall = sum(w for _, w in times)
    value = tot_samples / tot_wall
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * tot_wall / max(1, len(times)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if a.workload == "cfg4" else a.workload,
                       "note": "CPU arm: one serial tmg chain per host core (3 untimed burn-in transitions from the clamped MAP, "
                               f"then a timed bounded sample of {n_tr} transitions per core per step); no burn-in is charged to the CPU"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{len(times)} steps x {cores} chains x {n_tr} transitions (d={P['d']}, c={P['c']})"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    from lineqgpr_b200 import _lib as L
    from lineqgpr_b200 import build as B
    B.build()
    lib = L.lib()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; lineqgpr_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local) if world > 1 else None       # pinned buffers land next to the GPU's PCIe root
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    P = build_problem(a.workload)
    d = P["d"]
    n_chains = a.chains_per_gpu or lib.lqg_hmc_box_slots(d)
    n_per = -(-a.nsim_per_gpu // n_chains)
    n_gpu = n_chains * n_per
    burn = a.burn_in

    f = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    Ut = torch.from_numpy(np.asfortranarray(P["U"]).T.copy()).to(dev)        # column-major bytes
    St = torch.from_numpy(np.asfortranarray(P["Sigma"]).T.copy()).to(dev)
    mu_t, lb_t, ub_t, init_t = f(P["mu"]), f(P["lb"]), f(P["ub"]), f(P["init"])
    out_t = torch.empty((n_gpu, d), dtype=torch.float64, device=dev)         # d x N column-major
    flush_t = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    stream = torch.cuda.current_stream()

    def step_device(seed):
        st = L.HmcStats()
        o = L.make_opts(mem=L.MEM_DEVICE, stream=stream.cuda_stream, seed=seed, n_chains=n_chains,
                        chain_offset=rank * n_chains, prepass=a.prepass)
        rc = lib.lqg_hmc_box(d, mu_t.data_ptr(), Ut.data_ptr(), 1, St.data_ptr(), lb_t.data_ptr(), ub_t.data_ptr(),
                             init_t.data_ptr(), 0, n_per, burn, o, out_t.data_ptr(), None, st)
        L.check(rc, "lqg_hmc_box")
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # FP64 roofline denominators, measured here on this device
    dfma, dmma = ctypes.c_double(), ctypes.c_double()
    L.check(lib.lqg_measure_fp64_peaks(ctypes.byref(dfma), ctypes.byref(dmma), None), "peaks")

    # the same whitening factor computed by the device set-up stage (reported beside setup_s; the
    # timed steps below keep using the host factor U as an input, like tmg's pre-whitened f2)
    setup_dev = None
    if rank == 0:
        U_dev = np.empty((d, d), order="F")
        Sig_f = L.f64(P["Sigma"])                       # keep the column-major copy alive across the call
        kms = ctypes.c_double()
        for _ in range(2):
            t0 = time.perf_counter()
            L.check(lib.lqg_whiten_box(d, L.ptr(Sig_f), None, L.ptr(U_dev), ctypes.byref(kms)), "whiten")
            wall = time.perf_counter() - t0
        setup_dev = {"whiten_call_s": wall, "whiten_kernels_ms": kms.value,
                     "max_rel_diff_vs_host_factor_product": float(np.abs(U_dev @ U_dev.T - P["Sigma"]).max() / np.abs(P["Sigma"]).max())}
        del U_dev, Sig_f

    clocks = ClockSampler(local); clocks.start()        # started before warm-up: nvidia-smi start-up stays out of the timed window
    for w in range(a.warmup):
        flush_t.fill_(w)
        step_device(10 + w)
    barrier()
    lib.lqg_launch_count(1)
    t_win0 = time.perf_counter()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    stats = []
    t_wall = time.perf_counter()
    for s in range(a.steps):
        flush_t.fill_(s)                                   # L2 flush between timed iterations
        ev[s][0].record(stream)
        stats.append(step_device(100 + s))
        ev[s][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = int(lib.lqg_launch_count(0))
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    tot_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot_ms, op=dist.ReduceOp.MAX)
    tot_ms = float(tot_ms.item())
    value = world * n_gpu * a.steps / (tot_ms * 1e-3)

    # ---- e2e: C ABI with pinned host buffers, copies inside the timed region ----
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    hU = torch.from_numpy(np.asfortranarray(P["U"]).T.copy()).pin_memory()
    hS = torch.from_numpy(np.asfortranarray(P["Sigma"]).T.copy()).pin_memory()
    hmu, hlb, hub, hinit = pin(P["mu"]), pin(P["lb"]), pin(P["ub"]), pin(P["init"])
    hout = torch.empty((n_gpu, d), dtype=torch.float64).pin_memory()

    def step_host(seed):
        st = L.HmcStats()
        o = L.make_opts(mem=L.MEM_HOST, stream=stream.cuda_stream, seed=seed, n_chains=n_chains,
                        chain_offset=rank * n_chains, prepass=a.prepass)
        rc = lib.lqg_hmc_box(d, hmu.data_ptr(), hU.data_ptr(), 1, hS.data_ptr(), hlb.data_ptr(), hub.data_ptr(),
                             hinit.data_ptr(), 0, n_per, burn, o, hout.data_ptr(), None, st)
        L.check(rc, "lqg_hmc_box(host)")
        return st

    t_win1 = time.perf_counter()
    clocks.window(t_win0, t_win1)
    step_host(7); step_host(8)
    barrier()
    e2e_steps = max(2, min(a.steps, 5))
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        step_host(200 + s)
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * n_gpu * e2e_steps / float(e2e_t.item())
    clk = clocks.stop()
    viol_host = int((hout.numpy() < P["lb"][None, :]).sum() + (hout.numpy() > P["ub"][None, :]).sum())

    # ---- roofline of the dominant kernel (hmc_box_kernel): FP64 CUDA-core pipe ----
    c = P["c"]
    tr = sum(s.transitions for s in stats); bo = sum(s.bounces for s in stats); se = sum(s.hit_searches for s in stats)
    rs = sum(s.vel_restarts + s.chk_restarts for s in stats)
    chain_ms = sum(s.chain_ms for s in stats); pre_ms = sum(s.prepass_ms for s in stats)
    # SURVEY.md §8(d) convention: 40 flop per constraint per hit search, 8d per bounce, 3d per
    # transition, d^2 per in-kernel momentum (restarts); the first momentum of every transition
    # is the pre-pass GEMM (d^2 each, triangular U halves the executed work)
    w_chain = 40.0 * c * se + 8.0 * d * bo + 3.0 * d * tr + float(d) * d * rs
    w_pre = float(d) * d * tr
    traffic = traffic_detail = None
    try:      # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
        tj = json.load(open(ROOT / "profiles" / "r01_traffic.json"))["hmc_box_kernel"]
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        traffic_detail = {"algorithmic_bytes_per_launch": tj["algorithmic_bytes"], "launch": tj["launch"],
                          "source": "profiles/r01_hmc_box_dgemm_ncu_full.txt",
                          "note": "traffic ~ algorithmic bytes (pool columns in, samples out): no re-reads; HBM is not the bound"}
    except Exception:
        pass
    roof = {"bound": "fp64", "kernel": "hmc_box_kernel", "achieved": w_chain / (chain_ms * 1e-3) / 1e12,
            "peak": dfma.value, "peak_source": "lqg_measure_fp64_peaks DFMA micro-benchmark, this run",
            "unit": "TFLOP/s", "frac": w_chain / (chain_ms * 1e-3) / 1e12 / dfma.value, "traffic": traffic,
            "launch_ms": chain_ms / len(stats), "share_of_step": chain_ms / sum(step_ms),
            "traffic_detail": traffic_detail,
            "note": "bound = FP64 CUDA-core pipe (SURVEY.md par.8d): state in registers, Sigma/U L2-resident; see roofline_hbm for the HBM view"}
    # the same kernel against the HBM roofline: algorithmic bytes = momentum columns read + samples written
    hbm_peak = None
    try:
        hbm_peak = json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception:
        hbm_peak = 6650.0                                  # fallback of B200_PROFILING.md
    att = tr + rs
    bytes_chain = 8.0 * d * (att + sum(s.transitions for s in stats) * (n_per / (n_per + burn)))
    roof_hbm = {"bound": "hbm", "kernel": "hmc_box_kernel", "achieved": bytes_chain / (chain_ms * 1e-3) / 1e9,
                "peak": hbm_peak, "unit": "GB/s", "frac": bytes_chain / (chain_ms * 1e-3) / 1e9 / hbm_peak,
                "traffic": traffic,
                "note": "algorithmic bytes per step = 8 d (momenta consumed + samples emitted); a small fraction of the HBM peak: not HBM-bound"}
    roof_pre = {"bound": "tensor", "kernel": "dgemm_dmma_kernel (momentum pre-pass)",
                "achieved": w_pre / (pre_ms * 1e-3) / 1e12 if pre_ms > 0 else None, "peak": dmma.value,
                "unit": "TFLOP/s", "frac": (w_pre / (pre_ms * 1e-3) / 1e12 / dmma.value) if pre_ms > 0 else None,
                "note": "algorithmic d^2 flop per momentum (triangular product); DMMA.8x8x4 peak measured this run",
                "launch_ms": pre_ms / len(stats)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": tot_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if a.workload == "cfg4" else a.workload, "d": d, "c": c,
                       "samples_per_gpu": n_gpu, "chains_per_gpu": n_chains, "burn_in": burn,
                       "l2": "256 MiB buffer written between timed iterations (L2 flush)",
                       "parallelism": f"chains sharded over {world} GPU(s), no data-path collective"},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int((2 * d * d + 4 * d) * 8), "d2h_bytes_per_step": int(n_gpu * d * 8),
                    "steps": e2e_steps, "api": "lqg_hmc_box, LQG_MEM_HOST, pinned buffers"},
            "gpu_launches": launches, "clocks": clk, "roofline": roof, "roofline_hbm": roof_hbm, "roofline_prepass": roof_pre,
            "setup_s": P["setup_s"], "setup_device": setup_dev, "numa_bound_cpus": numa, "wall_s_timed_region": t_wall, "step_ms": [round(x, 3) for x in step_ms],
            "counters": {"transitions": tr, "bounces_per_transition": bo / tr, "hit_searches": se,
                         "restarts_per_transition": rs / tr, "exact_evals_per_search": sum(s.exact_evals for s in stats) / se,
                         "violations": sum(s.violations for s in stats) + viol_host,
                         "failed_chains": sum(s.failed_chains for s in stats),
                         "bounce_steps_per_s": se / (chain_ms * 1e-3) * world,
                         "transitions_per_s": world * tr / (tot_ms * 1e-3)},
            "fp64_peaks_measured": {"dfma_tflops": dfma.value, "dmma_tflops": dmma.value}}

    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_tr = a.cpu_transitions or 40
        v, kind, wall, res = cpu_reference_sample(P, n_tr, cores)
        v2, kind2, wall2, _ = cpu_reference_sample(P, max(4, n_tr // 2), cores, tuned=True)
        line["cpu_baseline_tuned"] = {"value": v2, "unit": UNIT, "cores": cores, "kind": kind2,
                                      "sample": f"same chain without tmg's per-constraint struct copies, {max(4, n_tr // 2)} timed transitions per core"}
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                                "sample": f"{cores} serial chains x {n_tr} timed transitions after 3 untimed burn-in transitions, "
                                          f"same canonical inputs (d={d}, c={c}), wall {wall:.1f} s; burn-in not charged"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)
