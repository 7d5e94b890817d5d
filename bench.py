#!/usr/bin/env python
"""Benchmark of the constrained-TMVN sampling hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU via torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU chain on the host cores
    python bench.py --workload cfg5 ...                      # config 5: 10^6 samples + bands of Phi xi at 10^5 test points
    python bench.py --workload cfg1|cfg2|cfg3 [--sampler HMC|RSM|ExpT]   # the small BASELINE configs

Default workload (config.workload): BASELINE.json configs[3] — 1-D bounded + monotone lineqGP, m = 1000
hat knots (d = 2000 eta coordinates, 2999 finite bound rows), HMC, 10^6 samples sharded over 8 GPUs,
i.e. 125 000 samples per GPU (weak scaling: per-GPU work fixed).  tmvrnorm.HMC defaults: burn.in = 100
per chain; one chain per resident chain slot of the device (444 on a B200 at d = 2000: 3 chains per SM) x 282 samples.
One "step" = one full pass: every chain's burn-in + its samples.  Synthetic data, seeded
(lineqgpr_b200.workloads).

value   = samples/s with all inputs resident in HBM (lqg_hmc_box, LQG_MEM_DEVICE), CUDA events.
e2e     = samples/s through ONE reference-facing call of the simulate tail, lqg_simulate, with pinned
          HOST buffers: H2D of mu, U, Sigma, lb, ub, init, Lambda and Phi (or the test points), the
          sampler, xi = qr.solve(Lambda, eta), Phi xi and its mean / quantile bands on the device, D2H of
          xi (m x N) and the bands — all inside the timed region.  With N > 1 GPUs the xi shards and the
          band pieces are all-gathered by NCCL inside the call (lqg_comm).
e2e_tmvrnorm = the round-1 end-to-end number: lqg_hmc_box with host buffers, i.e. the d x N matrix of
          eta returned to the host (twice the bytes of xi at config 4).
The one-off O(d^3) host set-up (whitening factor, Lambda^+) is outside all of them, as it is outside
the reference arm (which times the chain at the same .Call boundary); it is reported as setup_s.
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
WORKLOADS = {
    "cfg4": ("cfg4: 1-D bounded+monotone lineqGP, m=1000 knots (d=2000, 2999 bound rows), HMC T=pi/2, "
             "125000 samples per GPU (10^6 over 8), burn.in=100, one chain per resident chain slot"),
    "cfg5": ("cfg5: additive monotone lineqAGP D=100 x 10 knots (m=d=1000, 900 bound rows), HMC T=pi/2, 125000 samples per GPU "
             "(10^6 over 8), burn.in=100; e2e adds xi = qr.solve(Lambda, eta) and the bands of Phi xi at 10^5 test points"),
    "cfg1": "cfg1: 1-D monotone lineqGP, m=100 knots, 10 noisy points, 1000 samples via simulate",
    "cfg2": "cfg2: 2-D monotone MaxMod lineqGP (knots 5 x 2, d=20, 13 bound rows), 10^4 samples",
    "cfg3": "cfg3: 5-D additive monotone MaxMod lineqAGP (knots 6 + 4, d=10, 8 bound rows), 10^5 samples",
}
PROBS = (0.025, 0.975)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--sampler", default="HMC", choices=["HMC", "RSM", "ExpT"], help="cfg1-3 only")
    ap.add_argument("--nsim-per-gpu", type=int, default=0, help="0 = 125000 (cfg4/cfg5) or the config's own nsim")
    ap.add_argument("--chains-per-gpu", type=int, default=0, help="0 = lqg_hmc_box_slots(d)")
    ap.add_argument("--burn-in", type=int, default=100)
    ap.add_argument("--prepass", type=int, default=-1)
    ap.add_argument("--fork", type=int, default=6, help="chains per family sharing one root chain's burn-in (lqg_opts.fork; 1 = independent chains)")
    ap.add_argument("--fork-burn-in", type=int, default=5, help="transitions a forked chain runs before it emits")
    ap.add_argument("--n-test", type=int, default=0, help="test points of the e2e pipeline (0 = 1000 for cfg4, 100000 for cfg5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-transitions", type=int, default=0, help="transitions per core for the CPU sample (0 = auto)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def build_problem(name, sampler="HMC"):
    from lineqgpr_b200 import workloads as W
    from lineqgpr_b200.samplers import whiten_box, EPSILON
    from lineqgpr_b200.simulate import lambda_pinv
    if name == "cfg4":
        par, aux = W.cfg4(n_test=10)
    elif name == "cfg5":
        par, aux = W.cfg5(n_test=10)
    else:
        par, aux = W.CONFIGS[name](sampler=sampler)
    mu = np.asarray(par["mu"], float); Sigma = np.asarray(par["Sigma"], float)
    lb = np.asarray(par["lb"], float); ub = np.asarray(par["ub"], float)
    map_ = np.asarray(par.get("map", mu), float)
    init = np.minimum(np.maximum(map_, lb + EPSILON), ub - EPSILON)          # R/lineqGPsamplers.R:133-135
    t0 = time.perf_counter()
    U = whiten_box(mu, Sigma)                                                # one Cholesky (UL factor = tmg's Ri)
    setup_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    Lp = lambda_pinv(aux["Lambda"])                                          # the operator of qr.solve(Lambda, .)
    pinv_s = time.perf_counter() - t0
    return dict(par=par, aux=aux, mu=mu, map=map_, Sigma=Sigma, lb=lb, ub=ub, init=init, U=U, Lambda=aux["Lambda"], Lp=Lp,
                setup_s=setup_s, pinv_s=pinv_s, d=mu.size, m=aux["Lambda"].shape[1],
                c=int(np.isfinite(lb).sum() + np.isfinite(ub).sum()))


def test_points(P, name, n_t):
    """Test points of the e2e pipeline and the knot vectors of the hat basis (Phi is built on the
    device from them: basisCompute.lineqGP, R/lineqGP.R:29-60, R/lineqAGP.R:341-348)."""
    mdl = P["aux"]["model"]
    if name == "cfg5":
        ulist = [np.asarray(u, float).ravel() for u in mdl["ulist"]]
        xt = np.random.default_rng(4).uniform(size=(n_t, len(ulist)))         # uniform test points, seed 4
    else:
        ulist = [np.asarray(mdl["localParam"]["u"], float).ravel()]
        xt = np.linspace(0.0, 1.0, n_t).reshape(-1, 1)
    return xt, ulist


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
def cpu_reference_sample(P, n_tr, cores, seed0=100, tuned=False, count_bounces=True):
    """The reference chain on the host cores: one serial chain per core (tmg is one serial Markov
    chain by construction), canonical-frame inputs of the SAME workload, run by oracle/_ref (the
    UNMODIFIED tmg HmcSampler.cpp) when that library exists, else by the C++ restatement.
    tuned=True times the restatement WITHOUT tmg's per-constraint struct copies
    (tmg:src/HmcSampler.cpp:156,256), so that the speed-up is not inflated by an avoidable
    CPU inefficiency.
    Returns (samples_per_s, kind, wall_s, info); info carries bounces per transition: the compiled
    reference has no counters, so chain 0 is replayed, untimed, by the restatement (which is pinned to
    it bit for bit) with the same seed and start."""
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
        results[i] = (time.perf_counter() - t, z.shape[0], z[-1].copy())

    th = [threading.Thread(target=warm, args=(i,)) for i in range(cores)]      # ctypes drops the GIL
    [t.start() for t in th]; [t.join() for t in th]
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    [t.start() for t in th]; [t.join() for t in th]
    wall = time.perf_counter() - t0
    n = sum(r[1] for r in results)
    info = {}
    if count_bounces:
        z, ci = O.rtmg_canonical(n_tr, 7919 + seed0, starts[0], w["f2"], w["g2"], faithful_copy=False)
        info = {"bounces_per_transition": ci["bounces"] / max(1, n_tr), "hit_searches_per_transition": ci["hit_searches"] / max(1, n_tr),
                "restarts": ci["vel_restarts"] + ci["chk_restarts"],
                "replay_matches_timed_chain": bool(np.array_equal(z[-1], results[0][2])) if (use_ref and not tuned) else None}
        info["bounce_steps_per_s"] = info["hit_searches_per_transition"] * n / wall
    return n / wall, ("reference" if (use_ref and not tuned) else "port"), wall, info


def cpu_small_sample(P, a, budget_s=8.0):
    """CPU leg of the small configs for the samplers that have no compiled reference: the numpy
    restatements of R/lineqGPsamplers.R:36-61 (RSM) and of Botev's mvrandn (ExpT), one core."""
    from oracle import lineqgpr_oracle as O
    par = P["par"]
    d = P["d"]
    t0 = time.perf_counter()
    if a.sampler == "RSM":
        rng = np.random.default_rng(0)
        n = 200
        got = 0
        while time.perf_counter() - t0 < budget_s and got < 4000:
            x, info = O.tmvrnorm_RSM(par, n, rng.standard_normal((d, 40 * n)), rng.uniform(size=40 * n))
            got += x.shape[1]
        kind = "numpy restatement of tmvrnorm.RSM (eigen factor hoisted)"
    else:
        from oracle import expt_oracle as E
        mu = P["mu"]
        try:
            su = E.setup(P["lb"] - mu, P["ub"] - mu, P["Sigma"])
        except ValueError:                     # the checker's own root finder gives up at d = 100: take the product's tilt
            from lineqgpr_b200 import expt     # (host numpy set-up, excluded from the timing on both sides)
            su = expt.setup(P["lb"] - mu, P["ub"] - mu, P["Sigma"])
        got, seed = 0, 0
        while time.perf_counter() - t0 < budget_s and got < 2000:
            x, _, _ = E.mvrandn(P["lb"] - mu, P["ub"] - mu, P["Sigma"], 50, seed, su=su)
            got += x.shape[1]; seed += 1
        kind = "scalar numpy restatement of TruncatedNormal::mvrandn (Botev 2017), set-up excluded"
    wall = time.perf_counter() - t0
    return got / wall, kind, wall, got


def reference_line(a, value, ms_per_step, cpu):
    return {"impl": "reference", "metric": METRIC if a.workload in ("cfg4", "cfg5") else f"constrained TMVN samples/sec ({a.sampler}, {a.workload})",
            "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": WORKLOADS[a.workload]}, "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    P = build_problem(a.workload, a.sampler)
    cores = os.cpu_count() or 1
    if a.workload not in ("cfg4", "cfg5") and a.sampler != "HMC":
        vals = []
        for s in range(a.warmup + a.steps):
            v, kind, wall, got = cpu_small_sample(P, a, budget_s=2.0)
            if s >= a.warmup:
                vals.append((v, wall, got))
        value = sum(g for _, _, g in vals) / sum(w for _, w, _ in vals)
        line = reference_line(a, value, 1e3 * sum(w for _, w, _ in vals) / len(vals),
                              {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"{kind}; {len(vals)} steps of ~2 s"})
        print(json.dumps(line), flush=True)
        return
    n_tr = a.cpu_transitions or (8 if a.workload in ("cfg4", "cfg5") else 2000)
    times, info = [], {}
    for s in range(a.warmup + a.steps):
        v, kind, wall, info = cpu_reference_sample(P, n_tr, cores, seed0=1000 * s, count_bounces=(s == a.warmup + a.steps - 1))
        if s >= a.warmup:
            times.append((v, wall))
    tot_samples = len(times) * cores * n_tr
    tot_wall = sum(w for _, w in times)
    value = tot_samples / tot_wall
    line = reference_line(a, value, 1e3 * tot_wall / max(1, len(times)),
                          {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                           "sample": f"{len(times)} steps x {cores} chains x {n_tr} transitions (d={P['d']}, c={P['c']})",
                           "counters_chain0_last_step": info})
    line["config"]["note"] = ("CPU arm: one serial tmg chain per host core (3 untimed burn-in transitions from the clamped MAP, "
                              f"then a timed bounded sample of {n_tr} transitions per core per step); no burn-in is charged to the CPU")
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
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from lineqgpr_b200.dist import Comm
        comm = Comm.from_torch_distributed()                    # the library's own NCCL communicator (lqg_comm)
    dev = torch.device("cuda", local)
    if a.workload in ("cfg1", "cfg2", "cfg3"):
        return run_small(a, torch, dist, L, lib, rank, world, local, dev)

    P = build_problem(a.workload)
    d, m = P["d"], P["m"]
    nsim_gpu = a.nsim_per_gpu or 125000
    n_chains = a.chains_per_gpu or lib.lqg_hmc_box_slots(d)
    n_per = -(-nsim_gpu // n_chains)
    n_gpu = n_chains * n_per
    burn = a.burn_in
    n_t = a.n_test or (100000 if a.workload == "cfg5" else 1000)
    xt, ulist = test_points(P, a.workload, n_t)

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
                        chain_offset=rank * n_chains, prepass=a.prepass, fork=a.fork, fork_burn_in=a.fork_burn_in)
        rc = lib.lqg_hmc_box(d, mu_t.data_ptr(), Ut.data_ptr(), 1, St.data_ptr(), lb_t.data_ptr(), ub_t.data_ptr(),
                             init_t.data_ptr(), 0, n_per, burn, o, out_t.data_ptr(), None, st)
        L.check(rc, "lqg_hmc_box")
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
    tot_ms = max_over_ranks(sum(step_ms))
    value = world * n_gpu * a.steps / (tot_ms * 1e-3)
    t_win1 = time.perf_counter()
    clocks.window(t_win0, t_win1)

    # ---- e2e: ONE lqg_simulate call per step, pinned host buffers, copies inside the timed region ----
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    pinF = lambda x: torch.from_numpy(np.asfortranarray(x).T.copy()).pin_memory()      # column-major bytes
    hU, hS, hLam = pinF(P["U"]), pinF(P["Sigma"]), pinF(P["Lambda"])
    hmu, hlb, hub, hinit = pin(P["mu"]), pin(P["lb"]), pin(P["ub"]), pin(P["init"])
    hxt, hknots = pinF(xt), pin(np.concatenate(ulist))
    m_k = np.ascontiguousarray([u.size for u in ulist], dtype=np.int32)
    hprobs = pin(np.asarray(PROBS, float))
    hxi = torch.empty((n_gpu, m), dtype=torch.float64).pin_memory()                   # m x N_local column-major
    hmean = torch.empty(n_t, dtype=torch.float64).pin_memory()
    hq = torch.empty((n_t, len(PROBS)), dtype=torch.float64).pin_memory()
    sa = L.SimulateArgs()
    sa.d, sa.m = d, m
    sa.mu, sa.Sigma, sa.lb, sa.ub, sa.init, sa.U = (t.data_ptr() for t in (hmu, hS, hlb, hub, hinit, hU))
    sa.Lambda = hLam.data_ptr()            # qr.solve(Lambda, eta), R/lineqGP.R:639: the call probes Lambda's structure itself
    sa.n_t = n_t
    sa.basis_D, sa.basis_mk, sa.xtest, sa.ldx, sa.knots = len(ulist), m_k.ctypes.data, hxt.data_ptr(), n_t, hknots.data_ptr()
    sa.probs, sa.nprobs = hprobs.data_ptr(), len(PROBS)
    sa.n_per_chain, sa.burn_in = n_per, burn
    sa.xi_out, sa.mean, sa.qtls = hxi.data_ptr(), hmean.data_ptr(), hq.data_ptr()
    rows_rank = -(-n_t // world)
    h2d_bytes = int((2 * d * d + 4 * d + m * d + rows_rank * len(ulist) + m + len(PROBS)) * 8)
    d2h_bytes = int((n_gpu * m + n_t * (1 + len(PROBS))) * 8)

    def step_pipeline(seed):
        st = L.SimulateStats()
        o = L.make_opts(mem=L.MEM_HOST, stream=stream.cuda_stream, seed=seed, n_chains=n_chains, prepass=a.prepass, fork=a.fork, fork_burn_in=a.fork_burn_in)
        rc = lib.lqg_simulate(ctypes.byref(sa), o, comm.handle if comm is not None else None, st)
        L.check(rc, "lqg_simulate")
        return st

    def timed(fn, n_steps, seed0):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        out = []
        for s in range(n_steps):
            evs[s][0].record(stream)
            out.append(fn(seed0 + s))              # synchronous call: results are in host memory when it returns
            evs[s][1].record(stream)
        barrier()
        return max_over_ranks(sum(e0.elapsed_time(e1) for e0, e1 in evs)), out

    e2e_steps = max(2, min(a.steps, 5))
    step_pipeline(7); step_pipeline(8)
    barrier()
    e2e_ms, pst = timed(step_pipeline, e2e_steps, 200)
    e2e_value = world * n_gpu * e2e_steps / (e2e_ms * 1e-3)
    xi_np = hxi.numpy()                                  # (N, m) view of the m x N column-major result
    mean_np, q_np = hmean.numpy().copy(), hq.numpy().copy()
    # every rank must hold the same bands (they were all-gathered); xi finite
    bands_equal = True
    if world > 1:
        chk = torch.from_numpy(np.concatenate([mean_np, q_np.ravel()])).to(dev)
        ref = chk.clone()
        dist.broadcast(ref, src=0)
        ok_t = torch.tensor([1.0 if torch.equal(chk, ref) else 0.0], device=dev)
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        bands_equal = bool(ok_t.item() == 1.0)
    last = pst[-1]
    pipeline = {"api": "lqg_simulate, LQG_MEM_HOST, pinned buffers" + (", lqg_comm (NCCL all-gather of xi shards and band pieces)" if world > 1 else ""),
                "n_test": n_t, "rows_per_rank": rows_rank, "probs": list(PROBS), "ms_per_step": e2e_ms / e2e_steps,
                "stage_ms_last_step_rank0": {"chain": last.hmc.chain_ms, "momentum_pool": last.hmc.prepass_ms, "xi_products": last.xi_ms,
                                             "phi_xi_gemm": last.gemm_ms, "bands": last.bands_ms, "device_total": last.total_ms},
                "units_per_step": int(last.rounds), "bands_identical_on_all_ranks": bands_equal,
                "xi_finite": bool(np.isfinite(xi_np).all()), "bands_finite": bool(np.isfinite(mean_np).all() and np.isfinite(q_np).all()),
                "paths_per_s": world * n_gpu * e2e_steps / (e2e_ms * 1e-3),
                "band_outputs_per_s": float(n_t) * world * n_gpu * e2e_steps / (e2e_ms * 1e-3),
                "phi_xi_tflops": 2.0 * rows_rank * m * (world * n_gpu) / (last.gemm_ms * 1e-3) / 1e12 if last.gemm_ms > 0 else None,
                "d2h_gbs_per_gpu_inside_call": d2h_bytes * e2e_steps / (e2e_ms * 1e-3) / 1e9}

    # ---- the round-1 e2e: lqg_hmc_box with host buffers (eta d x N to the host) ----
    e2e_tmv = None
    if a.workload == "cfg4" or world == 1:
        hout = torch.empty((n_gpu, d), dtype=torch.float64).pin_memory()

        def step_host(seed):
            st = L.HmcStats()
            o = L.make_opts(mem=L.MEM_HOST, stream=stream.cuda_stream, seed=seed, n_chains=n_chains,
                            chain_offset=rank * n_chains, prepass=a.prepass, fork=a.fork, fork_burn_in=a.fork_burn_in)
            rc = lib.lqg_hmc_box(d, hmu.data_ptr(), hU.data_ptr(), 1, hS.data_ptr(), hlb.data_ptr(), hub.data_ptr(),
                                 hinit.data_ptr(), 0, n_per, burn, o, hout.data_ptr(), None, st)
            L.check(rc, "lqg_hmc_box(host)")
            return st

        step_host(7)
        barrier()
        n2 = max(2, min(a.steps, 3))
        ms2, _ = timed(step_host, n2, 300)
        viol_host = int((hout.numpy() < P["lb"][None, :]).sum() + (hout.numpy() > P["ub"][None, :]).sum())
        # raw pinned D2H of the same size on every rank at once: what the host can sink
        dsrc = out_t
        torch.cuda.synchronize(); barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(2):
            hout.copy_(dsrc, non_blocking=True)
        e1.record(stream)
        barrier()
        raw_ms = max_over_ranks(e0.elapsed_time(e1))
        e2e_tmv = {"value": world * n_gpu * n2 / (ms2 * 1e-3), "unit": UNIT, "api": "lqg_hmc_box, LQG_MEM_HOST, pinned buffers",
                   "h2d_bytes_per_step": int((2 * d * d + 4 * d) * 8), "d2h_bytes_per_step": int(n_gpu * d * 8), "steps": n2,
                   "d2h_gbs_per_gpu_inside_call": n_gpu * d * 8 * n2 / (ms2 * 1e-3) / 1e9,
                   "raw_pinned_d2h_gbs_per_gpu_all_ranks_at_once": 2 * n_gpu * d * 8 / (raw_ms * 1e-3) / 1e9,
                   "violations_in_host_copy": viol_host}
        del hout
    clk = clocks.stop()

    # ---- roofline of the dominant kernel (hmc_box_kernel): FP64 CUDA-core pipe ----
    c = P["c"]
    tr = sum(s.transitions for s in stats); bo = sum(s.bounces for s in stats); se = sum(s.hit_searches for s in stats)
    rs = sum(s.vel_restarts + s.chk_restarts for s in stats)
    chain_ms = sum(s.chain_ms for s in stats); pre_ms = sum(s.prepass_ms for s in stats)
    # SURVEY.md par.8(d) convention: 40 flop per constraint per hit search, 8d per bounce, 3d per
    # transition, d^2 per in-kernel momentum (restarts); the first momentum of every transition
    # is the pre-pass GEMM (d^2 each, triangular U halves the executed work)
    w_chain = 40.0 * c * se + 8.0 * d * bo + 3.0 * d * tr + float(d) * d * rs
    w_pre = float(d) * d * tr
    # forked families: the root phase (one chain per family burning in, at most one chain per SM: latency-bound) and
    # the sampling phase (every chain slot busy) of the same kernel, separately
    phases = None
    r_tr = sum(s.root_transitions for s in stats)
    if r_tr > 0:
        r_bo = sum(s.root_bounces for s in stats); r_se = sum(s.root_hit_searches for s in stats)
        r_rs = sum(s.root_restarts for s in stats); r_ms = sum(s.root_chain_ms for s in stats)
        w_root = 40.0 * c * r_se + 8.0 * d * r_bo + 3.0 * d * r_tr + float(d) * d * r_rs
        phases = {"root": {"transitions_per_step": r_tr / len(stats), "chain_ms_per_step": r_ms / len(stats),
                           "frac": w_root / (r_ms * 1e-3) / 1e12 / dfma.value},
                  "sampling": {"transitions_per_step": (tr - r_tr) / len(stats), "chain_ms_per_step": (chain_ms - r_ms) / len(stats),
                               "frac": (w_chain - w_root) / ((chain_ms - r_ms) * 1e-3) / 1e12 / dfma.value},
                  "note": "roofline.frac is over both phases; the root phase runs n_chains / fork chains (one per SM)"}
    traffic = traffic_detail = None
    try:      # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of THIS workload
        allt = json.load(open(ROOT / "profiles" / "traffic.json"))
        tj = allt[a.workload]["hmc_box_kernel"]
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        traffic_detail = {"algorithmic_bytes_per_launch": tj["algorithmic_bytes"], "launch": tj["launch"], "source": tj["source"],
                          "note": "per launch of the captured shape (one sampling round); traffic ~ algorithmic bytes (pool columns in, samples out): HBM is not the bound"}
    except Exception:
        pass
    roof = {"bound": "fp64", "kernel": "hmc_box_kernel", "achieved": w_chain / (chain_ms * 1e-3) / 1e12,
            "peak": dfma.value, "peak_source": "lqg_measure_fp64_peaks DFMA micro-benchmark, this run (MEASURED_PEAKS.json has no FP64 entry)",
            "unit": "TFLOP/s", "frac": w_chain / (chain_ms * 1e-3) / 1e12 / dfma.value, "traffic": traffic,
            "launch_ms": chain_ms / len(stats), "share_of_step": chain_ms / sum(step_ms),
            "traffic_detail": traffic_detail, "phases": phases,
            "note": "bound = FP64 CUDA-core pipe (SURVEY.md par.8d): state in registers, Sigma/U L2-resident; see roofline_hbm for the HBM view"}
    # the same kernel against the HBM roofline: algorithmic bytes = momentum columns read + samples written
    try:
        hbm_peak = json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md"
    att = tr + rs
    bytes_chain = 8.0 * d * (att + sum(s.transitions for s in stats) * (n_per / (n_per + burn)))
    roof_hbm = {"bound": "hbm", "kernel": "hmc_box_kernel", "achieved": bytes_chain / (chain_ms * 1e-3) / 1e9,
                "peak": hbm_peak, "peak_source": hbm_src, "unit": "GB/s", "frac": bytes_chain / (chain_ms * 1e-3) / 1e9 / hbm_peak,
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
            "config": {"workload": WORKLOADS[a.workload], "d": d, "c": c, "m": m,
                       "samples_per_gpu": n_gpu, "chains_per_gpu": n_chains, "burn_in": burn,
                       "burn_in_scheme": (f"families of {a.fork} chains share one root chain's {burn} burn-in transitions from the clamped MAP, "
                                          f"then each chain runs {a.fork_burn_in} more on its own stream before it emits "
                                          f"({burn + a.fork_burn_in} transitions between the start and any emitted sample)") if a.fork > 1
                                         else f"{burn} burn-in transitions per chain",
                       "l2": "256 MiB buffer written between timed iterations (L2 flush)",
                       "parallelism": f"chains sharded over {world} GPU(s); no collective while sampling; e2e: NCCL all-gather of xi shards and band pieces" if world > 1
                                      else "1 GPU"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "steps": e2e_steps, "api": pipeline["api"]},
            "e2e_tmvrnorm": e2e_tmv, "pipeline": pipeline,
            "gpu_launches": launches, "clocks": clk, "roofline": roof, "roofline_hbm": roof_hbm, "roofline_prepass": roof_pre,
            "setup_s": P["setup_s"], "setup_pinv_s": P["pinv_s"], "setup_device": setup_dev, "numa_bound_cpus": numa,
            "wall_s_timed_region": t_wall, "step_ms": [round(x, 3) for x in step_ms],
            "counters": {"transitions": tr, "bounces_per_transition": bo / tr, "hit_searches": se,
                         "restarts_per_transition": rs / tr, "exact_evals_per_search": sum(s.exact_evals for s in stats) / se,
                         "violations": sum(s.violations for s in stats) + sum(p.hmc.violations for p in pst) + (e2e_tmv or {}).get("violations_in_host_copy", 0),
                         "failed_chains": sum(s.failed_chains for s in stats),
                         "bounce_steps_per_s": se / (chain_ms * 1e-3) * world,
                         "transitions_per_s": world * tr / (tot_ms * 1e-3)},
            "fp64_peaks_measured": {"dfma_tflops": dfma.value, "dmma_tflops": dmma.value}}

    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_tr = a.cpu_transitions or 40
        v, kind, wall, info = cpu_reference_sample(P, n_tr, cores)
        v2, kind2, wall2, _ = cpu_reference_sample(P, max(4, n_tr // 2), cores, tuned=True, count_bounces=False)
        line["cpu_baseline_tuned"] = {"value": v2, "unit": UNIT, "cores": cores, "kind": kind2,
                                      "sample": f"same chain without tmg's per-constraint struct copies, {max(4, n_tr // 2)} timed transitions per core"}
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                                "sample": f"{cores} serial chains x {n_tr} timed transitions after 3 untimed burn-in transitions, "
                                          f"same canonical inputs (d={d}, c={c}), wall {wall:.1f} s; burn-in not charged",
                                "counters_chain0": info}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def run_small(a, torch, dist, L, lib, rank, world, local, dev):
    """BASELINE configs 1-3 (d = 100 / 20 / 10) with any of the three samplers.  These are latency /
    launch bound: every rank draws the config's own nsim (replicas, weak scaling).
    value = samples/s of the sampler call with device-resident inputs where the entry point has a device
    mode (HMC); e2e = the public call tmvrnorm(object, nsim, control) with host arrays in and out."""
    import lineqgpr_b200 as lqg
    P = build_problem(a.workload, a.sampler)
    par, aux = P["par"], P["aux"]
    d, c = P["d"], P["c"]
    nsim = a.nsim_per_gpu or int(aux["nsim"])
    obj = dict(par, **{"class": a.sampler})
    ctl = {"seed": 1}
    if a.sampler == "HMC":
        ctl.update({"burn.in": a.burn_in, "U": P["U"]})
    if a.sampler == "ExpT":
        from lineqgpr_b200 import expt
        t0 = time.perf_counter()
        ctl["setup"] = expt.setup(P["lb"] - P["mu"], P["ub"] - P["mu"], P["Sigma"])
        P["setup_s"] = time.perf_counter() - t0
    if a.sampler == "RSM":
        from lineqgpr_b200.samplers import mvrnorm_factor
        ctl["factor"] = mvrnorm_factor(P["Sigma"])
    stream = torch.cuda.current_stream()
    flush_t = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def call(seed):
        x = lqg.tmvrnorm(obj, nsim, dict(ctl, seed=seed))
        return x, dict(lqg.last_stats)

    clocks = ClockSampler(local); clocks.start()
    for w in range(a.warmup):
        call(10 + w)
    barrier()
    lib.lqg_launch_count(1)
    t_win0 = time.perf_counter()
    kern_ms, wall_ms, sts = 0.0, 0.0, []
    for s in range(a.steps):
        flush_t.fill_(s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        x, st = call(100 + s)
        e1.record(stream)
        torch.cuda.synchronize()
        wall_ms += e0.elapsed_time(e1)
        kern_ms += st.get("kernel_ms", 0.0) if a.sampler != "HMC" else st["chain_ms"] + st["prepass_ms"]
        sts.append(st)
    barrier()
    launches = int(lib.lqg_launch_count(0))
    clocks.window(t_win0, time.perf_counter())
    clk = clocks.stop()
    t = torch.tensor([kern_ms, wall_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    kern_ms, wall_ms = float(t[0]), float(t[1])
    value = world * nsim * a.steps / (kern_ms * 1e-3)
    e2e = world * nsim * a.steps / (wall_ms * 1e-3)
    viol = int((x < P["lb"][:, None]).sum() + (x > P["ub"][:, None]).sum())
    dfma, dmma = ctypes.c_double(), ctypes.c_double()
    L.check(lib.lqg_measure_fp64_peaks(ctypes.byref(dfma), ctypes.byref(dmma), None), "peaks")
    if a.sampler == "HMC":
        se = sum(s["hit_searches"] for s in sts); bo = sum(s["bounces"] for s in sts); tr = sum(s["transitions"] for s in sts)
        rs = sum(s["vel_restarts"] + s["chk_restarts"] for s in sts)
        work = 40.0 * c * se + 8.0 * d * bo + 3.0 * d * tr + float(d) * d * (tr + rs)
        counters = {"transitions": tr, "bounces_per_transition": bo / tr, "restarts_per_transition": rs / tr}
        kernel = "hmc_box_kernel"
    elif a.sampler == "RSM":
        pr = sum(s["proposals"] for s in sts)
        work = (2.0 * d * d + 4.0 * d) * pr                       # factor z + box test + dot
        counters = {"proposals": pr, "in_box": sum(s["in_box"] for s in sts), "acceptance": nsim * a.steps / pr}
        kernel = "rsm_eval_kernel + rsm_accept + rsm_emit"
    else:
        pr = sum(s["proposals"] for s in sts)
        work = (float(d) * d + 60.0 * d) * pr                     # triangular sweep d^2/2 mul-add + ~60 flop of 1-D truncated normal per coordinate
        counters = {"proposals": pr, "acceptance": nsim * a.steps / pr}
        kernel = "expt_propose_kernel"
    roof = {"bound": "fp64", "kernel": kernel, "achieved": work / (kern_ms * 1e-3) / 1e12, "peak": dfma.value, "unit": "TFLOP/s",
            "frac": work / (kern_ms * 1e-3) / 1e12 / dfma.value, "traffic": None,
            "note": "latency / launch bound at this size (d <= 100): the fraction is small by construction, see SURVEY.md par.8e"}
    line = {"metric": f"constrained TMVN samples/sec ({a.sampler}, {a.workload})", "value": value, "unit": UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": kern_ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[a.workload], "sampler": a.sampler, "d": d, "c": c, "nsim_per_gpu": nsim,
                       "l2": "256 MiB buffer written between timed iterations (L2 flush)",
                       "parallelism": f"{world} independent replica(s)"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int((2 * d * d + 5 * d) * 8), "d2h_bytes_per_step": int(nsim * d * 8),
                    "api": "lineqgpr_b200.tmvrnorm(object, nsim, control): host arrays in, d x nsim host matrix out"},
            "gpu_launches": launches, "clocks": clk, "roofline": roof, "setup_s": P["setup_s"], "counters": dict(counters, violations=viol),
            "fp64_peaks_measured": {"dfma_tflops": dfma.value, "dmma_tflops": dmma.value}}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        if a.sampler == "HMC":
            cores = os.cpu_count() or 1
            v, kind, wall, info = cpu_reference_sample(P, a.cpu_transitions or 4000, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "counters_chain0": info,
                                    "sample": f"{cores} serial chains x {a.cpu_transitions or 4000} transitions, same canonical inputs"}
        else:
            v, kind, wall, got = cpu_small_sample(P, a)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"{kind}; {got} samples in {wall:.1f} s"}
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
