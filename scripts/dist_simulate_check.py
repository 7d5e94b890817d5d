#!/usr/bin/env python
"""Multi-GPU check of lqg_simulate (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node=W --master-addr 127.0.0.1 \
        --master-port 29731 scripts/dist_simulate_check.py [--record profiles/x.json]

Every rank draws its share of the chains, the xi shards and the band pieces travel device to device
through NCCL all-gathers inside the library.  Rank 0 then runs the SAME job in one process
(W times the chains, chain offset 0) and checks: each rank's xi.sim equals the corresponding chains
of the single-process run bit for bit, the quantile bands are bit-identical, the means agree to
1e-15 (their summation is split differently), every rank returned the same full bands."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--record", default=None)
    ap.add_argument("--m", type=int, default=150)
    ap.add_argument("--chains", type=int, default=12)
    ap.add_argument("--per-chain", type=int, default=40)
    ap.add_argument("--n-test", type=int, default=203)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from lineqgpr_b200 import workloads as W
    from lineqgpr_b200.dist import Comm, simulate_rows
    from lineqgpr_b200.simulate import simulate_pipeline

    par, aux = W.cfg4(m=a.m, n_test=a.n_test)
    Lam, Phi = aux["Lambda"], aux["Phi_test"]
    comm = Comm.from_torch_distributed()
    probs = (0.025, 0.5, 0.975)
    ctl = {"burn.in": 7, "n.chains": a.chains, "seed": 31, "burn.in.exact": True}
    nsim = a.chains * a.per_chain
    t0 = time.perf_counter()
    r = simulate_pipeline(par, nsim, Lambda=Lam, Phi_test=Phi, probs=probs, control=ctl, comm=comm, return_ysim=True)
    wall = time.perf_counter() - t0
    assert r["rows"] == simulate_rows(a.n_test, rank, world)
    assert r["stats"]["world"] == world and r["stats"]["n_total"] == world * nsim

    # gather everything on rank 0 (test plumbing)
    box = [None] * world
    dist.gather_object(dict(xi=r["xi.sim"], mean=r["ymean"], q=r["qtls"], ysim=r["ysim"], rows=r["rows"],
                            rounds=r["stats"]["rounds"]), box if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        one = simulate_pipeline(par, world * nsim, Lambda=Lam, Phi_test=Phi, probs=probs,
                                control=dict(ctl, **{"n.chains": world * a.chains}), return_ysim=True)
        xi_job = np.concatenate([b["xi"] for b in box], axis=1)
        ok &= np.array_equal(xi_job, one["xi.sim"])                      # chains keyed by global id
        ysim_job = np.concatenate([b["ysim"] for b in box], axis=0)      # rank r holds rows r0..r1, all samples, job order
        ok &= np.array_equal(ysim_job, one["ysim"])
        for b in box:
            ok &= np.array_equal(b["q"], one["qtls"])                    # exact order statistics
            ok &= np.array_equal(b["q"], box[0]["q"]) and np.array_equal(b["mean"], box[0]["mean"])
            ok &= np.abs(b["mean"] - one["ymean"]).max() <= 1e-15 * max(1.0, np.abs(one["ymean"]).max())
        rec = {"world": world, "ok": bool(ok), "nsim_per_rank": nsim, "n_test": a.n_test, "d": int(par["mu"].size),
               "rounds_per_rank": [b["rounds"] for b in box], "wall_s_sharded_call": wall,
               "max_abs_mean_diff": float(max(np.abs(b["mean"] - one["ymean"]).max() for b in box))}
        print(json.dumps(rec))
        if a.record:
            Path(a.record).parent.mkdir(parents=True, exist_ok=True)
            Path(a.record).write_text(json.dumps(rec, indent=1) + "\n")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    comm.close()
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        raise SystemExit("sharded job differs from the single-process job")
    if rank == 0:
        print("DIST_SIMULATE_OK")


if __name__ == "__main__":
    main()
