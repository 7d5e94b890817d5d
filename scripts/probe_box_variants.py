"""Chain-kernel shapes side by side: throughput (hit searches per ms of chain kernel) and bitwise equality of the
samples with the default shape.  One subprocess per LQG_BOX_VARIANT (the library reads it once).
usage: python scripts/probe_box_variants.py [cfg4|cfg5] [n_per] [variant:chains ...]   e.g.  cfg4 64 -:296 a:740 b:740"""
import hashlib, json, os, subprocess, sys
sys.path.insert(0, ".")
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import numpy as np
    import lineqgpr_b200 as lqg
    from lineqgpr_b200 import workloads as W
    from lineqgpr_b200.samplers import whiten_box
    cfg, chains, n_per = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    par = W.cfg4()[0] if cfg == "cfg4" else W.cfg5(n_test=10)[0]
    U = whiten_box(np.asarray(par["mu"]), np.asarray(par["Sigma"]))
    best = None
    for rep in range(2):
        x = lqg.tmvrnorm(dict(par, **{"class": "HMC"}), chains * n_per, {"burn.in": 100, "n.chains": chains, "seed": 1, "U": U, "prepass": 1, "fork": int(os.environ.get("LQG_FORK", "1"))})
        st = dict(lqg.last_stats)
        if best is None or st["chain_ms"] < best["chain_ms"]:
            best = st
    # a fixed set of chains for the bitwise comparison, independent of how many chains ran
    xs = np.ascontiguousarray(x.reshape(x.shape[0], chains, n_per, order="F")[:, :64, :].copy())
    print(json.dumps({"variant": os.environ.get("LQG_BOX_VARIANT", "-"), "chains": chains, "chain_ms": best["chain_ms"], "prepass_ms": best["prepass_ms"],
                      "searches_per_ms": best["hit_searches"] / best["chain_ms"], "transitions_per_ms": best["transitions"] / best["chain_ms"],
                      "violations": best["violations"], "sha_first64": hashlib.sha1(xs.tobytes()).hexdigest()[:12]}))
    sys.exit(0)
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
n_per = int(sys.argv[2]) if len(sys.argv) > 2 else 64
for spec in (sys.argv[3:] or ["-:296"]):
    v, chains = spec.split(":")
    env = dict(os.environ)
    if v != "-":
        env["LQG_BOX_VARIANT"] = v
    else:
        env.pop("LQG_BOX_VARIANT", None)
    r = subprocess.run([sys.executable, __file__, "--child", cfg, chains, str(n_per)], env=env, capture_output=True, text=True, timeout=300)
    print(r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else f"{spec}: FAILED rc={r.returncode} {r.stderr[-400:]}")
