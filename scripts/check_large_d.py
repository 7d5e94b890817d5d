import sys, numpy as np
sys.path.insert(0, ".")
import lineqgpr_b200 as lqg
rng = np.random.default_rng(1)
for d in [int(a) for a in sys.argv[1:]] or (2500, 5000):
    B = rng.standard_normal((d, 64)); S = B @ B.T / 64 + np.eye(d)
    sd = np.sqrt(np.diag(S))
    par = dict(mu=np.zeros(d), Sigma=S, lb=-2.2 * sd, ub=2.6 * sd)
    par["class"] = "HMC"
    out = {}
    for mode in (0, 2, 1):
        try:
            x = lqg.tmvrnorm(par, 16, {"burn.in": 3, "n.chains": 8, "seed": 5, "burn.in.exact": True, "hmc.search": mode, "max.restarts": 2000})
        except Exception as e:
            print(d, "search mode", mode, "FAILED", str(e)[:120], {k: lqg.last_stats.get(k) for k in ("transitions", "bounces", "vel_restarts", "chk_restarts")}, flush=True)
            out[mode] = None
            continue
        st = dict(lqg.last_stats)
        out[mode] = x
        print(d, "search mode", mode, "violations", st["violations"], "bounces/transition", round(st["bounces"] / st["transitions"], 1),
              "in box", bool(np.all(x >= par["lb"][:, None]) and np.all(x <= par["ub"][:, None])), "chain_ms", round(st["chain_ms"], 2), flush=True)
    # first emitted sample of every chain: the three search modes decide the same events, so they agree closely
    if any(out[m] is None for m in (0, 2, 1)):
        continue
    first = [out[m].reshape(d, 8, 2, order="F")[:, :, 0] for m in (0, 2, 1)]
    print(d, "max |fast+windows - fast|", np.abs(first[0] - first[1]).max(), " max |fast - exact|", np.abs(first[1] - first[2]).max())
