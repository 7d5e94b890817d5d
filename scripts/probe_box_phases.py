"""Where does a hit search of the box kernel spend its cycles?  Diagnostic build (-DLQG_BOX_TIMING):
thread 0 of every chain reads the SM clock at the phase boundaries of the search loop.
usage: python scripts/probe_box_phases.py [cfg4|cfg5] [chains] [n_per]   (builds the variant library itself)"""
import ctypes, json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ["LQG_LIB_PATH"] = str(ROOT / "lineqgpr_b200" / "liblineqgpr_b200_timing.so")   # before the package is imported
from lineqgpr_b200 import build as B
assert str(B.build_variant("timing", ["-DLQG_BOX_TIMING"])) == os.environ["LQG_LIB_PATH"]
import numpy as np
import lineqgpr_b200 as lqg
from lineqgpr_b200 import _lib as L, workloads as W
from lineqgpr_b200.samplers import whiten_box
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 296
n_per = int(sys.argv[3]) if len(sys.argv) > 3 else 64
par = W.cfg4()[0] if cfg == "cfg4" else W.cfg5(n_test=10)[0]
U = whiten_box(np.asarray(par["mu"]), np.asarray(par["Sigma"]))
names = ["since last update (loop head)", "filter masks", "compaction", "barrier after pushes", "candidate evaluation",
         "warp arg-min + slot", "barrier of the slot exchange", "slot reduce + decision", "Sigma column (L2)", "rotate + reflect",
         "end of transition", "fresh momentum in place"]
buf = (ctypes.c_ulonglong * 16)()
for rep in range(2):
    L.lib().lqg_debug_box_phases(buf)
    x = lqg.tmvrnorm(dict(par, **{"class": "HMC"}), chains * n_per, {"burn.in": 20, "n.chains": chains, "seed": 1, "U": U, "burn.in.exact": True})
    st = dict(lqg.last_stats)
    L.lib().lqg_debug_box_phases(buf)
tot = sum(buf[i] for i in range(12))
se = st["hit_searches"]
out = {"cfg": cfg, "chains": chains, "searches": se, "chain_ms": st["chain_ms"], "cycles_per_search": tot / se,
       "phases_cycles_per_search": {names[i]: buf[i] / se for i in range(12)},
       "phases_share": {names[i]: buf[i] / tot for i in range(12)}}
print(json.dumps(out, indent=1))
