"""DFMA and DMMA side by side: do the two FP64 paths of a B200 SM add up?"""
import ctypes, json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from lineqgpr_b200 import _lib as L
lib = L.lib()
f, m, t, sh = (ctypes.c_double() for _ in range(4))
L.check(lib.lqg_measure_fp64_peaks(ctypes.byref(f), ctypes.byref(m), None), "peaks")
L.check(lib.lqg_measure_fp64_mixed(ctypes.byref(t), ctypes.byref(sh), None), "mixed")
print(json.dumps({"dfma_tflops": f.value, "dmma_tflops": m.value, "mixed_tflops": t.value, "mixed_dfma_share": sh.value}))
