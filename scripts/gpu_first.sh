set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -c "
from lineqgpr_b200 import _lib as L
import ctypes
a=ctypes.c_double(); b=ctypes.c_double()
print('peaks rc', L.lib().lqg_measure_fp64_peaks(ctypes.byref(a), ctypes.byref(b), None), 'dfma TF', a.value, 'dmma TF', b.value)
" 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -150
