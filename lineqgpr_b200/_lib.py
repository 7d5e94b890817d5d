"""ctypes binding of the C ABI declared in include/lineqgpr_b200.h.

There is no fallback: if liblineqgpr_b200.so is missing or a call reports LQG_ERR_CUDA the
error is raised to the caller."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

import os

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["LQG_LIB_PATH"]) if os.environ.get("LQG_LIB_PATH") else PKG / "liblineqgpr_b200.so"   # override: diagnostic builds

LQG_OK = 0
STATUS = {0: "LQG_OK", 1: "LQG_ERR_INVALID", 2: "LQG_ERR_UNSUPPORTED", 3: "LQG_ERR_INIT_INFEASIBLE",
          4: "LQG_ERR_CUDA", 5: "LQG_ERR_RESTART_CAP", 6: "LQG_ERR_DRAWS_EXHAUSTED", 7: "LQG_ERR_BUDGET",
          8: "LQG_ERR_NOT_PD", 9: "LQG_ERR_INTERNAL"}
MEM_HOST, MEM_DEVICE = 0, 1

# every symbol include/lineqgpr_b200.h declares
SYMBOLS = ["lqg_rtmg", "lqg_hmc_canonical", "lqg_hmc_canonical_trace", "lqg_hmc_box", "lqg_hmc_box_slots", "lqg_rsm", "lqg_expt", "lqg_expt_setup", "lqg_dgemm", "lqg_bands",
           "lqg_alloc_host", "lqg_free_host", "lqg_version", "lqg_last_error", "lqg_device_count", "lqg_measure_fp64_peaks",
           "lqg_launch_count", "lqg_guard_violations",
           "lqg_whiten_box", "lqg_chol", "lqg_solve_spd", "lqg_eta_params", "lqg_lambda_pinv", "lqg_solve_qp", "lqg_basis", "lqg_paths",
           "lqg_simulate", "lqg_comm_unique_id", "lqg_comm_init", "lqg_comm_rank", "lqg_comm_world", "lqg_comm_destroy",
           "lqg_comm_allgather", "lqg_measure_fp64_mixed", "lqg_hit_time_batch"]

_dp = C.POINTER(C.c_double)


class LqgError(RuntimeError):
    def __init__(self, status: int, where: str, text: str):
        self.status = status
        super().__init__(f"{where}: {STATUS.get(status, status)}: {text}")


class Opts(C.Structure):
    _fields_ = [("mem", C.c_int32), ("hmc_search", C.c_int32), ("stream", C.c_void_p),
                ("seed", C.c_uint64), ("n_chains", C.c_int32), ("max_restarts", C.c_int32),
                ("chain_offset", C.c_int64), ("injected", C.c_void_p), ("n_injected", C.c_int64),
                ("injected_u", C.c_void_p), ("n_injected_u", C.c_int64), ("prepass", C.c_int32),
                ("fork", C.c_int32), ("fork_burn_in", C.c_int32), ("canon_mode", C.c_int32)]


class HmcStats(C.Structure):
    _fields_ = [("transitions", C.c_int64), ("bounces", C.c_int64), ("hit_searches", C.c_int64),
                ("vel_restarts", C.c_int64), ("chk_restarts", C.c_int64), ("exact_evals", C.c_int64),
                ("violations", C.c_int64), ("failed_chains", C.c_int64), ("chain_ms", C.c_double),
                ("prepass_ms", C.c_double), ("slow_searches", C.c_int64), ("window_retries", C.c_int64),
                ("root_transitions", C.c_int64), ("root_bounces", C.c_int64), ("root_hit_searches", C.c_int64),
                ("root_restarts", C.c_int64), ("root_chain_ms", C.c_double), ("root_prepass_ms", C.c_double)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SimulateArgs(C.Structure):
    """lqg_simulate_args (include/lineqgpr_b200.h)."""
    _fields_ = [("d", C.c_int32), ("m", C.c_int32),
                ("mu", C.c_void_p), ("Sigma", C.c_void_p), ("lb", C.c_void_p), ("ub", C.c_void_p), ("init", C.c_void_p),
                ("U", C.c_void_p), ("Lambda", C.c_void_p), ("Lambda_pinv", C.c_void_p),
                ("n_t", C.c_int32), ("Phi", C.c_void_p), ("ldphi", C.c_int64),
                ("basis_D", C.c_int32), ("basis_mk", C.c_void_p), ("xtest", C.c_void_p), ("ldx", C.c_int64),
                ("knots", C.c_void_p), ("probs", C.c_void_p), ("nprobs", C.c_int32),
                ("n_per_chain", C.c_int32), ("burn_in", C.c_int32), ("slab_rows", C.c_int32), ("reserved0", C.c_int32),
                ("eta_out", C.c_void_p), ("xi_out", C.c_void_p), ("ysim_out", C.c_void_p), ("ldy", C.c_int64),
                ("mean", C.c_void_p), ("qtls", C.c_void_p)]


class SimulateStats(C.Structure):
    _fields_ = [("hmc", HmcStats), ("xi_ms", C.c_double), ("gemm_ms", C.c_double), ("bands_ms", C.c_double),
                ("total_ms", C.c_double), ("n_local", C.c_int64), ("n_total", C.c_int64),
                ("row_begin", C.c_int32), ("row_end", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
                ("rounds", C.c_int32), ("reserved0", C.c_int32)]

    def asdict(self):
        d = self.hmc.asdict()
        d.update({k: getattr(self, k) for k, _ in self._fields_ if k not in ("hmc", "reserved0")})
        return d


class ExptStats(C.Structure):
    _fields_ = [("proposals", C.c_int64), ("accepted", C.c_int64), ("rounds", C.c_int64), ("kernel_ms", C.c_double)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class RsmStats(C.Structure):
    _fields_ = [("proposals", C.c_int64), ("in_box", C.c_int64), ("accepted", C.c_int64),
                ("rounds", C.c_int64), ("kernel_ms", C.c_double)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib():
    """Loads the shared library (building nothing: see lineqgpr_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m lineqgpr_b200.build` "
            "(nvcc, sm_100a). lineqgpr_b200 has no CPU fallback.")
    L = C.CDLL(str(LIB_PATH))
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.lqg_rtmg.restype = C.c_int
    L.lqg_rtmg.argtypes = [i32, i32, vp, i32, i32, vp, vp, i32, vp, vp]
    L.lqg_hmc_canonical.restype = C.c_int
    L.lqg_hmc_canonical.argtypes = [i32, i32, vp, vp, vp, i64, i32, i32, C.POINTER(Opts), vp, vp,
                                    C.POINTER(HmcStats)]
    L.lqg_hmc_canonical_trace.restype = C.c_int
    L.lqg_hmc_canonical_trace.argtypes = [i32, i32, vp, vp, vp, i64, i32, C.POINTER(Opts), vp, i32, vp, vp, vp]
    L.lqg_hmc_box.restype = C.c_int
    L.lqg_hmc_box.argtypes = [i32, vp, vp, i32, vp, vp, vp, vp, i64, i32, i32, C.POINTER(Opts), vp, vp,
                              C.POINTER(HmcStats)]
    L.lqg_hmc_box_slots.restype = C.c_int
    L.lqg_hmc_box_slots.argtypes = [i32]
    L.lqg_rsm.restype = C.c_int
    L.lqg_rsm.argtypes = [i32, vp, vp, vp, vp, vp, i64, i64, C.POINTER(Opts), vp, C.POINTER(RsmStats)]
    L.lqg_expt.restype = C.c_int
    L.lqg_expt.argtypes = [i32, vp, vp, vp, vp, C.c_double, i64, i64, C.POINTER(Opts), vp, C.POINTER(ExptStats)]
    L.lqg_expt_setup.restype = C.c_int
    L.lqg_expt_setup.argtypes = [i32, vp, vp, vp, C.POINTER(Opts), vp, vp, vp, vp, vp, _dp, vp, vp, C.POINTER(C.c_int32)]
    L.lqg_dgemm.restype = C.c_int
    L.lqg_dgemm.argtypes = [i32, i32, i32, vp, i64, vp, i64, vp, i64, vp, C.POINTER(Opts), _dp]
    L.lqg_bands.restype = C.c_int
    L.lqg_bands.argtypes = [i32, i64, vp, i64, vp, i32, vp, vp, C.POINTER(Opts), _dp]
    L.lqg_alloc_host.restype = C.c_void_p
    L.lqg_alloc_host.argtypes = [C.c_uint64]
    L.lqg_free_host.restype = None
    L.lqg_free_host.argtypes = [C.c_void_p]
    L.lqg_version.restype = C.c_char_p
    L.lqg_last_error.restype = C.c_char_p
    L.lqg_device_count.restype = C.c_int
    L.lqg_measure_fp64_peaks.restype = C.c_int
    L.lqg_measure_fp64_peaks.argtypes = [_dp, _dp, C.POINTER(Opts)]
    L.lqg_measure_fp64_mixed.restype = C.c_int
    L.lqg_measure_fp64_mixed.argtypes = [_dp, _dp, C.POINTER(Opts)]
    L.lqg_hit_time_batch.restype = C.c_int
    L.lqg_hit_time_batch.argtypes = [i64, vp, vp, vp, C.c_double, vp, vp, vp, vp, vp]
    L.lqg_launch_count.restype = i64
    L.lqg_launch_count.argtypes = [i32]
    L.lqg_guard_violations.restype = C.c_int
    L.lqg_guard_violations.argtypes = []
    L.lqg_whiten_box.restype = C.c_int
    L.lqg_whiten_box.argtypes = [i32, vp, C.POINTER(Opts), vp, _dp]
    L.lqg_chol.restype = C.c_int
    L.lqg_chol.argtypes = [i32, vp, C.POINTER(Opts), vp, C.POINTER(i32)]
    L.lqg_solve_spd.restype = C.c_int
    L.lqg_solve_spd.argtypes = [i32, i32, vp, vp, C.POINTER(Opts), vp]
    L.lqg_eta_params.restype = C.c_int
    L.lqg_eta_params.argtypes = [i32, i32, vp, vp, vp, vp, C.c_double, C.POINTER(Opts), vp, vp, vp]
    L.lqg_lambda_pinv.restype = C.c_int
    L.lqg_lambda_pinv.argtypes = [i32, i32, vp, C.POINTER(Opts), vp]
    L.lqg_paths.restype = C.c_int
    L.lqg_paths.argtypes = [i32, i32, i64, vp, i64, vp, i64, vp, i32, vp, i64, vp, vp, i32, C.POINTER(Opts), _dp, _dp]
    L.lqg_basis.restype = C.c_int
    L.lqg_basis.argtypes = [i32, i32, vp, vp, i64, vp, C.POINTER(Opts), vp, i64]
    L.lqg_solve_qp.restype = C.c_int
    L.lqg_solve_qp.argtypes = [i32, i32, vp, vp, vp, vp, C.c_double, i32, C.POINTER(Opts), vp, C.POINTER(i32)]
    L.lqg_simulate.restype = C.c_int
    L.lqg_simulate.argtypes = [C.POINTER(SimulateArgs), C.POINTER(Opts), vp, C.POINTER(SimulateStats)]
    L.lqg_comm_unique_id.restype = C.c_int
    L.lqg_comm_unique_id.argtypes = [vp]
    L.lqg_comm_init.restype = C.c_int
    L.lqg_comm_init.argtypes = [vp, i32, i32, C.POINTER(vp)]
    L.lqg_comm_rank.restype = C.c_int
    L.lqg_comm_rank.argtypes = [vp]
    L.lqg_comm_world.restype = C.c_int
    L.lqg_comm_world.argtypes = [vp]
    L.lqg_comm_destroy.restype = C.c_int
    L.lqg_comm_destroy.argtypes = [vp]
    L.lqg_comm_allgather.restype = C.c_int
    L.lqg_comm_allgather.argtypes = [vp, vp, vp, i64, vp]
    _lib = L
    return L


def check(rc: int, where: str):
    if rc != LQG_OK:
        raise LqgError(rc, where, lib().lqg_last_error().decode())


def last_error() -> str:
    return lib().lqg_last_error().decode()


def f64(a, order="F"):
    """FP64 array in R (column-major) layout."""
    return np.require(np.asarray(a, dtype=np.float64), dtype=np.float64, requirements=[order.upper() + "_CONTIGUOUS", "ALIGNED"]) \
        if np.ndim(a) > 1 else np.ascontiguousarray(a, dtype=np.float64)


def ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a)          # raw device pointer (e.g. torch tensor.data_ptr())


def make_opts(mem=MEM_HOST, stream=None, seed=0, n_chains=1, max_restarts=0, chain_offset=0,
              injected=None, n_injected=0, injected_u=None, n_injected_u=0, prepass=-1, canon_mode=0, hmc_search=0,
              fork=0, fork_burn_in=0) -> Opts:
    o = Opts()
    o.mem = mem
    o.hmc_search = int(hmc_search)
    o.stream = stream
    o.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    o.n_chains = int(n_chains)
    o.max_restarts = int(max_restarts)
    o.chain_offset = int(chain_offset)
    o.injected = ptr(injected)
    o.n_injected = int(n_injected)
    o.injected_u = ptr(injected_u)
    o.n_injected_u = int(n_injected_u)
    o.prepass = int(prepass)
    o.canon_mode = int(canon_mode)
    o.fork = int(fork)
    o.fork_burn_in = int(fork_burn_in)
    return o


def empty_result(rows: int, cols: int) -> np.ndarray:
    """Column-major FP64 result matrix.  Large results are backed by page-locked memory
    (lqg_alloc_host) so that the library's D2H copies overlap the sampling and run at PCIe speed;
    the memory is released when the array is garbage collected."""
    import weakref
    nbytes = rows * cols * 8
    if nbytes < (8 << 20):
        return np.empty((rows, cols), order="F")
    p = lib().lqg_alloc_host(nbytes)
    if not p:
        return np.empty((rows, cols), order="F")
    buf = (C.c_double * (rows * cols)).from_address(p)
    arr = np.frombuffer(buf, dtype=np.float64).reshape((rows, cols), order="F")
    weakref.finalize(buf, lib().lqg_free_host, p)
    return arr
