"""Builds liblineqgpr_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "liblineqgpr_b200.so"
SOURCES = ["lqg_capi.cu", "lqg_hmc_box.cu", "lqg_hmc_canonical.cu", "lqg_dgemm.cu", "lqg_rsm.cu", "lqg_expt.cu",
           "lqg_bands.cu", "lqg_peaks.cu", "lqg_linalg.cu", "lqg_qp.cu", "lqg_basis.cu", "lqg_hostcopy.cu", "lqg_hmc_canonical_batched.cu", "lqg_comm.cu", "lqg_sparse.cu", "lqg_expt_setup.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: cannot build liblineqgpr_b200.so")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + \
        [PKG.parent / "include" / "lineqgpr_b200.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build_variant(name: str, extra_flags, verbose: bool = False) -> Path:
    """A diagnostic build of the same sources with extra nvcc flags (e.g. -DLQG_BOX_TIMING) into
    lineqgpr_b200/liblineqgpr_b200_<name>.so; load it with LQG_LIB_PATH=<that file>."""
    lib = PKG / f"liblineqgpr_b200_{name}.so"
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "lineqgpr_b200.h"]
    if lib.exists() and all(p.stat().st_mtime <= lib.stat().st_mtime for p in deps):
        return lib
    nvcc = _nvcc()
    objdir = PKG.parent / "build" / name
    objdir.mkdir(parents=True, exist_ok=True)
    objs, procs = [], []
    for src in SOURCES:
        obj = objdir / (Path(src).stem + ".o")
        procs.append((src, subprocess.Popen([nvcc, *NVCC_FLAGS, *extra_flags, "-c", str(CSRC / src), "-o", str(obj)],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    r = subprocess.run([nvcc, "-shared", "-o", str(lib), *objs, "-lcudart", "-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    import fcntl
    lock = open(PKG / ".build.lock", "w")
    fcntl.flock(lock, fcntl.LOCK_EX)          # several ranks may get here at once: one compiles
    try:
        if not force and not needs_build():
            return LIB
        return _build_locked(verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(verbose: bool) -> Path:
    nvcc = _nvcc()
    objdir = PKG.parent / "build"
    objdir.mkdir(exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = objdir / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out.strip():
            print(out)
    tmp = LIB.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-o", str(tmp), *objs, "-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)                      # atomic: a concurrent loader never sees a partial file
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
