"""lineqgpr_b200 — B200-native constrained-posterior simulation path of lineqGPR.

The product is liblineqgpr_b200.so (CUDA kernels behind the C ABI of include/lineqgpr_b200.h);
this package is the host-side mirror of the reference's R interface for that path:
tmvrnorm / simulate for lineqGP and lineqAGP models."""
from .samplers import tmvrnorm, tmvrnorm_HMC, tmvrnorm_RSM, tmvrnorm_ExpT, rtmg_canonical, rtmg_canonical_trace, last_stats  # noqa: F401
from .simulate import simulate, simulate_lineqGP, simulate_lineqAGP, bands, dgemm, paths, paths_bands  # noqa: F401
from .model import (create_lineqGP, create_lineqAGP, augment_lineqGP, augment_lineqAGP,  # noqa: F401
                    predict_lineqGP, predict_lineqAGP, basisCompute_lineqGP, lineqGPSys,
                    bounds2lineqSys, kernCompute)

__version__ = "0.1.0"
