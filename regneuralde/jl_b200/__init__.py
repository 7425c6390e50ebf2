"""regneuralde.jl_b200 -- B200-native hot path of avik-pal/RegNeuralDE.jl.

Batched adaptive Tsit5 integration of a small neural-network vector field with the
solver-heuristic regularisers and the discrete-adjoint backward pass, in hand-written
sm_100a CUDA kernels behind a C ABI (include/regnde.h, libregnde.so).  This package is
the host-side mirror of the reference's Julia surface for that path.  No CPU fallback.
"""
from . import _lib
from ._lib import RndeError, build, lib
from .node import (AutoTsit5, Dense, ERROR_ESTIMATE, ERROR_PLUS_STIFFNESS, MLPDynamics, STIFFNESS_ESTIMATE, STIFFNESS_SCALED,
                   SavedValues, SaveFunc, TDChain, Chain, TrackedNeuralODE, Tsit5, colmajor, from_colmajor, track, untrack,
                   solution, ODESolution, DEStats)
from .classifier import ADAMOptimiser, ClassifierNODE, Optimiser, update_parameters_
from .ffjord import CSQDynamics, ConcatSquashLinear, TrackedFFJORD
from . import ffjord
from .nsde import AutoSOSRI2, ClassifierNSDE, SOSRI, TrackedNeuralDSDE
from .latent import LatentGRU, LatentTimeSeriesModel, kl_divergence, latent_ode_model, log_likelihood, loss_function

__all__ = [
    "RndeError", "build", "lib", "AutoTsit5", "Tsit5", "Dense", "TDChain", "Chain", "MLPDynamics", "TrackedNeuralODE", "SavedValues", "SaveFunc",
    "ERROR_ESTIMATE", "STIFFNESS_ESTIMATE", "STIFFNESS_SCALED", "ERROR_PLUS_STIFFNESS", "ClassifierNODE", "Optimiser", "ADAMOptimiser", "ffjord",
    "update_parameters_", "track", "untrack", "colmajor", "from_colmajor", "solution", "ODESolution", "DEStats",
    "TrackedNeuralDSDE", "ClassifierNSDE", "SOSRI", "AutoSOSRI2", "TrackedFFJORD", "CSQDynamics", "ConcatSquashLinear", "LatentGRU", "LatentTimeSeriesModel", "latent_ode_model", "log_likelihood", "kl_divergence", "loss_function",
]
