"""ctypes binding of libregnde.so (include/regnde.h) + the in-tree nvcc build.

The product path has NO fallback: if the shared library is missing, or no CUDA
device is visible, every entry point raises.  Nothing in this package imports
``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_ROOT = _PKG.parent.parent
_CSRC = _PKG / "csrc"
_INCLUDE = _ROOT / "include"
LIB_PATH = _PKG / "libregnde.so"

NVCC_FLAGS = [
    "-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false",            # canonical arithmetic: no implicit contraction (include/regnde_canon.h)
]

# status codes / enums (mirror include/regnde.h)
OK, ERR_MAXITERS, ERR_DTMIN, ERR_NAN, ERR_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_TAPE_FULL, ERR_STATE = range(9)
ACT_IDENTITY, ACT_TANH = 0, 1
ALG_TSIT5, ALG_AUTO_TSIT5 = 0, 1
REG_NONE, REG_ERR_DT, REG_STIFF_DT_ABS, REG_STIFF_SCALED, REG_ERR_PLUS_STIFF = range(5)
KERNEL_AUTO, KERNEL_CTA, KERNEL_STREAM, KERNEL_CLUSTER, KERNEL_CLUSTER4, KERNEL_CHAIN, KERNEL_CHAIN8 = range(7)
DIST_SINGLE, DIST_EXACT, DIST_INDEPENDENT = range(3)
ARITH_FMA_CHAIN, ARITH_FIXED24, ARITH_SPLITK = 0, 1, 2
DETACH_ALL, DETACH_ALL_BUT_FIRST, DETACH_FIRST_TERM_ONLY = 0, 1, 2
SDE_SOSRI, SDE_AUTO_SOSRI2 = 0, 1

EXPORTS = [
    "rnde_version", "rnde_status_string", "rnde_device_count", "rnde_create", "rnde_destroy", "rnde_last_error",
    "rnde_num_params", "rnde_default_kblock", "rnde_kernel_variant", "rnde_launch_count", "rnde_set_tspan", "rnde_set_forced_steps", "rnde_set_detach", "rnde_set_reverse_time", "rnde_adam_update",
    "rnde_forward", "rnde_backward", "rnde_forward_host", "rnde_backward_host", "rnde_head_loss_grad", "rnde_get_steps",
    "rnde_opt_update", "rnde_test_tanh", "rnde_test_tanh_bits", "rnde_test_pow", "rnde_test_unary_bits", "rnde_test_csq_rhs", "rnde_debug_timeline", "rnde_debug_a6", "rnde_dist_export", "rnde_dist_import",
    "rnde_set_saveat", "rnde_forward_saveat", "rnde_backward_saveat", "rnde_set_noise",
    "rnde_gru_num_params", "rnde_gru_create", "rnde_gru_destroy", "rnde_gru_last_error", "rnde_gru_forward", "rnde_gru_backward",
    "rnde_gru_launch_count", "rnde_reg_agg", "rnde_last_stats", "rnde_allreduce_grads",
    "rnde_sde_num_params", "rnde_sde_create", "rnde_sde_destroy", "rnde_sde_last_error", "rnde_sde_forward", "rnde_sde_enable_tape", "rnde_sde_backward", "rnde_sde_get_log", "rnde_sde_launch_count",
]


class Config(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_int32), ("state_dim", C.c_int32), ("hidden_dim", C.c_int32), ("batch", C.c_int32),
        ("act_hidden", C.c_int32), ("act_out", C.c_int32), ("time_dep", C.c_int32), ("kblock", C.c_int32),
        ("alg", C.c_int32), ("reg_kind", C.c_int32), ("max_steps", C.c_int32), ("tape_capacity", C.c_int32),
        ("need_backward", C.c_int32), ("kernel_variant", C.c_int32), ("dist_mode", C.c_int32),
        ("rank", C.c_int32), ("nranks", C.c_int32),
        ("t0", C.c_float), ("t1", C.c_float), ("abstol", C.c_float), ("reltol", C.c_float), ("dtmin", C.c_float),
        ("max_saveat", C.c_int32), ("n_layers", C.c_int32),
        ("global_batch", C.c_int64),
        ("pre_act", C.c_int32), ("layer_width", C.c_int32 * 8), ("layer_act", C.c_int32 * 8), ("arith", C.c_int32),
        ("csq_extra", C.c_int32), ("reserved0", C.c_int32),
    ]


class GruConfig(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_int32), ("in_dim", C.c_int32), ("hidden_dim", C.c_int32), ("latent_dim", C.c_int32),
        ("batch", C.c_int32), ("seq_len", C.c_int32), ("need_backward", C.c_int32), ("reserved", C.c_int32),
    ]


class SdeConfig(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_int32), ("state_dim", C.c_int32), ("hidden_dim", C.c_int32), ("batch", C.c_int32), ("alg", C.c_int32),
        ("reg_kind", C.c_int32), ("max_steps", C.c_int32), ("max_saved", C.c_int32),
        ("t0", C.c_float), ("t1", C.c_float), ("abstol", C.c_float), ("reltol", C.c_float),
    ]


class SdeStats(C.Structure):
    _fields_ = [
        ("nfe1", C.c_int32), ("nfe2", C.c_int32), ("naccept", C.c_int32), ("nreject", C.c_int32), ("n_saved", C.c_int32), ("draws", C.c_int32),
        ("retcode", C.c_int32), ("reserved", C.c_int32),
        ("t_final", C.c_float), ("dt_init", C.c_float), ("dt_last", C.c_float), ("reserved2", C.c_float),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("nf", C.c_int32), ("naccept", C.c_int32), ("nreject", C.c_int32), ("n_saved", C.c_int32), ("retcode", C.c_int32),
        ("t_final", C.c_float), ("dt_last", C.c_float), ("dt_init", C.c_float),
    ]


def sources() -> list[Path]:
    return [_CSRC / "regnde.cu"]


def _stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = list(_CSRC.glob("*.cu")) + list(_CSRC.glob("*.cuh")) + list(_INCLUDE.glob("*.h"))
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/ for sm_100a with nvcc into regneuralde/jl_b200/libregnde.so (in-tree)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        if LIB_PATH.exists():
            return LIB_PATH          # GPU box without nvcc on PATH: use the prebuilt library that travelled
        raise RuntimeError("nvcc not found and no prebuilt libregnde.so")
    cmd = [nvcc, *NVCC_FLAGS, f"-I{_INCLUDE}", "-o", str(LIB_PATH)] + [str(s) for s in sources()]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load libregnde.so (building it first if nvcc is available and it is stale)."""
    global _lib
    if _lib is not None:
        return _lib
    build()          # no-op when fresh; rebuilds after any edit of csrc/ or include/ (falls back to the prebuilt library without nvcc)
    L = C.CDLL(str(LIB_PATH))
    vp, fp, ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)
    L.rnde_version.restype = C.c_int
    L.rnde_status_string.restype = C.c_char_p
    L.rnde_status_string.argtypes = [C.c_int]
    L.rnde_device_count.restype = C.c_int
    L.rnde_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.rnde_destroy.argtypes = [vp]
    L.rnde_destroy.restype = None
    L.rnde_last_error.argtypes = [vp]
    L.rnde_last_error.restype = C.c_char_p
    L.rnde_num_params.argtypes = [C.POINTER(Config)]
    L.rnde_num_params.restype = C.c_int64
    L.rnde_default_kblock.argtypes = [C.POINTER(Config)]
    L.rnde_kernel_variant.argtypes = [vp]
    L.rnde_launch_count.argtypes = [vp]
    L.rnde_launch_count.restype = C.c_int64
    L.rnde_set_tspan.argtypes = [vp, C.c_float, C.c_float]
    L.rnde_set_forced_steps.argtypes = [vp, fp, C.c_int32]
    L.rnde_set_detach.argtypes = [vp, C.c_int32]
    L.rnde_set_reverse_time.argtypes = [vp, C.c_int32]
    L.rnde_adam_update.argtypes = [vp, vp, vp, vp, vp, C.c_int64] + [C.c_float] * 7 + [vp]
    L.rnde_forward.argtypes = [vp, vp, vp, vp, vp, C.POINTER(Stats), vp]
    L.rnde_backward.argtypes = [vp, vp, vp, vp, vp, vp]
    L.rnde_forward_host.argtypes = [vp, vp, vp, vp, vp, C.POINTER(Stats)]
    L.rnde_backward_host.argtypes = [vp, vp, vp, vp, vp]
    L.rnde_head_loss_grad.argtypes = [vp, vp, vp, vp, C.c_int32, C.c_float, vp, vp, vp, vp, vp]
    L.rnde_get_steps.argtypes = [vp, vp, vp, vp, vp, C.c_int32]
    L.rnde_opt_update.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_float, C.c_float, C.c_float, vp]
    L.rnde_test_tanh.argtypes = [vp, vp, C.c_int64, vp]
    L.rnde_test_tanh_bits.argtypes = [C.c_uint32, C.c_int64, vp, vp]
    L.rnde_test_pow.argtypes = [vp, C.c_float, vp, vp, C.c_int64, vp]
    L.rnde_test_unary_bits.argtypes = [C.c_int32, C.c_uint32, C.c_int64, vp, vp]
    L.rnde_test_csq_rhs.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, C.c_float, vp, vp]
    L.rnde_debug_timeline.argtypes = [vp, vp, C.c_int]
    L.rnde_debug_a6.argtypes = [vp, fp]
    L.rnde_dist_export.argtypes = [vp, vp]
    L.rnde_dist_import.argtypes = [vp, vp, C.c_int32]
    L.rnde_set_saveat.argtypes = [vp, vp, C.c_int32]
    L.rnde_set_noise.argtypes = [vp, vp]
    L.rnde_last_stats.argtypes = [vp, C.POINTER(Stats)]
    L.rnde_allreduce_grads.argtypes = [vp, vp, C.c_int64, vp]
    L.rnde_reg_agg.argtypes = [vp, C.c_int32, C.c_float, C.c_float, vp, vp, vp, vp]
    L.rnde_gru_num_params.restype = C.c_int64
    L.rnde_gru_num_params.argtypes = [C.POINTER(GruConfig)]
    L.rnde_gru_create.argtypes = [C.POINTER(GruConfig), C.POINTER(vp)]
    L.rnde_gru_destroy.argtypes = [vp]
    L.rnde_gru_destroy.restype = None
    L.rnde_gru_last_error.restype = C.c_char_p
    L.rnde_gru_last_error.argtypes = [vp]
    L.rnde_gru_forward.argtypes = [vp, vp, vp, vp, vp]
    L.rnde_gru_backward.argtypes = [vp, vp, vp, vp]
    L.rnde_gru_launch_count.restype = C.c_int64
    L.rnde_gru_launch_count.argtypes = [vp]
    L.rnde_sde_num_params.argtypes = [C.POINTER(SdeConfig)]
    L.rnde_sde_num_params.restype = C.c_int64
    L.rnde_sde_create.argtypes = [C.POINTER(SdeConfig), C.POINTER(vp)]
    L.rnde_sde_destroy.argtypes = [vp]
    L.rnde_sde_destroy.restype = None
    L.rnde_sde_last_error.argtypes = [vp]
    L.rnde_sde_last_error.restype = C.c_char_p
    L.rnde_sde_forward.argtypes = [vp, vp, vp, vp, C.c_int32, vp, vp, C.POINTER(SdeStats), vp]
    L.rnde_sde_enable_tape.argtypes = [vp, C.c_int32]
    L.rnde_sde_backward.argtypes = [vp, vp, vp, vp, vp, vp]
    L.rnde_sde_get_log.argtypes = [vp, fp, C.c_int32]
    L.rnde_sde_launch_count.argtypes = [vp]
    L.rnde_sde_launch_count.restype = C.c_int64
    L.rnde_forward_saveat.argtypes = [vp, vp, vp, vp, vp, vp, C.POINTER(Stats), vp]
    L.rnde_backward_saveat.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    _lib = L
    return L


class RndeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"regnde error {code} ({lib().rnde_status_string(code).decode()}): {msg}")
        self.code = code


def require_device() -> None:
    if lib().rnde_device_count() <= 0:
        raise RuntimeError("regneuralde.jl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
