"""ctypes binding of the C oracle (oracle/rnde_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of rnde_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (regneuralde.jl_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_BUILD = _HERE / "_build"

ACT_ID, ACT_TANH = 0, 1
ALG_TSIT5, ALG_AUTO_TSIT5 = 0, 1
REG_NONE, REG_ERR_DT, REG_STIFF_DT_ABS, REG_STIFF_SCALED, REG_ERR_PLUS_STIFF = range(5)


class _Cfg(C.Structure):
    _fields_ = [
        ("D", C.c_int), ("H", C.c_int), ("B", C.c_int),
        ("act1", C.c_int), ("act2", C.c_int), ("time_dep", C.c_int),
        ("kblock1", C.c_int), ("kblock2", C.c_int),
        ("alg", C.c_int), ("reg_kind", C.c_int), ("max_steps", C.c_int), ("nthreads", C.c_int),
        ("t0", C.c_double), ("t1", C.c_double), ("abstol", C.c_double), ("reltol", C.c_double),
        ("dtmin", C.c_double),
        ("n_forced", C.c_int), ("forced_dt", C.POINTER(C.c_double)), ("forced_accept", C.POINTER(C.c_int)),
        ("n_saveat", C.c_int), ("saveat", C.POINTER(C.c_double)),
        ("n_layers", C.c_int), ("width", C.c_int * 8), ("act", C.c_int * 8), ("pre_act", C.c_int), ("arith", C.c_int),
        ("csq_extra", C.c_int), ("csq_noise", C.c_void_p),
    ]


class _Stats(C.Structure):
    _fields_ = [
        ("nf", C.c_int), ("naccept", C.c_int), ("nreject", C.c_int), ("n_saved", C.c_int), ("retcode", C.c_int),
        ("t_final", C.c_double), ("dt_last", C.c_double), ("dt_init", C.c_double),
    ]


def build(force: bool = False) -> None:
    """Compile the oracle with its Makefile (gcc; a few seconds)."""
    if force or not ((_BUILD / "liborc_f32.so").exists() and (_BUILD / "liborc_f64.so").exists()) or \
            (_BUILD / "liborc_f32.so").stat().st_mtime < max((_HERE / "rnde_oracle.c").stat().st_mtime, (_HERE / "rnde_oracle_bwd.inc").stat().st_mtime,
                                                             (_HERE.parent / "include" / "regnde_canon.h").stat().st_mtime):
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)


_libs: dict = {}


def _lib(f64: bool):
    key = "f64" if f64 else "f32"
    if key not in _libs:
        path = _BUILD / f"liborc_{key}.so"
        if not path.exists():
            build()
        _libs[key] = C.CDLL(str(path))
    return _libs[key]


@dataclass
class OracleConfig:
    D: int
    H: int
    B: int
    act1: int = ACT_TANH
    act2: int = ACT_TANH
    time_dep: bool = True
    kblock1: int = 0
    kblock2: int = 0
    alg: int = ALG_TSIT5
    reg_kind: int = REG_NONE
    max_steps: int = 0
    nthreads: int = 0
    t0: float = 0.0
    t1: float = 1.0
    abstol: float = float(np.float32(1.4e-8))
    reltol: float = float(np.float32(1.4e-8))
    dtmin: float = 0.0
    forced_dt: np.ndarray | None = None
    forced_accept: np.ndarray | None = None
    saveat: np.ndarray | None = None
    # chain field: widths of the Dense layers (last == D), their activations, tanh pre-activation flag
    widths: tuple | None = None
    acts: tuple | None = None
    pre_act: int = 0
    # FFJORD field: csq_extra = 1 or 3 augmented rows (D counts them), csq_noise = the Hutchinson noise (D - extra, B)
    csq_extra: int = 0
    csq_noise: np.ndarray | None = None
    arith: int = 0          # 1 = FIXED24 exact fixed-point layer products (tensor-core forward stepper); 2 = SPLITK fma-chain order of csrc/fwd4s_kernel.cuh

    @property
    def n_params(self) -> int:
        if self.csq_extra:
            Dz, H = self.D - self.csq_extra, self.H
            return (H * Dz + 4 * H) + (H * H + 4 * H) + (Dz * H + 4 * Dz)
        if self.widths is not None:
            n, K = 0, self.D
            for M in self.widths:
                n += M * K + M
                K = M
            return n
        td = 1 if self.time_dep else 0
        return self.H * (self.D + td) + self.H + self.D * (self.H + td) + self.D


@dataclass
class OracleResult:
    u: np.ndarray
    nf: int
    naccept: int
    nreject: int
    retcode: int
    saveval: np.ndarray
    dt_log: np.ndarray
    accept_log: np.ndarray
    eest_log: np.ndarray
    dt_init: float
    t_final: float
    steps: list = field(default_factory=list)   # (t, dt, EEst, eigen_est) per accepted step
    usave: np.ndarray | None = None              # (n_saveat, D, B) states at the saveat times


class Oracle:
    """Handle on one oracle solve: forward(), then backward()."""

    def __init__(self, cfg: OracleConfig, f64: bool = False):
        self.cfg = cfg
        self.f64 = f64
        self.dtype = np.float64 if f64 else np.float32
        self.pfx = "orc64_" if f64 else "orc32_"
        self.lib = _lib(f64)
        c = _Cfg()
        for name in ("D", "H", "B", "act1", "act2", "kblock1", "kblock2", "alg", "reg_kind", "max_steps", "nthreads"):
            setattr(c, name, int(getattr(cfg, name)))
        c.time_dep = 1 if cfg.time_dep else 0
        c.t0, c.t1, c.abstol, c.reltol, c.dtmin = cfg.t0, cfg.t1, cfg.abstol, cfg.reltol, cfg.dtmin
        self._keep = []
        if cfg.forced_dt is not None:
            fd = np.ascontiguousarray(cfg.forced_dt, dtype=np.float64)
            fa = np.ascontiguousarray(cfg.forced_accept, dtype=np.int32)
            self._keep += [fd, fa]
            c.n_forced = len(fd)
            c.forced_dt = fd.ctypes.data_as(C.POINTER(C.c_double))
            c.forced_accept = fa.ctypes.data_as(C.POINTER(C.c_int))
        if cfg.widths is not None:
            assert cfg.widths[-1] == cfg.D and len(cfg.widths) <= 8
            c.n_layers = len(cfg.widths)
            for i, wdt in enumerate(cfg.widths):
                c.width[i] = int(wdt)
                c.act[i] = int(cfg.acts[i]) if cfg.acts is not None else ACT_TANH
            c.pre_act = int(cfg.pre_act)
            c.time_dep = 0
            c.H = max(cfg.widths)
        c.arith = int(cfg.arith)
        if cfg.csq_extra:
            en = np.asfortranarray(cfg.csq_noise, dtype=self.dtype)
            assert en.shape == (cfg.D - cfg.csq_extra, cfg.B)
            self._keep.append(en)
            c.csq_extra = int(cfg.csq_extra)
            c.csq_noise = en.ctypes.data_as(C.c_void_p)
        if cfg.saveat is not None:
            sa = np.ascontiguousarray(cfg.saveat, dtype=np.float64)
            self._keep.append(sa)
            c.n_saveat = len(sa)
            c.saveat = sa.ctypes.data_as(C.POINTER(C.c_double))
        self._c = c
        self.h = C.c_void_p()
        rc = self._fn("create")(C.byref(c), C.byref(self.h))
        if rc != 0:
            raise ValueError(f"oracle create failed rc={rc}")

    def _fn(self, name):
        return getattr(self.lib, self.pfx + name)

    def close(self):
        if self.h:
            self._fn("destroy")(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ptr(self, a):
        return a.ctypes.data_as(C.c_void_p)

    def rhs(self, p, z, t):
        D, H, B = self.cfg.D, self.cfg.H, self.cfg.B
        p = np.ascontiguousarray(p, dtype=self.dtype)
        z = np.asfortranarray(z, dtype=self.dtype)
        k = np.zeros((D, B), dtype=self.dtype, order="F")
        h = np.zeros((H, B), dtype=self.dtype, order="F")
        f = self._fn("rhs")
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        f(C.byref(self._c), self._ptr(p), self._ptr(z), float(t), self._ptr(k), self._ptr(h))
        return k, h

    def forward(self, x, p) -> OracleResult:
        D, B = self.cfg.D, self.cfg.B
        x = np.asfortranarray(x, dtype=self.dtype)
        p = np.ascontiguousarray(p, dtype=self.dtype)
        assert x.shape == (D, B) and p.size == self.cfg.n_params
        u = np.zeros((D, B), dtype=self.dtype, order="F")
        st = _Stats()
        f = self._fn("forward")
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        f(self.h, self._ptr(x), self._ptr(p), self._ptr(u), C.byref(st))
        sv = np.zeros(max(st.n_saved, 1), dtype=self.dtype)
        g = self._fn("get_saveval")
        g.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        g(self.h, self._ptr(sv), len(sv))
        sv = sv[: st.n_saved]
        nlog = st.naccept + st.nreject + 1
        dts = np.zeros(nlog, dtype=np.float64)
        acc = np.zeros(nlog, dtype=np.int32)
        ee = np.zeros(nlog, dtype=np.float64)
        gl = self._fn("get_log")
        gl.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        n = gl(self.h, self._ptr(dts), self._ptr(acc), self._ptr(ee), nlog)
        self._naccept = st.naccept
        steps = []
        gs = self._fn("get_step")
        gs.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_double)] * 4
        for j in range(st.naccept):
            a = [C.c_double() for _ in range(4)]
            gs(self.h, j, *[C.byref(v) for v in a])
            steps.append(tuple(v.value for v in a))
        usave = None
        if self.cfg.saveat is not None:
            nsv = len(self.cfg.saveat)
            buf = np.zeros((nsv, B, D), dtype=self.dtype)
            gu = self._fn("get_usave"); gu.argtypes = [C.c_void_p, C.c_void_p]
            got = gu(self.h, self._ptr(buf))
            usave = np.ascontiguousarray(buf[:got].transpose(0, 2, 1))
        res = OracleResult(u=u, nf=st.nf, naccept=st.naccept, nreject=st.nreject, retcode=st.retcode, saveval=sv,
                            dt_log=dts[:n], accept_log=acc[:n], eest_log=ee[:n], dt_init=st.dt_init,
                            t_final=st.t_final, steps=steps)
        res.usave = usave
        return res

    def backward(self, du, dsaveval=None, hi: bool = False, dusave=None, first_dt_tracked: bool = True):
        """Discrete adjoint with frozen dt.  Returns dp, dx, dtbar[naccept], tbar[naccept].
        first_dt_tracked=True: Appendix A.6 with detach_dt = all_but_first -- the initial-dt heuristic
        (and through it the start time of every later step and the length of the last one) stays on the tape;
        first_dt_tracked="term" returns that extra term alone (the difference of the two modes, without their rounding).
        hi=True (FP32 oracle only): cotangents and accumulations in Float64 over the same FP32
        forward values -- the reference for judging the accuracy of FP32 adjoints."""
        D, B = self.cfg.D, self.cfg.B
        du = np.asfortranarray(du, dtype=self.dtype)
        if dsaveval is None:
            dsaveval = np.zeros(1, dtype=self.dtype)
        dsaveval = np.ascontiguousarray(dsaveval, dtype=self.dtype)
        dp = np.zeros(self.cfg.n_params, dtype=self.dtype)
        dx = np.zeros((D, B), dtype=self.dtype, order="F")
        nacc = max(self._naccept, 1)
        dtbar = np.zeros(nacc, dtype=np.float64)
        tbar = np.zeros(nacc, dtype=np.float64)
        sd = self._fn("set_detach"); sd.argtypes = [C.c_void_p, C.c_int]
        sd(self.h, 2 if first_dt_tracked == "term" else (1 if first_dt_tracked else 0))
        f = self._fn("backward_hi" if (hi and not self.f64) else "backward")
        f.argtypes = [C.c_void_p] * 8
        dus = None
        if dusave is not None:      # (n_saveat, D, B) -> per save a column-major D x B block
            dus = np.ascontiguousarray(np.asarray(dusave, dtype=self.dtype).transpose(0, 2, 1))
        rc = f(self.h, self._ptr(du), self._ptr(dsaveval), self._ptr(dp), self._ptr(dx), self._ptr(dtbar), self._ptr(tbar),
               self._ptr(dus) if dus is not None else None)
        if rc != 0:
            raise RuntimeError(f"oracle backward rc={rc}")
        return dp, dx, dtbar[: self._naccept], tbar[: self._naccept]


def glorot_params(rng: np.random.Generator, D: int, H: int, time_dep: bool = True, dtype=np.float32) -> np.ndarray:
    """Flux 0.11.6 glorot_uniform weights, zero biases, in Flux.destructure order
    (W1, b1, W2, b2; each W column-major out x in).  SURVEY.md section 8d."""
    td = 1 if time_dep else 0

    def glorot(out, inp):
        s = np.sqrt(6.0 / (inp + out))
        return rng.uniform(-s, s, size=(out, inp)).astype(dtype)

    W1 = glorot(H, D + td)
    W2 = glorot(D, H + td)
    return np.concatenate([W1.flatten(order="F"), np.zeros(H, dtype), W2.flatten(order="F"), np.zeros(D, dtype)]).astype(dtype)


def glorot_chain_params(rng: np.random.Generator, D: int, widths, dtype=np.float32, bias_scale: float = 0.0) -> np.ndarray:
    """Parameters of Chain(Dense(D,w0), Dense(w0,w1), ...) in Flux.destructure order (W_l column-major out x in, b_l);
    glorot_uniform weights; biases zero like Flux's default unless bias_scale is given (tests exercise db too)."""
    parts, K = [], D
    for M in widths:
        s = np.sqrt(6.0 / (K + M))
        parts.append(rng.uniform(-s, s, size=(M, K)).astype(dtype).flatten(order="F"))
        parts.append((bias_scale * rng.standard_normal(M)).astype(dtype))
        K = M
    return np.concatenate(parts).astype(dtype)
