"""Host-side mirror of the reference's neural-ODE layer surface over libregnde.so.

Same names, argument meaning and return tuples as the Julia reference, so the
parity tests read like the reference's own test script (test/test_node.jl):

    TDChain(Dense(3, 10, tanh), Dense(11, 2))            src/models/basic.jl:2-28
    MLPDynamics(784, 100)                                experiments/mnist_node.jl:41-54
    TrackedNeuralODE(model, tspan, time_dep, regularize, solver; reltol, abstol, ...)
                                                         src/models/neural_ode.jl:10-32
    node(x, p; func=..., tspan=...) -> (res, nfe, sv)    src/models/neural_ode.jl:48-144
    track / untrack                                      src/RegNeuralDE.jl:24-25

Arrays are torch CUDA tensors used as device buffers only (PyTorch is plumbing:
memory, streams, autograd glue).  A state of shape (D, B) is held column-major
(feature index fastest), i.e. the memory of a Julia ``D x B`` CuArray; helpers
``colmajor``/``from_colmajor`` convert.  All arithmetic runs in the CUDA library;
there is no CPU or eager fallback.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib as L


# ---- solver / regulariser selectors (the reference passes Julia objects/closures) ----
@dataclass(frozen=True)
class Tsit5:
    """OrdinaryDiffEq.Tsit5()"""
    alg: int = L.ALG_TSIT5


@dataclass(frozen=True)
class AutoTsit5:
    """OrdinaryDiffEq.AutoTsit5(Tsit5()) -- composite algorithm exposing eigen_est"""
    inner: Tsit5 = Tsit5()
    alg: int = L.ALG_AUTO_TSIT5


@dataclass(frozen=True)
class SaveFunc:
    """One of the `func(u, t, integrator)` closures the reference hands to SavingCallback.
    Closures cannot cross a C ABI, so the ones the reference uses are enumerated."""
    kind: int
    name: str


#: (u, t, integrator) -> integrator.EEst * integrator.dt      neural_ode.jl:116, mnist_node.jl:67
ERROR_ESTIMATE = SaveFunc(L.REG_ERR_DT, "error_estimate")
#: (u, t, integrator) -> abs(integrator.eigen_est * integrator.dt)      test/test_node.jl:75
STIFFNESS_ESTIMATE = SaveFunc(L.REG_STIFF_DT_ABS, "stiffness_estimate")
#: stability_size * |eigen_est| with the iszero/isnan guard              mnist_node.jl:76-79
STIFFNESS_SCALED = SaveFunc(L.REG_STIFF_SCALED, "stiff_est")
#: EEst*dt + 0.1*stability_size*eigen_est                               mnist_node.jl:88-97
ERROR_PLUS_STIFFNESS = SaveFunc(L.REG_ERR_PLUS_STIFF, "error_stiff_est")


def colmajor(x: torch.Tensor) -> torch.Tensor:
    """(D, B) tensor -> flat device buffer in Julia (column-major) order."""
    return x.t().contiguous().view(-1)


def from_colmajor(buf: torch.Tensor, D: int, B: int) -> torch.Tensor:
    """flat column-major buffer -> (D, B) tensor view (no copy)."""
    return buf.view(B, D).t()


# ---- layers (only what the hot path's vector fields need) ----------------------
class Dense:
    """Flux.Dense(in, out, σ): weight (out, in) glorot_uniform, zero bias (Flux 0.11.6)."""

    def __init__(self, inp: int, out: int, act=None, *, generator: Optional[torch.Generator] = None):
        self.inp, self.out = inp, out
        self.act = L.ACT_TANH if act in (torch.tanh, math.tanh, "tanh") else L.ACT_IDENTITY
        if act not in (None, "identity", torch.tanh, math.tanh, "tanh"):
            raise ValueError("only identity and tanh activations are supported by the CUDA field kernels")
        s = math.sqrt(6.0 / (inp + out))
        self.W = (torch.rand(out, inp, generator=generator, dtype=torch.float32) * 2 - 1) * s
        self.b = torch.zeros(out, dtype=torch.float32)

    def destructure(self) -> torch.Tensor:
        # Flux.destructure: vec(W) column-major, then b
        return torch.cat([self.W.t().contiguous().view(-1), self.b])


class TDChain:
    """Time-conditioned MLP: a 1xB row holding t is vcat'd onto the input of EVERY layer
    (src/models/basic.jl:16-28).  The CUDA kernels implement the 2-layer case."""

    def __init__(self, *layers: Dense):
        if len(layers) != 2:
            raise NotImplementedError("the fused field kernels implement 2-layer time-concatenated chains")
        l1, l2 = layers
        if l2.inp != l1.out + 1 or l2.out != l1.inp - 1:
            raise ValueError("TDChain(Dense(D+1,H,σ), Dense(H+1,D)) expected")
        self.layers = layers
        self.D, self.H = l2.out, l1.out

    def destructure(self) -> torch.Tensor:
        return torch.cat([l.destructure() for l in self.layers])


class Chain:
    """Flux.Chain of Dense layers, optionally led by the elementwise ``x -> tanh.(x)`` -- the Latent-ODE generator
    dynamics (experiments/latent_ode.jl:109-121).  Used with ``time_dep=False``: the field is ``re(p)(u)``
    (src/models/neural_ode.jl:57).  Up to 8 layers; the last one maps back to the state dimension."""

    def __init__(self, *layers, pre_act=None):
        layers = list(layers)
        if layers and not isinstance(layers[0], Dense):        # Chain(x -> tanh.(x), Dense..., ) spelling of the reference
            if layers[0] not in (torch.tanh, math.tanh, "tanh"):
                raise ValueError("only an elementwise tanh may lead the chain")
            pre_act = "tanh"
            layers = layers[1:]
        if not 1 <= len(layers) <= 8:
            raise NotImplementedError("chain fields hold 1..8 Dense layers")
        for a, b in zip(layers[:-1], layers[1:]):
            if b.inp != a.out:
                raise ValueError("Dense layer sizes do not chain")
        if layers[-1].out != layers[0].inp:
            raise ValueError("a vector field maps the state dimension onto itself")
        self.layers = tuple(layers)
        self.pre_act = L.ACT_TANH if pre_act in (torch.tanh, math.tanh, "tanh") else L.ACT_IDENTITY
        self.D = layers[0].inp
        self.H = max(l.out for l in layers)

    def destructure(self) -> torch.Tensor:
        return torch.cat([l.destructure() for l in self.layers])


def MLPDynamics(inp: int, hidden: int, *, generator: Optional[torch.Generator] = None) -> TDChain:
    """experiments/mnist_node.jl:41-54: Dense(in+1, hidden, tanh), Dense(hidden+1, in, tanh)."""
    return TDChain(Dense(inp + 1, hidden, "tanh", generator=generator), Dense(hidden + 1, inp, "tanh", generator=generator))


def track(p: torch.Tensor) -> torch.Tensor:
    """RegNeuralDE.track: mark an array as a differentiable parameter (Tracker.param)."""
    return p.detach().clone().requires_grad_(True)


def untrack(p: torch.Tensor) -> torch.Tensor:
    """RegNeuralDE.untrack: Tracker.data"""
    return p.detach()


class SavedValues:
    """DiffEqCallbacks.SavedValues as returned by the regularised functor: .t, .saveval"""

    def __init__(self, t: torch.Tensor, saveval: torch.Tensor):
        self.t = t
        self.saveval = saveval

    def __len__(self):
        return int(self.saveval.numel())


class _Handle:
    """Owns one rnde_handle (fixed batch size / regulariser / solver)."""

    def __init__(self, cfg: L.Config):
        L.require_device()
        self.lib = L.lib()
        self.cfg = cfg
        self.h = C.c_void_p()
        self.serial = 0          # number of taped forwards: a backward may only consume the latest tape
        rc = self.lib.rnde_create(C.byref(cfg), C.byref(self.h))
        if rc != L.OK:
            raise L.RndeError(rc, f"rnde_create(D={cfg.state_dim}, H={cfg.hidden_dim}, B={cfg.batch})")

    def check(self, rc: int, what: str):
        if rc != L.OK:
            raise L.RndeError(rc, what + ": " + self.lib.rnde_last_error(self.h).decode())

    def check_tape(self, serial: int):
        """The tape lives in the handle (one per batch size / regulariser): a second forward overwrites it."""
        if serial != self.serial:
            raise RuntimeError("backward through a solve whose tape was overwritten by a later forward of the same node and "
                               "batch size; call backward before the next forward (or use a second TrackedNeuralODE)")

    def __del__(self):
        try:
            if self.h:
                self.lib.rnde_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Solve(torch.autograd.Function):
    """Glue between torch autograd and rnde_forward / rnde_backward (the role Tracker's
    custom-gradient hook plays in the Julia wrapper, julia/RegNeuralDEB200.jl)."""

    @staticmethod
    def forward(ctx, xbuf: torch.Tensor, p: torch.Tensor, node: "TrackedNeuralODE", hd: _Handle):
        cfg = hd.cfg
        D, B = cfg.state_dim, cfg.batch
        u = torch.empty(D * B, device=xbuf.device, dtype=torch.float32)
        sv = torch.zeros(cfg.tape_capacity + 1, device=xbuf.device, dtype=torch.float32)
        st = L.Stats()
        rc = hd.lib.rnde_forward(hd.h, xbuf.data_ptr(), p.data_ptr(), u.data_ptr(), sv.data_ptr(), C.byref(st), _stream_ptr())
        node.last_stats = st
        hd.check(rc, "rnde_forward")
        hd.serial += 1
        ctx.serial = hd.serial
        ctx.hd = hd
        ctx.n_saved = st.n_saved
        ctx.p_ref = p            # keeps the parameter buffer alive until backward
        return u, sv[: st.n_saved]

    @staticmethod
    def backward(ctx, du: torch.Tensor, dsv: torch.Tensor):
        hd = ctx.hd
        hd.check_tape(ctx.serial)
        cfg = hd.cfg
        D, B = cfg.state_dim, cfg.batch
        du = du.contiguous() if du is not None else torch.zeros(D * B, device=ctx.p_ref.device)
        dsv_full = torch.zeros(cfg.tape_capacity + 1, device=du.device, dtype=torch.float32)
        if dsv is not None and ctx.n_saved > 0:
            dsv_full[: ctx.n_saved] = dsv
        dp = torch.empty(ctx.p_ref.numel(), device=du.device, dtype=torch.float32)
        dx = torch.empty(D * B, device=du.device, dtype=torch.float32)
        rc = hd.lib.rnde_backward(hd.h, du.data_ptr(), dsv_full.data_ptr(), dp.data_ptr(), dx.data_ptr(), _stream_ptr())
        hd.check(rc, "rnde_backward")
        return dx, dp, None, None


class _SolveSaveat(torch.autograd.Function):
    """Multi-save functors (neural_ode.jl:79-108, :146-180): the result is the state at every saveat time."""

    @staticmethod
    def forward(ctx, xbuf: torch.Tensor, p: torch.Tensor, node: "TrackedNeuralODE", hd: _Handle, nsave: int):
        cfg = hd.cfg
        D, B = cfg.state_dim, cfg.batch
        us = torch.zeros(D * nsave * B, device=xbuf.device, dtype=torch.float32)
        sv = torch.zeros(cfg.tape_capacity + 1, device=xbuf.device, dtype=torch.float32)
        st = L.Stats()
        rc = hd.lib.rnde_forward_saveat(hd.h, xbuf.data_ptr(), p.data_ptr(), None, us.data_ptr(), sv.data_ptr(), C.byref(st), _stream_ptr())
        node.last_stats = st
        hd.check(rc, "rnde_forward_saveat")
        hd.serial += 1
        ctx.hd, ctx.n_saved, ctx.p_ref, ctx.serial = hd, st.n_saved, p, hd.serial
        return us, sv[: st.n_saved]

    @staticmethod
    def backward(ctx, dus: torch.Tensor, dsv: torch.Tensor):
        hd = ctx.hd
        hd.check_tape(ctx.serial)
        cfg = hd.cfg
        D, B = cfg.state_dim, cfg.batch
        dev = ctx.p_ref.device
        dus = dus.contiguous() if dus is not None else None
        dsv_full = torch.zeros(cfg.tape_capacity + 1, device=dev, dtype=torch.float32)
        if dsv is not None and ctx.n_saved > 0:
            dsv_full[: ctx.n_saved] = dsv
        if dus is None:
            dus = torch.zeros(1, device=dev)    # unreachable in practice: the saved states are the primary output
        dp = torch.empty(ctx.p_ref.numel(), device=dev, dtype=torch.float32)
        dx = torch.empty(D * B, device=dev, dtype=torch.float32)
        rc = hd.lib.rnde_backward_saveat(hd.h, None, dus.data_ptr(), dsv_full.data_ptr(), dp.data_ptr(), dx.data_ptr(), _stream_ptr())
        hd.check(rc, "rnde_backward_saveat")
        return dx, dp, None, None, None


def resolve_arith(model, kernel_variant: int = L.KERNEL_AUTO, saveat: bool = False) -> int:
    """Default canonical arithmetic of a node (mirrors v5_shape_ok / try_variant in csrc/regnde.cu): the split-K stepper for
    2-layer fields with 128 < D/4 <= 224 rows per CTA, D % 8 == 0 and 4 <= H <= 112, on the AUTO or cluster-4 variant."""
    if isinstance(model, Chain) or saveat or kernel_variant not in (L.KERNEL_AUTO, L.KERNEL_CLUSTER4):
        return L.ARITH_FMA_CHAIN
    if os.environ.get("RNDE_ARITH"):          # A/B of the steppers (tools/fwd_time.py): 0 = FMA_CHAIN (fwd4_kernel), 2 = SPLITK (fwd4s_kernel)
        return int(os.environ["RNDE_ARITH"])
    D, H = model.D, model.H
    if D % 8 == 0 and 128 < D // 4 <= 224 and 4 <= H <= 112:
        return L.ARITH_SPLITK
    return L.ARITH_FMA_CHAIN


class TrackedNeuralODE:
    """src/models/neural_ode.jl:1-33.  ``TrackedNeuralODE(model, tspan, time_dep, regularize,
    solver; reltol, abstol, save_everystep=false, save_start=false)``."""

    def __init__(self, model: TDChain, tspan: Sequence[float], time_dep: bool, regularize: bool, solver=Tsit5(), *,
                 reltol: float = 1.4e-8, abstol: float = 1.4e-8, save_everystep: bool = False, save_start: bool = False,
                 saveat=None, maxiters: int = 0, tape_capacity: int = 256, kernel_variant: int = L.KERNEL_AUTO,
                 kblock: int = 0, device: str = "cuda", dist_mode: int = L.DIST_SINGLE, rank: int = 0, world: int = 1,
                 arith: Optional[int] = None, detach_dt: str = "all_but_first"):
        if save_everystep:
            raise NotImplementedError("save_everystep=true has no call site in the reference; use saveat")
        # return_multiple = haskey(kwargs, :saveat)  (neural_ode.jl:11): fixes which functor the object dispatches to
        self.return_multiple = saveat is not None
        self.saveat = None if saveat is None else [float(v) for v in saveat]
        if isinstance(model, Chain) == bool(time_dep):
            raise NotImplementedError("TDChain fields are time dependent (basic.jl:16-28), Chain fields are not (latent_ode.jl:109-121)")
        L.require_device()
        self.model = model
        self.device = torch.device(device)
        self.p = model.destructure().to(self.device)
        self.tspan = (float(tspan[0]), float(tspan[1]))
        self.time_dep, self.regularize = bool(time_dep), bool(regularize)
        self.solver = solver
        self.reltol, self.abstol = float(reltol), float(abstol)
        # normalised once: every buffer below is sized tape_capacity + 1 (the library's default for <= 0 is 256)
        self.maxiters, self.tape_capacity = maxiters, (int(tape_capacity) if int(tape_capacity) > 0 else 256)
        self.kernel_variant, self.kblock = kernel_variant, kblock
        # data parallel: DIST_EXACT shares the step sequence of the global batched solve across ranks (x holds this
        # rank's columns, all shards equal); DIST_INDEPENDENT / DIST_SINGLE integrate the local columns on their own
        self.dist_mode, self.rank, self.world = dist_mode, rank, world
        # canonical arithmetic of the layer products: ARITH_FMA_CHAIN (all variants), ARITH_FIXED24 (exact integer tensor-core
        # MMAs) or ARITH_SPLITK (the 8x8-tile FFMA2 stepper of the cluster-4 variant).  None = SPLITK where that stepper
        # applies (MNIST-shaped 2-layer fields on the AUTO / cluster-4 variants without saveat), FMA_CHAIN elsewhere.
        self.arith = resolve_arith(model, kernel_variant, self.return_multiple) if arith is None else int(arith)
        # what the backward differentiates (utils.jl:21-23 makes tspan tracked): "all_but_first" = the
        # frozen-step discrete adjoint plus the gradient through the first step size (initial-dt heuristic), the recalled
        # upstream behaviour; "all" = the frozen-step adjoint only
        if detach_dt not in ("all", "all_but_first", "first_term_only"):      # the last one is a test diagnostic: only the extra term
            raise ValueError('detach_dt must be "all" or "all_but_first"')
        self.detach_dt = detach_dt
        self._handles: dict = {}
        self.last_stats: Optional[L.Stats] = None

    # Flux.trainable-style access
    def parameters(self) -> torch.Tensor:
        return self.p

    def _handle(self, B: int, reg_kind: int, need_backward: bool, max_saveat: int = 0) -> _Handle:
        key = (B, reg_kind, need_backward, max_saveat)
        if key not in self._handles:
            cfg = L.Config()
            cfg.struct_bytes = C.sizeof(L.Config)
            cfg.state_dim, cfg.hidden_dim, cfg.batch = self.model.D, self.model.H, B
            if isinstance(self.model, Chain):
                cfg.n_layers, cfg.pre_act, cfg.time_dep = len(self.model.layers), self.model.pre_act, 0
                for i, lay in enumerate(self.model.layers):
                    cfg.layer_width[i], cfg.layer_act[i] = lay.out, lay.act
            else:
                cfg.act_hidden, cfg.act_out = self.model.layers[0].act, self.model.layers[1].act
                cfg.time_dep = 1
            cfg.kblock = self.kblock
            cfg.alg = self.solver.alg
            cfg.reg_kind = reg_kind
            cfg.max_steps = self.maxiters
            cfg.tape_capacity = self.tape_capacity
            cfg.need_backward = 1 if need_backward else 0
            cfg.kernel_variant = self.kernel_variant
            cfg.dist_mode, cfg.rank, cfg.nranks = self.dist_mode, self.rank, self.world
            cfg.t0, cfg.t1 = self.tspan
            cfg.abstol, cfg.reltol, cfg.dtmin = self.abstol, self.reltol, 0.0
            cfg.max_saveat = max_saveat
            cfg.arith = self.arith
            cfg.global_batch = B * (self.world if self.dist_mode == L.DIST_EXACT else 1)
            hd = _Handle(cfg)
            hd.check(hd.lib.rnde_set_detach(hd.h, {"all": L.DETACH_ALL, "all_but_first": L.DETACH_ALL_BUT_FIRST, "first_term_only": L.DETACH_FIRST_TERM_ONLY}[self.detach_dt]), "rnde_set_detach")
            if self.dist_mode == L.DIST_EXACT and self.world > 1:
                from .parallel import exchange_ipc_handles
                mine = (C.c_ubyte * 64)()
                hd.check(hd.lib.rnde_dist_export(hd.h, mine), "rnde_dist_export")
                allh = exchange_ipc_handles(bytes(mine), self.world)
                buf = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(allh))
                hd.check(hd.lib.rnde_dist_import(hd.h, buf, self.world), "rnde_dist_import")
            self._handles[key] = hd
        return self._handles[key]

    def __call__(self, x: torch.Tensor, p: Optional[torch.Tensor] = None, *, func: Optional[SaveFunc] = None, tspan=None,
                 saveat=None):
        """-> (res, nfe, sv): final state (D, B) -- or, for a node built with ``saveat``, the states at the save
        times as (D, nsave, B) (diffeqsol_to_3dtrackedarray, utils.jl) -- then sol.destats.nf, SavedValues or None.
        The per-call ``saveat`` keyword replaces the stored times for this call (update_saveat!, neural_ode.jl:35-45)."""
        if saveat is not None and not self.return_multiple:
            raise ValueError("this node was built without saveat: its functor returns the final state only (neural_ode.jl:11)")
        p = self.p if p is None else p
        D = self.model.D
        if x.dim() != 2 or x.shape[0] != D:
            raise ValueError(f"x must be ({D}, B)")
        if not x.is_cuda or not p.is_cuda:
            raise RuntimeError("regneuralde.jl_b200 runs on CUDA tensors only (no CPU fallback)")
        if p.dtype != torch.float32 or p.dim() != 1 or p.numel() != self.p.numel():
            raise ValueError(f"p must be the flat Float32 parameter vector of length {self.p.numel()} (Flux.destructure order)")
        B = x.shape[1]
        if self.regularize:
            func = ERROR_ESTIMATE if func is None else func     # default of neural_ode.jl:116
            reg_kind = func.kind
        else:
            reg_kind = L.REG_NONE                               # {false,*}: func is ignored (neural_ode.jl:51)
        need_bwd = torch.is_grad_enabled() and (x.requires_grad or p.requires_grad)
        times = None
        if self.return_multiple:
            times = self.saveat if saveat is None else [float(v) for v in (saveat.tolist() if torch.is_tensor(saveat) else saveat)]
        hd = self._handle(B, reg_kind, need_bwd, 0 if times is None else 64 * ((len(times) + 63) // 64))
        t0, t1 = self.tspan if tspan is None else (float(tspan[0]), float(tspan[1]))
        hd.check(hd.lib.rnde_set_tspan(hd.h, t0, t1), "rnde_set_tspan")
        xbuf = colmajor(x.to(torch.float32))
        if times is None:
            ubuf, saveval = _Solve.apply(xbuf, p.contiguous(), self, hd)
            res = from_colmajor(ubuf, D, B)
        else:
            arr = (C.c_float * len(times))(*times)
            hd.check(hd.lib.rnde_set_saveat(hd.h, arr, len(times)), "rnde_set_saveat")
            usbuf, saveval = _SolveSaveat.apply(xbuf, p.contiguous(), self, hd, len(times))
            res = usbuf.view(B, len(times), D).permute(2, 1, 0)        # feat x nsave x batch
        nfe = int(self.last_stats.nf)
        if not self.regularize:
            return res, nfe, None
        n = int(self.last_stats.naccept)
        # times of the saved values: t0, then t after each accepted step
        tcpu = (C.c_float * max(n, 1))()
        dcpu = (C.c_float * max(n, 1))()
        hd.check(hd.lib.rnde_get_steps(hd.h, tcpu, dcpu, None, None, n), "rnde_get_steps")
        # Float32 arithmetic like the stepper's own t + dt (a Float64 sum could differ from it in the last bit)
        ts = np.concatenate([np.asarray([t0], np.float32), np.asarray(tcpu[:n], np.float32) + np.asarray(dcpu[:n], np.float32)])
        return res, nfe, SavedValues(torch.from_numpy(ts), saveval)

    def steps(self, B: int, reg_kind: int = L.REG_NONE, need_backward: bool = False):
        """(t, dt, EEst, eigen_est) of every accepted step of the last solve on that handle."""
        hd = next(h for k, h in self._handles.items() if k[:3] == (B, reg_kind, need_backward))
        n = int(self.last_stats.naccept)
        arrs = [(C.c_float * max(n, 1))() for _ in range(4)]
        hd.check(hd.lib.rnde_get_steps(hd.h, *arrs, n), "rnde_get_steps")
        return [list(a)[:n] for a in arrs]

    def allreduce_(self, *tensors: torch.Tensor) -> None:
        """Sum flat Float32 tensors over the ranks of a reference-exact group, in place, through the library's one-shot
        all-reduce over NVLink peer memory (rnde_allreduce_grads; no NCCL call; bitwise identical result on every rank)."""
        if self.dist_mode != L.DIST_EXACT or self.world <= 1:
            raise RuntimeError("allreduce_ needs dist_mode=DIST_EXACT with world > 1")
        hd = next(iter(self._handles.values()))
        for t in tensors:
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
                raise ValueError("contiguous CUDA Float32 tensors only")
            hd.check(hd.lib.rnde_allreduce_grads(hd.h, t.data_ptr(), t.numel(), _stream_ptr()), "rnde_allreduce_grads")

    def launch_count(self) -> int:
        return sum(int(h.lib.rnde_launch_count(h.h)) for h in self._handles.values())


@dataclass
class DEStats:
    """sol.destats of OrdinaryDiffEq (the fields the reference reads: nf; plus the accept/reject counts)."""
    nf: int
    naccept: int
    nreject: int


@dataclass
class ODESolution:
    """What ``solution(n, x, p; ...)`` hands back (src/models/neural_ode.jl:182-210): the time points and states the
    solver saved, the statistics and the return code.  Without ``saveat`` the node's kwargs (``save_everystep=false``,
    ``save_start=false``) leave exactly the final state; with ``saveat`` the states at those times."""
    t: torch.Tensor
    u: list
    destats: DEStats
    retcode: str
    step_t: torch.Tensor        # end time of every accepted step (not part of the reference object; handy for plots)

    def __getitem__(self, i):
        return self.u[i]

    def __len__(self):
        return len(self.u)


def solution(n: TrackedNeuralODE, x: torch.Tensor, p: Optional[torch.Tensor] = None, *, solver=None, tspan=None, saveat=None) -> ODESolution:
    """src/models/neural_ode.jl:182-210: solve without the regulariser callback, optionally with another solver
    (``Tsit5()`` / ``AutoTsit5(Tsit5())``) or time span, and return the solution object instead of the functor's tuple."""
    times = saveat if saveat is not None else n.saveat
    node = TrackedNeuralODE(n.model, n.tspan if tspan is None else tspan, n.time_dep, False, n.solver if solver is None else solver,
                            reltol=n.reltol, abstol=n.abstol, maxiters=n.maxiters, tape_capacity=n.tape_capacity,
                            kernel_variant=n.kernel_variant, kblock=n.kblock, device=str(n.device),
                            **({"saveat": list(times)} if times is not None else {}))
    with torch.no_grad():
        res, nfe, _ = node(x, n.p if p is None else p)
    st = node.last_stats
    t0, t1 = node.tspan
    hd = next(iter(node._handles.values()))
    k = int(st.naccept)
    tb, db = (C.c_float * max(k, 1))(), (C.c_float * max(k, 1))()
    hd.check(hd.lib.rnde_get_steps(hd.h, tb, db, None, None, k), "rnde_get_steps")
    step_t = torch.tensor([t0] + [tb[i] + db[i] for i in range(k)], dtype=torch.float32)
    if times is not None:
        ts = torch.tensor(list(times), dtype=torch.float32)
        us = [res[:, i, :] for i in range(res.shape[1])]
    else:
        ts = torch.tensor([t1], dtype=torch.float32)
        us = [res]
    return ODESolution(t=ts, u=us, destats=DEStats(int(st.nf), int(st.naccept), int(st.nreject)),
                       retcode="Success" if st.retcode == L.OK else L.lib().rnde_status_string(st.retcode).decode(), step_t=step_t)
