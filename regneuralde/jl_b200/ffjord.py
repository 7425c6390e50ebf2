"""FFJORD model surface over libregnde.so: log-density and, since round 2, its gradient.

    TrackedFFJORD(model, tspan, time_dep, regularize, solver; dynamics, reltol, abstol, ...)   src/models/ffjord.jl:1-51
    ffjord(x, p, e; regularize=false) -> (logpx, lambda1, lambda2, nfe, sv)                    src/models/ffjord.jl:68-137
    ConcatSquashLinear / MLPDynamics(in, hidden)                                               experiments/ffjord_tabular.jl:47-90

The augmented-state solve (z; delta_logp (; ||f||^2; ||e^T J||^2)) runs in the CUDA library (csrc/csq.cuh inside the generic
Tsit5 stepper); the standard-normal log-density of z(1) is the host glue the reference keeps in Julia.  With gradients enabled
the solve is taped and torch autograd reaches it through node._Solve: the reverse sweep differentiates the field by hand
(csrc/csq_bwd.cuh: reverse mode through forw_n_back, i.e. the second-order terms Tracker would produce) and the parameter
gradients are contracted from the tape in Float64 (dense_wgrad_kernel).  The step sizes are frozen (detach_dt = "all")."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import torch

from . import _lib as L
from .node import ERROR_ESTIMATE, SavedValues, Tsit5, _Handle, _Solve, _stream_ptr, colmajor, from_colmajor


class ConcatSquashLinear:
    """experiments/ffjord_tabular.jl:49-63: glorot_uniform layer / bias / gate weights, zero biases."""

    def __init__(self, inp: int, out: int, *, generator: Optional[torch.Generator] = None):
        self.inp, self.out = inp, out
        u = lambda r, c: (torch.rand(r, c, generator=generator, dtype=torch.float32) * 2 - 1) * math.sqrt(6.0 / (r + c))
        self.layer_W, self.layer_B = u(out, inp), torch.zeros(out, 1)
        self.bias_W, self.bias_B = u(out, 1), torch.zeros(out, 1)
        self.gate_W = u(out, 1)

    def destructure(self) -> torch.Tensor:
        return torch.cat([self.layer_W.t().contiguous().view(-1), self.layer_B.view(-1), self.bias_W.view(-1), self.bias_B.view(-1),
                          self.gate_W.view(-1)])


class CSQDynamics:
    """MLPDynamics(in_dims, hsize) of the tabular experiment (ffjord_tabular.jl:76-90): three ConcatSquash layers, softplus between."""

    def __init__(self, in_dims: int, hsize: int, *, generator: Optional[torch.Generator] = None):
        self.D, self.H = in_dims, hsize
        self.layers = (ConcatSquashLinear(in_dims, hsize, generator=generator), ConcatSquashLinear(hsize, hsize, generator=generator),
                       ConcatSquashLinear(hsize, in_dims, generator=generator))

    def destructure(self) -> torch.Tensor:
        return torch.cat([l.destructure() for l in self.layers])


class TrackedFFJORD:
    """src/models/ffjord.jl:1-51.  ``regularize`` is the type parameter R: True attaches the SavingCallback of EEst*dt."""

    def __init__(self, model: CSQDynamics, tspan: Sequence[float], time_dep: bool, regularize: bool, solver=Tsit5(), *, reltol: float = 1.4e-8,
                 abstol: float = 1.4e-8, save_everystep: bool = False, save_start: bool = False, dynamics=None, maxiters: int = 0,
                 tape_capacity: int = 256, device: str = "cuda"):
        if not isinstance(model, CSQDynamics) or not time_dep:
            raise NotImplementedError("the CUDA field is the time-dependent ConcatSquash MLPDynamics with its own forw_n_back "
                                      "(experiments/ffjord_tabular.jl:92-101)")
        if save_everystep:
            raise NotImplementedError("save_everystep=true has no call site in the reference")
        L.require_device()
        self.model, self.tspan, self.regularize, self.solver = model, (float(tspan[0]), float(tspan[1])), bool(regularize), solver
        self.reltol, self.abstol, self.maxiters, self.tape_capacity = float(reltol), float(abstol), maxiters, tape_capacity
        self.device = torch.device(device)
        self.p = model.destructure().to(self.device)
        self._handles: dict = {}
        self.last_stats = None

    def _handle(self, B: int, extra: int, reg_kind: int, need_backward: bool = False) -> _Handle:
        key = (B, extra, reg_kind, need_backward)
        if key not in self._handles:
            cfg = L.Config()
            cfg.struct_bytes = C.sizeof(L.Config)
            cfg.state_dim, cfg.hidden_dim, cfg.batch = self.model.D + extra, self.model.H, B
            cfg.time_dep, cfg.alg, cfg.reg_kind = 1, self.solver.alg, reg_kind
            cfg.max_steps, cfg.tape_capacity, cfg.need_backward = self.maxiters, self.tape_capacity, (1 if need_backward else 0)
            cfg.t0, cfg.t1 = self.tspan
            cfg.abstol, cfg.reltol = self.abstol, self.reltol
            cfg.global_batch, cfg.csq_extra = B, extra
            self._handles[key] = _Handle(cfg)
        return self._handles[key]

    def __call__(self, x: torch.Tensor, p: Optional[torch.Tensor] = None, e: Optional[torch.Tensor] = None, *, regularize: bool = False):
        """-> (logpx (B,), lambda1, lambda2, nfe, sv).  x, e: (D, B).  ``regularize`` (keyword of the {false} functor) appends the
        kinetic rows; the {true} functor ignores it (ffjord.jl:121)."""
        p = self.p if p is None else p
        D = self.model.D
        if x.dim() != 2 or x.shape[0] != D:
            raise ValueError(f"x must be ({D}, B)")
        if not x.is_cuda or not p.is_cuda:
            raise RuntimeError("regneuralde.jl_b200 runs on CUDA tensors only (no CPU fallback)")
        need_bwd = torch.is_grad_enabled() and (p.requires_grad or x.requires_grad)
        B = x.shape[1]
        e = torch.randn(D, B, device=x.device) if e is None else e          # CUDA.randn(Float32, size(x)...)  (ffjord.jl:71)
        extra = 3 if (regularize and not self.regularize) else 1
        hd = self._handle(B, extra, ERROR_ESTIMATE.kind if self.regularize else L.REG_NONE, need_bwd)
        ebuf = colmajor(e.to(torch.float32))
        ubuf0 = colmajor(torch.cat([x.to(torch.float32), torch.zeros(extra, B, device=x.device)], 0))
        u = torch.empty((D + extra) * B, device=x.device, dtype=torch.float32)
        sv = torch.zeros(hd.cfg.tape_capacity + 1, device=x.device, dtype=torch.float32)
        st = L.Stats()
        hd.check(hd.lib.rnde_set_noise(hd.h, ebuf.data_ptr()), "rnde_set_noise")
        self._noise_keepalive = ebuf          # the library keeps the pointer until the backward pass has run
        if need_bwd:
            u, sv_all = _Solve.apply(ubuf0, p.contiguous(), self, hd)      # sets self.last_stats
            st = self.last_stats
            sv = sv_all
        else:
            rc = hd.lib.rnde_forward(hd.h, ubuf0.data_ptr(), p.contiguous().data_ptr(), u.data_ptr(), sv.data_ptr(), C.byref(st), _stream_ptr())
            self.last_stats = st
            hd.check(rc, "rnde_forward")
        pred = from_colmajor(u, D + extra, B)
        z, delta_logp = pred[:D], pred[D]
        zero = torch.zeros(B, device=x.device)
        lam1, lam2 = (pred[D + 1], pred[D + 2]) if extra == 3 else (zero, zero)
        logpz = (-(math.log(2 * math.pi) + z * z) / 2).sum(0)
        svals = SavedValues(torch.zeros(0), sv[: st.n_saved]) if self.regularize else None
        return logpz - delta_logp, lam1, lam2, int(st.nf), svals


def sample(n: TrackedFFJORD, indims: int, p: Optional[torch.Tensor] = None, *, nsamples: int = 1, z: Optional[torch.Tensor] = None):
    """sample(n, indims, p; nsamples) (src/models/ffjord.jl:160-167): draw z ~ N(0, I) and integrate the flow BACKWARDS over
    [tspan[2], tspan[1]] -> generated data (indims, nsamples).  The reference integrates the deterministic augmented field
    (exact trace) and drops the last row; the data rows do not depend on the trace row, so the library solves the same z-dynamics
    with the trace estimator switched off (zero noise) through rnde_set_reverse_time.  ``z`` overrides the draw (tests)."""
    p = n.p if p is None else p
    if indims != n.model.D:
        raise ValueError(f"indims must be {n.model.D}")
    dev = p.device
    zs = torch.randn(indims, nsamples, device=dev) if z is None else z.to(dev)
    B = zs.shape[1]
    hd = n._handle(B, 1, L.REG_NONE, False)
    e0 = torch.zeros(indims * B, device=dev)
    ubuf0 = colmajor(torch.cat([zs.to(torch.float32), torch.zeros(1, B, device=dev)], 0))
    u = torch.empty((indims + 1) * B, device=dev, dtype=torch.float32)
    sv = torch.zeros(hd.cfg.tape_capacity + 1, device=dev, dtype=torch.float32)
    st = L.Stats()
    hd.check(hd.lib.rnde_set_noise(hd.h, e0.data_ptr()), "rnde_set_noise")
    hd.check(hd.lib.rnde_set_reverse_time(hd.h, 1), "rnde_set_reverse_time")
    try:
        rc = hd.lib.rnde_forward(hd.h, ubuf0.data_ptr(), p.detach().contiguous().data_ptr(), u.data_ptr(), sv.data_ptr(), C.byref(st), _stream_ptr())
    finally:
        hd.lib.rnde_set_reverse_time(hd.h, 0)
    n.last_stats = st
    hd.check(rc, "rnde_forward")
    return from_colmajor(u, indims + 1, B)[:indims]


def loss_and_gradient(n: TrackedFFJORD, x: torch.Tensor, p: torch.Tensor, e: Optional[torch.Tensor] = None, *, lam: float = 1.0e2):
    """loss_function of experiments/ffjord_tabular.jl:137-141 and its Tracker.gradient: -mean(logpx) (+ lam * mean(sv.saveval) for the
    regularised model).  Returns dict(loss, nll, reg, nfe, g)."""
    pg = p.detach().requires_grad_(True)
    logpx, _, _, nfe, sv = n(x, pg, e)
    nll = -logpx.mean()
    reg = lam * sv.saveval.mean() if n.regularize else torch.zeros((), device=x.device)
    loss = nll + reg
    (g,) = torch.autograd.grad(loss, [pg])
    return {"loss": loss.detach(), "nll": nll.detach(), "reg": reg.detach(), "nfe": nfe, "g": g}
