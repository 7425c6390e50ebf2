"""Neural-SDE model surface over libregnde.so: forward solves and, since the end of round 2, their gradient.

    TrackedNeuralDSDE(model1, model2, tspan, regularize, solver; reltol, abstol, ...)      src/models/neural_sde.jl:1-42
    (n::TrackedNeuralDSDE{R,false})(x, p; func) -> (res, nfe1, nfe2, sv)                   src/models/neural_sde.jl:84-146
    ClassifierNSDE(presde, nsde, postsde); m(x, p1, p2, p3; trajectories, func) -> (z, nfe1, nfe2, sv)
                                                                                           src/models/supervised_classification.jl:50-103
    experiments/mnist_nsde.jl:44-84: drift Chain(Dense(32,64,tanh), Dense(64,32)), diagonal diffusion Dense(32,32),
    SOSRI() or AutoSOSRI2(SOSRI2()), reltol = abstol = 1.4f-1.

The adaptive SOSRI / SOSRI2 solve with the RSwM3 noise bookkeeping is ONE persistent kernel of the CUDA library
(csrc/sde_kernel.cuh).  The reference draws its Wiener increments from Julia's MersenneTwister, which cannot be reproduced:
the functor takes the standard normals as the ``noise`` keyword ((n_draws, D, B) tensor; default: torch.randn on the device) and
consumes them in the order documented in include/regnde.h.  The pre / post Dense layers and the mean over trajectories are the
host glue the reference keeps in Flux.  With gradients enabled the solve tapes its accepted steps and torch autograd reaches it
through _SdeSolve: the reverse sweep (csrc/sde_bwd.cuh) is the discrete adjoint of those steps with step sizes and Wiener
increments frozen -- Tracker.gradient through solve(...; sensealg = SensitivityADPassThrough()) (mnist_nsde.jl:201-204), with both
regularisers of the experiment (EEst * dt, and the scaled stiffness estimate of AutoSOSRI2) differentiated."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib as L
from .node import Chain, Dense, ERROR_ESTIMATE, STIFFNESS_SCALED, SaveFunc, SavedValues, _stream_ptr, colmajor, from_colmajor


class SOSRI:
    alg = L.SDE_SOSRI


class AutoSOSRI2:
    """AutoSOSRI2(SOSRI2()): SOSRI2 with the stiffness estimate exposed as integrator.eigen_est (mnist_nsde.jl:55,61)"""
    alg = L.SDE_AUTO_SOSRI2


class _SdeHandle:
    def __init__(self, cfg: L.SdeConfig):
        L.require_device()
        self.lib = L.lib()
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.lib.rnde_sde_create(C.byref(cfg), C.byref(self.h))
        if rc != L.OK:
            raise L.RndeError(rc, f"rnde_sde_create(D={cfg.state_dim}, H={cfg.hidden_dim}, B={cfg.batch})")

    def check(self, rc: int, what: str):
        if rc != L.OK:
            raise L.RndeError(rc, what + ": " + self.lib.rnde_sde_last_error(self.h).decode())

    def __del__(self):
        try:
            if self.h:
                self.lib.rnde_sde_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass


class _SdeSolve(torch.autograd.Function):
    """Glue between torch autograd and rnde_sde_forward / rnde_sde_backward."""

    @staticmethod
    def forward(ctx, xbuf: torch.Tensor, p: torch.Tensor, node, hd, nz: torch.Tensor):
        D, B = node.D, hd.cfg.batch
        u = torch.empty(D * B, device=xbuf.device, dtype=torch.float32)
        sv = torch.zeros(hd.cfg.max_saved if hd.cfg.max_saved > 0 else 1024, device=xbuf.device, dtype=torch.float32)
        st = L.SdeStats()
        rc = hd.lib.rnde_sde_forward(hd.h, xbuf.data_ptr(), p.data_ptr(), nz.data_ptr(), nz.shape[0], u.data_ptr(), sv.data_ptr(), C.byref(st), _stream_ptr())
        node.last_stats, node.last_handle = st, hd
        hd.check(rc, "rnde_sde_forward")
        hd.serial = getattr(hd, "serial", 0) + 1
        ctx.serial, ctx.hd, ctx.n_saved, ctx.p_ref, ctx.shape = hd.serial, hd, int(st.n_saved), p, (D, B)
        return u, sv[: st.n_saved]

    @staticmethod
    def backward(ctx, du, dsv):
        hd = ctx.hd
        if ctx.serial != hd.serial:
            raise RuntimeError("the tape of this SDE solve was overwritten by a later solve on the same handle")
        D, B = ctx.shape
        dev = ctx.p_ref.device
        du = du.contiguous() if du is not None else torch.zeros(D * B, device=dev)
        dsv_full = torch.zeros(max(ctx.n_saved, 1), device=dev, dtype=torch.float32)
        if dsv is not None and ctx.n_saved > 0:
            dsv_full[: ctx.n_saved] = dsv
        dp = torch.empty(ctx.p_ref.numel(), device=dev, dtype=torch.float32)
        dx = torch.empty(D * B, device=dev, dtype=torch.float32)
        hd.check(hd.lib.rnde_sde_backward(hd.h, du.data_ptr(), dsv_full.data_ptr() if ctx.n_saved > 0 else None, dp.data_ptr(), dx.data_ptr(), _stream_ptr()),
                 "rnde_sde_backward")
        return dx, dp, None, None, None


class TrackedNeuralDSDE:
    """src/models/neural_sde.jl:1-42.  model1 = drift Chain(Dense(D,H,tanh), Dense(H,D)), model2 = diffusion Dense(D,D);
    p = vcat(p1, p2) (neural_sde.jl:16-18)."""

    def __init__(self, model1: Chain, model2: Dense, tspan: Sequence[float], regularize: bool, solver=SOSRI(), *, reltol: float = 1.4e-1,
                 abstol: float = 1.4e-1, save_everystep: bool = False, save_start: bool = False, maxiters: int = 0, max_saved: int = 1024,
                 device: str = "cuda", tape_capacity: int = 1024):
        if save_everystep:
            raise NotImplementedError("the multi-save functors {R,true} have no call site in the reference experiments")
        if not (isinstance(model1, Chain) and len(model1.layers) == 2 and model1.pre_act == 0 and model1.layers[0].act == L.ACT_TANH
                and model1.layers[1].act == L.ACT_IDENTITY and isinstance(model2, Dense) and model2.act == L.ACT_IDENTITY):
            raise NotImplementedError("the CUDA stepper implements the experiment's drift Chain(Dense(D,H,tanh), Dense(H,D)) and linear "
                                      "diagonal diffusion Dense(D,D) (experiments/mnist_nsde.jl:73-74)")
        D, H = model1.layers[0].inp, model1.layers[0].out
        if model1.layers[1].out != D or model2.inp != D or model2.out != D:
            raise ValueError("drift and diffusion must map the state dimension to itself")
        L.require_device()
        self.D, self.H = D, H
        self.model1, self.model2 = model1, model2
        self.device = torch.device(device)
        self.p = torch.cat([model1.destructure(), model2.destructure()]).to(self.device)
        self.len = model1.destructure().numel()
        self.tspan = (float(tspan[0]), float(tspan[1]))
        self.regularize, self.solver = bool(regularize), solver
        self.reltol, self.abstol, self.maxiters, self.max_saved = float(reltol), float(abstol), maxiters, max_saved
        self.tape_capacity = int(tape_capacity)
        self._handles: dict = {}
        self.last_stats: Optional[L.SdeStats] = None
        self.last_handle: Optional[_SdeHandle] = None

    def _handle(self, B: int, reg_kind: int, need_backward: bool = False) -> _SdeHandle:
        key = (B, reg_kind, need_backward)
        if key not in self._handles:
            cfg = L.SdeConfig()
            cfg.struct_bytes = C.sizeof(L.SdeConfig)
            cfg.state_dim, cfg.hidden_dim, cfg.batch = self.D, self.H, B
            cfg.alg, cfg.reg_kind, cfg.max_steps, cfg.max_saved = self.solver.alg, reg_kind, self.maxiters, self.max_saved
            cfg.t0, cfg.t1 = self.tspan
            cfg.abstol, cfg.reltol = self.abstol, self.reltol
            hd = _SdeHandle(cfg)
            if need_backward:
                hd.check(hd.lib.rnde_sde_enable_tape(hd.h, self.tape_capacity), "rnde_sde_enable_tape")
            self._handles[key] = hd
        return self._handles[key]

    def __call__(self, x: torch.Tensor, p: Optional[torch.Tensor] = None, *, func: Optional[SaveFunc] = None, noise: Optional[torch.Tensor] = None):
        """-> (res (D, B), nfe1, nfe2, sv)"""
        p = self.p if p is None else p
        if x.dim() != 2 or x.shape[0] != self.D:
            raise ValueError(f"x must be ({self.D}, B)")
        if not x.is_cuda or not p.is_cuda:
            raise RuntimeError("regneuralde.jl_b200 runs on CUDA tensors only (no CPU fallback)")
        need_bwd = torch.is_grad_enabled() and (p.requires_grad or x.requires_grad)
        B = x.shape[1]
        if self.regularize:
            func = ERROR_ESTIMATE if func is None else func          # default of neural_sde.jl:119
            if func.kind not in (L.REG_ERR_DT, L.REG_STIFF_SCALED):
                raise NotImplementedError("SDE regularisers of the experiment: EEst*dt (mnist_nsde.jl:48) or the scaled stiffness estimate (:52-56)")
            reg_kind = func.kind
        else:
            reg_kind = L.REG_NONE
        if noise is None:
            noise = torch.randn(256, self.D, B, device=x.device, dtype=torch.float32)
        if tuple(noise.shape[1:]) != (self.D, B) or noise.dtype != torch.float32 or not noise.is_cuda:
            raise ValueError(f"noise must be a CUDA Float32 tensor of shape (n_draws, {self.D}, {B})")
        hd = self._handle(B, reg_kind, need_bwd)
        xbuf = colmajor(x.to(torch.float32))
        if need_bwd:
            ubuf, saveval = _SdeSolve.apply(xbuf, p.contiguous(), self, hd, noise.contiguous())
            st = self.last_stats
            res = from_colmajor(ubuf, self.D, B)
            if not self.regularize:
                return res, int(st.nfe1), int(st.nfe2), None
            return res, int(st.nfe1), int(st.nfe2), SavedValues(torch.zeros(0), saveval)
        u = torch.empty(self.D * B, device=x.device, dtype=torch.float32)
        sv = torch.zeros(hd.cfg.max_saved if hd.cfg.max_saved > 0 else 1024, device=x.device, dtype=torch.float32)
        st = L.SdeStats()
        nz = noise.contiguous()
        rc = hd.lib.rnde_sde_forward(hd.h, xbuf.data_ptr(), p.contiguous().data_ptr(), nz.data_ptr(), nz.shape[0], u.data_ptr(), sv.data_ptr(),
                                     C.byref(st), _stream_ptr())
        self.last_stats, self.last_handle = st, hd
        hd.check(rc, "rnde_sde_forward")
        res = from_colmajor(u, self.D, B)
        if not self.regularize:
            return res, int(st.nfe1), int(st.nfe2), None
        return res, int(st.nfe1), int(st.nfe2), SavedValues(torch.zeros(0), sv[: st.n_saved])

    def attempts(self):
        """(dt, EEst, accepted) of every attempt of the last solve (test introspection)"""
        st, hd = self.last_stats, self.last_handle
        n = int(st.naccept + st.nreject)
        buf = (C.c_float * (3 * max(n, 1)))()
        hd.check(hd.lib.rnde_sde_get_log(hd.h, buf, max(n, 1)), "rnde_sde_get_log")
        return [(buf[3 * i], buf[3 * i + 1], bool(buf[3 * i + 2])) for i in range(n)]

    def launch_count(self) -> int:
        return sum(int(h.lib.rnde_sde_launch_count(h.h)) for h in self._handles.values())


class ClassifierNSDE:
    """supervised_classification.jl:50-103: Dense(784,32) pre-net, the SDE solve on `trajectories` replicas of the batch, Dense(32,10)
    post-net, mean over the trajectories."""

    def __init__(self, presde: Dense, nsde: TrackedNeuralDSDE, postsde: Dense):
        self.presde, self.nsde, self.postsde = presde, nsde, postsde
        dev = nsde.device
        self.p1 = presde.destructure().to(dev)
        self.p2 = nsde.p
        self.p3 = postsde.destructure().to(dev)

    def trainable(self):
        return (self.p1, self.p2, self.p3)

    @staticmethod
    def _dense(p: torch.Tensor, out: int, inp: int, x: torch.Tensor) -> torch.Tensor:
        W = p[: out * inp].view(inp, out).t()
        return W @ x + p[out * inp:][:, None]

    def __call__(self, x: torch.Tensor, p1=None, p2=None, p3=None, *, trajectories: int = 10, **nsde_kwargs):
        """x: (784, B) -> (z (10, B), nfe1, nfe2, sv)"""
        p1 = self.p1 if p1 is None else p1
        p2 = self.p2 if p2 is None else p2
        p3 = self.p3 if p3 is None else p3
        bsize = x.shape[-1]
        xr = x.repeat(1, trajectories)                                   # _expand (supervised_classification.jl:102-103)
        h = self._dense(p1, self.presde.out, self.presde.inp, xr)
        u, nfe1, nfe2, sv = self.nsde(h, p2, **nsde_kwargs)
        z = self._dense(p3, self.postsde.out, self.postsde.inp, u)
        z = z.reshape(z.shape[0], trajectories, bsize).mean(dim=1)
        return z, nfe1, nfe2, sv

    def loss_and_gradient(self, x: torch.Tensor, y_onehot: torch.Tensor, *, lam: float = 1.0e2, trajectories: int = 1, func: Optional[SaveFunc] = None,
                          noise: Optional[torch.Tensor] = None):
        """loss_function of experiments/mnist_nsde.jl:89-110 and its Tracker.gradient (:191-204): logitcrossentropy(model(x), y) +
        lam * mean(sv.saveval).  The pre / post Dense layers and the trajectory mean are torch operations; the SDE solve and its
        reverse sweep run in the library.  Returns dict(loss, ce, reg, nfe1, nfe2, g1, g2, g3)."""
        ps = [q.detach().requires_grad_(True) for q in (self.p1, self.p2, self.p3)]
        kw = {} if noise is None else {"noise": noise}
        z, nfe1, nfe2, sv = self(x, *ps, trajectories=trajectories, func=func, **kw)
        ce = -(torch.log_softmax(z, dim=0) * y_onehot).sum(0).mean()          # Flux.Losses.logitcrossentropy: mean over the batch
        reg = lam * sv.saveval.mean() if sv is not None else torch.zeros((), device=x.device)
        loss = ce + reg
        g1, g2, g3 = torch.autograd.grad(loss, ps)
        return {"loss": loss.detach(), "ce": ce.detach(), "reg": reg.detach(), "nfe1": nfe1, "nfe2": nfe2, "g1": g1, "g2": g2, "g3": g3}
