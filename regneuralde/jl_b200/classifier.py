"""ClassifierNODE, the experiment's loss and update_parameters! over libregnde.so.

    ClassifierNODE(preode, node, postode)            src/models/supervised_classification.jl:2-46
    loss_function(x, y, model, p1, p2, p3; λ)        experiments/mnist_node.jl:132-152
    update_parameters!(ps, gs, opt)                  src/utils.jl:149-156
    Optimiser(InvDecay(1e-5), Momentum(0.1, 0.9))    experiments/mnist_node.jl:130

`loss_and_gradient` is the fused training-step path (what Tracker.gradient of the loss
computes, experiments/mnist_node.jl:229-232): forward solve -> head + cross-entropy
forward/backward -> reverse sweep + weight-gradient contractions, all in the library's
own kernels with no torch compute kernels in between.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from .node import (ERROR_ESTIMATE, Dense, SaveFunc, SavedValues, TrackedNeuralODE, _stream_ptr, colmajor, from_colmajor)


_AGG = {"mean": 0, "maximum": 1, "sum": 2}


class ClassifierNODE:
    """pre-net (reshape to 784 x B) -> NODE -> post-net Dense(784, 10), three flat parameter vectors."""

    def __init__(self, preode, node: TrackedNeuralODE, postode: Dense):
        self.preode = preode                       # Chain(x -> reshape(x, 784, :)): no parameters
        self.node = node
        self.postode = postode
        dev = node.device
        self.p1 = torch.zeros(0, device=dev)       # Flux.destructure of a parameter-free chain
        self.p2 = node.p
        self.p3 = postode.destructure().to(dev)
        self.n_classes = postode.out

    def trainable(self):
        """Flux.trainable(m) = (p1, p2, p3)"""
        return (self.p1, self.p2, self.p3)

    def __call__(self, x, p1=None, p2=None, p3=None, **node_kwargs):
        """-> (logits (C, B), nfe, sv); differentiable through torch autograd."""
        p2 = self.p2 if p2 is None else p2
        p3 = self.p3 if p3 is None else p3
        x = self.preode(x) if self.preode is not None else x
        u, nfe, sv = self.node(x, p2, **node_kwargs)
        Cn, D = self.n_classes, self.node.model.D
        W3 = p3[: Cn * D].view(D, Cn).t()
        b3 = p3[Cn * D:]
        return W3 @ u + b3[:, None], nfe, sv

    # ---- fused training-step path -------------------------------------------------
    def loss_and_gradient(self, x: torch.Tensor, y_onehot: torch.Tensor, *, lam: float = 1.0e2, func: Optional[SaveFunc] = None,
                          agg: str = "mean", tspan=None, ce_scale: float = 1.0, reg_scale: float = 1.0):
        """loss = logitcrossentropy(model(x), y) + λ*agg(sv.saveval) and its gradient w.r.t. (p2, p3).
        Returns dict(loss, ce, reg, nfe, naccept, g2, g3, logits).
        ce_scale / reg_scale: weights of this rank's shard in a data-parallel global loss (exact mode: the cross-entropy
        mean runs over world*B samples -> ce_scale = 1/world; the regulariser is global, every rank back-propagates the
        full cotangent through its own columns -> reg_scale = 1; gradients are then SUMMED over ranks)."""
        node = self.node
        D, Cn = node.model.D, self.n_classes
        x = self.preode(x) if self.preode is not None else x
        B = x.shape[1]
        dev = x.device
        if agg not in _AGG:
            raise ValueError(f"agg must be one of {sorted(_AGG)}")
        if x.dim() != 2 or x.shape[0] != D or tuple(y_onehot.shape) != (Cn, B):
            raise ValueError(f"x must be ({D}, B) and y_onehot ({Cn}, B)")
        if not x.is_cuda or not y_onehot.is_cuda:
            raise RuntimeError("regneuralde.jl_b200 runs on CUDA tensors only (no CPU fallback)")
        reg_kind = (ERROR_ESTIMATE if func is None else func).kind if node.regularize else L.REG_NONE
        hd = node._handle(B, reg_kind, True)
        lib = hd.lib
        # a per-call tspan must not outlive the call: without the keyword the node's own tspan applies (neural_ode.jl:53,58)
        t0, t1 = node.tspan if tspan is None else (float(tspan[0]), float(tspan[1]))
        hd.check(lib.rnde_set_tspan(hd.h, t0, t1), "rnde_set_tspan")
        ws = self._workspace(B, dev, hd.cfg.tape_capacity)
        xbuf = colmajor(x.to(torch.float32))
        ybuf = colmajor(y_onehot.to(torch.float32))
        stream = _stream_ptr()
        # forward solve, head and regulariser aggregation are queued back to back without a host round trip; the one
        # stream synchronisation of the step happens inside rnde_backward (it needs the accepted-step count)
        hd.check(lib.rnde_forward(hd.h, xbuf.data_ptr(), self.p2.data_ptr(), ws["u"].data_ptr(), ws["sv"].data_ptr(), None, stream),
                 "rnde_forward")
        hd.serial += 1      # the handle's tape now belongs to this step (node.py _Handle.check_tape)
        hd.check(lib.rnde_head_loss_grad(hd.h, ws["u"].data_ptr(), self.p3.data_ptr(), ybuf.data_ptr(), Cn, float(ce_scale), ws["loss"].data_ptr(),
                                         ws["logits"].data_ptr(), ws["du"].data_ptr(), ws["g3"].data_ptr(), stream), "rnde_head_loss_grad")
        regularized = reg_kind != L.REG_NONE
        if regularized:
            hd.check(lib.rnde_reg_agg(hd.h, _AGG[agg], float(lam), float(reg_scale), ws["sv"].data_ptr(), ws["dsv"].data_ptr(),
                                      ws["reg"].data_ptr(), stream), "rnde_reg_agg")
        hd.check(lib.rnde_backward(hd.h, ws["du"].data_ptr(), ws["dsv"].data_ptr() if regularized else None, ws["g2"].data_ptr(), None, stream),
                 "rnde_backward")
        st = L.Stats()
        hd.check(lib.rnde_last_stats(hd.h, C.byref(st)), "rnde_last_stats")
        node.last_stats = st
        ce = ws["loss"][0]
        reg = ws["reg"][0] if regularized else None
        loss = ce + reg if reg is not None else ce
        return {"loss": loss, "ce": ce, "reg": reg, "nfe": int(st.nf), "naccept": int(st.naccept), "nreject": int(st.nreject),
                "g2": ws["g2"], "g3": ws["g3"], "logits": from_colmajor(ws["logits"], Cn, B), "n_saved": int(st.n_saved)}

    def _workspace(self, B, dev, cap):
        key = (B, str(dev))
        if getattr(self, "_ws_key", None) != key:
            D, Cn = self.node.model.D, self.n_classes
            f = lambda n: torch.empty(n, device=dev, dtype=torch.float32)
            self._ws = {"u": f(D * B), "sv": torch.zeros(cap + 1, device=dev), "dsv": torch.zeros(cap + 1, device=dev), "loss": f(1), "reg": f(1),
                        "logits": f(Cn * B), "du": f(D * B), "g3": f(Cn * D + Cn), "g2": f(self.p2.numel())}
            self._ws_key = key
        return self._ws


class Optimiser:
    """Flux.Optimise.Optimiser(InvDecay(γ), Momentum(η, ρ)) -- experiments/mnist_node.jl:130."""

    def __init__(self, gamma: float = 1.0e-5, eta: float = 0.1, rho: float = 0.9):
        self.gamma, self.eta, self.rho = gamma, eta, rho
        self.state: dict = {}

    def slot(self, p: torch.Tensor):
        k = p.data_ptr()
        if k not in self.state:
            self.state[k] = {"n": 1, "v": torch.zeros_like(p)}
        return self.state[k]


class ADAMOptimiser:
    """Flux.Optimise.Optimiser(WeightDecay(wd), ADAM(eta, beta)) -- experiments/ffjord_tabular.jl:128 (Flux 0.11.6: eps = 1e-8, the
    running powers beta^t live in the per-parameter state)."""

    def __init__(self, weight_decay: float = 1.0e-5, eta: float = 1.0e-2, beta=(0.9, 0.999), eps: float = 1.0e-8, inv_decay: float = 0.0):
        """inv_decay > 0: Optimiser(InvDecay(inv_decay), ADAM(eta)) of experiments/mnist_nsde.jl:87 (use weight_decay = 0 there)."""
        self.weight_decay, self.eta, self.beta, self.eps, self.inv_decay = weight_decay, eta, (float(beta[0]), float(beta[1])), eps, inv_decay
        self.state: dict = {}

    def slot(self, p: torch.Tensor):
        k = p.data_ptr()
        if k not in self.state:
            self.state[k] = {"m": torch.zeros_like(p), "v": torch.zeros_like(p), "bp": [self.beta[0], self.beta[1]], "n": 1}
        return self.state[k]


def update_parameters_(ps, gs, opt) -> None:
    """update_parameters!(ps, gs, opt): in-place on raw arrays, skipping empty parameter vectors
    (src/utils.jl:149-156)."""
    lib = L.lib()
    for p, g in zip(ps, gs):
        if p.numel() == 0:
            continue
        if isinstance(opt, ADAMOptimiser):
            s = opt.slot(p)
            if opt.inv_decay > 0:          # InvDecay: delta / (1 + gamma n), n = 1-based update count (Flux 0.11.6)
                g = g * (1.0 / (1.0 + opt.inv_decay * s["n"]))
            s["n"] += 1
            rc = lib.rnde_adam_update(None, p.data_ptr(), g.data_ptr(), s["m"].data_ptr(), s["v"].data_ptr(), p.numel(), opt.eta, opt.beta[0],
                                      opt.beta[1], s["bp"][0], s["bp"][1], opt.eps, opt.weight_decay, _stream_ptr())
            if rc != L.OK:
                raise L.RndeError(rc, "rnde_adam_update")
            s["bp"][0] *= opt.beta[0]; s["bp"][1] *= opt.beta[1]
            continue
        s = opt.slot(p)
        scale = 1.0 / (1.0 + opt.gamma * s["n"])
        rc = lib.rnde_opt_update(None, p.data_ptr(), g.data_ptr(), s["v"].data_ptr(), p.numel(), scale, opt.eta, opt.rho, _stream_ptr())
        if rc != L.OK:
            raise L.RndeError(rc, "rnde_opt_update")
        s["n"] += 1
