"""Latent-ODE model surface over libregnde.so.

    LatentGRU(in_dim, h_dim, latent_dim)                       experiments/latent_ode.jl:39-99
    LatentTimeSeriesModel(rnn, enc, node, dec)                 src/models/time_series.jl:1-70
    model(x, p1, p2, p3, p4; func, saveat) -> (result, mu0, logvar, nfe, sv)
    log_likelihood / kl_divergence / loss_function             experiments/latent_ode.jl:212-262

The two sequential hot loops run in the CUDA library: the recognition RNN (rnde_gru_forward / _backward, one
persistent kernel per direction) and the generator ODE solve with `saveat` (chain field, rnde_forward_saveat).
`rec_to_gen` (Dense 100->50->40), the reparametrisation sample, the decoder Dense(20, 37) and the masked Gaussian
likelihood are the host glue the reference keeps in Flux; here they are torch ops on the same device buffers."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import torch

from . import _lib as L
from .node import Chain, Dense, SaveFunc, TrackedNeuralODE, _stream_ptr


class _GruHandle:
    def __init__(self, cfg: L.GruConfig):
        L.require_device()
        self.lib = L.lib()
        self.cfg = cfg
        self.h = C.c_void_p()
        self.serial = 0          # taped forwards so far: the tape in the handle serves the latest one only
        rc = self.lib.rnde_gru_create(C.byref(cfg), C.byref(self.h))
        if rc != L.OK:
            raise L.RndeError(rc, f"rnde_gru_create(I={cfg.in_dim}, H={cfg.hidden_dim}, L={cfg.latent_dim}, B={cfg.batch}, T={cfg.seq_len})")

    def check(self, rc: int, what: str):
        if rc != L.OK:
            raise L.RndeError(rc, what + ": " + self.lib.rnde_gru_last_error(self.h).decode())

    def __del__(self):
        try:
            if self.h:
                self.lib.rnde_gru_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass


class _GruRun(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xbuf: torch.Tensor, p: torch.Tensor, hd: _GruHandle):
        cfg = hd.cfg
        out = torch.empty(2 * cfg.latent_dim * cfg.batch, device=xbuf.device, dtype=torch.float32)
        hd.check(hd.lib.rnde_gru_forward(hd.h, xbuf.data_ptr(), p.data_ptr(), out.data_ptr(), _stream_ptr()), "rnde_gru_forward")
        hd.serial += 1
        ctx.hd, ctx.p_ref, ctx.serial = hd, p, hd.serial
        return out

    @staticmethod
    def backward(ctx, dout: torch.Tensor):
        hd = ctx.hd
        if ctx.serial != hd.serial:
            raise RuntimeError("backward through a LatentGRU run whose tape was overwritten by a later forward of the same "
                               "batch size and sequence length; call backward before the next forward")
        dp = torch.empty(ctx.p_ref.numel(), device=dout.device, dtype=torch.float32)
        hd.check(hd.lib.rnde_gru_backward(hd.h, dout.contiguous().data_ptr(), dp.data_ptr(), _stream_ptr()), "rnde_gru_backward")
        return None, dp, None


class LatentGRU:
    """experiments/latent_ode.jl:39-60.  Parameters: Flux.destructure order (update_gate, reset_gate, new_state)."""

    def __init__(self, in_dim: int, h_dim: int, latent_dim: int, *, generator: Optional[torch.Generator] = None, device: str = "cuda"):
        self.in_dim, self.h_dim, self.latent_dim = in_dim, h_dim, latent_dim
        c = latent_dim * 2 + in_dim * 2 + 1
        nets = [(Dense(c, h_dim, "tanh", generator=generator), Dense(h_dim, m, None, generator=generator))
                for m in (latent_dim, latent_dim, 2 * latent_dim)]
        self.device = torch.device(device)
        self.p = torch.cat([l.destructure() for pair in nets for l in pair]).to(self.device)
        self._handles: dict = {}

    def _handle(self, B: int, T: int, need_backward: bool) -> _GruHandle:
        key = (B, T, need_backward)
        if key not in self._handles:
            cfg = L.GruConfig()
            cfg.struct_bytes = C.sizeof(L.GruConfig)
            cfg.in_dim, cfg.hidden_dim, cfg.latent_dim = self.in_dim, self.h_dim, self.latent_dim
            cfg.batch, cfg.seq_len, cfg.need_backward = B, T, 1 if need_backward else 0
            self._handles[key] = _GruHandle(cfg)
        return self._handles[key]

    def __call__(self, x: torch.Tensor, p: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x: (2*in_dim + 1, T, B) = vcat(data, mask, delta-t) -> vcat(y_mean, y_std): (2*latent_dim, B)."""
        p = self.p if p is None else p
        X, T, B = x.shape
        if X != 2 * self.in_dim + 1:
            raise ValueError(f"x must have {2 * self.in_dim + 1} rows (data, mask, time)")
        if not x.is_cuda or not p.is_cuda:
            raise RuntimeError("regneuralde.jl_b200 runs on CUDA tensors only (no CPU fallback)")
        if p.dtype != torch.float32 or p.dim() != 1 or p.numel() != self.p.numel():
            raise ValueError(f"p must be the flat Float32 parameter vector of length {self.p.numel()} (Flux.destructure order)")
        need_bwd = torch.is_grad_enabled() and p.requires_grad
        hd = self._handle(B, T, need_bwd)
        xbuf = x.detach().to(torch.float32).permute(2, 1, 0).contiguous().view(-1)       # Julia layout X x T x B
        out = _GruRun.apply(xbuf, p.contiguous(), hd)
        return out.view(B, 2 * self.latent_dim).t()

    def launch_count(self) -> int:
        return sum(int(h.lib.rnde_gru_launch_count(h.h)) for h in self._handles.values())


def _dense_chain(p: torch.Tensor, x: torch.Tensor, layers: Sequence[Dense]) -> torch.Tensor:
    """re(p)(x) for a Chain of Dense layers (host glue: rec_to_gen, gen_to_data)."""
    o = 0
    for lay in layers:
        W = p[o:o + lay.out * lay.inp].view(lay.inp, lay.out).t(); o += lay.out * lay.inp
        b = p[o:o + lay.out]; o += lay.out
        x = W @ x + b[:, None]
        if lay.act == L.ACT_TANH:
            x = torch.tanh(x)
    return x


class LatentTimeSeriesModel:
    """src/models/time_series.jl:1-70: rnn -> enc -> (mu0, logvar) -> z0 -> node(saveat) -> dec."""

    def __init__(self, rnn: LatentGRU, enc: Sequence[Dense], node: TrackedNeuralODE, dec: Sequence[Dense]):
        self.rnn, self.enc, self.node, self.dec = rnn, tuple(enc), node, tuple(dec)
        dev = rnn.device
        self.p1 = rnn.p
        self.p2 = torch.cat([l.destructure() for l in self.enc]).to(dev)
        self.p3 = node.p
        self.p4 = torch.cat([l.destructure() for l in self.dec]).to(dev)

    def trainable(self):
        return self.p1, self.p2, self.p3, self.p4

    def __call__(self, x: torch.Tensor, p1=None, p2=None, p3=None, p4=None, *, func: Optional[SaveFunc] = None, saveat=None,
                 sample: Optional[torch.Tensor] = None):
        """x: (2I+1, T, B).  `sample` replaces CUDA.randn(size(mu0)) (time_series.jl:47) so that tests are deterministic."""
        p1 = self.p1 if p1 is None else p1
        p2 = self.p2 if p2 is None else p2
        p3 = self.p3 if p3 is None else p3
        p4 = self.p4 if p4 is None else p4
        out = self.rnn(x, p1)
        out = _dense_chain(p2, out, self.enc)
        latent = out.shape[0] // 2
        mu0, logvar = out[:latent], out[latent:]
        if sample is None:
            sample = torch.randn_like(mu0)
        z0 = sample * torch.exp(logvar / 2) + mu0
        res, nfe, sv = self.node(z0, p3, func=func, saveat=saveat)          # feat x nsave x B
        D, S, B = res.shape
        result = _dense_chain(p4, res.reshape(D, S * B), self.dec).reshape(-1, S, B)
        return result, mu0, logvar, nfe, sv


def log_likelihood(dpred: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """experiments/latent_ode.jl:212-220 (sigma = 0.01; constants summed over all entries, normalised by sum(mask))."""
    s = 0.01
    sample = -dpred.pow(2) / (2 * s * s) - math.log(s) - math.log(2 * math.pi) / 2
    return (sample.sum(dim=(0, 1)) / mask.sum(dim=(0, 1))).reshape(-1)


def kl_divergence(mu: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
    """experiments/latent_ode.jl:223-224 (standard Gaussian prior)."""
    return ((torch.exp(logvar) + mu.pow(2) - 1 - logvar).mean(dim=0) / 2).reshape(-1)


def loss_function(data, mask, t_row, model: LatentTimeSeriesModel, p1, p2, p3, p4, *, func: Optional[SaveFunc], regularize: bool,
                  lam_r: float = 1.0e2, lam_k: float = 1.0, agg: str = "mean", saveat=None, sample=None):
    """experiments/latent_ode.jl:226-262 -> (total_loss, nfe, parts)."""
    x = torch.cat([data, mask, t_row], 0)
    result, mu0, logvar, nfe, sv = model(x, p1, p2, p3, p4, func=func, saveat=saveat, sample=sample)
    dpred = result * mask - data * mask
    ll = log_likelihood(dpred, mask)
    kl = lam_k * kl_divergence(mu0, logvar)
    reg = torch.zeros((), device=data.device)
    if regularize:
        if agg not in ("mean", "maximum", "sum"):
            raise ValueError("agg must be mean, maximum or sum")
        reg = lam_r * {"mean": torch.mean, "maximum": torch.max, "sum": torch.sum}[agg](sv.saveval)
    total = -(ll - kl).mean() + reg
    return total, nfe, {"nll": -ll.mean().detach(), "kl": kl.mean().detach(), "reg": reg.detach()}


def latent_ode_model(in_dim: int = 37, h_dim: int = 40, latent_dim: int = 50, gen_dim: int = 20, gen_hidden: int = 50, *, saveat,
                     regularize: bool, solver, generator: Optional[torch.Generator] = None, **node_kw) -> LatentTimeSeriesModel:
    """The model assembled at experiments/latent_ode.jl:105-150."""
    rnn = LatentGRU(in_dim, h_dim, latent_dim, generator=generator)
    enc = (Dense(2 * latent_dim, latent_dim, "tanh", generator=generator), Dense(latent_dim, 2 * gen_dim, None, generator=generator))
    layers = []
    for _ in range(4):
        layers += [Dense(gen_dim, gen_hidden, "tanh", generator=generator), Dense(gen_hidden, gen_dim, "tanh", generator=generator)]
    node = TrackedNeuralODE(Chain("tanh", *layers), [0.0, 1.0], False, regularize, solver, saveat=saveat, reltol=1.4e-8, abstol=1.4e-8, **node_kw)
    dec = (Dense(gen_dim, in_dim, None, generator=generator),)
    return LatentTimeSeriesModel(rnn, enc, node, dec)
