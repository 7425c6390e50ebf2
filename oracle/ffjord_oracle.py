"""CPU restatement of the reference's FFJORD path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED against Julia (no toolchain here, the reference holds no golden vectors for this path); pinned instead
against closed forms and exact Jacobians in tests/test_ffjord_oracle.py.

No CUDA path exists for this row yet: this file fixes WHAT the kernels will have to reproduce, so the
next round can start from a checked oracle.  Restated, with plain torch ops on the CPU so autograd differentiates through
every solver operation (the reference uses Tracker + SensitivityADPassThrough, src/models/ffjord.jl:72,117):

* experiments/ffjord_tabular.jl:39-45      numerically stable sigmoid / softplus
* experiments/ffjord_tabular.jl:47-74      ConcatSquashLinear: (W x + B) * sigmoid(G t) + (bW t + bB), and its x-VJP
* experiments/ffjord_tabular.jl:76-105     MLPDynamics(in, hidden): three ConcatSquash layers, softplus between; forw_n_back
                                           returns (f(z, t), e^T df/dz) without a second pass through the network
* src/models/ffjord.jl:53-66               _ffjord: d/dt [z; logdet (; ||f||^2; ||e^T J||^2)] = [f; -sum(eJ .* e) (; ...)]
* src/models/ffjord.jl:68-137              the two functors: Hutchinson noise e fixed per call, augmented state solved with
                                           Tsit5 at reltol = abstol = 1.4e-8, logp(x) = logN(z(1)) - delta_logp;
                                           {true}: SavingCallback of EEst*dt
* src/models/ffjord.jl:139-167             jacobian_fn / _deterministic_ffjord: exact trace (used by `sample`)
* experiments/ffjord_tabular.jl:143-147    loss = -mean(logpx) + lambda_r * mean(sv.saveval)

The time stepping itself is oracle/torch_oracle.py's solve (Tsit5, PI controller, initial dt: SURVEY.md Appendix A.1-A.8) on
the augmented state; the solver's RMS norms therefore run over all (D + 1) x B entries, as in the reference.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import torch_oracle as T


def sigmoid(x):
    """ffjord_tabular.jl:39-42"""
    t = torch.exp(-torch.abs(x))
    return torch.where(x >= 0, 1 / (1 + t), t / (1 + t))


def softplus(x):
    """ffjord_tabular.jl:44"""
    return torch.where(x > 0, x + torch.log1p(torch.exp(-torch.abs(x))), torch.log1p(torch.exp(-torch.abs(x))))


def layer_sizes(D: int, H: int):
    return [(D, H), (H, H), (H, D)]


def n_params(D: int, H: int) -> int:
    return sum(o * i + 4 * o for i, o in layer_sizes(D, H))


def unpack(p, D: int, H: int):
    """Flux.destructure order of MLPDynamics (ffjord_tabular.jl:49-55, :76-80): per layer layer_W (out x in, column-major),
    layer_B, bias_W, bias_B, gate_W (each out x 1)."""
    out, o = [], 0
    for i, m in layer_sizes(D, H):
        W = p[o:o + m * i].reshape(i, m).t(); o += m * i
        cols = []
        for _ in range(4):
            cols.append(p[o:o + m].reshape(m, 1)); o += m
        out.append((W, *cols))
    return out


def glorot_params(rng: np.random.Generator, D: int, H: int, dtype=np.float32, bias_scale: float = 0.0) -> np.ndarray:
    """ConcatSquashLinear(in, out) initialisation (ffjord_tabular.jl:57-63): glorot_uniform weights, zero biases
    (bias_scale > 0 randomises them so that tests exercise every term)."""
    parts = []
    for i, m in layer_sizes(D, H):
        s = math.sqrt(6.0 / (i + m))
        parts.append(rng.uniform(-s, s, size=(m, i)).flatten(order="F"))
        parts.append(bias_scale * rng.standard_normal(m))                    # layer_B
        s1 = math.sqrt(6.0 / (1 + m))
        parts.append(rng.uniform(-s1, s1, size=m))                           # bias_W
        parts.append(bias_scale * rng.standard_normal(m))                    # bias_B
        parts.append(rng.uniform(-s1, s1, size=m))                           # gate_W
    return np.concatenate(parts).astype(dtype)


def dynamics(p, z, t, D: int, H: int):
    """MLPDynamics(x, t) (ffjord_tabular.jl:87-90)."""
    x = z
    L = unpack(p, D, H)
    for n, (W, B, bW, bB, G) in enumerate(L):
        x = (W @ x + B) * sigmoid(G * t) + (bW * t + bB)
        if n < 2:
            x = softplus(x)
    return x


def forw_n_back(p, z, t, e, D: int, H: int):
    """(f(z, t), J^T e) the way the experiment computes it (ffjord_tabular.jl:65-72, :92-101): one forward pass keeping
    the gates, then the chain of transposed products -- no second traversal of the network."""
    (W1, B1, bW1, bB1, G1), (W2, B2, bW2, bB2, G2), (W3, B3, bW3, bB3, G3) = unpack(p, D, H)
    g1, g2, g3 = sigmoid(G1 * t), sigmoid(G2 * t), sigmoid(G3 * t)
    r1 = (W1 @ z + B1) * g1 + (bW1 * t + bB1)
    r2 = (W2 @ softplus(r1) + B2) * g2 + (bW2 * t + bB2)
    r3 = (W3 @ softplus(r2) + B3) * g3 + (bW3 * t + bB3)
    back = lambda W, g, v: (W * g).t() @ v
    eJ = back(W1, g1, sigmoid(r1) * back(W2, g2, sigmoid(r2) * back(W3, g3, e)))
    return r3, eJ


def ffjord_rhs(u, p, t, e, D: int, H: int, regularize: bool):
    """_ffjord (src/models/ffjord.jl:53-66)."""
    extra = 3 if regularize else 1
    z = u[: u.shape[0] - extra]
    mz, eJ = forw_n_back(p, z, t, e, D, H)
    rows = [mz, -(eJ * e).sum(0, keepdim=True)]
    if regularize:
        rows += [(mz * mz).sum(0, keepdim=True), (eJ * eJ).sum(0, keepdim=True)]       # norm_batched(eJ).^2 (utils.jl:25)
    return torch.cat(rows, 0)


def deterministic_rhs(u, p, t, D: int, H: int):
    """_deterministic_ffjord (src/models/ffjord.jl:153-159): exact Jacobian trace, one VJP per output row."""
    z = u[:-1].detach().requires_grad_(True)
    with torch.enable_grad():
        y = dynamics(p, z, t, D, H)
        tr = torch.zeros(1, z.shape[1], dtype=u.dtype)
        for i in range(D):
            gi, = torch.autograd.grad(y[i].sum(), z, retain_graph=True)
            tr = tr + gi[i:i + 1]
    return torch.cat([y.detach(), -tr.detach()], 0)


class FfjordResult:
    def __init__(self, logpx, lam1, lam2, nfe, saveval, sol):
        self.logpx, self.lam1, self.lam2, self.nfe, self.saveval, self.sol = logpx, lam1, lam2, nfe, saveval, sol


def ffjord(x, p, e, *, D: int, H: int, regularized_functor: bool, regularize: bool = False, t0=0.0, t1=1.0, abstol=1.4e-8,
           reltol=1.4e-8, **solve_kw) -> FfjordResult:
    """The two functors of TrackedFFJORD (src/models/ffjord.jl:68-137) -> (logpx, lambda1, lambda2, nfe, sv).
    regularized_functor = the type parameter R (SavingCallback of EEst*dt, one extra row); `regularize` = the keyword of the
    {false} functor that appends the two kinetic rows (||f||^2, ||e^T J||^2)."""
    B = x.shape[1]
    if regularized_functor:
        regularize = False                                                     # ffjord.jl:121 passes `false`
    extra = 3 if regularize else 1
    u0 = torch.cat([x, torch.zeros(extra, B, dtype=x.dtype)], 0)
    rhs = lambda u, t: ffjord_rhs(u, p, t, e, D, H, regularize)
    sol = T.solve(u0, p, D=D + extra, H=H, t0=t0, t1=t1, abstol=abstol, reltol=reltol, rhs=rhs,
                  reg_kind=T.REG_ERR_DT if regularized_functor else T.REG_NONE, **solve_kw)
    pred = sol.u
    z = pred[:D]
    delta_logp = pred[D]
    zero = torch.zeros(B, dtype=x.dtype)
    lam1, lam2 = (pred[D + 1], pred[D + 2]) if regularize else (zero, zero)
    logpz = (-(math.log(2 * math.pi) + z * z) / 2).sum(0)
    sv = torch.stack(sol.saveval) if regularized_functor else None
    return FfjordResult(logpz - delta_logp, lam1, lam2, sol.nf, sv, sol)


def loss_function(x, p, e, *, D: int, H: int, regularized_functor: bool, lam_r: float = 1.0e2, **kw):
    """experiments/ffjord_tabular.jl:143-147 -> (total, FfjordResult)."""
    r = ffjord(x, p, e, D=D, H=H, regularized_functor=regularized_functor, **kw)
    total = -r.logpx.mean()
    if regularized_functor:
        total = total + lam_r * r.saveval.mean()
    return total, r
