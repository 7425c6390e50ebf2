"""gru_oracle.py -- CPU restatement (torch, autograd) of the Latent-ODE recognition RNN.  TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench legs may import it (see oracle/rnde_oracle.c header).
PARITY UNPINNED against Julia (no toolchain here); it follows the reference source line by line:

  LatentGRU(in_dim, h_dim, latent_dim)          /root/reference/experiments/latent_ode.jl:46-60
      update_gate = Chain(Dense(2L + 2I + 1, h, tanh), Dense(h, L, sigmoid))
      reset_gate  = Chain(Dense(2L + 2I + 1, h, tanh), Dense(h, L, sigmoid))
      new_state   = Chain(Dense(2L + 2I + 1, h, tanh), Dense(h, 2L))
  single_run(p, y_mean, y_std, x)               latent_ode.jl:64-90
  (p::LatentGRU)(x)  -- runs t = T:-1:1         latent_ode.jl:92-99

Parameters are one flat vector in Flux.destructure order: update_gate (W1,b1,W2,b2), reset_gate, new_state; every
W is out x in, column-major.  x is (2I+1, T, B): data, observation mask, delta-t row (time_series data format)."""
from __future__ import annotations

import numpy as np
import torch


def n_params(I: int, H: int, L: int) -> int:
    C = 2 * L + 2 * I + 1
    return 2 * (H * C + H + L * H + L) + (H * C + H + 2 * L * H + 2 * L)


def unpack(p, I, H, L):
    C = 2 * L + 2 * I + 1
    o = 0
    out = []
    for M2 in (L, L, 2 * L):
        W1 = p[o:o + H * C].reshape(C, H).T; o += H * C
        b1 = p[o:o + H]; o += H
        W2 = p[o:o + M2 * H].reshape(H, M2).T; o += M2 * H
        b2 = p[o:o + M2]; o += M2
        out.append((W1, b1, W2, b2))
    return out


def glorot_params(rng: np.random.Generator, I: int, H: int, L: int, dtype=np.float32, bias_scale: float = 0.0) -> np.ndarray:
    C = 2 * L + 2 * I + 1
    parts = []
    for M2 in (L, L, 2 * L):
        for (M, K) in ((H, C), (M2, H)):
            s = np.sqrt(6.0 / (M + K))
            parts.append(rng.uniform(-s, s, size=(M, K)).astype(dtype).flatten(order="F"))
            parts.append((bias_scale * rng.standard_normal(M)).astype(dtype))
    return np.concatenate(parts).astype(dtype)


def single_run(params, y_mean, y_std, x, L):
    (Wu1, bu1, Wu2, bu2), (Wr1, br1, Wr2, br2), (Wn1, bn1, Wn2, bn2) = params
    y_concat = torch.cat([y_mean, y_std, x], 0)                                   # latent_ode.jl:68
    update_gate = torch.sigmoid(Wu2 @ torch.tanh(Wu1 @ y_concat + bu1[:, None]) + bu2[:, None])
    reset_gate = torch.sigmoid(Wr2 @ torch.tanh(Wr1 @ y_concat + br1[:, None]) + br2[:, None])
    concat = torch.cat([y_mean * reset_gate, y_std * reset_gate, x], 0)           # :73
    new_state = Wn2 @ torch.tanh(Wn1 @ concat + bn1[:, None]) + bn2[:, None]
    new_state_mean, new_state_std = new_state[:L], new_state[L:]                  # :76-79
    new_y_mean = (1 - update_gate) * new_state_mean + update_gate * y_mean        # :81-82
    new_y_std = (1 - update_gate) * new_state_std + update_gate * y_std
    mask = (x[x.shape[0] // 2:].sum(0, keepdim=True) > 0).to(x.dtype)            # :84  x[(size(x,1)÷2+1):end, :]
    new_y_mean = mask * new_y_mean + (1 - mask) * y_mean                          # :86-87
    new_y_std = mask * new_y_std + (1 - mask) * y_std
    return new_y_mean, new_y_std


def forward(p: torch.Tensor, x: torch.Tensor, I: int, H: int, L: int) -> torch.Tensor:
    """x: (2I+1, T, B) -> vcat(y_mean, y_std): (2L, B); the sequence is consumed backwards in time (:95)."""
    params = unpack(p, I, H, L)
    B = x.shape[2]
    y_mean = torch.zeros(L, B, dtype=x.dtype)
    y_std = torch.zeros(L, B, dtype=x.dtype)
    for t in range(x.shape[1] - 1, -1, -1):
        y_mean, y_std = single_run(params, y_mean, y_std, x[:, t, :], L)
    return torch.cat([y_mean, y_std], 0)
