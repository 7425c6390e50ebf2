"""CPU checks of the Latent-ODE recognition RNN restatement (oracle/gru_oracle.py, following
experiments/latent_ode.jl:39-99): parameter layout, mask gating, reverse-time order, gradient by finite differences."""
import numpy as np
import torch

from oracle import gru_oracle as G


def _inputs(rng, I, T, B):
    x = rng.standard_normal((2 * I + 1, T, B))
    x[I:2 * I] = rng.random((I, T, B)) < 0.3
    x[2 * I] = rng.random((T, B)) * 0.1
    return x


def test_param_count_matches_flux_destructure():
    # LatentGRU(37, 40, 50): three Chains Dense(175,40), Dense(40,50|50|100)        latent_ode.jl:46-60,105
    assert G.n_params(37, 40, 50) == 2 * (175 * 40 + 40 + 40 * 50 + 50) + (175 * 40 + 40 + 40 * 100 + 100) == 29320
    assert G.glorot_params(np.random.default_rng(0), 37, 40, 50).size == 29320


def test_unobserved_steps_leave_the_state_unchanged_and_order_is_reverse_time():
    rng = np.random.default_rng(1)
    I, H, L, T, B = 3, 5, 4, 6, 2
    p = torch.tensor(G.glorot_params(rng, I, H, L, dtype=np.float64, bias_scale=0.1))
    x = _inputs(rng, I, T, B)
    full = G.forward(p, torch.tensor(x), I, H, L)
    x2 = np.concatenate([x[:, :2], np.zeros((2 * I + 1, 1, B)), x[:, 2:]], 1)     # insert a step with mask = dt = 0
    assert torch.allclose(G.forward(p, torch.tensor(x2), I, H, L), full, atol=1e-14)
    # the sequence is consumed from the last time index to the first (latent_ode.jl:95)
    params = G.unpack(p, I, H, L)
    ym = torch.zeros(L, B, dtype=torch.float64); ys = torch.zeros(L, B, dtype=torch.float64)
    for t in range(T - 1, -1, -1):
        ym, ys = G.single_run(params, ym, ys, torch.tensor(x[:, t, :]), L)
    assert torch.equal(torch.cat([ym, ys], 0), full)


def test_gradient_finite_differences():
    rng = np.random.default_rng(2)
    I, H, L, T, B = 2, 3, 2, 4, 3
    p0 = G.glorot_params(rng, I, H, L, dtype=np.float64, bias_scale=0.1)
    x = torch.tensor(_inputs(rng, I, T, B))
    w = torch.tensor(rng.standard_normal((2 * L, B)))
    f = lambda pv: float((G.forward(torch.tensor(pv), x, I, H, L) * w).sum())
    p = torch.tensor(p0, requires_grad=True)
    (G.forward(p, x, I, H, L) * w).sum().backward()
    for i in rng.choice(p0.size, 12, replace=False):
        e = np.zeros_like(p0); e[i] = 1e-6
        fd = (f(p0 + e) - f(p0 - e)) / 2e-6
        assert abs(fd - float(p.grad[i])) <= 1e-6 * max(1.0, abs(fd))
