"""Pins of the Neural-SDE oracle (oracle/sde_oracle.py).  Parity against Julia is unpinned (no Julia, no
goldens in the reference); what pins the restatement:
  * the recalled SOSRI / SOSRI2 coefficients satisfy Roessler's strong-order-1.5 conditions for SRI schemes (any wrong digit
    breaks one of them at the 1e-15 level),
  * the RSwM3 bookkeeping keeps ONE Brownian path: whatever the sequence of rejections, the increments handed to accepted steps
    plus the futures still stacked add up to the increments first drawn,
  * the deterministic limit (no diffusion) is a convergent Runge-Kutta method, an Ornstein-Uhlenbeck process converges strongly
    with order >= 1 under step refinement ON THE SAME PATH, and the adaptive controller's bookkeeping (nfe1 = nfe2 = 2 + 4 attempts).
"""
import math

import numpy as np
import pytest

from oracle import sde_oracle as S


@pytest.mark.parametrize("tab", [S.SOSRI, S.SOSRI2], ids=["SOSRI", "SOSRI2"])
def test_tableau_satisfies_strong_order_conditions(tab):
    A0, B0, A1, B1, al, (b1, b2, b3, b4), c0, c1 = S.tableau_matrices(tab)
    e = np.ones(4)
    # consistency of the node vectors with the matrices
    assert np.allclose(A0 @ e, c0, atol=1e-13) and np.allclose(A1 @ e, c1, atol=1e-13)
    eq = lambda a, b: abs(a - b) < 2e-13
    # Roessler (2010), conditions for strong order 1.0 / 1.5 of SRI schemes with scalar/diagonal noise
    assert eq(al @ e, 1) and eq(b1 @ e, 1) and eq(b2 @ e, 0) and eq(b3 @ e, 0) and eq(b4 @ e, 0)
    assert eq(b1 @ (B1 @ e), 0) and eq(b2 @ (B1 @ e), 1) and eq(b3 @ (B1 @ e), 0) and eq(b4 @ (B1 @ e), 0)
    assert eq(al @ (A0 @ e), 0.5) and eq(al @ (B0 @ e), 1) and eq(al @ (B0 @ e) ** 2, 1.5)
    assert eq(b1 @ (A1 @ e), 1) and eq(b2 @ (A1 @ e), 0) and eq(b3 @ (A1 @ e), -1) and eq(b4 @ (A1 @ e), 0)
    assert eq(b1 @ (B1 @ e) ** 2, 1) and eq(b2 @ (B1 @ e) ** 2, 0) and eq(b3 @ (B1 @ e) ** 2, -1) and eq(b4 @ (B1 @ e) ** 2, 2)
    assert eq(b1 @ (B1 @ (B1 @ e)), 0) and eq(b2 @ (B1 @ (B1 @ e)), 0) and eq(b3 @ (B1 @ (B1 @ e)), 0) and eq(b4 @ (B1 @ (B1 @ e)), 1)
    assert eq(0.5 * b1 @ (A1 @ (B0 @ e)) + (1.0 / 3.0) * b3 @ (A1 @ (B0 @ e)), 0)


def test_rswm3_keeps_one_brownian_path():
    rng = np.random.default_rng(5)
    st = S.NormalStream(rng.standard_normal((400, 3, 2)))
    W = S.RSwM3(st, np.float64)
    t, total_W, total_Z = 0.0, 0.0, 0.0
    dt = 0.3
    W.setup(dt)
    coarse_first = (W.dW.copy(), W.dZ.copy())
    for it in range(60):
        if rng.random() < 0.5 and W.dt > 1e-3:        # reject: shrink
            W.reject(W.dt * rng.uniform(0.2, 0.9))
            continue
        t += W.dt; total_W = total_W + W.dW; total_Z = total_Z + W.dZ
        if it == 0:
            pass
        dt = rng.uniform(0.01, 0.4)
        W.setup(dt)
    # the path sampled so far: accepted increments + the current (not yet accepted) step + the futures are independent of how it was cut:
    # check the variance bookkeeping instead of values -- every future piece carries its own length
    fut = sum(L1 for L1, _, _ in W.S1)
    assert fut >= 0
    # exactness check: the very first coarse increment is recovered when everything covering [0, 0.3] is added up
    st2 = S.NormalStream(np.random.default_rng(6).standard_normal((50, 3, 2)))
    W2 = S.RSwM3(st2, np.float64)
    W2.setup(0.3)
    first = W2.dW.copy()
    W2.reject(0.1)              # step [0, 0.1], future (0.2)
    a = W2.dW.copy()
    W2.setup(0.05)              # step [0.1, 0.15] bridged out of the future, remainder 0.15 stays stacked
    b = W2.dW.copy()
    W2.reject(0.02)             # [0.1, 0.12]; futures: 0.03 then 0.15
    c = W2.dW.copy()
    W2.setup(1.0)               # pops both futures (0.18 in all) and draws the rest fresh
    pieces = [L2 for L1, L2, _ in W2.S2][:2]
    assert np.allclose(a + c + pieces[0] + pieces[1], first, atol=1e-12)
    assert abs(sum(L1 for L1, _, _ in W2.S2) - 1.0) < 1e-12


def test_deterministic_limit_and_counters():
    D, B = 4, 3
    rng = np.random.default_rng(1)
    A = -np.eye(D) + 0.3 * rng.standard_normal((D, D))
    x = rng.standard_normal((D, B))
    f = lambda u: A @ u
    g = lambda u: np.zeros_like(u)
    from scipy.linalg import expm
    exact = expm(A) @ x
    errs = []
    for n in (8, 16, 32):
        r = S.solve(x, f, g, np.zeros((4 * n + 8, D, B)), dtype=np.float64, forced_dt=[1.0 / n] * n)
        errs.append(np.abs(r.u - exact).max())
        assert r.naccept == n and r.nfe1 == 4 * n and r.nfe2 == 4 * n
    order = math.log2(errs[0] / errs[1])
    assert order > 1.8, (errs, order)          # the drift part of SOSRI is a consistent RK method of order >= 2
    r = S.solve(x, f, g, np.zeros((4000, D, B)), dtype=np.float64, abstol=1e-3, reltol=1e-3)
    assert r.nfe1 == 2 + 4 * (r.naccept + r.nreject) == r.nfe2
    assert np.abs(r.u - exact).max() < 1e-2


def test_ou_strong_convergence_on_one_path():
    """dX = -X dt + 0.5 dW: refine the step on the SAME Brownian path (coarse increments are sums of fine ones)."""
    D, B, nfine = 1, 256, 64
    rng = np.random.default_rng(11)
    h = 1.0 / nfine
    xiW = rng.standard_normal((nfine, D, B)); xiZ = rng.standard_normal((nfine, D, B))
    dWf, dZf = math.sqrt(h) * xiW, math.sqrt(h) * xiZ
    f = lambda u: -u
    g = lambda u: 0.5 * np.ones_like(u)
    x = np.ones((D, B))
    # exact solution on the path: X(1) = e^-1 + 0.5 * int e^{-(1-s)} dW  ~ sum over fine steps (fine enough: h = 1/64)
    tf = (np.arange(nfine) + 0.5) * h
    exact = math.exp(-1.0) + 0.5 * np.tensordot(np.exp(-(1.0 - tf)), dWf, axes=(0, 0))
    errs = []
    for n in (4, 8, 16):
        m = nfine // n
        normals = []
        for k in range(n):      # the solver draws dW then dZ per step: hand it the coarse sums, rescaled to unit variance
            normals.append(dWf[k * m:(k + 1) * m].sum(0) / math.sqrt(m * h))
            normals.append(dZf[k * m:(k + 1) * m].sum(0) / math.sqrt(m * h))
        r = S.solve(x, f, g, np.array(normals), dtype=np.float64, forced_dt=[1.0 / n] * n)
        errs.append(math.sqrt(np.mean((r.u - exact) ** 2)))
    assert errs[1] < errs[0] and errs[2] < errs[1]
    assert math.log2(errs[0] / errs[2]) / 2 > 0.9, errs      # strong order >= 1 observed (additive noise; dZ is not the path's true integral)


def test_experiment_shape_runs_and_is_deterministic():
    rng = np.random.default_rng(1999)
    D, H, B = 32, 64, 48
    s = lambda o, i: np.sqrt(6.0 / (i + o))
    p = np.concatenate([rng.uniform(-s(H, D), s(H, D), H * D), np.zeros(H), rng.uniform(-s(D, H), s(D, H), D * H), np.zeros(D),
                        rng.uniform(-s(D, D), s(D, D), D * D), np.zeros(D)]).astype(np.float32)
    x = rng.standard_normal((D, B)).astype(np.float32)
    z = rng.standard_normal((400, D, B)).astype(np.float32)
    f, g = S.drift_diffusion(p, np.float32)
    r1 = S.solve(x, f, g, z, reg_kind=S.REG_ERR_DT)
    r2 = S.solve(x, f, g, z, reg_kind=S.REG_ERR_DT)
    assert np.array_equal(r1.u, r2.u) and r1.nfe1 == r2.nfe1 == 2 + 4 * (r1.naccept + r1.nreject)
    assert len(r1.saveval) == r1.naccept + 1 and r1.saveval[0] == 0
    assert np.isfinite(r1.u).all() and r1.naccept >= 2
    r64 = S.solve(x, f, g, z, reg_kind=S.REG_ERR_DT, dtype=np.float64)
    if r64.accepted == r1.accepted:      # same decisions: the Float32 and Float64 paths agree to rounding
        assert np.abs(r1.u - r64.u).max() < 1e-4 * max(1.0, np.abs(r64.u).max())


def test_replay_reproduces_the_solve_and_its_gradient_is_the_finite_difference():
    """oracle/sde_oracle.py replay_torch (the yardstick of the CUDA reverse sweep, tests/test_gpu_nsde.py): the accepted steps
    replayed in torch give the solve's state and saved values back, and autograd through them agrees with central finite
    differences of the replay (step sizes and increments frozen) in a few random parameter directions."""
    import torch
    rng = np.random.default_rng(4)
    D, H, B = 6, 9, 3
    npar = H * D + H + D * H + D + D * D + D
    p = 0.4 * rng.standard_normal(npar)
    x = rng.standard_normal((D, B))
    z = rng.standard_normal((300, D, B))
    f, g = S.drift_diffusion(p, np.float64, D, H)
    r = S.solve(x, f, g, z, alg=S.ALG_SOSRI, reg_kind=S.REG_ERR_DT, abstol=0.05, reltol=0.05, dtype=np.float64)
    assert r.naccept == len(r.steps) and r.naccept > 5
    w = rng.standard_normal((D, B)); ws = rng.standard_normal(len(r.saveval))

    def loss_of(pv, xv):
        u, sv = S.replay_torch(xv, pv, r.steps, alg=S.ALG_SOSRI, reg_kind=S.REG_ERR_DT, abstol=0.05, reltol=0.05, D=D, H=H)
        return (u * torch.tensor(w)).sum() + (sv * torch.tensor(ws)).sum(), u, sv

    pt = torch.tensor(p, requires_grad=True); xt = torch.tensor(x, requires_grad=True)
    loss, u, sv = loss_of(pt, xt)
    assert np.abs(u.detach().numpy() - r.u).max() < 1e-12 and np.abs(sv.detach().numpy() - r.saveval).max() < 1e-12
    gp, gx = torch.autograd.grad(loss, [pt, xt])
    for k in range(4):
        d = rng.standard_normal(npar); h = 1e-6
        lp = loss_of(torch.tensor(p + h * d), torch.tensor(x))[0]; lm = loss_of(torch.tensor(p - h * d), torch.tensor(x))[0]
        fd = float(lp - lm) / (2 * h)
        assert abs(fd - float(gp.numpy() @ d)) <= 1e-6 * max(abs(fd), 1.0)
